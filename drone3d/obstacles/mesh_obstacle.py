''' drone3d/obstacles/mesh_obstacle.py of the reference '''
from aircraft_trajectory_optimization_b200.obstacles import MeshObstacle               # noqa: F401
from aircraft_trajectory_optimization_b200.raceline import ObstacleFreeTube            # noqa: F401
