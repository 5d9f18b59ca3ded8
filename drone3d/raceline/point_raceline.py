''' drone3d/raceline/point_raceline.py of the reference '''
from aircraft_trajectory_optimization_b200.raceline import GlobalPointRaceline, ParametricPointRaceline, \
    ParametricObstaclePointRaceline                                   # noqa: F401
