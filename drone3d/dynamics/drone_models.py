''' drone3d/dynamics/drone_models.py of the reference '''
from aircraft_trajectory_optimization_b200.models import DroneModel, ParametricDroneModel  # noqa: F401
