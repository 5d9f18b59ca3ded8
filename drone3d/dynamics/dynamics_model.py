''' drone3d/dynamics/dynamics_model.py of the reference '''
from aircraft_trajectory_optimization_b200.models import DynamicsModel  # noqa: F401
