''' drone3d/dynamics/point_model.py of the reference '''
from aircraft_trajectory_optimization_b200.models import PointModel, ParametricPointModel  # noqa: F401
