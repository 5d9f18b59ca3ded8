''' drone3d/centerlines/base_centerline.py of the reference '''
from aircraft_trajectory_optimization_b200.centerlines import GateShape, BaseCenterlineConfig, BaseCenterline  # noqa: F401
