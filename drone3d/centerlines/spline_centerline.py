''' drone3d/centerlines/spline_centerline.py of the reference '''
from aircraft_trajectory_optimization_b200.centerlines import SplineCenterline, SplineCenterlineConfig, \
    SplineRyFitOptions                                                # noqa: F401
