''' drone3d/utils/discretization_utils.py of the reference '''
from aircraft_trajectory_optimization_b200.collocation import get_collocation_coefficients, \
    get_intermediate_collocation_coefficients                         # noqa: F401
from aircraft_trajectory_optimization_b200.interpolation import *    # noqa: F401,F403
