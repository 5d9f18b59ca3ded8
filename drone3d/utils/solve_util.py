''' drone3d/utils/solve_util.py of the reference '''
from aircraft_trajectory_optimization_b200.solve_util import solve_util  # noqa: F401
