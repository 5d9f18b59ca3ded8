'''
racetrack racelines, global frame and curvilinear (non-Euclidean) frame, RK4 shooting -- headless mirror of
the reference's scripts/race.py:12-49 (waypoints, SQUARE gates, N = 70, use_rk4, quaternion, warm start).
The comparison against the external CPC csv and the viewer are out of scope.
'''
import numpy as np

from _common import print_table
from aircraft_trajectory_optimization_b200.centerlines import GateShape, SplineCenterline, SplineCenterlineConfig
from aircraft_trajectory_optimization_b200.solve_util import solve_util


def _main(verbose=False):
    x = np.array([-1.1, 9.2, 9.2, -4.5, -4.5, 4.75, -2.8])
    y = np.array([-1.6, 6.6, -4, -6, -6, -0.9, 6.8])
    z = np.array([3.6, 1.0, 1.2, 3.5, 0.8, 1.2, 1.2])
    config = SplineCenterlineConfig(x=np.array([x, y, z]))
    config.closed = True
    config.gate_shape = GateShape.SQUARE
    line = SplineCenterline(config)
    N = len(x) * 10
    global_solver, global_raceline = solve_util(line=line, global_frame=True, drone=True, use_ws=True,
                                                use_quaternion=True, use_rk4=True, N=N, verbose=verbose)
    parametric_solver, parametric_raceline = solve_util(line=line, global_frame=False, drone=True, use_ws=True,
                                                        use_quaternion=True, use_rk4=True, N=N, verbose=verbose)
    solvers = [global_solver, parametric_solver, global_solver.ws_solver, parametric_solver.ws_solver]
    results = [global_raceline, parametric_raceline, global_solver.ws_raceline, parametric_solver.ws_raceline]
    print_table(solvers, results)
    return solvers, results


if __name__ == '__main__':
    import sys
    _main(verbose='-v' in sys.argv)
