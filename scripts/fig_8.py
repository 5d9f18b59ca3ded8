'''
racelines through a loop to demonstrate warmstart and nonconvexity -- headless mirror of the reference's
scripts/fig_8.py (same waypoints, gate shape, N and the same four solve_util calls; the OpenGL window is
out of scope).  Runs on the GPU through the B200 path.
'''
import numpy as np

from _common import print_table
from aircraft_trajectory_optimization_b200.centerlines import GateShape, SplineCenterline, SplineCenterlineConfig
from aircraft_trajectory_optimization_b200.solve_util import solve_util


def _main(verbose=False, skip_coldstarts=False):
    x = np.array([0, 5, 0, -5, 0, 5, 0, -5])
    y = np.array([0, 1, 2, 1, 0, -1, -2, -1])
    z = np.array([10, 5, 0, -5, -10, -5, 0, 5])
    N = 50
    config = SplineCenterlineConfig(x=np.array([x, y, z], dtype=float))
    config.closed = True
    config.gate_shape = GateShape.CIRCLE
    line = SplineCenterline(config)
    results, solvers = [], []

    solver, raceline = solve_util(line=line, global_frame=False, drone=True, use_quaternion=True, global_r=True,
                                  use_ws=True, N=N, verbose=verbose)
    raceline.label = 'Drone with warmstart'
    solver.ws_raceline.label = 'Point Mass WS'
    results += [raceline, solver.ws_raceline]
    solvers += [solver, solver.ws_solver]

    if not skip_coldstarts:
        baseline_solver, baseline_raceline = solve_util(line=line, global_frame=False, drone=True, use_quaternion=True,
                                                        global_r=True, use_ws=False, N=N, verbose=verbose)
        baseline_raceline.label = 'Drone coldstart'
        baseline_solver_euler, baseline_raceline_euler = solve_util(line=line, global_frame=False, drone=True,
                                                                    use_quaternion=False, global_r=True, use_ws=False,
                                                                    N=N, verbose=verbose)
        baseline_raceline_euler.label = 'Drone coldstart (Euler)'
        results += [baseline_raceline, baseline_raceline_euler]
        solvers += [baseline_solver, baseline_solver_euler]

    global_solver, global_raceline = solve_util(line=line, global_frame=True, drone=True, use_quaternion=True,
                                                global_r=True, use_ws=True, N=N, verbose=verbose)
    global_raceline.label = 'Global Drone with ws'
    results.append(global_raceline)
    solvers.append(global_solver)
    print_table(solvers, results)
    return solvers, results


if __name__ == '__main__':
    import sys
    _main(verbose='-v' in sys.argv, skip_coldstarts='--skip-coldstarts' in sys.argv)
