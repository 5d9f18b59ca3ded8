'''
Stand-in for the reference's scripts/fig_8_cpc.py.  The reference script only *displays* a time-optimal
trajectory produced by an external complementary-progress-constraint (CPC) solver
(drone3d/utils/cpc_utils.py reads assets/cpc_warmstart_raceline.csv); it holds no CPC formulation to rebuild.
What this script does instead ("stand-in, no parity claim"): the same figure-eight
track with the same narrow gates (gate_ri = 0.6, gate_ro = 0.8, scripts/fig_8_cpc.py:18-19), solved by the
parametric collocation raceline with many intervals (N = 200, K = 7: nw = 33 800), on the GPU path.
'''
import sys

import numpy as np

from _common import print_table
from aircraft_trajectory_optimization_b200.centerlines import GateShape, SplineCenterline, SplineCenterlineConfig
from aircraft_trajectory_optimization_b200.solve_util import solve_util


def _main(N=200, verbose=False):
    x = np.array([0, 5, 0, -5, 0, 5, 0, -5])
    y = np.array([0, 1, 2, 1, 0, -1, -2, -1])
    z = np.array([10, 5, 0, -5, -10, -5, 0, 5])
    config = SplineCenterlineConfig(x=np.array([x, y, z], dtype=float))
    config.closed = True
    config.gate_shape = GateShape.CIRCLE
    config.gate_ri = 0.6
    config.gate_ro = 0.8
    line = SplineCenterline(config)
    solver, raceline = solve_util(line=line, global_frame=False, drone=True, use_quaternion=True, global_r=True,
                                  use_ws=True, N=N, verbose=verbose)
    raceline.label = f'Narrow-gate fig-8 N={N}'
    solver.ws_raceline.label = 'Point Mass WS'
    print_table([solver, solver.ws_solver], [raceline, solver.ws_raceline])
    return solver, raceline


if __name__ == '__main__':
    _main(N=int(sys.argv[1]) if len(sys.argv) > 1 else 200)
