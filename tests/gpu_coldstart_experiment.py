''' ad-hoc: the quaternion cold start of scripts/fig_8.py (reference scripts/fig_8.py:21-29) with the iteration log (not a pytest file) '''
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from cases import make_line
from aircraft_trajectory_optimization_b200.solve_util import solve_util
line = make_line('fig8')
quat = '--euler' not in sys.argv
t0 = time.time()
solver, res = solve_util(line=line, global_frame=False, drone=True, use_quaternion=quat, global_r=True, use_ws=False,
                         N=int(os.environ.get('N', '50')), verbose=True)
s = solver.solver.stats()
print('lap', res.time, s['return_status'], 'iters', s['iter_count'], f'wall {time.time() - t0:.1f}s')
