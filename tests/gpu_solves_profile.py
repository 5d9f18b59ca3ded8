''' ad-hoc: host-side profile of the batched solves leg of bench.py (not a pytest file) '''
import sys, time, os, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from aircraft_trajectory_optimization_b200.models import vehicle_params
from aircraft_trajectory_optimization_b200.ipm import IpmOptions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
window = int(sys.argv[2]) if len(sys.argv) > 2 else 888
refine = int(sys.argv[3]) if len(sys.argv) > 3 else 4
prod = bench.build_c2_with_warm_start()
st = prod.structure
X0, VP = bench.multistart_inputs(st, vehicle_params(prod.vehicle_config), B, seed0=0)
prod.solver.verbose = False
prod.solver.max_iter = 300
prod.solver.options = IpmOptions(window=window, refine_steps=refine)
pr = cProfile.Profile()
torch.cuda.synchronize()
t0 = time.time()
pr.enable()
sol = prod.solver(x0=X0, lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg, p=VP)
torch.cuda.synchronize()
pr.disable()
dt = time.time() - t0
s = prod.solver.stats()
print('B', B, 'window', window, 'refine', refine, 'wall %.1fs' % dt, 'converged', int(s['success_each'].sum()), 'solves/s %.1f' % (s['success_each'].sum() / dt),
      't_eval %.1f t_kkt %.1f' % (s['t_wall_nlp_hess_l'], s['t_wall_linear_solver']), 'sweeps', prod.solver.result.n_iter, 'n_factor', s['n_factor'])
o = io.StringIO()
pstats.Stats(pr, stream=o).sort_stats('tottime').print_stats(30)
print(o.getvalue()[:7000])
