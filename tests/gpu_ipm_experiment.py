''' ad-hoc GPU experiment: interior-point solves through the product API (not a pytest file) '''
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from cases import build_product

name = sys.argv[1] if len(sys.argv) > 1 else 'race_param_rk4_point'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 7
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
verbose = B == 1
prod = build_product(name, N=N)
prod.solver.verbose = verbose
st = prod.structure
t0 = time.time()
if B == 1:
    res = prod.solve()
    print('lap', res.time, 'feasible', res.feasible, 'solve_time', prod.solve_time, 'stats', {k: v for k, v in prod.solver.stats().items() if not k.endswith('_each')})
else:
    rng = np.random.default_rng(0)
    X0 = st.w0[None, :] + 0.02 * rng.standard_normal((B, st.nw)) * (np.arange(B)[:, None] > 0)
    sol = prod.solver(x0=X0, lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg)
    s = prod.solver.stats()
    laps = sol['x'][:, :prod.config.N].sum(1)
    print('B', B, 'time', time.time() - t0, 'success', int(s['success_each'].sum()), 'iters', s['iterations_each'].min(), s['iterations_each'].max(),
          'laps', np.round(np.sort(laps)[:8], 5), 't_eval', s['t_wall_nlp_hess_l'], 't_kkt', s['t_wall_linear_solver'], 'n_factor', s['n_factor'], 'n_eval', s['n_eval'])
