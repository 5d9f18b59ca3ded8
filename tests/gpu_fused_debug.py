''' GPU experiment (not collected by pytest): per-instance comparison of the torch / fused sweeps (test_gpu_ipm) '''
import sys
import numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from cases import build_product   # noqa: E402
from aircraft_trajectory_optimization_b200.ipm import IpmOptions   # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else 'race_param_rk4_point'
prod = build_product(name, N=7)
st = prod.structure
prod.solver.verbose = False
rng = np.random.default_rng(2)
B = 6
X0 = np.tile(st.w0, (B, 1))
X0[1:] += 0.03 * rng.standard_normal((B - 1, st.nw)) * (np.abs(st.w0) > 0)
X0 = np.clip(X0, st.lbw, st.ubw)
kw = dict(lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg)
out = {}
for key, opts in (('torch', IpmOptions(use_glue=False, compact=False)), ('glue', IpmOptions(use_glue=True, compact=False))):
    prod.solver.options = opts
    sol = prod.solver(x0=X0, **kw)
    out[key] = (sol, prod.solver.result.status.cpu().numpy().copy(), prod.solver.result.iterations.cpu().numpy().copy())
    print(key, 'status', out[key][1], 'iters', out[key][2], 'laps', sol['x'][:, :st.N].sum(1))
for k in ('x', 'lam_g'):
    d = np.abs(out['glue'][0][k] - out['torch'][0][k]).max(1)
    print(k, 'max diff per instance', d, 'scale', np.abs(out['torch'][0][k]).max())
