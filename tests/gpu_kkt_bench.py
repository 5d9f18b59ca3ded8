''' ad-hoc GPU timing of the KKT kernels on the C2 structure (not a pytest file) '''
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from cases import build_product, eval_point
from aircraft_trajectory_optimization_b200.kkt import KktSolver
from aircraft_trajectory_optimization_b200.models import vehicle_params

name = sys.argv[1] if len(sys.argv) > 1 else 'race_param_rk4_drone'
Bs = [int(b) for b in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 64, 256, 512]
prod = build_product(name)
st, F = prod.structure, prod.functions
K = KktSolver(st, condensed=False) if os.environ.get('RB_KKT_BIG') else KktSolver(st)
ks = K.ks
if K.cs is not None:
    print('condensed: NI', K.cs.NI, 'amax', K.cs.amax, 'smax', K.cs.smax)
print(name, 'N', ks.N, 'bmax', ks.bmax, 'nb', ks.nb, 'mmax', ks.mmax, 'qmax', ks.qmax)
dev = torch.device('cuda', 0)
x, lam = eval_point(st, 0)
out = F.eval(x, lam_f=1.0, lam_g=lam, want=('jac', 'hess'))
rng = np.random.default_rng(0)
for B in Bs:
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, (B,) + a.shape))).to(dev).contiguous()
    th, tj = t(out['hess']), t(out['jac'])
    tdx = t(1.0 + rng.uniform(0, 1, st.nw)); tnd = t(-np.where(st.lbg == st.ubg, 0.0, rng.uniform(0.1, 1, st.ng)))
    tr = t(rng.standard_normal(st.nw + st.ng))
    def timeit(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    tf = timeit(lambda: K.factor_solve(th, tj, tdx, tnd, tr))
    sol, _ = K.factor_solve(th, tj, tdx, tnd, tr)
    ts = timeit(lambda: K.resolve(th, tj, tdx, tnd, tr))
    tm = timeit(lambda: K.matvec(th, tj, tdx, tnd, sol))
    print(f'B={B}: factor_solve {tf:.2f} ms, resolve {ts:.2f} ms, matvec {tm:.3f} ms')
