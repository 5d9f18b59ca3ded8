''' ad-hoc GPU debug: per-block inverses of the factorisation kernel vs the numpy twin (not a pytest file) '''
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from cases import build_product, eval_point
from aircraft_trajectory_optimization_b200.kkt import KktSolver
from oracle.kkt_blocks_ref import _gather, sym_invert_bp

name = sys.argv[1] if len(sys.argv) > 1 else 'race_param_rk4_point'
prod = build_product(name, small=True)
st, F = prod.structure, prod.functions
K = KktSolver(st); ks = K.ks
x, lam = eval_point(st, 0)
out = F.eval(x, lam_f=1.0, lam_g=lam, want=('jac', 'hess'))
rng = np.random.default_rng(0)
dxd = 1.0 + rng.uniform(0, 1, st.nw); D = np.where(st.lbg == st.ubg, 0.0, rng.uniform(0.1, 1, st.ng)); rhs = rng.standard_normal(st.nw + st.ng)
dev = torch.device('cuda', 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a[None])).to(dev)
sol, status = K.factor_solve(t(out['hess']), t(out['jac']), t(dxd), t(-D), t(rhs))
torch.cuda.synchronize()
print('status', status.cpu().numpy(), 'want neg', st.ng)
fac = K._factors.cpu().numpy()
N, bmax = ks.N, ks.bmax
Sinv_g = fac[:N * bmax * bmax].reshape(N, bmax, bmax)
hess, jac, neg_d = out['hess'], out['jac'], -D
nw = st.nw
carry = None
for n in range(N):
    u = ks.unk[ks.blk_ptr[n]:ks.blk_ptr[n + 1]]; b = len(u)
    M = np.zeros((bmax, bmax)); a, e = ks.dA_ptr[n], ks.dA_ptr[n + 1]
    M.ravel()[ks.dA_pos[a:e]] = _gather(ks, ks.dA_src[a:e], hess, jac, dxd, neg_d)
    M[np.arange(b), np.arange(b)] += np.where(u < nw, dxd[np.minimum(u, nw - 1)], neg_d[np.maximum(u - nw, 0)])
    if n > 0:
        cr = ks.cr[ks.cr_ptr[n - 1]:ks.cr_ptr[n]]
        M[np.ix_(cr, cr)] -= carry
    A = M[:b, :b]
    G = Sinv_g[n, :b, :b]
    err = np.abs(G @ A - np.eye(b)).max()
    inv, neg = sym_invert_bp(A)
    print(n, 'b', b, 'gpu inv err', f'{err:.2e}', 'twin err', f'{np.abs(inv @ A - np.eye(b)).max():.2e}', 'nan' if np.isnan(G).any() else '')
    if err > 1e-4:
        bad = np.argwhere(np.abs(G @ A - np.eye(b)) > 1e-4)
        print('  bad entries (row, col) of G A - I:', bad[:10].tolist(), 'diag zero idx', np.nonzero(np.diag(A) == 0)[0].tolist())
        break
    if n < N - 1:
        cc = ks.cc[ks.cc_ptr[n]:ks.cc_ptr[n + 1]]; m = ks.cr_ptr[n + 1] - ks.cr_ptr[n]
        Lc = np.zeros((max(m, 1), ks.qmax)); a, e = ks.cL_ptr[n], ks.cL_ptr[n + 1]
        Lc.ravel()[ks.cL_pos[a:e]] = _gather(ks, ks.cL_src[a:e], hess, jac, dxd, neg_d)
        Lc = Lc[:m, :len(cc)]
        carry = Lc @ (G[:, cc] @ Lc.T)[cc]
