import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from cases import build_product
from test_gpu_kkt import _inputs, KKT_CASES
from oracle.kkt_blocks_ref import sparse_solve
from aircraft_trajectory_optimization_b200.kkt import KktSolver
dev = torch.device('cuda', 0)
for name in KKT_CASES:
    name, _, variant = name.partition('@')
    prod = build_product(name, small=True)
    st, F = prod.structure, prod.functions
    B = 3
    hess, jac, dxd, D, rhs = _inputs(st, F, B)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    K = KktSolver(st, condensed=False) if variant == 'big' else KktSolver(st)
    sol, status = K.factor_solve(t(hess), t(jac), t(dxd), t(-D), t(rhs))
    sol = sol.cpu().numpy()
    errs = []
    for b in range(B):
        ref = sparse_solve(st, hess[b], jac[b], dxd[b], D[b], rhs[b])
        errs.append(np.abs(sol[b] - ref).max() / np.abs(ref).max())
    print(name, variant, ' '.join('%.1e' % e for e in errs))
