'''
GPU parity of the batched KKT factor / solve (through the C ABI) against scipy's sparse LU on the
same matrices: values of jac_g / hess_l from the CUDA evaluation at seeded points, random positive
diagonals.  fp64; the block solve is followed by one refinement step through rb_kkt_resolve.
'''
import numpy as np
import pytest

from cases import CASES, build_case, build_product, eval_point

RK4_CASES = [c for c, v in CASES.items() if v[3]]
# collocation structures: interiors condensed (csrc/kkt_condense.cuh) + chain kernels on the reduced system; the
# uncondensed large-block kernel (csrc/kkt_big.cuh) is kept and tested on one structure ('@big')
KKT_CASES = RK4_CASES + ['fig8_global_colloc_point', 'fig8_global_colloc_drone', 'fig8_param_colloc_drone',
                         'fig8_param_colloc_drone_euler', 'obs_param_colloc_point', 'fig8_global_colloc_point@big',
                         # open racelines / skew closure (point mass: the end rows of open drone racelines are rank
                         # deficient by construction, tests/test_open_tracks.py)
                         'race_global_rk4_point_open', 'fig8_global_colloc_point_open', 'race_param_rk4_point_skew']


def _inputs(st, F, B, seed=0):
    rng = np.random.default_rng(seed)
    X = np.stack([eval_point(st, seed + b)[0] for b in range(B)])
    L = np.stack([eval_point(st, seed + b)[1] for b in range(B)])
    out = F.eval(X, lam_f=np.ones(B), lam_g=L, want=('jac', 'hess'))
    dxd = 1.0 + rng.uniform(0, 1, (B, st.nw))
    D = np.where(st.lbg == st.ubg, 0.0, rng.uniform(0.1, 1, (B, st.ng)))
    rhs = rng.standard_normal((B, st.nw + st.ng))
    return out['hess'], out['jac'], dxd, D, rhs


@pytest.mark.gpu
@pytest.mark.parametrize('name', KKT_CASES)
def test_kkt_solve_matches_sparse_lu(name, built_library):
    import torch
    from oracle.kkt_blocks_ref import sparse_solve, kkt_matrix
    from aircraft_trajectory_optimization_b200.kkt import KktSolver
    name, _, variant = name.partition('@')
    prod = build_product(name, small=True)
    st, F = prod.structure, prod.functions
    B = 3
    hess, jac, dxd, D, rhs = _inputs(st, F, B)
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    K = KktSolver(st, condensed=False) if variant == 'big' else KktSolver(st)
    th, tj, tdx, tnd, tr = t(hess), t(jac), t(dxd), t(-D), t(rhs)
    sol, status = K.factor_solve(th, tj, tdx, tnd, tr)
    assert int(status[:, 0].abs().sum()) == 0
    # K v product against scipy
    kv = K.matvec(th, tj, tdx, tnd, sol).cpu().numpy()
    # iterative refinement with the stored factors
    sol2 = sol.clone()
    for _ in range(3):
        sol2 = sol2 + K.resolve(th, tj, tdx, tnd, tr - K.matvec(th, tj, tdx, tnd, sol2))
    sol, sol2, status = sol.cpu().numpy(), sol2.cpu().numpy(), status.cpu().numpy()
    for b in range(B):
        Km = kkt_matrix(st, hess[b], jac[b], dxd[b], D[b])
        ref = sparse_solve(st, hess[b], jac[b], dxd[b], D[b], rhs[b])
        scale = np.abs(ref).max()
        # K v product: bit-for-bit the same sum up to ordering
        assert np.abs(Km @ sol[b] - kv[b]).max() <= 1e-9 * max(1.0, np.abs(kv[b]).max())
        # the unrefined block solve (explicit block inverses): measured 1e-14 .. 3.4e-9 relative on these matrices
        # (tests/gpu_kkt_accuracy.py); the interior-point matrices are worse conditioned, hence the driver's refinement
        assert np.abs(sol[b] - ref).max() <= 1e-6 * scale, (np.abs(sol[b] - ref).max(), scale)
        # refined: residual at fp64 level, solution equal to SuperLU's within conditioning
        assert np.abs(Km @ sol2[b] - rhs[b]).max() <= 1e-9 * max(1.0, np.abs(rhs[b]).max())
        assert np.abs(sol2[b] - ref).max() <= 1e-7 * scale, (np.abs(sol2[b] - ref).max(), scale)
        # inertia: exactly ng negative eigenvalues here (dx_diag > 0 dominates W on these instances?) -- compare
        # with the count from a dense eigendecomposition
        neg_ref = int((np.linalg.eigvalsh(Km.toarray()) < 0).sum())
        assert status[b, 1] == neg_ref, (status[b], neg_ref)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['race_param_rk4_drone', 'fig8_global_colloc_drone', 'obs_param_colloc_drone',
                                  'fig8cpc_param_colloc_drone'])
def test_kkt_full_size_residual(name, built_library):
    ''' full-size C2 (N = 490), C1 (N = 56), C3 (N = 100), C4 stand-in (N = 200) structures: residual of the refined solution '''
    import torch
    from aircraft_trajectory_optimization_b200.kkt import KktSolver
    prod = build_product(name)
    st, F = prod.structure, prod.functions
    B = 2
    hess, jac, dxd, D, rhs = _inputs(st, F, B, seed=11)
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    K = KktSolver(st)
    th, tj, tdx, tnd, tr = t(hess), t(jac), t(dxd), t(-D), t(rhs)
    sol, status = K.factor_solve(th, tj, tdx, tnd, tr)
    for _ in range(2):
        sol = sol + K.resolve(th, tj, tdx, tnd, tr - K.matvec(th, tj, tdx, tnd, sol))
    res = (tr - K.matvec(th, tj, tdx, tnd, sol)).abs().max().item()
    assert int(status[:, 0].abs().sum()) == 0
    assert res <= 1e-8 * max(1.0, float(np.abs(rhs).max()))


@pytest.mark.gpu
def test_resolve_rows_equals_full_resolve(built_library):
    ''' re-solve for a subset of the factorised instances (rb_kkt_resolve_rows): bit-identical to the rows of the full
    re-solve, whatever the order of the subset '''
    import torch
    from aircraft_trajectory_optimization_b200.kkt import KktSolver
    prod = build_product('race_param_rk4_drone', small=True)
    st, F = prod.structure, prod.functions
    B = 6
    hess, jac, dxd, D, rhs = _inputs(st, F, B, seed=3)
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    K = KktSolver(st)
    assert K.can_resolve_rows
    th, tj, tdx, tnd, tr = t(hess), t(jac), t(dxd), t(-D), t(rhs)
    K.factor_solve(th, tj, tdx, tnd, tr)
    r2 = t(np.random.default_rng(5).standard_normal(rhs.shape))
    full = K.resolve(th, tj, tdx, tnd, r2).cpu().numpy()
    idx = torch.tensor([4, 1, 5], dtype=torch.int32, device=dev)
    sub = K.resolve_rows(idx, r2[idx.long()].contiguous()).cpu().numpy()
    assert np.array_equal(sub, full[[4, 1, 5]])
