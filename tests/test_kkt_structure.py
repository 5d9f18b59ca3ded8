'''
CPU tests of the KKT block structure (aircraft_trajectory_optimization_b200/kkt.py): the tables are
walked in numpy (oracle/kkt_blocks_ref.block_solve, the twin of csrc/kkt_blocks.cuh) and compared with
scipy's sparse LU and with a dense eigendecomposition (inertia) on matrices built from the oracle's
jac_g / hess_l values at seeded points.
'''
import numpy as np
import pytest

from cases import CASES, build_case, build_product, eval_point
from aircraft_trajectory_optimization_b200.kkt import build_kkt_structure

RK4_CASES = [c for c, v in CASES.items() if v[3]]


def _kkt_inputs(st, nlp, seed):
    x, lam = eval_point(st, seed)
    _, jac = nlp.nlp_jac_g(x)
    hess = nlp.nlp_hess_l(x, 1.0, lam)
    rng = np.random.default_rng(seed + 100)
    # bounds far away: tiny barrier diagonal; a few variables without any (0) -- as in a real iteration
    dxd = np.where(rng.uniform(size=st.nw) < 0.2, 0.0, 10.0 ** rng.uniform(-6, 2, st.nw)) + 1.0
    D = np.where(st.lbg == st.ubg, 0.0, 10.0 ** rng.uniform(-6, 1, st.ng))
    rhs = rng.standard_normal(st.nw + st.ng)
    return hess, jac, dxd, D, rhs


@pytest.mark.parametrize('name', RK4_CASES + ['fig8_global_colloc_point', 'fig8_param_colloc_point'])
def test_block_tables_reproduce_the_sparse_solve(name):
    from oracle.nlp_functions import OracleNLP
    from oracle.kkt_blocks_ref import kkt_matrix, sparse_solve, block_solve, chain_factor, chain_solve
    prod, ref = build_case(name, small=True)
    st = prod.structure
    ks = build_kkt_structure(st)
    nlp = OracleNLP(ref)
    # every unknown appears exactly once
    assert np.array_equal(np.sort(ks.unk), np.arange(st.nw + st.ng))
    assert ks.blk_ptr[-1] == st.nw + st.ng and ks.nb == ks.blk_ptr[-1] - ks.blk_ptr[-2]
    for seed in range(2 if st.K == 0 else 1):
        hess, jac, dxd, D, rhs = _kkt_inputs(st, nlp, seed)
        K = kkt_matrix(st, hess, jac, dxd, D)
        ref_sol = sparse_solve(st, hess, jac, dxd, D, rhs)
        sol, neg = block_solve(ks, hess, jac, dxd, D, rhs, with_inertia=True)
        for _ in range(2):
            sol = sol + block_solve(ks, hess, jac, dxd, D, rhs - K @ sol)
        assert np.abs(K @ sol - rhs).max() <= 1e-9 * max(1.0, np.abs(rhs).max())
        assert np.abs(sol - ref_sol).max() <= 1e-7 * np.abs(ref_sol).max()
        assert neg == int((np.linalg.eigvalsh(K.toarray()) < 0).sum())
        # interface form of the border columns (the twin of csrc/kkt_chain.cuh): same answer, same inertia
        F, neg2 = chain_factor(ks, hess, jac, dxd, D)
        sol2 = chain_solve(ks, F, rhs)
        for _ in range(2):
            sol2 = sol2 + chain_solve(ks, F, rhs - K @ sol2)
        assert neg2 == neg
        assert np.abs(K @ sol2 - rhs).max() <= 1e-9 * max(1.0, np.abs(rhs).max())
        assert np.abs(sol2 - ref_sol).max() <= 1e-7 * np.abs(ref_sol).max()


def test_full_size_structure_shapes():
    ''' C2 (race.py, parametric RK4 drone, N = 490): block sizes that the shared-memory solver is sized for '''
    prod = build_product('race_param_rk4_drone')
    ks = build_kkt_structure(prod.structure)
    assert ks.N == 489 and ks.bmax <= 48 and ks.nb <= 48 and ks.mmax <= 20
    # chain couplings only between neighbouring triples
    assert len(ks.cr_ptr) == ks.N + 1 and ks.cr_ptr[-1] == len(ks.cr)


def test_collocation_structure_builds():
    ''' collocation intervals give much larger triples (interior points belong to c_n) '''
    prod = build_product('fig8_global_colloc_point', small=True)
    ks = build_kkt_structure(prod.structure)
    assert ks.bmax > 100 and ks.nb < 64      # c_{N-1} stays in the chain: the border is x_0 + closure rows
