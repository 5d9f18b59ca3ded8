'''
CPU tests of the condensed KKT structure (aircraft_trajectory_optimization_b200/kkt_condensed.py): the interior
and reduced-chain tables are walked in numpy (oracle/kkt_blocks_ref.condensed_factor / condensed_solve, the twin of
csrc/kkt_condense.cuh + csrc/kkt_chain.cuh) and compared with scipy's sparse LU and a dense eigendecomposition
(inertia) on matrices built from the oracle's jac_g / hess_l values at seeded points.
'''
import numpy as np
import pytest

from cases import build_case, build_product
from aircraft_trajectory_optimization_b200.kkt_condensed import build_condensed_structure
from test_kkt_structure import _kkt_inputs

COLLOC_CASES = ['fig8_global_colloc_point', 'fig8_global_colloc_drone', 'fig8_param_colloc_drone',
                'fig8_param_colloc_drone_euler', 'obs_param_colloc_point']


@pytest.mark.parametrize('name', COLLOC_CASES)
def test_condensed_tables_reproduce_the_sparse_solve(name):
    from oracle.nlp_functions import OracleNLP
    from oracle.kkt_blocks_ref import kkt_matrix, sparse_solve, condensed_factor, condensed_solve
    prod, ref = build_case(name, small=True)
    st = prod.structure
    cs = build_condensed_structure(st)
    # every unknown is either interior to exactly one interval or part of the reduced chain / border
    ch = cs.chain
    red = ch.unk[:ch.blk_ptr[ch.N + 1]]
    assert np.array_equal(np.sort(np.concatenate([cs.iunk, red])), np.arange(st.nw + st.ng))
    assert ch.bmax <= 64 and ch.nb <= 64 and cs.smax <= 40
    nlp = OracleNLP(ref)
    hess, jac, dxd, D, rhs = _kkt_inputs(st, nlp, 0)
    K = kkt_matrix(st, hess, jac, dxd, D)
    ref_sol = sparse_solve(st, hess, jac, dxd, D, rhs)
    F, neg = condensed_factor(cs, hess, jac, dxd, D)
    sol = condensed_solve(cs, F, rhs)
    assert np.abs(sol - ref_sol).max() <= 1e-6 * np.abs(ref_sol).max()       # before any refinement
    for _ in range(2):
        sol = sol + condensed_solve(cs, F, rhs - K @ sol)
    assert np.abs(K @ sol - rhs).max() <= 1e-9 * max(1.0, np.abs(rhs).max())
    assert np.abs(sol - ref_sol).max() <= 1e-8 * np.abs(ref_sol).max()
    assert neg == int((np.linalg.eigvalsh(K.toarray()) < 0).sum())


def test_full_size_condensed_shapes():
    ''' C1 (fig_8.py, N = 56), C3 (obstacles.py, N = 100): the reduced system fits the shared-memory chain kernels '''
    for name, N in (('fig8_global_colloc_drone', 56), ('obs_param_colloc_drone', 100)):
        cs = build_condensed_structure(build_product(name).structure)
        assert cs.NI == N and cs.amax <= 300 and cs.smax <= 36
        assert cs.chain.bmax <= 44 and cs.chain.nb <= 48 and cs.chain.mmax <= 20
