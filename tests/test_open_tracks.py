'''
Open (non-periodic) racelines and skewly closed centerlines, host logic (no GPU): SURVEY.md s8 a8 / a7.  The product's builders reproduce the oracle's
rows, bounds and CCS patterns index for index, and the levelised tape of the end rows (tail.py) -- interpreted here
with numpy, on the GPU by csrc/tail_tape.cuh -- reproduces the oracle's values of those rows, of their Jacobian
entries and of the Hessian entries only they touch.  Reference: drone3d/raceline/base_raceline.py:516-543, :914-918,
drone_raceline.py:110-148, point_raceline.py:15-45; skew closure :1208-1227 (two closure rows with two partners each
go through the same expression rows).
'''
import numpy as np
import pytest

from cases import OPEN_CASES, SKEW_CASES, build_case, eval_point


@pytest.mark.parametrize('name', list(OPEN_CASES) + list(SKEW_CASES))
def test_open_structure_and_tape_match_oracle(name, built_library):
    from oracle.nlp_functions import OracleNLP
    from aircraft_trajectory_optimization_b200.tail import TailRows, T_STORE_G, T_STORE_J, T_STORE_H, T_ADD_H
    from aircraft_trajectory_optimization_b200.models import vehicle_params
    prod, ref = build_case(name, small=True)
    st = prod.structure
    nlp = OracleNLP(ref)
    assert (st.nw, st.ng) == (nlp.nw, nlp.ng)
    assert np.array_equal(st.lbg, ref.lbg) and np.array_equal(st.ubg, ref.ubg)
    assert np.allclose(st.w0, ref.w0, rtol=1e-12, atol=1e-12)
    assert np.array_equal(st.jac_colind, nlp.jac_colind) and np.array_equal(st.jac_row, nlp.jac_row)
    assert np.array_equal(st.hess_colind, nlp.hess_colind) and np.array_equal(st.hess_row, nlp.hess_row)
    tp = st.tail
    assert tp is not None
    ins, lp = tp['ins'], tp['lvl_ptr']
    # levels: no instruction reads a slot written in its own level, stores close every phase
    for l in range(len(lp) - 1):
        blk = ins[lp[l]:lp[l + 1]]
        arith = blk[(blk[:, 0] >= 4) & (blk[:, 0] < T_STORE_G)]
        written = set(blk[blk[:, 0] < T_STORE_G][:, 3].tolist())
        reads = set(arith[:, 1].tolist()) | set(arith[(arith[:, 0] <= 7), 2].tolist())
        assert not (written & reads)
    codes = ins[:, 0]
    # every CCS slot has exactly one assigning producer: cells, simple rows or the tape
    jslots = np.concatenate([st.cell['jslot'][st.cell['jslot'] >= 0].ravel(), st.srow['jslot'], ins[codes == T_STORE_J, 2]])
    assert len(jslots) == st.nnz_jac and len(np.unique(jslots)) == st.nnz_jac
    hs = st.cell['hslot'][st.cell['hslot'] >= 0].ravel()
    extra = st.shess['slot'][st.shess['add'] == 0]
    assigned = np.concatenate([hs, extra, ins[codes == T_STORE_H, 2]])
    assert len(np.unique(assigned)) == len(assigned) == st.nnz_hess
    assert np.isin(ins[codes == T_ADD_H, 2], np.concatenate([hs, st.shess['slot']])).all()

    vp = vehicle_params(prod.vehicle_config)
    for seed in range(2):
        x, lam = eval_point(st, seed)
        g_ref, j_ref = nlp.nlp_jac_g(x)
        h_ref = nlp.nlp_hess_l(x, 1.0, lam)
        g = np.full(st.ng, np.nan)
        jac = np.full(st.nnz_jac, np.nan)
        hess = np.full(st.nnz_hess, np.nan)
        TailRows.run_tape(tp, x, vp, lam, g, jac, hess)
        m, mj = ~np.isnan(g), ~np.isnan(jac)
        assert m.sum() == np.count_nonzero(codes == T_STORE_G) and mj.sum() == np.count_nonzero(codes == T_STORE_J)
        assert np.max(np.abs(g[m] - g_ref[m]) / np.maximum(1, np.abs(g_ref[m]))) <= 1e-12
        assert np.max(np.abs(jac[mj] - j_ref[mj]) / np.maximum(1, np.abs(j_ref[mj]))) <= 1e-12
        only = ins[codes == T_STORE_H, 2]
        if len(only):
            assert np.max(np.abs(hess[only] - h_ref[only]) / np.maximum(1, np.abs(h_ref[only]))) <= 1e-12
        # the g-only phase gives the same rows
        g1 = np.full(st.ng, np.nan)
        TailRows.run_tape(tp, x, vp, None, g1, None, None, phase=0)
        assert np.array_equal(g1[m], g[m])


def test_open_parametric_raises(built_library):
    ''' the reference cannot open a parametric raceline either (its helper functions would need the spline at a
    symbolic path length); the builder says so instead of producing something else '''
    from aircraft_trajectory_optimization_b200 import raceline as RL
    import cases
    line = cases.make_line('race')
    cfg = RL.ParametricRacelineConfig(N=7, use_rk4=True, closed=False, verbose=False)
    with pytest.raises(NotImplementedError):
        RL.ParametricPointRaceline(line, cfg, cases.vehicle_config('point'))


# lap times of the open point-mass racelines, measured with the interior-point driver on the CPU oracle backend
# (test_open_point_mass_solve_on_the_cpu_backend below); the GPU solves must reach the same minima
OPEN_LAPS = {('race_global_rk4_point_open', 7): 5.992371263281883, ('fig8_global_colloc_point_open', 8): 4.803612974951485}


def test_open_point_mass_solve_on_the_cpu_backend(built_library):
    ''' the open NLP is solvable: the driver on the oracle's functions converges to a KKT point of the oracle's NLP '''
    import torch
    from oracle.nlp_functions import OracleNLP
    from oracle.cpu_backend import OracleBackend
    from aircraft_trajectory_optimization_b200.ipm import InteriorPoint, IpmOptions
    from aircraft_trajectory_optimization_b200.kkt import build_kkt_structure
    from test_ipm_cpu import _kkt_conditions
    name, N = 'race_global_rk4_point_open', 7
    prod, ref = build_case(name, N=N)
    st = prod.structure
    nlp = OracleNLP(ref)
    be = OracleBackend(nlp, nlp, ks=build_kkt_structure(st))
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=float))
    r = InteriorPoint(be, IpmOptions(max_iter=300)).solve(T(st.w0)[None], T(st.lbw), T(st.ubw), T(st.lbg), T(st.ubg))
    assert r.success.all() and (r.status == 0).all()
    lap = float(r.x[0, :st.N].sum())
    assert abs(lap - OPEN_LAPS[(name, N)]) <= 1e-9 * lap
    x = r.x[0].numpy()
    _kkt_conditions(st, nlp, x, r.lam_g[0].numpy(), r.lam_x[0].numpy())
    # the end rows hold: at rest at both ends (vg'vg <= 0 relaxed by the bound relaxation), vertical thrust
    g = nlp.nlp_g(x)
    ins = st.tail['ins']
    rows = ins[ins[:, 0] == 14, 2]
    assert (g[rows] <= st.ubg[rows] + 1e-7).all() and (g[rows] >= st.lbg[rows] - 1e-7).all()


@pytest.mark.parametrize('name', ['race_global_rk4_point_open', 'race_global_rk4_drone_open', 'fig8_global_colloc_point_open',
                                  'race_param_rk4_point_skew'])
def test_open_kkt_block_tables_reproduce_the_sparse_solve(name, built_library):
    '''
    the KKT block analysis (kkt.py) needs no special case for open tracks: x_0 and the end rows join the border, the
    chain is walked by the numpy twin of csrc/kkt_chain.cuh and compared with SuperLU and a dense eigendecomposition.
    The end rows of open DRONE racelines are rank deficient by construction (R(qF)[:,2] = e3 on the renormalised
    quaternion: the gradient of the third row vanishes on the feasible set), so the equality rows carry the
    delta_c > 0 the interior-point driver gives them in that case.
    '''
    from oracle.nlp_functions import OracleNLP
    from oracle.kkt_blocks_ref import kkt_matrix, sparse_solve, chain_factor, chain_solve
    from aircraft_trajectory_optimization_b200.kkt import build_kkt_structure
    from test_kkt_structure import _kkt_inputs
    prod, ref = build_case(name, small=True)
    st = prod.structure
    ks = build_kkt_structure(st)
    assert np.array_equal(np.sort(ks.unk), np.arange(st.nw + st.ng))
    assert ks.bmax <= 64 or st.K > 0
    nlp = OracleNLP(ref)
    hess, jac, dxd, D, rhs = _kkt_inputs(st, nlp, 0)
    D = np.maximum(D, 1e-5)
    K = kkt_matrix(st, hess, jac, dxd, D)
    ref_sol = sparse_solve(st, hess, jac, dxd, D, rhs)
    F, neg = chain_factor(ks, hess, jac, dxd, D)
    sol = chain_solve(ks, F, rhs)
    for _ in range(2):
        sol = sol + chain_solve(ks, F, rhs - K @ sol)
    assert np.abs(K @ sol - rhs).max() <= 1e-9 * max(1.0, np.abs(rhs).max())
    assert np.abs(sol - ref_sol).max() <= 1e-7 * np.abs(ref_sol).max()
    assert neg == int((np.linalg.eigvalsh(K.toarray()) < 0).sum())
