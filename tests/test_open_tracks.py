'''
Open (non-periodic) racelines, host logic (no GPU): SURVEY.md s8 a8.  The product's builders reproduce the oracle's
rows, bounds and CCS patterns index for index, and the levelised tape of the end rows (tail.py) -- interpreted here
with numpy, on the GPU by csrc/tail_tape.cuh -- reproduces the oracle's values of those rows, of their Jacobian
entries and of the Hessian entries only they touch.  Reference: drone3d/raceline/base_raceline.py:516-543, :914-918,
drone_raceline.py:110-148, point_raceline.py:15-45.
'''
import numpy as np
import pytest

from cases import OPEN_CASES, build_case, build_product, eval_point


@pytest.mark.parametrize('name', list(OPEN_CASES))
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
