'''
Host logic (no GPU): the product's structured descriptor reproduces the oracle's NLP layout
index for index -- sizes, bounds, initial guess, jac_g and hess_l CCS patterns.
'''
import hashlib
import os

import numpy as np
import pytest

from cases import CASES, VARIANT_CASES, OPEN_CASES, SKEW_CASES, build_case, build_product

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name', list(CASES) + list(VARIANT_CASES))
def test_structure_matches_oracle(name, built_library):
    from oracle.nlp_functions import OracleNLP
    prod, ref = build_case(name, small=True)
    st = prod.structure
    nlp = OracleNLP(ref, build_hess=True)
    assert (st.nw, st.ng) == (nlp.nw, nlp.ng)
    assert np.array_equal(st.lbg, ref.lbg) and np.array_equal(st.ubg, ref.ubg)
    assert np.allclose(st.lbw, ref.lbw, rtol=1e-13, atol=1e-13)
    assert np.allclose(st.ubw, ref.ubw, rtol=1e-13, atol=1e-13)
    assert np.allclose(st.w0, ref.w0, rtol=1e-12, atol=1e-12)
    assert np.array_equal(st.jac_colind, nlp.jac_colind) and np.array_equal(st.jac_row, nlp.jac_row)
    assert np.array_equal(st.hess_colind, nlp.hess_colind) and np.array_equal(st.hess_row, nlp.hess_row)
    # every CCS slot is written by exactly one producer
    slots = np.concatenate([st.cell['jslot'][st.cell['jslot'] >= 0].ravel(), st.srow['jslot']])
    assert len(slots) == st.nnz_jac and len(np.unique(slots)) == st.nnz_jac
    hs = st.cell['hslot'][st.cell['hslot'] >= 0].ravel()
    assert len(np.unique(hs)) == len(hs)
    extra = st.shess['slot'][st.shess['add'] == 0]
    assert len(np.unique(np.concatenate([hs, extra]))) == st.nnz_hess


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int64).tobytes()).hexdigest()


@pytest.mark.parametrize('name', list(CASES) + list(OPEN_CASES) + list(SKEW_CASES))
def test_full_size_structure_matches_golden(name, built_library):
    ''' BASELINE.json sizes: patterns (sha256), bounds and initial guess against the oracle fixtures '''
    path = os.path.join(GOLDEN_DIR, f'{name}.npz')
    if not os.path.exists(path):
        pytest.skip('fixture not generated')
    G = np.load(path)
    st = build_product(name).structure
    assert (st.nw, st.ng, st.nnz_jac, st.nnz_hess) == tuple(int(G[k]) for k in ('nw', 'ng', 'nnz_jac', 'nnz_hess'))
    assert _sha(np.concatenate([st.jac_colind, st.jac_row])) == str(G['jac_sha'])
    assert _sha(np.concatenate([st.hess_colind, st.hess_row])) == str(G['hess_sha'])
    assert np.array_equal(st.lbg, G['lbg']) and np.array_equal(st.ubg, G['ubg'])
    assert np.allclose(st.w0, G['w0'], rtol=1e-12, atol=1e-12)
    assert np.allclose(st.lbw, G['lbw'], rtol=1e-13, atol=1e-13)


def test_reference_config_mutation_quirk():
    ''' use_rk4 rewrites N, K, h0 of the passed config in place '''
    prod = build_product('race_param_rk4_drone', small=True)
    assert prod.config.K == 0 and prod.config.N == 49 and abs(prod.config.h0 - 1 / 7) < 1e-15


def test_model_api_and_kinematics_property():
    ''' product models: reference tests/test_kinematics.py property incl. the TORSION_FREE fit '''
    from aircraft_trajectory_optimization_b200.centerlines import SplineCenterline, SplineCenterlineConfig, \
        SplineRyFitOptions
    from aircraft_trajectory_optimization_b200.pytypes import DroneConfig, PointConfig
    from aircraft_trajectory_optimization_b200.models import ParametricDroneModel, DroneModel, \
        ParametricPointModel, PointModel
    for planar in (True, False):
        cfg = SplineCenterlineConfig(x=np.array([[0, 10, 0], [0, 10, 20], [0, 5, 10]], dtype=float))
        cfg.closed = False
        cfg.ry_fit_method = SplineRyFitOptions.PLANAR if planar else SplineRyFitOptions.TORSION_FREE
        cent = SplineCenterline(cfg)
        for make_par, make_glob, vc in (
                (ParametricPointModel, PointModel, PointConfig(g=1, global_r=True)),
                (ParametricDroneModel, DroneModel, DroneConfig(g=1, global_r=True, use_quat=True)),
                (ParametricDroneModel, DroneModel, DroneConfig(g=1, global_r=True, use_quat=False))):
            model, gmodel = make_par(vc, cent), make_glob(vc)
            state, gstate = model.get_empty_state(), gmodel.get_empty_state()
            gstate.v.from_vec(cent.p2es(0))
            state.v.from_vec(cent.p2es(0))
            for _ in range(15):
                gmodel.step(gstate)
                model.step(state)
                assert np.linalg.norm(state.x.to_vec() - gstate.x.to_vec()) < 1e-4
