'''
Pins for the CPU oracle (oracle/): the reference's only test property (tests/test_kinematics.py:
a parametric model and the global model started on the centerline trace the same path), internal
consistency of its derivatives by finite differences, and the committed golden fixtures.
'''
import os

import numpy as np
import pytest
import scipy.sparse as sp

from cases import CASES, build_oracle, eval_point

DIST_TOL = 1e-4      # reference tests/test_kinematics.py:11


def _open_line(cls, planar=True):
    from aircraft_trajectory_optimization_b200.centerlines import SplineCenterlineConfig, SplineRyFitOptions
    cfg = SplineCenterlineConfig(x=np.array([[0, 10, 0], [0, 10, 20], [0, 5, 10]], dtype=float))
    cfg.closed = False
    cfg.ry_fit_method = SplineRyFitOptions.PLANAR if planar else SplineRyFitOptions.TORSION_FREE
    return cls(cfg)


def _simulate(rhs, z0, steps=100, dt=0.1):
    from scipy.integrate import solve_ivp
    z = np.array(z0, dtype=float)
    out = []
    for _ in range(steps):
        z = solve_ivp(lambda t, y: rhs(y), (0, dt), z, rtol=1e-9, atol=1e-11).y[:, -1]
        out.append(z.copy())
    return np.array(out)


@pytest.mark.parametrize('vehicle,global_r,use_quat', [('point', True, True), ('point', False, True),
                                                      ('drone', True, True), ('drone', False, True),
                                                      ('drone', True, False), ('drone', False, False)])
def test_oracle_kinematics_property(vehicle, global_r, use_quat):
    ''' reference tests/test_kinematics.py:13-102 restated for the oracle's models (PLANAR fit) '''
    from aircraft_trajectory_optimization_b200 import symbolic as sx
    from aircraft_trajectory_optimization_b200.pytypes import DroneConfig, PointConfig
    from oracle.ref_centerline import RefSplineCenterline
    from oracle.ref_models import RefDroneModel, RefPointModel
    line = _open_line(RefSplineCenterline)
    cfg = DroneConfig(g=1, global_r=global_r, use_quat=use_quat) if vehicle == 'drone' \
        else PointConfig(g=1, global_r=global_r)
    cls = RefDroneModel if vehicle == 'drone' else RefPointModel
    par, glob = cls(cfg, line), cls(cfg, None)
    nz, nu = par.nz, par.nu
    nr = nz - 9 if vehicle == 'drone' else 0

    def numeric_rhs(model):
        def rhs(z):
            sx.new_graph()
            pt = model.line.f_param_terms(z[0]) if model.parametric else None
            zs = np.array([sx.SX.const(v) for v in z], dtype=object)
            us = np.array([sx.SX.const(0.0)] * nu, dtype=object)
            return np.array([e.value() for e in model.zdot(zs, us, pt)])
        return rhs

    es0 = line.p2es(0)
    zp, zg = np.zeros(nz), np.zeros(nz)
    if use_quat and vehicle == 'drone':
        zp[6] = zg[6] = 1.0                   # identity quaternion [0, 0, 0, 1]
    zg[3 + nr:6 + nr] = es0
    zp[3 + nr:6 + nr] = es0 if global_r else [1, 0, 0]
    if vehicle == 'drone' and not global_r:
        # body frame aligned with the centerline frame: global model needs that orientation
        from aircraft_trajectory_optimization_b200.pytypes import matrix_to_quat
        Rp = line.p2Rp(0)
        if use_quat:
            zg[3:7] = matrix_to_quat(Rp)
        else:
            from scipy.spatial.transform import Rotation
            zg[3:6] = Rotation.from_matrix(Rp).as_euler('xyz')[::-1]
        zg[3 + nr:6 + nr] = [1, 0, 0]
    traj_p = _simulate(numeric_rhs(par), zp, steps=30)
    traj_g = _simulate(numeric_rhs(glob), zg, steps=30)
    xp = np.array([line.p2x(*z[:3]) for z in traj_p])
    assert np.max(np.linalg.norm(xp - traj_g[:, :3], axis=1)) < DIST_TOL


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_derivatives_by_finite_differences(name, built_library):
    from oracle.nlp_functions import OracleNLP
    ref = build_oracle(name, small=True)
    nlp = OracleNLP(ref)
    x, lam = eval_point(NS_from(ref, nlp), 0)
    v = np.random.default_rng(1).standard_normal(nlp.nw)
    eps = 1e-6
    g0, jv = nlp.nlp_jac_g(x)
    J = nlp.jac_csc(jv)
    fd = (nlp.nlp_g(x + eps * v) - nlp.nlp_g(x - eps * v)) / (2 * eps)
    assert np.max(np.abs(J @ v - fd)) < 1e-6 * max(1, np.max(np.abs(fd)))
    H = nlp.hess_csc(nlp.nlp_hess_l(x, 0.7, lam))
    H = H + H.T - sp.diags(H.diagonal())

    def grad_lag(xx):
        _, gf = nlp.nlp_grad_f(xx)
        _, jj = nlp.nlp_jac_g(xx)
        return 0.7 * gf + nlp.jac_csc(jj).T @ lam
    fd = (grad_lag(x + eps * v) - grad_lag(x - eps * v)) / (2 * eps)
    assert np.max(np.abs(H @ v - fd)) < 1e-6 * max(1, np.max(np.abs(fd)))
    # structural sanity: rows ascending within columns, upper triangle only
    for c in range(nlp.nw):
        r = nlp.hess_row[nlp.hess_colind[c]:nlp.hess_colind[c + 1]]
        assert (np.diff(r) > 0).all() and (r <= c).all()


class NS_from:
    ''' adapter: eval_point wants w0 / bounds / sizes '''

    def __init__(self, ref, nlp):
        self.w0, self.lbw, self.ubw, self.nw, self.ng = ref.w0, ref.lbw, ref.ubw, nlp.nw, nlp.ng


def test_collocation_coefficients():
    ''' discretization_utils.py:6-34: Legendre nodes, integral weights, differentiation matrix '''
    from oracle.ref_discretization import get_collocation_coefficients as ref_cc, \
        get_intermediate_collocation_coefficients as ref_ic
    from aircraft_trajectory_optimization_b200.collocation import get_collocation_coefficients as cc, \
        get_intermediate_collocation_coefficients as ic
    for K in (1, 3, 7):
        a, b = ref_cc(K), cc(K)
        for u, v in zip(a, b):
            assert np.allclose(u, v, rtol=1e-12, atol=1e-12)
        tau, B, C, D = a
        # the reference's poly1d expansion carries ~1e-12 .. 1e-10 of rounding for K = 7; kept as is
        assert abs(B.sum() - 1) < 1e-11 and abs(D.sum() - 1) < 1e-10
        assert np.allclose(C.sum(axis=0), 0, atol=1e-8)           # derivative of a constant
        assert np.allclose(tau @ C, 1, atol=1e-8)                # derivative of t
    assert np.array_equal(ref_ic(7, 0.0), np.eye(8)[0]) and np.array_equal(ic(7, 0.0), np.eye(8)[0])
    assert np.allclose(ref_ic(7, 0.3), ic(7, 0.3), rtol=1e-12, atol=1e-13)


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.slow
@pytest.mark.parametrize('name', ['race_param_rk4_point', 'fig8_global_colloc_point'])
def test_oracle_reproduces_golden(name, built_library):
    ''' the committed full-size fixtures are what the oracle produces (cheap cases only on CPU) '''
    from oracle.nlp_functions import OracleNLP
    path = os.path.join(GOLDEN_DIR, f'{name}.npz')
    if not os.path.exists(path):
        pytest.skip('fixture not generated')
    G = np.load(path)
    ref = build_oracle(name)
    nlp = OracleNLP(ref)
    assert (nlp.nw, nlp.ng, nlp.nnz_jac, nlp.nnz_hess) == tuple(int(G[k]) for k in ('nw', 'ng', 'nnz_jac', 'nnz_hess'))
    rng = np.random.default_rng(0)
    x = np.clip(ref.w0 + 1e-2 * rng.standard_normal(nlp.nw), ref.lbw, ref.ubw)
    g, jv = nlp.nlp_jac_g(x)
    assert np.array_equal(g, G['g_0']) and np.array_equal(jv[G['jac_idx_0']], G['jac_val_0'])
