'''
GPU parity: the CUDA path (through the C ABI, host-buffer entry points) against the CPU oracle on
the same seeded inputs.  fp64; tolerance 1e-10 relative (BASELINE.json north_star), indices bit-exact.
'''
import numpy as np
import pytest

from cases import CASES, VARIANT_CASES, OPEN_CASES, SKEW_CASES, build_case, eval_point

RTOL = 1e-10
RK4_CASES = [c for c, v in CASES.items() if v[3]]
COLLOC_CASES = [c for c, v in CASES.items() if not v[3]]


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.maximum(1.0, np.abs(b))
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def _check_case(name):
    from oracle.nlp_functions import OracleNLP
    prod, ref = build_case(name, small=True)
    st = prod.structure
    nlp = OracleNLP(ref)
    # indices: bit-exact
    assert st.nw == nlp.nw and st.ng == nlp.ng
    assert np.array_equal(st.jac_colind, nlp.jac_colind) and np.array_equal(st.jac_row, nlp.jac_row)
    assert np.array_equal(st.hess_colind, nlp.hess_colind) and np.array_equal(st.hess_row, nlp.hess_row)
    F = prod.functions
    for seed in range(3):
        x, lam = eval_point(st, seed)
        sigma = 1.0 if seed == 0 else 0.37
        out = F.eval(x, lam_f=sigma, lam_g=lam)
        f_ref, gf_ref = nlp.nlp_grad_f(x)
        g_ref, j_ref = nlp.nlp_jac_g(x)
        h_ref = nlp.nlp_hess_l(x, sigma, lam)
        assert abs(out['f'] - f_ref) <= RTOL * max(1.0, abs(f_ref))
        assert _rel(out['grad_f'], gf_ref) <= RTOL
        assert _rel(out['g'], g_ref) <= RTOL
        assert _rel(out['jac'], j_ref) <= RTOL
        assert _rel(out['hess'], h_ref) <= RTOL
    # CasADi-shaped single-function calls agree with the all-in-one call
    x, lam = eval_point(st, 7)
    g1 = F.nlp_g(x, F.vp)
    g2, j2 = F.nlp_jac_g(x, F.vp)
    assert np.array_equal(g1, g2)
    assert F.nlp_jac_g.sparsity_out(1).nnz() == st.nnz_jac
    assert F.nlp_hess_l.sparsity_out(0).colind() == st.hess_colind.tolist()


@pytest.mark.gpu
@pytest.mark.parametrize('name', RK4_CASES)
def test_rk4_parity(name, built_library):
    _check_case(name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', COLLOC_CASES)
def test_collocation_parity(name, built_library):
    _check_case(name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(VARIANT_CASES))
def test_variant_parity(name, built_library):
    ''' frame-relative orientation (global_r=False) and linear-drag variants '''
    _check_case(name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(OPEN_CASES) + list(SKEW_CASES))
def test_open_track_parity(name, built_library):
    ''' open racelines and skewly closed centerlines (base_raceline.py:1208-1227): the expression
    rows -- initial / terminal rows, end gate, skew closure rows -- evaluated from the tape '''
    _check_case(name)
    # the g-only and g + Jacobian calls stop the tape early: same values as the all-in-one call
    prod, _ = build_case(name, small=True)
    st, F = prod.structure, prod.functions
    x, lam = eval_point(st, 11)
    full = F.eval(x, lam_f=1.0, lam_g=lam)
    assert np.array_equal(F.nlp_g(x, F.vp), full['g'])
    g2, j2 = F.nlp_jac_g(x, F.vp)
    assert np.array_equal(g2, full['g']) and np.array_equal(j2, full['jac'])
    assert np.array_equal(F.nlp_hess_l(x, F.vp, 1.0, lam), full['hess'])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['race_param_rk4_drone', 'race_global_rk4_drone_open', 'fig8_global_colloc_point_open'])
def test_batch_matches_single(name, built_library):
    ''' a batch of B instances with per-instance vehicle parameters == B single calls (open racelines: the tape
    kernel runs one CTA per instance and reads the instance's own vehicle parameters) '''
    from aircraft_trajectory_optimization_b200.models import vehicle_params
    prod, _ = build_case(name, small=True)
    st, F = prod.structure, prod.functions
    B = 5
    rng = np.random.default_rng(3)
    X = np.stack([eval_point(st, s)[0] for s in range(B)])
    L = np.stack([eval_point(st, s)[1] for s in range(B)])
    VP = np.tile(vehicle_params(prod.vehicle_config), (B, 1)) * rng.uniform(0.9, 1.1, size=(B, F.nvp))
    batch = F.eval(X, p=VP, lam_f=np.ones(B), lam_g=L)
    for b in range(B):
        one = F.eval(X[b], p=VP[b], lam_f=1.0, lam_g=L[b])
        for k in ('grad_f', 'g', 'jac', 'hess'):
            assert np.array_equal(batch[k][b], one[k]), k
        assert batch['f'][b] == one['f']


@pytest.mark.gpu
def test_structural_zero_slots_and_linearity(built_library):
    ''' size-independent properties on the full-size C2 instance: hess_l is linear in (lam_f, lam_g) '''
    prod, _ = build_case('race_param_rk4_drone', small=False)
    st, F = prod.structure, prod.functions
    x, lam = eval_point(st, 0)
    _, lam2 = eval_point(st, 1)
    h1 = F.eval(x, lam_f=1.0, lam_g=lam, want='hess')['hess']
    h2 = F.eval(x, lam_f=0.5, lam_g=lam2, want='hess')['hess']
    h12 = F.eval(x, lam_f=1.5, lam_g=lam + lam2, want='hess')['hess']
    assert np.max(np.abs(h12 - (h1 + h2))) <= 1e-9 * max(1.0, np.max(np.abs(h12)))
    # J v against central differences of g (property check at full size, no oracle needed)
    import scipy.sparse as sp
    out = F.eval(x, want=('g', 'jac'))
    J = sp.csc_matrix((out['jac'], st.jac_row, st.jac_colind), shape=(st.ng, st.nw))
    v = np.random.default_rng(5).standard_normal(st.nw)
    eps = 1e-6
    fd = (F.eval(x + eps * v, want='g')['g'] - F.eval(x - eps * v, want='g')['g']) / (2 * eps)
    assert np.max(np.abs(J @ v - fd)) <= 1e-6 * max(1.0, np.max(np.abs(fd)))
