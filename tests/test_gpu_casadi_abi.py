''' the CasADi generated-code shaped symbols (int F(arg, res, iw, w, mem) + companions) against the oracle '''
import ctypes
import numpy as np
import pytest

from cases import build_case, eval_point

pytestmark = pytest.mark.gpu
TOL = 1e-10   # BASELINE.json north_star: 1e-10 relative in fp64


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0


def _call(lib, name, args, outs):
    ''' args: list of float64 arrays or None; outs: list of float64 arrays or None '''
    dp = ctypes.POINTER(ctypes.c_double)
    a = (dp * len(args))(*[None if v is None else v.ctypes.data_as(dp) for v in args])
    r = (dp * len(outs))(*[None if v is None else v.ctypes.data_as(dp) for v in outs])
    f = getattr(lib, name)
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    rc = f(a, r, None, None, 0)
    assert rc == 0, lib.rb_last_error().decode()


def _sparsity(lib, name, i):
    f = getattr(lib, name)
    f.restype, f.argtypes = ctypes.POINTER(ctypes.c_longlong), [ctypes.c_longlong]
    p = f(i)
    nrow, ncol = p[0], p[1]
    colind = np.array([p[2 + k] for k in range(ncol + 1)])
    row = np.array([p[2 + ncol + 1 + k] for k in range(colind[-1])])
    return nrow, ncol, colind, row


@pytest.mark.parametrize('name', ['race_param_rk4_drone', 'fig8_global_colloc_point'])
def test_casadi_shaped_calls(name):
    from oracle.nlp_functions import OracleNLP
    prod, ref = build_case(name, small=True)
    nlp = OracleNLP(ref)
    F, st = prod.functions, prod.structure
    F.bind_casadi_symbols()
    lib = F.lib
    x, lam = eval_point(st, 3)
    vp = np.ascontiguousarray(F.vp, dtype=np.float64)
    one = np.array([1.0])

    # sparsity_out is the structure's CCS pattern, bit-exact, in CasADi's compressed form
    nrow, ncol, colind, row = _sparsity(lib, 'nlp_jac_g_sparsity_out', 1)
    assert (nrow, ncol) == (st.ng, st.nw)
    assert np.array_equal(colind, st.jac_colind) and np.array_equal(row, st.jac_row)
    nrow, ncol, colind, row = _sparsity(lib, 'nlp_hess_l_sparsity_out', 0)
    assert (nrow, ncol) == (st.nw, st.nw)
    assert np.array_equal(colind, st.hess_colind) and np.array_equal(row, st.hess_row)
    assert np.all(row <= np.repeat(np.arange(st.nw), np.diff(colind)))          # upper triangle
    nrow, ncol, colind, row = _sparsity(lib, 'nlp_g_sparsity_out', 0)
    assert (nrow, ncol) == (st.ng, 1) and colind[-1] == st.ng

    f = np.zeros(1); g = np.zeros(st.ng); gf = np.zeros(st.nw)
    jac = np.zeros(st.nnz_jac); hess = np.zeros(st.nnz_hess)
    _call(lib, 'nlp_f', [x, vp], [f])
    _call(lib, 'nlp_g', [x, vp], [g])
    f2 = np.zeros(1)
    _call(lib, 'nlp_grad_f', [x, vp], [f2, gf])
    g2 = np.zeros(st.ng)
    _call(lib, 'nlp_jac_g', [x, vp], [g2, jac])
    _call(lib, 'nlp_hess_l', [x, vp, one, lam], [hess])
    g_ref, j_ref = nlp.nlp_jac_g(x)
    f_ref, gf_ref = nlp.nlp_grad_f(x)
    h_ref = nlp.nlp_hess_l(x, 1.0, lam)
    assert f[0] == f2[0] and np.array_equal(g, g2)
    errs = dict(f=_rel(f, np.array([f_ref])), g=_rel(g, g_ref), gf=_rel(gf, gf_ref), jac=_rel(jac, j_ref),
                hess=_rel(hess, h_ref))
    assert all(e < TOL for e in errs.values()), errs

    # NULL conventions: p omitted -> the bound vehicle parameters; an output omitted -> not computed / not written
    g3 = np.zeros(st.ng)
    _call(lib, 'nlp_jac_g', [x, None], [g3, None])
    assert np.array_equal(g3, g)
    # lam_f omitted reads as zero: Hessian of lam' g alone; by linearity the two pieces add up
    h_g = np.zeros(st.nnz_hess); h_f = np.zeros(st.nnz_hess)
    _call(lib, 'nlp_hess_l', [x, vp, None, lam], [h_g])
    _call(lib, 'nlp_hess_l', [x, vp, one, None], [h_f])
    assert _rel(h_g + h_f, hess) < 1e-12
    # unbinding makes the symbols fail loudly again
    lib.rb_casadi_bind(None, None)
    dp = ctypes.POINTER(ctypes.c_double)
    assert lib.nlp_f((dp * 2)(), (dp * 1)(), None, None, 0) != 0
