'''
GPU parity at BASELINE.json's FULL sizes (C1 N=56, C2 N=490, C3 N=100, C4 stand-in N=200 and the warm-start
NLPs): the CUDA path through the C ABI (rb_nlp_eval_all, host buffers) against the committed oracle fixtures
tests/golden/<case>.npz (made by tests/golden/make_golden.py).  Row order: drone3d/raceline/base_raceline.py:232-239
at the sizes scripts/fig_8.py:55-62, scripts/race.py:42-49 and scripts/obstacles.py use.

fp64; tolerance 1e-10 relative (BASELINE.json north_star); CCS index arrays bit-exact (sha256).
'''
import hashlib
import os

import numpy as np
import pytest

from cases import CASES, OPEN_CASES, SKEW_CASES, build_product

RTOL = 1e-10
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int64).tobytes()).hexdigest()


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(CASES) + list(OPEN_CASES) + list(SKEW_CASES))
def test_full_size_values_match_golden(name, built_library):
    ''' open racelines / skew closures: the fixtures' samples hold EVERY Jacobian / Hessian
    entry of the expression rows (tests/golden/make_golden.py) next to the 4000 random ones '''
    path = os.path.join(GOLDEN_DIR, f'{name}.npz')
    if not os.path.exists(path):
        pytest.skip('fixture not generated')
    G = np.load(path)
    prod = build_product(name)
    st, F = prod.structure, prod.functions
    assert (st.nw, st.ng, st.nnz_jac, st.nnz_hess) == tuple(int(G[k]) for k in ('nw', 'ng', 'nnz_jac', 'nnz_hess'))
    assert _sha(np.concatenate([st.jac_colind, st.jac_row])) == str(G['jac_sha'])
    assert _sha(np.concatenate([st.hess_colind, st.hess_row])) == str(G['hess_sha'])
    # what the C ABI itself reports for the patterns (rb_sparsity_get), not only the host-side structure
    sj, sh = F.nlp_jac_g.sparsity_out(1), F.nlp_hess_l.sparsity_out(0)
    assert _sha(np.concatenate([sj.colind(), sj.row()])) == str(G['jac_sha'])
    assert _sha(np.concatenate([sh.colind(), sh.row()])) == str(G['hess_sha'])
    for seed in (0, 1):
        rng = np.random.default_rng(seed)
        x = np.clip(G['w0'] + 1e-2 * rng.standard_normal(st.nw), G['lbw'], G['ubw'])
        lam = rng.standard_normal(st.ng)
        sigma = float(G[f'sigma_{seed}'])
        out = F.eval(x, lam_f=sigma, lam_g=lam)
        f_ref = float(G[f'f_{seed}'])
        assert abs(out['f'] - f_ref) <= RTOL * max(1.0, abs(f_ref))
        assert _rel(out['grad_f'], G[f'grad_f_{seed}']) <= RTOL
        assert _rel(out['g'], G[f'g_{seed}']) <= RTOL
        assert _rel(out['jac'][G[f'jac_idx_{seed}']], G[f'jac_val_{seed}']) <= RTOL
        assert _rel(out['hess'][G[f'hess_idx_{seed}']], G[f'hess_val_{seed}']) <= RTOL
        # checksums over ALL entries: sums relative to the sum of magnitudes
        for k in ('jac', 'hess'):
            s_abs = float(G[f'{k}_abs_{seed}'])
            assert abs(out[k].sum() - float(G[f'{k}_sum_{seed}'])) <= RTOL * max(1.0, s_abs)
            assert abs(np.abs(out[k]).sum() - s_abs) <= RTOL * max(1.0, s_abs)
