'''
Generates tests/golden/<case>.npz from the CPU oracle at BASELINE.json's FULL sizes.

Run in the build container:  python tests/golden/make_golden.py [case ...]
Each fixture holds, for the oracle's restatement of the reference NLP:
  sizes, sha256 of the CCS index arrays of jac_g / hess_l, bounds checksums, and for seeds 0..1 the
  full g vector plus sums and 4000 seeded samples of the jac_g / hess_l value arrays.
The oracle itself is pinned only by property tests (see oracle/__init__.py: parity unpinned vs CasADi).
'''
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import CASES, OPEN_CASES, SKEW_CASES, build_oracle, build_product  # noqa: E402
from oracle.nlp_functions import OracleNLP  # noqa: E402

NSAMPLE = 4000


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int64).tobytes()).hexdigest()


def make(name):
    t0 = time.time()
    ref = build_oracle(name)
    nlp = OracleNLP(ref)
    out = dict(nw=nlp.nw, ng=nlp.ng, nnz_jac=nlp.nnz_jac, nnz_hess=nlp.nnz_hess,
               jac_sha=sha(np.concatenate([nlp.jac_colind, nlp.jac_row])),
               hess_sha=sha(np.concatenate([nlp.hess_colind, nlp.hess_row])),
               w0=ref.w0, lbw=ref.lbw, ubw=ref.ubw, lbg=ref.lbg, ubg=ref.ubg)
    # open racelines / skew closures: EVERY entry the expression rows (tail.py) produce is part of the samples --
    # they are a few hundred of ~10^5 entries, random samples would hardly meet them
    extra_j = extra_h = np.zeros(0, dtype=np.int64)
    if name in OPEN_CASES or name in SKEW_CASES:
        ins = build_product(name).structure.tail['ins']
        extra_j = ins[ins[:, 0] == 15, 2].astype(np.int64)
        extra_h = ins[(ins[:, 0] == 16) | (ins[:, 0] == 17), 2].astype(np.int64)
        out['tail_rows'] = np.sort(ins[ins[:, 0] == 14, 2].astype(np.int64))
    for seed in (0, 1):
        rng = np.random.default_rng(seed)
        x = np.clip(ref.w0 + 1e-2 * rng.standard_normal(nlp.nw), ref.lbw, ref.ubw)
        lam = rng.standard_normal(nlp.ng)
        sigma = 1.0 if seed == 0 else 0.37
        f, gf = nlp.nlp_grad_f(x)
        g, jv = nlp.nlp_jac_g(x)
        hv = nlp.nlp_hess_l(x, sigma, lam)
        srng = np.random.default_rng(1000 + seed)
        ji = np.unique(np.concatenate([srng.choice(len(jv), size=min(NSAMPLE, len(jv)), replace=False), extra_j]))
        hi = np.unique(np.concatenate([srng.choice(len(hv), size=min(NSAMPLE, len(hv)), replace=False), extra_h]))
        out.update({f'f_{seed}': f, f'grad_f_{seed}': gf, f'g_{seed}': g, f'sigma_{seed}': sigma,
                    f'jac_idx_{seed}': ji, f'jac_val_{seed}': jv[ji], f'jac_sum_{seed}': jv.sum(),
                    f'jac_abs_{seed}': np.abs(jv).sum(),
                    f'hess_idx_{seed}': hi, f'hess_val_{seed}': hv[hi], f'hess_sum_{seed}': hv.sum(),
                    f'hess_abs_{seed}': np.abs(hv).sum()})
    np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **out)
    print(f'{name}: nw={nlp.nw} ng={nlp.ng} nnzJ={nlp.nnz_jac} nnzH={nlp.nnz_hess} in {time.time() - t0:.0f}s',
          flush=True)


if __name__ == '__main__':
    for nme in (sys.argv[1:] or list(CASES)):
        make(nme)
