'''
Golden interior-point solutions, produced on the CPU: the product's driver (ipm.InteriorPoint) run on
the oracle backend (oracle/cpu_backend.py: OracleNLP tapes + numpy block walk).  The GPU tests start
from the same w0 and must reach the same lap time.  Run from the repo root:
    python tests/golden/make_golden_ipm.py
'''
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from cases import build_case                                                     # noqa: E402
from oracle.nlp_functions import OracleNLP                                       # noqa: E402
from oracle.cpu_backend import OracleBackend                                     # noqa: E402
from aircraft_trajectory_optimization_b200.ipm import InteriorPoint, IpmOptions  # noqa: E402
from aircraft_trajectory_optimization_b200.kkt import build_kkt_structure        # noqa: E402

CASES = [('race_param_rk4_point', 7), ('race_global_rk4_point', 7), ('fig8_global_colloc_point', 8)]


def solve_cpu(name, N, max_iter=400):
    prod, ref = build_case(name, N=N)
    st = prod.structure
    nlp = OracleNLP(ref)
    be = OracleBackend(nlp, nlp, ks=build_kkt_structure(st))
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=float))
    r = InteriorPoint(be, IpmOptions(max_iter=max_iter)).solve(T(st.w0)[None, :], T(st.lbw), T(st.ubw), T(st.lbg), T(st.ubg))
    return prod, r


if __name__ == '__main__':
    for name, N in CASES:
        prod, r = solve_cpu(name, N)
        st = prod.structure
        assert int(r.status[0]) == 0, (name, r.status)
        x = r.x[0].numpy()
        out = os.path.join(ROOT, 'tests', 'golden', f'ipm_{name}_N{N}.npz')
        np.savez_compressed(out, x=x, lam_g=r.lam_g[0].numpy(), lam_x=r.lam_x[0].numpy(), f=float(r.f[0]),
                            lap=float(x[:st.N].sum()), iterations=int(r.iterations[0]))
        print(name, 'lap', float(x[:st.N].sum()), 'iterations', int(r.iterations[0]), '->', out)
