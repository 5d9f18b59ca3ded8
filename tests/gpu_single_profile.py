''' ad-hoc: host-side profile of ONE single-instance solve (not a pytest file) '''
import sys, time, os, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from cases import build_product
name = sys.argv[1] if len(sys.argv) > 1 else 'fig8_param_colloc_point'
prod = build_product(name)
prod.solve()          # warm-up (library load, allocations)
t0 = time.time()
pr = cProfile.Profile()
pr.enable()
res = prod.solve()
torch.cuda.synchronize()
pr.disable()
dt = time.time() - t0
stt = prod.solver.stats()
print(name, 'lap', res.time, stt['return_status'], 'iters', stt['iter_count'], 'wall %.2fs' % dt, '%.1f ms/iter' % (1e3 * dt / max(1, stt['iter_count'])))
print({k: v for k, v in stt.items() if k.startswith('t_') or k.startswith('n_')})
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28)
print(s.getvalue()[:6000])
