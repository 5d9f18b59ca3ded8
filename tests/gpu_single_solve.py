''' ad-hoc: one interior-point solve of a case from its warm start, with the time split (not a pytest file) '''
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from cases import build_product
name = sys.argv[1]
prod = build_product(name)
t0 = time.time()
res = prod.solve()
t1 = time.time()
s = prod.solver.stats()
print(name, 'lap', res.time, s['return_status'], 'iters', s['iter_count'], f'wall {t1 - t0:.2f}s',
      {k: round(v, 3) for k, v in s.items() if k.startswith('t_')})
