''' GPU experiment (not collected by pytest): one evaluation of an open drone raceline (for the ncu launch list) '''
import sys
import numpy as np
sys.path.insert(0, 'tests')
sys.path.insert(0, '.')
from cases import build_product, eval_point   # noqa: E402
for name in ('race_global_rk4_drone_open', 'fig8_global_colloc_drone_open'):
    prod = build_product(name)
    st, F = prod.structure, prod.functions
    x, lam = eval_point(st, 0)
    for _ in range(3):
        out = F.eval(x, lam_f=1.0, lam_g=lam)
    print(name, 'N', st.N, 'nw', st.nw, 'tape', len(st.tail['ins']), 'levels', st.tail['n_levels'], 'slots', st.tail['n_slots'])
