'''
GPU experiment (not collected by pytest): what IPOPT's gradient-based NLP scaling (nlp_scaling_method = gradient-based,
nlp_scaling_max_gradient = 100) would do on the BASELINE configurations: rows of jac_g whose largest entry at the
starting point exceeds 100 are scaled down to 100, likewise the objective gradient.
'''
import sys
import numpy as np
sys.path.insert(0, 'tests')
sys.path.insert(0, '.')
from cases import build_product   # noqa: E402

for name in ('race_param_rk4_point', 'race_param_rk4_drone', 'fig8_global_colloc_drone', 'fig8_param_colloc_drone',
             'obs_param_colloc_drone', 'race_global_rk4_drone'):
    prod = build_product(name)
    st, F = prod.structure, prod.functions
    x = np.clip(st.w0, st.lbw, st.ubw)
    out = F.eval(x, lam_f=1.0, lam_g=np.zeros(st.ng))
    rows = np.asarray(st.jac_row)
    rmax = np.zeros(st.ng)
    np.maximum.at(rmax, rows, np.abs(out['jac']))
    gf = np.abs(out['grad_f']).max()
    big = rmax > 100
    print(f'{name}: N={st.N} ng={st.ng} |grad f|_inf={gf:.3g} rows with max|J| > 100: {int(big.sum())} '
          f'(largest {rmax.max():.4g}, median of all rows {np.median(rmax):.3g}); 99.9th percentile {np.percentile(rmax, 99.9):.4g}', flush=True)
