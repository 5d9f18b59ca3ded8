''' ad-hoc: effect of the second-order correction on iteration counts (not a pytest file) '''
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from aircraft_trajectory_optimization_b200.models import vehicle_params
from aircraft_trajectory_optimization_b200.ipm import IpmOptions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prod = bench.build_c2_with_warm_start()
st = prod.structure
X0, VP = bench.multistart_inputs(st, vehicle_params(prod.vehicle_config), B, seed0=0)
prod.solver.verbose = False
prod.solver.max_iter = 300
for soc in (0, 4):
    prod.solver.options = IpmOptions(window=888, max_soc=soc)
    torch.cuda.synchronize(); t0 = time.time()
    sol = prod.solver(x0=X0, lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg, p=VP)
    torch.cuda.synchronize(); dt = time.time() - t0
    s = prod.solver.stats()
    its = s['iterations_each']; ok = s['success_each']
    print(f'max_soc={soc}: wall {dt:.1f}s converged {int(ok.sum())}/{B} iterations median {np.median(its[ok]):.0f} p90 {np.percentile(its[ok], 90):.0f} '
          f'nominal {its[0]} sweeps {prod.solver.result.n_iter} n_eval {s["n_eval"]} n_soc {s.get("n_soc")} lap0 {sol["x"][0, :st.N].sum():.6f} '
          f't_kkt {s["t_wall_linear_solver"]:.1f}')
