''' ad-hoc: one instance of the C5 batch, verbose (not a pytest file) '''
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from aircraft_trajectory_optimization_b200.models import vehicle_params
from aircraft_trajectory_optimization_b200.ipm import IpmOptions
idx = int(sys.argv[1])
prod = bench.build_c2_with_warm_start()
st = prod.structure
X0, VP = bench.multistart_inputs(st, vehicle_params(prod.vehicle_config), idx + 1, seed0=0)
prod.solver.verbose = True
prod.solver.max_iter = 300
prod.solver.options = IpmOptions(window=0, restoration=True, verbose=True)
sol = prod.solver(x0=X0[idx], lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg, p=VP[idx])
print(prod.solver.stats()['return_status'], prod.solver.result.n_restorations)
