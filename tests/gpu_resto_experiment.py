''' ad-hoc: which instances of the C5 multi-start batch visit the restoration and how they end (not a pytest file) '''
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from aircraft_trajectory_optimization_b200.models import vehicle_params
from aircraft_trajectory_optimization_b200.ipm import IpmOptions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prod = bench.build_c2_with_warm_start()
st = prod.structure
X0, VP = bench.multistart_inputs(st, vehicle_params(prod.vehicle_config), B, seed0=0)
for resto in (False, True):
    prod.solver.verbose = False
    prod.solver.max_iter = 300
    prod.solver.options = IpmOptions(window=0, restoration=resto, verbose=False)
    t0 = time.time()
    sol = prod.solver(x0=X0, lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg, p=VP)
    s = prod.solver.stats()
    r = prod.solver.result
    stt = r.status.cpu().numpy()
    print('restoration', resto, 'time %.1f' % (time.time() - t0), 'status counts', np.unique(stt, return_counts=True),
          'visits', r.n_restorations)
    print('  failed ids', np.nonzero(stt >= 2)[0][:40].tolist())
    print('  their iterations', r.iterations.cpu().numpy()[stt >= 2][:40].tolist())
