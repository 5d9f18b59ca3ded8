''' GPU experiment (not collected by pytest): converged solves of open racelines, point mass and warm-started drone '''
import sys
import time

import numpy as np

sys.path.insert(0, 'tests')
sys.path.insert(0, '.')
from cases import build_case, make_line, vehicle_config   # noqa: E402
from test_ipm_cpu import _kkt_conditions                   # noqa: E402
from oracle.nlp_functions import OracleNLP                 # noqa: E402
from aircraft_trajectory_optimization_b200 import raceline as RL   # noqa: E402

for name, N in [('race_global_rk4_point_open', 7), ('fig8_global_colloc_point_open', 8)]:
    prod, ref = build_case(name, N=N)
    t = time.time()
    res = prod.solve()
    s = prod.solver.stats()
    print(name, s['return_status'], 'iter', s.get('iter_count'), 'lap', res.time, '%.2fs' % (time.time() - t), flush=True)
    try:
        _kkt_conditions(prod.structure, OracleNLP(ref, build_hess=False), prod.sol['x'], prod.sol['lam_g'], prod.sol['lam_x'])
        print('   KKT conditions OK (oracle functions)')
    except AssertionError:
        print('   KKT conditions FAILED')

for track, rk4, N in [('race', True, 14), ('fig8', False, 56), ('fig8', False, 16)]:
    line = make_line(track)
    cfg = RL.GlobalRacelineConfig(N=N, use_rk4=rk4, closed=False, verbose=False, gate_xi=line.config.x[0],
                                  gate_xj=line.config.x[1], gate_xk=line.config.x[2])
    t = time.time()
    try:
        solver = RL.GlobalDroneRaceline(line, cfg, vehicle_config('drone', True))
        print(track, 'warm start', solver.ws_solver.solver.stats()['return_status'], 'lap', solver.ws_raceline.time, flush=True)
        res = solver.solve()
        s = solver.solver.stats()
        st = solver.structure
        g = solver.sol['g']
        viol = max(float(np.max(st.lbg - g)), float(np.max(g - st.ubg)), 0.0)
        print(track, 'open drone', s['return_status'], 'iter', s.get('iter_count'), 'lap', res.time, 'viol', viol,
              '%.2fs' % (time.time() - t), flush=True)
    except Exception as e:   # noqa: BLE001
        import traceback
        traceback.print_exc()
