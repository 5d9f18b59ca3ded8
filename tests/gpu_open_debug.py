''' GPU experiment (not collected by pytest): iteration log of an open drone raceline '''
import sys
import numpy as np
sys.path.insert(0, 'tests')
sys.path.insert(0, '.')
from cases import make_line, vehicle_config   # noqa: E402
from aircraft_trajectory_optimization_b200 import raceline as RL   # noqa: E402
quat = len(sys.argv) < 2 or sys.argv[1] != 'euler'
line = make_line('race')
cfg = RL.GlobalRacelineConfig(N=7, use_rk4=True, closed=False, verbose=False, gate_xi=line.config.x[0],
                              gate_xj=line.config.x[1], gate_xk=line.config.x[2])
solver = RL.GlobalDroneRaceline(line, cfg, vehicle_config('drone', quat))
print('warm start', solver.ws_solver.solver.stats()['return_status'], solver.ws_raceline.time)
st = solver.structure
F = solver.functions
out = F.eval(st.w0, lam_f=1.0, lam_g=np.zeros(st.ng))
g = out['g']
viol = np.maximum(np.maximum(st.lbg - g, g - st.ubg), 0)
print('initial violation max', viol.max(), 'rows', np.argsort(-viol)[:10], 'ng', st.ng, 'nan in w0', np.isnan(st.w0).any())
print('bounds violated in w0', np.maximum(st.lbw - st.w0, st.w0 - st.ubw).max())
solver.solver.verbose = True
solver.config.verbose = True
res = solver.solve()
print(solver.solver.stats()['return_status'], res.time)
