''' ad-hoc GPU experiment: scripts/race.py chain (point-mass warm start -> drone) headless '''
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from cases import make_line, TRACKS
from aircraft_trajectory_optimization_b200 import raceline as RL
from aircraft_trajectory_optimization_b200.pytypes import DroneConfig, PointConfig

track = sys.argv[1] if len(sys.argv) > 1 else 'race'
frame = sys.argv[2] if len(sys.argv) > 2 else 'parametric'
N = int(sys.argv[3]) if len(sys.argv) > 3 else 70
rk4 = (sys.argv[4] != 'colloc') if len(sys.argv) > 4 else True
verbose = (sys.argv[5] == 'v') if len(sys.argv) > 5 else False
line = make_line(track)
vc = DroneConfig(global_r=True, use_quat=True)
t0 = time.time()
if frame == 'global':
    cfg = RL.GlobalRacelineConfig(N=N, use_rk4=rk4, closed=True, verbose=verbose, gate_xi=line.config.x[0],
                                  gate_xj=line.config.x[1], gate_xk=line.config.x[2])
    solver = RL.GlobalDroneRaceline(line, cfg, vc)
else:
    cfg = RL.ParametricRacelineConfig(N=N, use_rk4=rk4, closed=True, verbose=verbose)
    cfg.fixed_gates = line.config.s[:-1]
    solver = RL.ParametricDroneRaceline(line, cfg, vc)
print('setup (incl. warm-start solve)', time.time() - t0)
ws = solver.ws_solver
print('WS: lap', ws.sol['x'][:ws.config.N].sum(), {k: v for k, v in ws.solver.stats().items() if not k.endswith('_each')})
res = solver.solve()
print('DRONE: lap', res.time, 'feasible', res.feasible, 'solve_time', solver.solve_time,
      {k: v for k, v in solver.solver.stats().items() if not k.endswith('_each')})
