''' ad-hoc GPU experiment: scripts/obstacles.py pattern (synthetic tube) headless '''
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from cases import make_line, synthetic_tube_arrays
from aircraft_trajectory_optimization_b200 import raceline as RL
from aircraft_trajectory_optimization_b200.pytypes import DroneConfig

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
line = make_line('obs')
vc = DroneConfig(global_r=True, use_quat=True, collision_radius=0.4)
cfg = RL.ParametricRacelineConfig(N=N, closed=True, verbose=False)
K = 7
tau = RL.get_collocation_coefficients(K)[0]
ds = (line.s_max() - line.s_min()) / N
s_all = np.array([line.s_min() + ds * (n + tau[k]) for n in range(N) for k in range(K + 1)])
bp, br = synthetic_tube_arrays(s_all)
tube = RL.ObstacleFreeTube(bp, br, vc.collision_radius)
t0 = time.time()
solver = RL.ParametricObstacleDroneRaceline(line, cfg, vc, None, tube)
print('setup (incl. warm-start solve)', time.time() - t0)
ws = solver.ws_solver
print('WS: lap', ws.sol['x'][:ws.config.N].sum(), {k: v for k, v in ws.solver.stats().items() if not k.endswith('_each') and not k.startswith('t_wall_nlp')})
res = solver.solve()
print('DRONE: lap', res.time, 'feasible', res.feasible, 'solve_time', solver.solve_time,
      {k: v for k, v in solver.solver.stats().items() if not k.endswith('_each') and not k.startswith('t_wall_nlp')})
g = solver.sol['g']; st = solver.structure
print('max constraint violation', float(np.maximum(st.lbg - g, g - st.ubg).max()))
