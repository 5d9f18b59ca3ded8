''' ad-hoc: drone warm-start chain at reduced size with / without the fused kernels (not a pytest file) '''
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from cases import make_line
from aircraft_trajectory_optimization_b200 import raceline as RL
from aircraft_trajectory_optimization_b200.pytypes import DroneConfig
from aircraft_trajectory_optimization_b200.ipm import IpmOptions
line = make_line('race')
for glue, resto, refine, NN in ((True, True, 4, 10), (False, True, 4, 10), (True, True, 4, 14), (False, True, 4, 14), (True, True, 4, 20), (True, False, 1, 20)):
    cfg = RL.ParametricRacelineConfig(N=NN, use_rk4=True, closed=True, verbose=False)
    cfg.fixed_gates = line.config.s[:-1]
    solver = RL.ParametricDroneRaceline(line, cfg, DroneConfig(global_r=True, use_quat=True))
    solver.solver.options = IpmOptions(use_glue=glue, restoration=resto, refine_steps=refine, verbose='-v' in sys.argv)
    solver.solver.verbose = '-v' in sys.argv
    res = solver.solve()
    s = solver.solver.stats()
    print('N', NN, 'glue', glue, 'resto', resto, 'refine', refine, s['return_status'], s['iter_count'], 'lap', res.time, 'restorations', solver.solver.result.n_restorations)
