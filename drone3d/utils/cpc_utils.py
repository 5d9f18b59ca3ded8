'''
drone3d/utils/cpc_utils.py of the reference: packages a trajectory computed by an EXTERNAL complementary-progress
solver (a CSV asset) for display next to the racelines (scripts/race.py:51-54, scripts/fig_8_cpc.py).  Display data only,
outside the hot path; here the lap time is extracted the way the reference does (cpc_utils.py:30-53) so that the
result table of the scripts prints, and no model is built (there is no viewer).
'''
import os

import numpy as np

from aircraft_trajectory_optimization_b200.raceline import RacelineResults


def package_cpc_data_as_raceline(file: str, line, clip: bool = True):
    res = RacelineResults(solve_time=-1, ipopt_time=-1, feval_time=-1, feasible=None, label='CPC', global_frame=True,
                          time=float('nan'), periodic=True, color=[0, 0, 1, 1])
    if not os.path.exists(file):
        res.label = 'CPC (asset missing)'
        return res, None
    data = np.genfromtxt(file, delimiter=',')[1:]
    t, x = data[:, 0], data[:, 1:4]
    lap = float(t[-1])
    if clip:
        x0 = np.asarray(line.p2xc(line.s_min())).ravel()
        i0 = int(np.linalg.norm(x - x0[None], axis=1).argmin())
        tp = np.linspace(t[min(i0 + 10, len(t) - 1)], t.max(), 1000)
        xi = np.stack([np.interp(tp, t, x[:, k]) for k in range(3)], axis=1)
        tf = tp[np.linalg.norm(xi - x[i0][None], axis=1).argmin()]
        lap = float(tf - t[i0])
    res.time = lap
    return res, None
