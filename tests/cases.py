'''
The BASELINE.json problem instances, built twice: by the product builders and by
the oracle's literal restatement of the reference.  Waypoints are the ones in the reference's
scripts (scripts/fig_8.py:10-12, scripts/race.py:12-14, scripts/obstacles.py:14-16).
'''
from types import SimpleNamespace as NS

import numpy as np

from aircraft_trajectory_optimization_b200.pytypes import DroneConfig, PointConfig
from aircraft_trajectory_optimization_b200.centerlines import SplineCenterlineConfig, GateShape, \
    SplineCenterline
from aircraft_trajectory_optimization_b200 import raceline as RL

# gate radii that differ from the centerline defaults (reference scripts/fig_8_cpc.py:18-19)
TRACK_GATES = {'fig8cpc': dict(gate_ri=0.6, gate_ro=0.8)}
TRACKS = {
    'fig8cpc': (np.array([[0, 5, 0, -5, 0, 5, 0, -5], [0, 1, 2, 1, 0, -1, -2, -1],
                          [10, 5, 0, -5, -10, -5, 0, 5]], dtype=float), GateShape.CIRCLE),
    'fig8': (np.array([[0, 5, 0, -5, 0, 5, 0, -5], [0, 1, 2, 1, 0, -1, -2, -1],
                       [10, 5, 0, -5, -10, -5, 0, 5]], dtype=float), GateShape.CIRCLE),
    'race': (np.array([[-1.1, 9.2, 9.2, -4.5, -4.5, 4.75, -2.8], [-1.6, 6.6, -4, -6, -6, -0.9, 6.8],
                       [3.6, 1.0, 1.2, 3.5, 0.8, 1.2, 1.2]]), GateShape.SQUARE),
    'obs': (np.array([[-5, -2.75, -0.66, 2.95, 8.67, 9.2, 1.57, -2.39, -4.7, -2.39, 4.23, -2.66],
                      [4.5, -0.08, -1.36, 1.25, 6.69, -3.6, -6.43, -6, -6.43, -6.23, -0.66, 6.66],
                      [1.2, 2.815, 3.9, 2.815, 1.0, 1.0, 2.815, 3.9, 2.815, 1.0, 1.0, 1.0]]), GateShape.CIRCLE),
}


def make_line(track, cls=SplineCenterline):
    x, shape = TRACKS[track]
    cfg = SplineCenterlineConfig(x=x.copy())
    cfg.closed = True
    cfg.gate_shape = shape
    for k, v in TRACK_GATES.get(track, {}).items():
        setattr(cfg, k, v)
    line = cls(cfg)
    if track == 'obs':
        line.config.gate_s = None        # scripts/obstacles.py:22-23
    return line


def make_skew_line(track, cls=SplineCenterline, twist=0.6):
    '''
    the same waypoints with a provided lateral direction that twists about the tangent along the lap, so that the
    frame at s_max is rotated against the frame at s_min: a SKEWLY closed centerline (base_raceline.py:1208-1227)
    '''
    base = make_line(track)
    s = np.asarray(base.config.s, dtype=float)
    ry = []
    for sk in s:
        ey, en = base.p2ey(min(sk, base.s_max() - 1e-9)), base.p2en(min(sk, base.s_max() - 1e-9))
        th = twist * (sk - s[0]) / (s[-1] - s[0])
        ry.append(np.cos(th) * ey + np.sin(th) * en)
    x, shape = TRACKS[track]
    cfg = SplineCenterlineConfig(x=x.copy(), ry=np.array(ry))
    cfg.closed = True
    cfg.gate_shape = shape
    line = cls(cfg)
    assert not line.cleanly_closed
    return line


def synthetic_tube_arrays(s):
    ''' stand-in tube while trimesh is unavailable '''
    s = np.asarray(s, dtype=float)
    ball_p = np.stack([s, 0.3 * np.sin(2 * np.pi * s / 12), 0.2 * np.cos(2 * np.pi * s / 12)], axis=1)
    return ball_p, 0.9 * np.ones(len(s))


# name -> (track, frame, vehicle, rk4, N full, N small, quat, tube)
CASES = {
    'fig8_global_colloc_drone': ('fig8', 'global', 'drone', False, 50, 8, True, False),      # C1
    'fig8_global_colloc_point': ('fig8', 'global', 'point', False, 50, 8, True, False),      # WS of C1
    'fig8_param_colloc_drone': ('fig8', 'parametric', 'drone', False, 50, 8, True, False),
    'fig8_param_colloc_drone_euler': ('fig8', 'parametric', 'drone', False, 50, 8, False, False),
    'fig8_param_colloc_point': ('fig8', 'parametric', 'point', False, 50, 8, True, False),
    'race_global_rk4_drone': ('race', 'global', 'drone', True, 70, 7, True, False),
    'race_global_rk4_point': ('race', 'global', 'point', True, 70, 7, True, False),
    'race_param_rk4_drone': ('race', 'parametric', 'drone', True, 70, 7, True, False),       # C2 / C5
    'race_param_rk4_point': ('race', 'parametric', 'point', True, 70, 7, True, False),       # WS of C2
    'race_param_rk4_drone_euler': ('race', 'parametric', 'drone', True, 70, 7, False, False),
    'obs_param_colloc_drone': ('obs', 'parametric', 'drone', False, 100, 12, True, True),    # C3
    'obs_param_colloc_point': ('obs', 'parametric', 'point', False, 100, 12, True, True),    # WS of C3
    # C4 stand-in: the reference holds no complementary-progress formulation, only its result
    # CSV; the fig-8 track with the narrow gates of scripts/fig_8_cpc.py, parametric collocation, N = 200
    'fig8cpc_param_colloc_drone': ('fig8cpc', 'parametric', 'drone', False, 200, 8, True, False),
}


# variants of the cases above with other vehicle settings: name -> (base case, vehicle keyword arguments).
# `global_r=False`: orientation relative to the centerline frame (drone3d/dynamics/drone_models.py:249-292,
# point_model.py:149-213; exposed by drone3d/utils/solve_util.py:11-24); b != 0: linear drag changes the pattern.
VARIANT_CASES = {
    'fig8_param_colloc_drone_lr': ('fig8_param_colloc_drone', dict(global_r=False)),
    'race_param_rk4_drone_lr': ('race_param_rk4_drone', dict(global_r=False)),
    'race_param_rk4_drone_euler_lr': ('race_param_rk4_drone_euler', dict(global_r=False)),
    'race_param_rk4_point_lr': ('race_param_rk4_point', dict(global_r=False)),
    'fig8_param_colloc_point_lr': ('fig8_param_colloc_point', dict(global_r=False)),
    'race_param_rk4_drone_drag': ('race_param_rk4_drone', dict(b1=0.05, b2=0.07, b3=0.02)),
    'fig8_global_colloc_drone_drag': ('fig8_global_colloc_drone', dict(b1=0.05, b2=0.07, b3=0.02)),
}


# open (non-periodic) racelines, global frame (base_raceline.py:516-543, drone_raceline.py:110-148,
# point_raceline.py:15-45): initial / terminal rows and the gate on the end state instead of the loop closure
OPEN_CASES = {
    'race_global_rk4_point_open': 'race_global_rk4_point',
    'race_global_rk4_drone_open': 'race_global_rk4_drone',
    'fig8_global_colloc_point_open': 'fig8_global_colloc_point',
    'fig8_global_colloc_drone_open': 'fig8_global_colloc_drone',
    'race_global_rk4_drone_euler_open': 'race_global_rk4_drone',
}


# skewly closed centerlines (base_raceline.py:1208-1227): parametric point-mass racelines on make_skew_line tracks
SKEW_CASES = {
    'race_param_rk4_point_skew': 'race_param_rk4_point',
    'fig8_param_colloc_point_skew': 'fig8_param_colloc_point',
}


def _closed(name):
    return name not in OPEN_CASES


def _resolve(name, vehicle_kw):
    if name in OPEN_CASES:
        name = OPEN_CASES[name]
    if name in SKEW_CASES:
        name = SKEW_CASES[name]
    if name in VARIANT_CASES:
        base, kw = VARIANT_CASES[name]
        return base, {**kw, **(vehicle_kw or {})}
    return name, vehicle_kw


def vehicle_config(vehicle, quat=True, tube=False, **kw):
    rc = dict(collision_radius=0.4) if tube else {}
    kw = {'global_r': True, **kw}
    if vehicle == 'drone':
        return DroneConfig(use_quat=quat, **rc, **kw)
    return PointConfig(**rc, **kw)


def build_product(name, small=False, N=None, vehicle_kw=None):
    closed = _closed(name)
    quat_override = False if name.endswith('_euler_open') else None
    skew = name in SKEW_CASES
    name, vehicle_kw = _resolve(name, vehicle_kw)
    track, frame, vehicle, rk4, n_full, n_small, quat, tube = CASES[name]
    quat = quat if quat_override is None else quat_override
    N = N or (n_small if small else n_full)
    line = make_skew_line(track) if skew else make_line(track)
    vc = vehicle_config(vehicle, quat, tube, **(vehicle_kw or {}))
    if frame == 'global':
        cfg = RL.GlobalRacelineConfig(N=N, use_rk4=rk4, closed=closed, verbose=False,
                                      gate_xi=line.config.x[0], gate_xj=line.config.x[1],
                                      gate_xk=line.config.x[2])
        cls = RL.GlobalDroneRaceline if vehicle == 'drone' else RL.GlobalPointRaceline
        args = (line, cfg, vc)
    else:
        cfg = RL.ParametricRacelineConfig(N=N, use_rk4=rk4, closed=True, verbose=False)
        if track != 'obs':
            cfg.fixed_gates = line.config.s[:-1]          # drone3d/utils/solve_util.py:56-57
        if tube:
            K = 7
            tau = RL.get_collocation_coefficients(K)[0]
            ds = (line.s_max() - line.s_min()) / N
            s_all = np.array([line.s_min() + ds * (n + tau[k]) for n in range(N) for k in range(K + 1)])
            bp, br = synthetic_tube_arrays(s_all)
            t = RL.ObstacleFreeTube(bp, br, vc.collision_radius)
            cls = RL.ParametricObstacleDroneRaceline if vehicle == 'drone' else RL.ParametricObstaclePointRaceline
            args = (line, cfg, vc, None, t)
        else:
            cls = RL.ParametricDroneRaceline if vehicle == 'drone' else RL.ParametricPointRaceline
            args = (line, cfg, vc)
    if vehicle == 'drone':
        return cls(*args, generate_ws=False)
    return cls(*args)


def build_oracle(name, small=False, N=None, vehicle_kw=None, ws=None):
    from oracle.ref_centerline import RefSplineCenterline
    from oracle.ref_raceline import RefRaceline, RefTube
    closed = _closed(name)
    quat_override = False if name.endswith('_euler_open') else None
    skew = name in SKEW_CASES
    name, vehicle_kw = _resolve(name, vehicle_kw)
    track, frame, vehicle, rk4, n_full, n_small, quat, tube = CASES[name]
    quat = quat if quat_override is None else quat_override
    N = N or (n_small if small else n_full)
    line = make_skew_line(track, RefSplineCenterline) if skew else make_line(track, RefSplineCenterline)
    vc = vehicle_config(vehicle, quat, tube, **(vehicle_kw or {}))
    cfg = NS(N=N, K=7, use_rk4=rk4, R=1e-7, dR=1e-7, h0=1, v0=1, closed=closed, fix_gate_center=False)
    rt = None
    if frame == 'global':
        cfg.gate_xi, cfg.gate_xj, cfg.gate_xk = line.config.x
    else:
        cfg.fixed_gates = line.config.s[:-1] if track != 'obs' else None
        cfg.force_regularity = True
        if tube:
            from oracle.ref_discretization import get_collocation_coefficients
            tau = get_collocation_coefficients(7)[0]
            ds = (line.s_max() - line.s_min()) / N
            s_all = np.array([line.s_min() + ds * (n + tau[k]) for n in range(N) for k in range(8)])
            bp, br = synthetic_tube_arrays(s_all)
            rt = RefTube(bp, br, vc.collision_radius)
    return RefRaceline(line, cfg, vc, frame, vehicle, tube=rt, ws=ws)


def build_case(name, small=False, N=None, vehicle_kw=None):
    return build_product(name, small, N, vehicle_kw), build_oracle(name, small, N, vehicle_kw)


def eval_point(st, seed):
    ''' SURVEY.md s8d eval-point distribution '''
    rng = np.random.default_rng(seed)
    x = np.clip(st.w0 + 1e-2 * rng.standard_normal(st.nw), st.lbw, st.ubw)
    lam = rng.standard_normal(st.ng)
    return x, lam
