'''
GPU parity of the batched warm-start chain (csrc/warm_start.cuh through rb_ws_drone_guess) against the oracle's
point-by-point restatement of DroneRaceline._guess_z / _guess_u (oracle/ref_raceline.py:524-586, following
drone3d/raceline/drone_raceline.py:158-277 with scipy Rotation): B point-mass "solutions" (a smooth lap around the
track, perturbed per instance) -> B drone initial guesses.  fp64, tolerance 1e-12 relative.
'''
import numpy as np
import pytest

from cases import CASES, build_oracle

# (drone case whose guess is built, point-mass case that provides the warm start)
PAIRS = [('race_param_rk4_drone', 'race_param_rk4_point'), ('race_global_rk4_drone', 'race_global_rk4_point'),
         ('fig8_global_colloc_drone', 'fig8_global_colloc_point'), ('fig8_param_colloc_drone', 'fig8_param_colloc_point'),
         ('fig8_param_colloc_drone_euler', 'fig8_param_colloc_point'), ('fig8_param_colloc_drone_lr', 'fig8_param_colloc_point_lr')]


def _point_mass_laps(pm, B, seed=0):
    ''' plausible point-mass laps: the oracle's own initial guess (velocity along the track) with thrust that holds the
    vehicle against gravity plus smooth per-instance variations, so that T x dT, the yaw and the quaternion sign all move '''
    rng = np.random.default_rng(seed)
    N, P = pm.config.N, (pm.config.K + 1)
    H, Z, U, dU = pm.unpack(pm.w0)
    out = []
    tt = np.linspace(0, 2 * np.pi, N * P, endpoint=False).reshape(N, P)
    for b in range(B):
        a = rng.uniform(0.5, 2.0, 6) * (0.1 if b < 2 else 1.0)
        ph = rng.uniform(0, 2 * np.pi, 6)
        Zb, Ub, dUb = Z.copy(), U.copy(), dU.copy()
        Zb[..., 3:6] = Z[..., 3:6] * (2.0 + np.sin(tt + ph[0]))[..., None] + (0.03 if b < 2 else 0.3) * np.stack(
            [np.sin(2 * tt + ph[1]), np.cos(3 * tt + ph[2]), np.sin(tt + ph[3])], -1)
        Ub[...] = np.stack([a[0] * np.sin(2 * tt + ph[3]), a[1] * np.cos(tt + ph[4]), 9.81 + a[2] * np.sin(3 * tt + ph[5])], -1)
        dUb[...] = np.stack([2 * a[0] * np.cos(2 * tt + ph[3]), -a[1] * np.sin(tt + ph[4]), 3 * a[2] * np.cos(3 * tt + ph[5])], -1)
        Hb = H * rng.uniform(0.8, 1.2, N)
        out.append(np.concatenate([Hb, np.concatenate([Zb, Ub, dUb], -1).reshape(-1)]))
    return np.stack(out)


@pytest.mark.gpu
@pytest.mark.parametrize('drone_case,pm_case', PAIRS)
def test_batched_drone_guess_matches_the_point_by_point_mapping(drone_case, pm_case, built_library):
    from aircraft_trajectory_optimization_b200.warm_start import drone_guess_batch
    pm = build_oracle(pm_case, small=True)
    B = 5
    Wpm = _point_mass_laps(pm, B)
    ref0 = build_oracle(drone_case, small=True)
    N, K = ref0.config.N, ref0.config.K
    quat, global_r = ref0.model.config.use_quat, ref0.model.config.global_r
    fc = None
    if ref0.parametric:
        s_all = np.array([ref0._get_s(n, k) for n in range(N) for k in range(K + 1)])
        Rp = np.stack([np.asarray(ref0.line.p2Rp(s)) for s in s_all]).reshape(-1, 9)
        fc = np.concatenate([Rp, np.zeros((len(s_all), 4))], axis=1)
    w0, info = drone_guess_batch(Wpm, N, K, quat=quat, closed=True, global_r=global_r, fc=fc)
    w0, info = w0.cpu().numpy(), info.cpu().numpy()
    n_failed = 0
    for b in range(B):
        # the oracle builds the drone NLP around this warm start; its w0 is the reference guess
        try:
            ref = build_oracle(drone_case, small=True, ws=pm.warmstart_from(Wpm[b]))
        except NotImplementedError:
            # 'Warmstart continuity failed for euler angles' (drone_raceline.py:236-238): the kernel reports it
            assert not quat and info[b, 2] > 0
            n_failed += 1
            continue
        assert ref.w0.shape == w0[b].shape
        err = np.abs(w0[b] - ref.w0).max() / max(1.0, np.abs(ref.w0).max())
        assert err <= 1e-12, (drone_case, b, err)
        if quat:
            flipped = np.linalg.norm(ref._first_ws_r - ref._last_ws_r) > 1
            assert bool(info[b, 0]) == bool(flipped)
        else:
            assert info[b, 1] == int(np.round((ref._last_ws_r - ref._first_ws_r)[0] / 2 / np.pi))
        assert info[b, 2] == 0
    assert n_failed < B


@pytest.mark.gpu
def test_nominal_guess_and_batched_chain(built_library):
    ''' (1) the batched mapping of the nominal warm-start solution reproduces the start point the builder made point
    by point on the host (raceline.py::DroneRaceline._guess_z); (2) solve_batch: every vehicle variant is warm-started
    from its own point-mass solve and converges to a lap close to the nominal one '''
    from cases import make_line
    from aircraft_trajectory_optimization_b200 import raceline as RL
    from aircraft_trajectory_optimization_b200.pytypes import DroneConfig
    from aircraft_trajectory_optimization_b200.models import vehicle_params
    line = make_line('race')
    cfg = RL.ParametricRacelineConfig(N=14, use_rk4=True, closed=True, verbose=False)
    cfg.fixed_gates = line.config.s[:-1]
    solver = RL.ParametricDroneRaceline(line, cfg, DroneConfig(global_r=True, use_quat=True))
    st = solver.structure
    w0, info = solver.guess_batch(solver.ws_solver.sol['x'])
    w0 = w0.cpu().numpy()[0]
    assert np.abs(w0 - st.w0).max() <= 1e-9 * max(1.0, np.abs(st.w0).max())
    vp0 = vehicle_params(solver.vehicle_config)
    B = 4
    VP = np.tile(vp0, (B, 1))
    VP[1:, 0] *= np.array([0.9, 1.05, 1.1])          # mass
    VP[1:, 2:5] *= np.array([[1.1], [0.95], [0.9]])  # inertia
    sol = solver.solve_batch(VP)
    assert sol['ws_success'].all() and sol['success'].all() and not sol['closure_mismatch'].any()
    nominal = solver.solve()
    assert abs(sol['lap_time'][0] - nominal.time) <= 1e-6 * nominal.time
    # heavier vehicles are slower with the same rotor thrust limits
    assert sol['lap_time'][3] > sol['lap_time'][0] > sol['lap_time'][1]
    assert np.abs(sol['lap_time'] - nominal.time).max() < 0.5
