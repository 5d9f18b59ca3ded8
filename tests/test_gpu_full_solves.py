'''
GPU: converged solves at BASELINE.json's FULL sizes, checked INDEPENDENTLY of the product's own functions: the
first-order optimality conditions of the returned (x, lam_g, lam_x) are evaluated with the CPU oracle's nlp_grad_f /
nlp_jac_g of the SAME warm-started problem (the oracle is built from the GPU's point-mass solution through its own
restatement of the reference's warm-start mapping, drone3d/raceline/drone_raceline.py:158-277, so the closure rows
-- quaternion sign / yaw wraps, :47-104 -- and the start point are the reference's).

  C2: scripts/race.py:42-49   parametric quaternion drone, RK4, N = 490 (solve_util N = 70 x 7 gates)
  C1: scripts/fig_8.py:55-62  global-frame quaternion drone, Legendre collocation N = 56, K = 7
  C3: scripts/obstacles.py:14-44  parametric quaternion drone inside the (synthetic) obstacle-free tube, N = 100, K = 7

The reference's own golden for these numbers would be an IPOPT run (CasADi is not in the image: parity unpinned).
'''
import numpy as np
import pytest

from cases import CASES, build_oracle, make_line
from test_ipm_cpu import _kkt_conditions


def _solve_obstacles():
    ''' scripts/obstacles.py:14-44 with the synthetic tube of SURVEY.md s8d (C3) '''
    from cases import synthetic_tube_arrays, vehicle_config
    from aircraft_trajectory_optimization_b200 import raceline as RL
    line = make_line('obs')
    cfg = RL.ParametricRacelineConfig(N=100, use_rk4=False, closed=True, verbose=False)
    vc = vehicle_config('drone', True, True)
    tau = RL.get_collocation_coefficients(cfg.K)[0]
    ds = (line.s_max() - line.s_min()) / cfg.N
    s_all = np.array([line.s_min() + ds * (n + tau[k]) for n in range(cfg.N) for k in range(cfg.K + 1)])
    tube = RL.ObstacleFreeTube(*synthetic_tube_arrays(s_all), vc.collision_radius)
    solver = RL.ParametricObstacleDroneRaceline(line, cfg, vc, None, tube, generate_ws=True)
    return solver, solver.solve()


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['race_param_rk4_drone', 'fig8_global_colloc_drone', 'obs_param_colloc_drone'])
def test_full_size_drone_solve_is_a_kkt_point_of_the_oracle_nlp(name, built_library):
    from oracle.nlp_functions import OracleNLP
    from aircraft_trajectory_optimization_b200.solve_util import solve_util
    track, frame, vehicle, rk4, n_full, _, quat, tube = CASES[name]
    line = make_line(track)
    if tube:
        solver, res = _solve_obstacles()
    else:
        solver, res = solve_util(line, global_frame=(frame == 'global'), drone=True, use_quaternion=quat, use_ws=True,
                                 use_rk4=rk4, N=n_full, verbose=False)
    st = solver.structure
    stats = solver.solver.stats()
    assert stats['success'], stats['return_status']
    assert solver.ws_solver.solver.stats()['success']
    # the oracle's restatement of the same warm-started NLP
    pm = build_oracle(name.replace('_drone', '_point'))
    ref = build_oracle(name, ws=pm.warmstart_from(np.asarray(solver.ws_solver.sol['x'])))
    nlp = OracleNLP(ref, build_hess=False)
    assert (nlp.nw, nlp.ng) == (st.nw, st.ng)
    assert np.array_equal(nlp.jac_colind, st.jac_colind) and np.array_equal(nlp.jac_row, st.jac_row)
    assert np.abs(ref.w0 - st.w0).max() <= 1e-9 * max(1.0, np.abs(ref.w0).max())
    x, lam_g, lam_x = (np.asarray(solver.sol[k], dtype=float) for k in ('x', 'lam_g', 'lam_x'))
    _kkt_conditions(st, nlp, x, lam_g, lam_x)
    # the scaled stationarity residual IPOPT's tol = 1e-8 refers to (s_d = max(100, mean |multipliers|) / 100)
    _, gf = nlp.nlp_grad_f(x)
    _, jv = nlp.nlp_jac_g(x)
    stat = gf + nlp.jac_csc(jv).T @ lam_g + lam_x
    s_d = max(100.0, (np.abs(lam_g).sum() + np.abs(lam_x).sum()) / (len(lam_g) + len(lam_x))) / 100.0
    print(f'{name}: lap {res.time:.6f} s, {stats.get("iter_count")} iterations, |grad L|_inf = {np.abs(stat).max():.2e} '
          f'(scaled {np.abs(stat).max() / s_d:.2e})')
    assert np.abs(stat).max() / s_d <= 1e-7
    assert abs(float(np.sum(x[:st.N])) - res.time) <= 1e-12
