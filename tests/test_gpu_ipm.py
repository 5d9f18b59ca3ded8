'''
GPU: converged raceline solves through the product API (builders -> nlpsol-shaped solver -> batched
interior-point driver -> CUDA evaluation + KKT kernels), checked against
  * the golden solutions the same driver produced on the CPU oracle backend (tests/golden/make_golden_ipm.py):
    lap time within 1e-6 relative (BASELINE.json north_star tolerance for converged results),
  * the first-order optimality conditions, evaluated independently with the oracle's functions.
'''
import os

import numpy as np
import pytest

from cases import build_case, build_product, make_line

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
LAP_RTOL = 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize('name,N', [('race_param_rk4_point', 7), ('race_global_rk4_point', 7),
                                    ('fig8_global_colloc_point', 8)])
def test_point_mass_solve_matches_cpu_golden(name, N, built_library):
    from oracle.nlp_functions import OracleNLP
    from test_ipm_cpu import _kkt_conditions
    prod, ref = build_case(name, N=N)
    st = prod.structure
    res = prod.solve()
    stats = prod.solver.stats()
    assert stats['success'] and stats['return_status'] == 'Solve_Succeeded'
    gold = np.load(os.path.join(GOLD, f'ipm_{name}_N{N}.npz'))
    assert abs(res.time - float(gold['lap'])) <= LAP_RTOL * float(gold['lap'])
    assert np.abs(prod.sol['x'] - gold['x']).max() <= 1e-5
    _kkt_conditions(st, OracleNLP(ref, build_hess=False), prod.sol['x'], prod.sol['lam_g'], prod.sol['lam_x'])
    assert res.feasible and len(res.states) == st.N * (st.K + 1) and abs(sum(res.step_sizes) - res.time) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('name,N', [('race_global_rk4_point_open', 7), ('fig8_global_colloc_point_open', 8)])
def test_open_track_point_mass_solves(name, N, built_library):
    ''' open racelines: the GPU path converges to the minimum the driver finds on the CPU oracle
    backend (tests/test_open_tracks.py), a KKT point of the oracle's NLP '''
    from oracle.nlp_functions import OracleNLP
    from test_ipm_cpu import _kkt_conditions
    from test_open_tracks import OPEN_LAPS
    prod, ref = build_case(name, N=N)
    res = prod.solve()
    stats = prod.solver.stats()
    assert stats['success'] and stats['return_status'] == 'Solve_Succeeded'
    assert abs(res.time - OPEN_LAPS[(name, N)]) <= LAP_RTOL * res.time
    _kkt_conditions(prod.structure, OracleNLP(ref, build_hess=False), prod.sol['x'], prod.sol['lam_g'], prod.sol['lam_x'])
    assert res.feasible and not res.periodic


@pytest.mark.gpu
def test_batched_multistart_solves(built_library):
    ''' B instances in lock step: identical starts give identical answers, perturbed starts still converge '''
    prod = build_product('race_param_rk4_point', N=7)
    st = prod.structure
    prod.solver.verbose = False
    rng = np.random.default_rng(0)
    B = 6
    X0 = np.tile(st.w0, (B, 1))
    X0[3:] += 0.02 * rng.standard_normal((B - 3, st.nw))
    sol = prod.solver(x0=X0, lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg)
    s = prod.solver.stats()
    assert s['success_each'].all()
    assert np.array_equal(sol['x'][0], sol['x'][1]) and np.array_equal(sol['x'][0], sol['x'][2])
    gold = np.load(os.path.join(GOLD, 'ipm_race_param_rk4_point_N7.npz'))
    laps = sol['x'][:, :st.N].sum(1)
    assert abs(laps[0] - float(gold['lap'])) <= LAP_RTOL * float(gold['lap'])
    g = sol['g']
    assert (g >= st.lbg - 1e-6).all() and (g <= st.ubg + 1e-6).all()


@pytest.mark.gpu
def test_drone_warm_start_chain(built_library):
    ''' scripts/race.py pattern at reduced size: point-mass solve -> drone initial guess -> drone solve '''
    from aircraft_trajectory_optimization_b200 import raceline as RL
    from aircraft_trajectory_optimization_b200.pytypes import DroneConfig
    line = make_line('race')
    # (N = 7, i.e. 49 shooting intervals, is too coarse for this track: the iteration path becomes erratic -- 250 to
    # 1000 iterations depending on rounding; 98 intervals converge in ~90 iterations on either implementation of the sweep)
    cfg = RL.ParametricRacelineConfig(N=14, use_rk4=True, closed=True, verbose=False)
    cfg.fixed_gates = line.config.s[:-1]
    solver = RL.ParametricDroneRaceline(line, cfg, DroneConfig(global_r=True, use_quat=True))
    assert solver.ws_solver.solver.stats()['success']
    res = solver.solve()
    st = solver.structure
    assert solver.solver.stats()['success'] and res.feasible
    g = solver.sol['g']
    assert (g >= st.lbg - 1e-6).all() and (g <= st.ubg + 1e-6).all()
    assert 3.0 < res.time < solver.ws_raceline.time + 1.0
    # quaternions on the unit sphere at every node after the first (continuity renormalises)
    q = np.array([s.q.to_vec() for s in res.states[1:]]) if hasattr(res.states[0], 'q') else None
    if q is not None:
        assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 1e-6


@pytest.mark.gpu
def test_speculative_candidates_and_batch_composition_do_not_change_results(built_library):
    ''' the spare slots of a factorisation wave try further delta_w candidates: same iterates, fewer sweeps; and an
    instance's answer does not depend on which other instances share its batch '''
    from aircraft_trajectory_optimization_b200.ipm import IpmOptions
    prod = build_product('race_global_rk4_point', N=7)
    st = prod.structure
    prod.solver.verbose = False
    rng = np.random.default_rng(1)
    X0 = np.tile(st.w0, (3, 1))
    X0[1] += 0.05 * rng.standard_normal(st.nw)
    X0[2] += 0.10 * rng.standard_normal(st.nw)
    kw = dict(lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg)
    prod.solver.options = IpmOptions(speculate=0)
    base = prod.solver(x0=X0, **kw)
    r0 = prod.solver.result
    assert r0.n_speculated == 0 and (r0.factorisations_each > r0.iterations.cpu().numpy()).any()
    prod.solver.options = IpmOptions(speculate=2)
    spec = prod.solver(x0=X0, **kw)
    r1 = prod.solver.result
    assert r1.n_speculated > 0 and r1.n_factor < r0.n_factor
    assert np.array_equal(r0.factorisations_each, r1.factorisations_each)
    for k in ('x', 'lam_g', 'lam_x'):
        assert np.array_equal(base[k], spec[k]), k
    # the hardest start alone gives the same answer as inside the batch
    alone = prod.solver(x0=X0[2], **kw)
    assert np.array_equal(alone['x'], spec['x'][2])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['race_global_rk4_point', 'race_param_rk4_point'])
def test_fused_sweep_kernels_match_the_torch_arithmetic(name, built_library):
    ''' csrc/ipm_glue.cuh against the same formulas in torch (ipm.py, the path the CPU tests exercise): same iteration
    counts and statuses, solutions equal to rounding; a window + compaction run (evaluation data must
    survive compaction) gives the same answers as the plain run '''
    from aircraft_trajectory_optimization_b200.ipm import IpmOptions
    prod = build_product(name, N=7)
    st = prod.structure
    prod.solver.verbose = False
    rng = np.random.default_rng(2)
    B = 6
    X0 = np.tile(st.w0, (B, 1))
    X0[1:] += 0.03 * rng.standard_normal((B - 1, st.nw)) * (np.abs(st.w0) > 0)
    X0 = np.clip(X0, st.lbw, st.ubw)
    kw = dict(lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg)
    out = {}
    for key, opts in (('torch', IpmOptions(use_glue=False, compact=False)), ('glue', IpmOptions(use_glue=True, compact=False)),
                      ('glue_compact', IpmOptions(use_glue=True, compact=True, compact_min=1, compact_frac=0.7, window=3))):
        prod.solver.options = opts
        sol = prod.solver(x0=X0, **kw)
        out[key] = (sol, prod.solver.result.status.cpu().numpy().copy(), prod.solver.result.iterations.cpu().numpy().copy())
    l0 = prod.functions.launch_count()
    # the hard starts (hundreds of iterations) are sensitive to the summation order of the reductions -- in the glue
    # kernels and in the KKT solve kernels alike: one of them may take another path, even run out of iterations
    conv_t = out['torch'][1] <= 1
    assert conv_t.sum() >= B - 1, out['torch'][1]
    for key in ('glue', 'glue_compact'):
        conv_k = out[key][1] <= 1
        assert conv_k.sum() >= B - 1, (key, out[key][1])
        # instances that converge quickly follow the same path to rounding; the perturbed starts may end in another of
        # the track's local minima (laps 5.797 .. 5.820 s) when a path forks: compared are the instances that reach the
        # same minimum in about the same number of iterations
        laps_t = out['torch'][0]['x'][:, :st.N].sum(1)
        laps_k = out[key][0]['x'][:, :st.N].sum(1)
        same = (np.abs(out[key][2] - out['torch'][2]) <= 2) & conv_t & conv_k & (np.abs(laps_k - laps_t) <= 1e-6)
        assert np.array_equal(out[key][1][same], out['torch'][1][same]), key
        assert same[0] and same.sum() >= 2, (key, out[key][2], out["torch"][2])
        for k in ('x', 'lam_g'):
            scale = max(1.0, np.abs(out['torch'][0][k]).max())
            assert np.abs(out[key][0][k][same] - out['torch'][0][k][same]).max() <= 1e-6 * scale, (key, k)
        laps = out[key][0]['x'][:, :st.N].sum(1)
        assert np.abs(laps - out['torch'][0]['x'][:, :st.N].sum(1))[same].max() <= 1e-8 * laps.max()
    # the compacted / windowed run reproduces the plain fused run exactly (per-instance arithmetic is independent)
    assert np.array_equal(out['glue'][0]['x'], out['glue_compact'][0]['x'])
    assert l0 > 0


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['race_param_rk4_drone', 'fig8_global_colloc_drone'])
def test_fused_sweep_kernels_first_iterations_on_drone_problems(name, built_library):
    ''' the drone NLPs are non-convex and their iteration paths are sensitive to rounding, so the two implementations of
    the sweep are compared over the first iterations from the same start: iterates equal to 1e-9 '''
    from aircraft_trajectory_optimization_b200.ipm import IpmOptions
    prod = build_product(name, N=7 if 'rk4' in name else 8)
    st = prod.structure
    prod.solver.verbose = False
    rng = np.random.default_rng(3)
    X0 = np.clip(np.tile(st.w0, (3, 1)) + 0.01 * rng.standard_normal((3, st.nw)), st.lbw, st.ubw)
    kw = dict(lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg)
    sols = {}
    for glue in (False, True):
        prod.solver.options = IpmOptions(use_glue=glue)
        prod.solver.max_iter = 8
        sols[glue] = prod.solver(x0=X0, **kw)
    for k in ('x', 'lam_g', 'lam_x', 'g'):
        scale = max(1.0, np.abs(sols[False][k]).max())
        assert np.abs(sols[True][k] - sols[False][k]).max() <= 1e-9 * scale, k


@pytest.mark.gpu
def test_second_order_correction_reaches_the_same_minimum(built_library):
    ''' IPOPT's second-order correction (IpmOptions.max_soc, off by default): corrected steps are taken on a perturbed
    multi-start batch and the results are KKT points '''
    from aircraft_trajectory_optimization_b200.ipm import IpmOptions
    prod = build_product('race_param_rk4_point', N=7)
    st = prod.structure
    prod.solver.verbose = False
    rng = np.random.default_rng(3)
    X0 = np.clip(np.tile(st.w0, (6, 1)) + 0.02 * rng.standard_normal((6, st.nw)) * np.maximum(np.abs(st.w0), 0.1), st.lbw, st.ubw)
    X0[0] = st.w0
    kw = dict(lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg)
    prod.solver.options = IpmOptions(max_soc=0)
    base = prod.solver(x0=X0, **kw)
    ok0 = prod.solver.stats()['success_each']
    prod.solver.options = IpmOptions(max_soc=4)
    soc = prod.solver(x0=X0, **kw)
    ok1 = prod.solver.stats()['success_each']
    assert prod.solver.stats()['n_soc'] > 0
    assert ok0.sum() >= 5 and ok1.sum() >= 5
    lap0, lap1 = base['x'][:, :st.N].sum(1), soc['x'][:, :st.N].sum(1)
    # this coarse, non-convex problem has several neighbouring local minima (laps 5.80 .. 5.82): the corrected steps
    # change the path and may end in another one of them (nominal start: 5.8204 without, 5.8108 with the correction), so
    # the check is that every result is a KKT point of the same NLP, verified independently with the oracle's functions
    assert np.abs(lap1[ok1] - lap0[0]).max() < 0.05 and np.abs(lap0[ok0] - lap0[0]).max() < 0.05
    from oracle.nlp_functions import OracleNLP
    from test_ipm_cpu import _kkt_conditions
    from cases import build_oracle
    nlp = OracleNLP(build_oracle('race_param_rk4_point', N=7), build_hess=False)
    for b in np.nonzero(ok1)[0][:3]:
        _kkt_conditions(st, nlp, soc['x'][b], soc['lam_g'][b], soc['lam_x'][b])
