'''
The interior-point driver (aircraft_trajectory_optimization_b200/ipm.py) on the CPU oracle backend:
convergence on a small race.py point-mass instance, the committed golden solution, and an independent
check of the first-order optimality conditions with the oracle's own functions.
'''
import os

import numpy as np
import pytest
import torch

from cases import build_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _kkt_conditions(st, nlp, x, lam_g, lam_x, tol=1e-6):
    ''' CasADi sign convention: grad f + J' lam_g + lam_x = 0, lam >= 0 at upper / <= 0 at lower bounds '''
    f, gf = nlp.nlp_grad_f(x)
    g, jv = nlp.nlp_jac_g(x)
    J = nlp.jac_csc(jv)
    stat = gf + J.T @ lam_g + lam_x
    assert np.abs(stat).max() <= tol * max(1.0, np.abs(lam_g).max() / 100)
    assert (g >= st.lbg - tol).all() and (g <= st.ubg + tol).all()
    assert (x >= st.lbw - tol).all() and (x <= st.ubw + tol).all()
    # complementarity: a multiplier is non-zero only at an active bound
    act_g = np.minimum(g - st.lbg, st.ubg - g)
    free_g, free_x = np.isinf(act_g), np.isinf(np.minimum(x - st.lbw, st.ubw - x))
    assert np.abs(lam_g[free_g]).max(initial=0.0) == 0 and np.abs(lam_x[free_x]).max(initial=0.0) == 0
    ineq = (st.lbg < st.ubg) & ~free_g
    assert np.abs(lam_g[ineq] * act_g[ineq]).max(initial=0.0) <= 1e-5
    act_x = np.minimum(x - st.lbw, st.ubw - x)
    assert np.abs(lam_x[~free_x] * act_x[~free_x]).max(initial=0.0) <= 1e-5


def test_point_mass_solve_matches_golden():
    from oracle.nlp_functions import OracleNLP
    from oracle.cpu_backend import OracleBackend
    from aircraft_trajectory_optimization_b200.ipm import InteriorPoint, IpmOptions
    from aircraft_trajectory_optimization_b200.kkt import build_kkt_structure
    name, N = 'race_global_rk4_point', 7
    prod, ref = build_case(name, N=N)
    st = prod.structure
    nlp = OracleNLP(ref)
    be = OracleBackend(nlp, nlp, ks=build_kkt_structure(st))
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=float))
    X0 = torch.stack([T(st.w0), T(st.w0)])           # two identical instances advance in lock step
    r = InteriorPoint(be, IpmOptions(max_iter=200)).solve(X0, T(st.lbw), T(st.ubw), T(st.lbg), T(st.ubg))
    assert r.success.all() and (r.status == 0).all()
    assert torch.equal(r.x[0], r.x[1])
    gold = np.load(os.path.join(GOLD, f'ipm_{name}_N{N}.npz'))
    lap = float(r.x[0, :st.N].sum())
    assert abs(lap - float(gold['lap'])) <= 1e-9 * lap
    assert np.abs(r.x[0].numpy() - gold['x']).max() <= 1e-7
    _kkt_conditions(st, nlp, r.x[0].numpy(), r.lam_g[0].numpy(), r.lam_x[0].numpy())


def test_golden_solutions_are_kkt_points():
    ''' the committed golden solutions satisfy the optimality conditions of the oracle's NLP '''
    from oracle.nlp_functions import OracleNLP
    for name, N in (('race_param_rk4_point', 7), ('race_global_rk4_point', 7), ('fig8_global_colloc_point', 8)):
        prod, ref = build_case(name, N=N)
        gold = np.load(os.path.join(GOLD, f'ipm_{name}_N{N}.npz'))
        _kkt_conditions(prod.structure, OracleNLP(ref, build_hess=False), gold['x'], gold['lam_g'], gold['lam_x'])


def test_compaction_and_window_do_not_change_results():
    ''' instances that finish early are compacted away / refilled from the queue; every instance's answer is unchanged '''
    from oracle.nlp_functions import OracleNLP
    from oracle.cpu_backend import OracleBackend
    from aircraft_trajectory_optimization_b200.ipm import InteriorPoint, IpmOptions
    from aircraft_trajectory_optimization_b200.kkt import build_kkt_structure
    prod, ref = build_case('race_global_rk4_point', N=7)
    st = prod.structure
    nlp = OracleNLP(ref)
    be = OracleBackend(nlp, nlp, ks=build_kkt_structure(st))
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=float))
    rng = np.random.default_rng(0)
    X0 = np.tile(st.w0, (3, 1))
    X0[1] += 0.02 * rng.standard_normal(st.nw)
    X0[2] += 0.05 * rng.standard_normal(st.nw)
    args = (T(X0), T(st.lbw), T(st.ubw), T(st.lbg), T(st.ubg))
    base = InteriorPoint(be, IpmOptions(max_iter=150, compact=False)).solve(*args)
    comp = InteriorPoint(be, IpmOptions(max_iter=150, compact=True, compact_min=1, compact_frac=0.7, window=2)).solve(*args)
    assert base.success.all() and len(set(base.iterations.tolist())) > 1      # they finish at different times
    assert torch.equal(base.status, comp.status) and torch.equal(base.iterations, comp.iterations)
    assert torch.equal(base.x, comp.x) and torch.equal(base.lam_g, comp.lam_g) and torch.equal(base.lam_x, comp.lam_x)


def test_speculative_regularisation_candidates_do_not_change_results():
    ''' spare slots of a factorisation wave try the next delta_w candidates: fewer sweeps, identical iterates '''
    from oracle.nlp_functions import OracleNLP
    from oracle.cpu_backend import OracleBackend
    from aircraft_trajectory_optimization_b200.ipm import InteriorPoint, IpmOptions
    from aircraft_trajectory_optimization_b200.kkt import build_kkt_structure
    prod, ref = build_case('race_global_rk4_point', N=7)
    st = prod.structure
    nlp = OracleNLP(ref)
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=float))
    rng = np.random.default_rng(1)
    X0 = np.tile(st.w0, (3, 1))
    X0[1] += 0.05 * rng.standard_normal(st.nw)
    X0[2] += 0.10 * rng.standard_normal(st.nw)
    args = (T(X0), T(st.lbw), T(st.ubw), T(st.lbg), T(st.ubg))
    be = OracleBackend(nlp, nlp, ks=build_kkt_structure(st))
    base = InteriorPoint(be, IpmOptions(max_iter=150, speculate=0)).solve(*args)
    assert base.n_speculated == 0
    # the delta_w escalation is exercised: some instance needed more factorisations than iterations
    assert (base.factorisations_each > base.iterations.numpy()).any()
    be.kkt_wave = 8                       # 3 instances in flight: 5 spare slots -> 2 instances x 2 candidates
    spec = InteriorPoint(be, IpmOptions(max_iter=150, speculate=2)).solve(*args)
    be.kkt_wave = 0
    assert spec.n_speculated > 0 and spec.n_factor < base.n_factor
    assert torch.equal(base.status, spec.status) and torch.equal(base.iterations, spec.iterations)
    assert np.array_equal(base.factorisations_each, spec.factorisations_each)
    assert torch.equal(base.x, spec.x) and torch.equal(base.lam_g, spec.lam_g) and torch.equal(base.lam_x, spec.lam_x)


@pytest.mark.parametrize('name,N,pert', [('race_global_rk4_point', 7, 1e-2), ('fig8_global_colloc_point', 8, 1e-3)])
def test_independent_solver_confirms_golden_minimum(name, N, pert):
    '''
    SURVEY.md s8c(5): an independent solver on the oracle's functions.  scipy's trust-constr (a different algorithm and
    code base: trust-region SQP / barrier with its own linear algebra), started from a perturbation of the committed
    golden solution, comes back to the same lap time and objective to 1e-6 relative -- the golden point is a local
    minimiser of the restated NLP, not merely a point where this package's driver stops.
    '''
    import scipy.optimize as so
    import scipy.sparse as sp
    from oracle.nlp_functions import OracleNLP
    prod, ref = build_case(name, N=N)
    st = prod.structure
    nlp = OracleNLP(ref)
    gold = np.load(os.path.join(GOLD, f'ipm_{name}_N{N}.npz'))

    def sym(hv):
        M = sp.csc_matrix((hv, nlp.hess_row, nlp.hess_colind), shape=(st.nw, st.nw))
        return M + sp.triu(M, 1).T

    cons = so.NonlinearConstraint(lambda x: nlp.nlp_jac_g(x)[0], st.lbg, st.ubg,
                                  jac=lambda x: nlp.jac_csc(nlp.nlp_jac_g(x)[1]),
                                  hess=lambda x, v: sym(nlp.nlp_hess_l(x, 0.0, v)))
    x0 = np.clip(gold['x'] + pert * np.random.default_rng(0).standard_normal(st.nw), st.lbw, st.ubw)
    res = so.minimize(lambda x: nlp.nlp_grad_f(x)[0], x0, jac=lambda x: nlp.nlp_grad_f(x)[1],
                      hess=lambda x: sym(nlp.nlp_hess_l(x, 1.0, np.zeros(st.ng))), method='trust-constr',
                      constraints=[cons], bounds=so.Bounds(st.lbw, st.ubw),
                      options=dict(gtol=1e-10, xtol=1e-12, barrier_tol=1e-10, maxiter=800, initial_barrier_parameter=0.1))
    assert res.status in (1, 2) and res.constr_violation <= 1e-8
    lap_gold = float(gold['lap'])
    assert abs(res.x[:st.N].sum() - lap_gold) <= 1e-6 * lap_gold
    f_gold = nlp.nlp_grad_f(gold['x'])[0]
    assert abs(res.fun - f_gold) <= 1e-6 * abs(f_gold)
