'''
Pins for the shared expression engine (aircraft_trajectory_optimization_b200.symbolic): its
elementary derivative rules and folding are checked against sympy and finite differences, so
that oracle and code generator do not merely agree with each other.
'''
import numpy as np
import pytest

from aircraft_trajectory_optimization_b200 import symbolic as sx
from aircraft_trajectory_optimization_b200.models import Variant, zdot, NFC
from aircraft_trajectory_optimization_b200.codegen import ALL_VARIANTS, PointFunctionGraphs


def test_folding_rules():
    g = sx.new_graph()
    x, y = sx.SX.sym('x'), sx.SX.sym('y')
    assert (0 * x).is_const() and (0 * x).value() == 0.0
    assert (x + 0).i == x.i and (1 * x).i == x.i and (x - 0).i == x.i
    assert (x - x).is_const() and (x / x).value() == 1.0
    assert (x * y).i == (y * x).i                      # hash-consing of commutative ops
    assert (-(-x)).i == x.i
    assert (-0.0 * x).is_const()                       # -0.0 folds like 0 (b1 = -0 drag terms)
    assert len(g) < 30


def test_elementary_partials_against_sympy():
    sympy = pytest.importorskip('sympy')
    g = sx.new_graph()
    a, b = sx.SX.sym('a'), sx.SX.sym('b')
    sa, sb = sympy.symbols('a b')
    exprs = [
        (a * b + a / b - b * b, sa * sb + sa / sb - sb * sb),
        (sx.sqrt(a * a + b * b) * sx.sin(a) - sx.cos(b) * sx.tan(a), sympy.sqrt(sa ** 2 + sb ** 2) * sympy.sin(sa)
         - sympy.cos(sb) * sympy.tan(sa)),
        ((a - b) ** 3 / (1 + a ** 2), (sa - sb) ** 3 / (1 + sa ** 2)),
    ]
    pt = {sa: 0.7, sb: -1.3}
    for e, se in exprs:
        wrt = {a.i: 0, b.i: 1}
        d1 = g.forward_sparse([e.i], wrt)[0]
        adj = g.reverse([e.i], [g.one])
        for k, (node, sym) in enumerate(((a, sa), (b, sb))):
            exact = float(sympy.diff(se, sym).subs(pt))
            fwd = g.evaluate([d1[k]], [0.7, -1.3])[0]
            rev = g.evaluate([adj[node.i]], [0.7, -1.3])[0]
            assert abs(fwd - exact) < 1e-12 * max(1, abs(exact))
            assert abs(rev - exact) < 1e-12 * max(1, abs(exact))
        # second derivatives through reverse-then-forward
        grads = [adj[a.i], adj[b.i]]
        h = g.forward_sparse(grads, wrt)
        for i, si in enumerate((sa, sb)):
            for j, sj in enumerate((sa, sb)):
                exact = float(sympy.diff(se, si, sj).subs(pt))
                got = g.evaluate([h[i].get(j, g.zero)], [0.7, -1.3])[0]
                assert abs(got - exact) < 1e-11 * max(1, abs(exact))


@pytest.mark.parametrize('variant', ALL_VARIANTS, ids=lambda v: v.name)
def test_generated_point_functions_against_finite_differences(variant):
    ''' J, W, jvp, hvp graphs of every model variant agree with central differences of f '''
    pf = PointFunctionGraphs(variant)
    g = pf.g
    rng = np.random.default_rng(0)
    nz, nu = variant.nz, variant.nu
    nx = nz + nu
    nvp = len(variant.vp_names)
    x0 = rng.standard_normal(nx) * 0.5
    if variant.vehicle == 'drone' and variant.orient == 'quat':
        x0[3:7] = [0.3, -0.2, 0.1, 0.9]
    fc0 = np.concatenate([np.linalg.qr(rng.standard_normal((3, 3)))[0].ravel(), [0.1, -0.2, 0.15, 1.3]])
    vp0 = rng.uniform(0.5, 1.5, nvp)
    dx0, kb0, dkb0 = rng.standard_normal(nx), rng.standard_normal(nz), rng.standard_normal(nz)

    def inputs(x, dx=dx0, kb=kb0, dkb=dkb0):
        return [*x, *fc0, *vp0, *dx, *kb, *dkb]

    def f(x):
        return np.array(g.evaluate(pf.f, inputs(x)), dtype=float)

    eps = 1e-6
    Jfd = np.stack([(f(x0 + eps * np.eye(nx)[j]) - f(x0 - eps * np.eye(nx)[j])) / (2 * eps) for j in range(nx)], 1)
    J = np.zeros((nz, nx))
    vals = g.evaluate(pf.J_nodes, inputs(x0))
    for (r, c), v in zip(pf.J, vals):
        J[r, c] = v
    assert np.max(np.abs(J - Jfd)) < 1e-7 * max(1, np.max(np.abs(Jfd)))
    # jvp
    df = np.array(g.evaluate(pf.df, inputs(x0)), dtype=float)
    assert np.max(np.abs(df - J @ dx0)) < 1e-11 * max(1, np.max(np.abs(df)))
    # xb = J' kb ; W = d(J'kb)/dx ; dxb = W dx + J' dkb
    xb = np.array(g.evaluate(pf.xb, inputs(x0)), dtype=float)
    assert np.max(np.abs(xb - J.T @ kb0)) < 1e-11 * max(1, np.max(np.abs(xb)))

    def vjp(x):
        return np.array(g.evaluate(pf.xb, inputs(x)), dtype=float)
    Wfd = np.stack([(vjp(x0 + eps * np.eye(nx)[j]) - vjp(x0 - eps * np.eye(nx)[j])) / (2 * eps) for j in range(nx)], 1)
    W = np.zeros((nx, nx))
    for (r, c), v in zip(pf.W, g.evaluate(pf.W_nodes, inputs(x0)) if pf.W else []):
        W[r, c] = v
        W[c, r] = v
    assert np.max(np.abs(W - Wfd)) < 1e-6 * max(1, np.max(np.abs(Wfd)))
    dxb = np.array(g.evaluate(pf.dxb, inputs(x0)), dtype=float)
    assert np.max(np.abs(dxb - (W @ dx0 + J.T @ dkb0))) < 1e-10 * max(1, np.max(np.abs(dxb)))
