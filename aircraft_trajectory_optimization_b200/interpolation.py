'''
Trajectory interpolants of a solved raceline (numeric).

Reference: drone3d/utils/discretization_utils.py:53-137 builds these as CasADi functions of
(t, w) from pw_const / pw_lin pieces; `_unpack_soln` (drone3d/raceline/base_raceline.py:824-842)
then fixes w.  Here they are plain numpy closures over the solution arrays:
  collocation: the degree-K Lagrange polynomial of the interval that contains t,
  RK4:         linear interpolation between interval start states.
'''
import numpy as np


def _lagrange_weights(tau, rel):
    ''' l_j(rel) for nodes tau '''
    n = len(tau)
    w = np.ones(n)
    for j in range(n):
        for r in range(n):
            if r != j:
                w[j] *= (rel - tau[r]) / (tau[j] - tau[r])
    return w


def make_interpolants(H, Z, U, dU, colloc=None):
    '''
    H (N,), Z (N, P, nz), U/dU (N, P, nu).  colloc = (tau, D) for collocation, None for RK4.
    Returns three callables t -> vector.
    '''
    H = np.asarray(H, dtype=float)
    tp = np.concatenate([[0.0], np.cumsum(H)])
    N = len(H)

    def build(X):
        X = np.asarray(X, dtype=float)
        if colloc is None:
            knots = tp[:-1]
            vals = X[:, 0, :]

            def f(t):
                # pw_lin over the interval start times; linear extrapolation beyond the ends
                i = int(np.clip(np.searchsorted(knots, t, side='right') - 1, 0, max(N - 2, 0)))
                if N == 1:
                    return vals[0].copy()
                a = (t - knots[i]) / (knots[i + 1] - knots[i])
                return vals[i] + a * (vals[i + 1] - vals[i])
            return f
        tau, D = colloc
        x_end = np.tensordot(D, X[-1], axes=(0, 0))

        def f(t):
            if t < tp[0]:
                return X[0, 0].copy()
            if t >= tp[-1]:
                return x_end.copy()
            n = int(np.searchsorted(tp, t, side='right') - 1)
            rel = (t - tp[n]) / H[n]
            return _lagrange_weights(tau, rel) @ X[n]
        return f

    return build(Z), build(U), build(dU)
