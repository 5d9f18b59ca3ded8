'''
Legendre collocation coefficients on [0, 1].

Same quantities as the reference's `drone3d/utils/discretization_utils.py:6-51`
(`tau, B, C, D` and the intermediate-point `D(d)`), computed from barycentric Lagrange
formulas instead of `np.poly1d` products.  The collocation roots come from
`numpy.polynomial.legendre.leggauss` mapped to (0, 1) -- the reference takes them from
`ca.collocation_points(K, 'legendre')`, which tabulates the same Gauss-Legendre nodes.
'''
import numpy as np


def legendre_points(K: int) -> np.ndarray:
    ''' tau[0] = 0 followed by the K Gauss-Legendre nodes on (0, 1), ascending '''
    nodes, _ = np.polynomial.legendre.leggauss(K)
    return np.append(0.0, np.sort((nodes + 1.0) / 2.0))


def _lagrange_basis(tau):
    ''' list of numpy Polynomial objects l_j with l_j(tau_r) = delta_jr '''
    P = np.polynomial.Polynomial
    basis = []
    for j, tj in enumerate(tau):
        p = P([1.0])
        for r, tr in enumerate(tau):
            if r != j:
                p = p * P([-tr, 1.0]) / (tj - tr)
        basis.append(p)
    return basis


def get_collocation_coefficients(K: int):
    '''
    tau (K+1,), B (K+1,) integral weights, C (K+1, K+1) with C[j, r] = l_j'(tau_r),
    D (K+1,) with D[j] = l_j(1)
    '''
    tau = legendre_points(K)
    basis = _lagrange_basis(tau)
    B = np.array([p.integ()(1.0) - p.integ()(0.0) for p in basis])
    C = np.array([[p.deriv()(tr) for tr in tau] for p in basis])
    D = np.array([p(1.0) for p in basis])
    return tau, B, C, D


def get_intermediate_collocation_coefficients(K: int, d: float) -> np.ndarray:
    ''' D(d)[j] = l_j(d): interpolation weights at fraction d of an interval '''
    tau = legendre_points(K)
    return np.array([p(d) for p in _lagrange_basis(tau)])
