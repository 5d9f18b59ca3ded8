'''
Legendre collocation coefficients on [0, 1].

Same quantities as the reference's `drone3d/utils/discretization_utils.py:6-51`
(`tau, B, C, D` and the intermediate-point `D(d)`).  The Lagrange basis polynomials are expanded
with `np.poly1d` in the same factor order as the reference: the expansion carries ~1e-11 of
rounding in C for K = 7, and results are only identical to the reference's if that rounding is the
same.  The roots come from `numpy.polynomial.legendre.leggauss` mapped to (0, 1) -- the reference
takes them from `ca.collocation_points(K, 'legendre')`, which tabulates the same Gauss-Legendre
nodes.
'''
import numpy as np


def legendre_points(K: int) -> np.ndarray:
    ''' tau[0] = 0 followed by the K Gauss-Legendre nodes on (0, 1), ascending '''
    nodes, _ = np.polynomial.legendre.leggauss(K)
    return np.append(0.0, np.sort((nodes + 1.0) / 2.0))


def lagrange_basis(tau):
    ''' np.poly1d l_j with l_j(tau_r) = delta_jr, factors multiplied in ascending r '''
    basis = []
    for j, tj in enumerate(tau):
        lj = np.poly1d([1])
        for r, tr in enumerate(tau):
            if r == j:
                continue
            lj = lj * (np.poly1d([1, -tr]) / (tj - tr))
        basis.append(lj)
    return basis


def get_collocation_coefficients(K: int):
    '''
    tau (K+1,), B (K+1,) integral weights, C (K+1, K+1) with C[j, r] = l_j'(tau_r),
    D (K+1,) with D[j] = l_j(1)
    '''
    tau = legendre_points(K)
    basis = lagrange_basis(tau)
    B = np.array([np.polyint(lj)(1.0) for lj in basis])
    C = np.array([[np.polyder(lj)(tr) for tr in tau] for lj in basis])
    D = np.array([lj(1.0) for lj in basis])
    return tau, B, C, D


def get_intermediate_collocation_coefficients(K: int, d: float) -> np.ndarray:
    ''' D(d)[j] = l_j(d): interpolation weights at fraction d of an interval '''
    return np.array([lj(d) for lj in lagrange_basis(legendre_points(K))])
