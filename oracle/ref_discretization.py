'''
Oracle restatement of drone3d/utils/discretization_utils.py:6-51 (collocation coefficients).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`ca.collocation_points(K, 'legendre')` [third party] returns the K Gauss-Legendre nodes on
(0, 1); numpy's leggauss gives the same nodes.  The polynomial recipe is the reference's own
`np.poly1d` one so that exact-zero coefficients (d == 0 -> D = e0) come out
the same way.
'''
import numpy as np


def collocation_points(K):
    x, _ = np.polynomial.legendre.leggauss(K)
    return np.sort((x + 1.0) / 2.0)


def _basis_poly(tau, j):
    # discretization_utils.py:20-24
    p = np.poly1d([1])
    for r in range(len(tau)):
        if r != j:
            p *= np.poly1d([1, -tau[r]]) / (tau[j] - tau[r])
    return p


def get_collocation_coefficients(K):
    # discretization_utils.py:8-34
    tau = np.append(0, collocation_points(K))
    B = np.zeros(K + 1)
    C = np.zeros((K + 1, K + 1))
    D = np.zeros(K + 1)
    for j in range(K + 1):
        p = _basis_poly(tau, j)
        B[j] = np.polyint(p)(1.0)
        dp = np.polyder(p)
        for r in range(K + 1):
            C[j, r] = dp(tau[r])
        D[j] = p(1.0)
    return tau, B, C, D


def get_intermediate_collocation_coefficients(K, d):
    # discretization_utils.py:36-51
    tau = np.append(0, collocation_points(K))
    return np.array([_basis_poly(tau, j)(d) for j in range(K + 1)])
