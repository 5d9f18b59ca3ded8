'''
CPU checker for the mesh distance queries.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

The reference gets signed distances and closest points from trimesh.proximity (drone3d/obstacles/mesh_obstacle.py:38-42,
:86-145) [third party: trimesh, un-pinned in the reference's setup.py, not installed here -- parity with trimesh itself
is unpinned].  This restates the definition in numpy: distance to the closest point of the closest triangle (exact
Voronoi-region projection, vectorised over triangles), sign from the parity of ray crossings -- positive outside, the
reference's convention.
'''
import numpy as np


def closest_points_on_triangles(p, T):
    ''' p (3,), T [nt, 3, 3] -> closest point on every triangle [nt, 3] '''
    a, b, c = T[:, 0], T[:, 1], T[:, 2]
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = np.einsum('ij,ij->i', ab, ap), np.einsum('ij,ij->i', ac, ap)
    bp = p - b
    d3, d4 = np.einsum('ij,ij->i', ab, bp), np.einsum('ij,ij->i', ac, bp)
    cp = p - c
    d5, d6 = np.einsum('ij,ij->i', ab, cp), np.einsum('ij,ij->i', ac, cp)
    vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
    out = np.empty_like(a)
    done = np.zeros(len(a), dtype=bool)

    def put(mask, val):
        m = mask & ~done
        out[m] = val[m]
        done[m] = True

    with np.errstate(divide='ignore', invalid='ignore'):
        put((d1 <= 0) & (d2 <= 0), a)
        put((d3 >= 0) & (d4 <= d3), b)
        put((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + ab * (d1 / (d1 - d3))[:, None])
        put((d6 >= 0) & (d5 <= d6), c)
        put((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + ac * (d2 / (d2 - d6))[:, None])
        put((va <= 0) & (d4 - d3 >= 0) & (d5 - d6 >= 0), b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[:, None])
        den = 1.0 / (va + vb + vc)
        put(np.ones(len(a), dtype=bool), a + ab * (vb * den)[:, None] + ac * (vc * den)[:, None])
    return out


def ray_crossings(p, d, T):
    a, b, c = T[:, 0], T[:, 1], T[:, 2]
    e1, e2 = b - a, c - a
    h = np.cross(d, e2)
    det = np.einsum('ij,ij->i', e1, h)
    with np.errstate(divide='ignore', invalid='ignore'):
        inv = 1.0 / det
        s = p - a
        u = np.einsum('ij,ij->i', s, h) * inv
        q = np.cross(s, e1)
        v = (q @ d) * inv
        t = np.einsum('ij,ij->i', e2, q) * inv
    hit = (np.abs(det) >= 1e-300) & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 0)
    return int(hit.sum())


def signed_distance(T, X, direction=(0.5773502691896258, 0.7071067811865476 * 0.8164965809277260, 0.40824829046386296)):
    ''' T [nt, 3, 3], X [n, 3] -> (signed distance positive outside [n], closest point [n, 3]) '''
    T = np.asarray(T, dtype=float).reshape(-1, 3, 3)
    d = np.asarray(direction, dtype=float)
    dist, close = np.empty(len(X)), np.empty((len(X), 3))
    for i, p in enumerate(np.asarray(X, dtype=float)):
        cp = closest_points_on_triangles(p, T)
        d2 = np.einsum('ij,ij->i', p - cp, p - cp)
        k = int(np.argmin(d2))
        r = np.sqrt(d2[k])
        dist[i] = -r if ray_crossings(p, d, T) & 1 else r
        close[i] = cp[k]
    return dist, close


def box_triangles(lo, hi):
    ''' the 12 triangles of an axis-aligned box (outward orientation) '''
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    v = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]])
    f = [[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [1, 2, 6], [1, 6, 5], [2, 3, 7], [2, 7, 6],
         [3, 0, 4], [3, 4, 7]]
    return v[np.array(f)]


def box_sdf(lo, hi, X):
    ''' analytic signed distance to an axis-aligned box, positive outside '''
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    c, e = 0.5 * (lo + hi), 0.5 * (hi - lo)
    q = np.abs(np.asarray(X, float) - c) - e
    return np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(q.max(axis=1), 0)
