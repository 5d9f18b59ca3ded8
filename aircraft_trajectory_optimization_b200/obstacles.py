'''
Mesh obstacles and the obstacle-free tube, with the reference's Python API:
drone3d/obstacles/mesh_obstacle.py -- `MeshObstacle` (:18-145: signed_distance, check_line_for_collisions,
compute_plannning_tube [sic], check_for_collisions, search_largest_sphere) and `ObstacleFreeTube` (:200-237, kept in
raceline.py next to the row it produces).

The reference loads the mesh with trimesh and queries its r-tree accelerated proximity functions on the CPU; trimesh is
not a dependency here.  The mesh is read from the Wavefront .obj directly and every query goes to the brute-force CUDA
kernel `mesh_sdf_kernel` (csrc/mesh_sdf.cuh) through the C ABI `rb_mesh_sdf`: the tube of scripts/obstacles.py is
800 collocation points x 41 samples against 8884 faces.  No CPU fallback.
'''
import ctypes
import os
from typing import List

import numpy as np

from .raceline import ObstacleFreeTube

_DEFAULT_FILENAME = 'arena_track_obstacles_multistory.obj'


def load_obj_triangles(path) -> np.ndarray:
    ''' triangles [nt, 3, 3] of a Wavefront .obj (polygons are fanned; negative indices are relative) '''
    verts, faces = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith('v '):
                verts.append([float(t) for t in line.split()[1:4]])
            elif line.startswith('f '):
                idx = [int(t.split('/')[0]) for t in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    V = np.asarray(verts, dtype=float)
    return V[np.asarray(faces, dtype=np.int64)]


class MeshObstacle:
    ''' mesh-based obstacle; distance queries run on the GPU '''

    def __init__(self, filename: str = _DEFAULT_FILENAME, color: List[float] = None, triangles: np.ndarray = None,
                 device=None):
        self.filename = filename
        self.color = color if color is not None else [1, 0, 0, 1]
        if triangles is None:
            path = filename if os.path.isabs(filename) and os.path.exists(filename) else _find_asset(filename)
            triangles = load_obj_triangles(path)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 9)
        self.device = device
        self._tri_dev = None

    # ---- queries -------------------------------------------------------------------------------------
    def _query(self, x: np.ndarray, want_closest=False):
        import torch
        from .functions import load_library, _check
        if not torch.cuda.is_available():
            raise RuntimeError('mesh distance queries run only on a CUDA device (no CPU fallback)')
        lib = load_library()
        dev = torch.device('cuda', torch.cuda.current_device() if self.device is None else self.device)
        if self._tri_dev is None or self._tri_dev.device != dev:
            self._tri_dev = torch.from_numpy(self.triangles).to(dev)
        pts = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)).to(dev)
        npts = pts.shape[0]
        dist = torch.empty(npts, dtype=torch.float64, device=dev)
        closest = torch.empty(npts, 3, dtype=torch.float64, device=dev) if want_closest else None
        vp = ctypes.c_void_p
        lib.rb_mesh_sdf.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, vp, vp, vp]
        _check(lib.rb_mesh_sdf(vp(self._tri_dev.data_ptr()), self.triangles.shape[0], vp(pts.data_ptr()), npts,
                               vp(dist.data_ptr()), vp(closest.data_ptr()) if want_closest else None,
                               vp(torch.cuda.current_stream(dev).cuda_stream)), 'rb_mesh_sdf')
        return dist.cpu().numpy(), (closest.cpu().numpy() if want_closest else None)

    def signed_distance(self, x: np.ndarray) -> np.ndarray:
        ''' signed distance, positive outside of the object (mesh_obstacle.py:38-42) '''
        return self._query(x)[0]

    def closest_point(self, x: np.ndarray) -> np.ndarray:
        return self._query(x, want_closest=True)[1]

    def check_line_for_collisions(self, line, n: int = 1000) -> float:
        ''' closest distance of a centerline to the mesh (mesh_obstacle.py:44-48) '''
        s = np.linspace(line.s_min(), line.s_max(), n)
        return float(self.signed_distance(np.array([np.asarray(line.p2xc(sk)).ravel() for sk in s])).min())

    def check_for_collisions(self, x: np.ndarray, sep_radius: float = 0.3) -> bool:
        ''' x: (n, 3) points (mesh_obstacle.py:78-84) '''
        return bool((self.signed_distance(x) >= sep_radius).all())

    def search_largest_sphere(self, x0: np.ndarray, ey: np.ndarray, en: np.ndarray, r_max: float = 0.5, nr: int = 5,
                              nth: int = 8):
        ''' brute-force search for the largest empty sphere near every point (mesh_obstacle.py:110-145) '''
        x0, ey, en = (np.asarray(a, dtype=float).reshape(-1, 3) for a in (x0, ey, en))
        d0 = self.signed_distance(x0)
        r = np.linspace(r_max, 0, nr, endpoint=False)
        th = np.linspace(0, 2 * np.pi, nth, endpoint=False)
        R, TH = np.meshgrid(r, th)
        R, TH = R.reshape(-1), TH.reshape(-1)
        X = x0[:, None, :] + ey[:, None, :] * (R * np.cos(TH))[None, :, None] \
            + en[:, None, :] * (R * np.sin(TH))[None, :, None]                       # [n, nr*nth, 3]
        D = self.signed_distance(X.reshape(-1, 3)).reshape(len(x0), -1).T               # [nr*nth, n]
        idxs = D.argmax(axis=0)
        dn = D.max(axis=0)
        rn, thn = R[idxs], TH[idxs]
        xn = x0 + rn[:, None] * (ey * np.cos(thn[:, None]) + en * np.sin(thn[:, None]))
        x, rr = xn.copy(), dn.copy()
        x[d0 >= dn] = x0[d0 >= dn]
        rr[d0 >= dn] = d0[d0 >= dn]
        pts = self.closest_point(x)
        return x, rr, pts

    def compute_plannning_tube(self, line, s: np.ndarray, collision_r: float) -> ObstacleFreeTube:
        ''' obstacle-free tube along a centerline at path lengths s (mesh_obstacle.py:50-76) '''
        s = np.asarray(s, dtype=float)
        x = np.array([np.asarray(line.p2xc(sk)).ravel() for sk in s])
        ey = np.array([np.asarray(line.p2ey(sk)).ravel() for sk in s])
        en = np.array([np.asarray(line.p2en(sk)).ravel() for sk in s])
        ball_center, ball_r, ball_tangent_pts = self.search_largest_sphere(x, ey, en)
        ball_y = np.sum((ball_center - x) * ey, axis=1)
        ball_n = np.sum((ball_center - x) * en, axis=1)
        return ObstacleFreeTube(np.array([s, ball_y, ball_n]).T, ball_r, collision_r, line=line,
                                ball_center=ball_center, ball_tangent_pts=ball_tangent_pts)


def _find_asset(filename):
    here = os.path.dirname(os.path.abspath(__file__))
    for folder in (os.path.join(os.path.dirname(here), 'drone3d', 'assets'), os.environ.get('RACELINE_ASSETS', ''),
                   '/root/reference/drone3d/assets'):
        if folder and os.path.exists(os.path.join(folder, filename)):
            return os.path.join(folder, filename)
    raise FileNotFoundError(f'{filename}: the mesh is an asset of the reference repository and is not shipped here; put '
                            f'it under drone3d/assets/ or point RACELINE_ASSETS at the reference\'s drone3d/assets')
