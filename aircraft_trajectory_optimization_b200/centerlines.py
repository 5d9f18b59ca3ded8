'''
Numeric cubic-spline centerline with a Darboux frame.

Mirrors the public surface of the reference's `drone3d/centerlines/base_centerline.py`
(BaseCenterlineConfig :26-54, p2x/p2xc/p2es/... :300-312, gate_position/orientation :314-336,
x2p/g2lx :231-264) and `spline_centerline.py` (SplineCenterlineConfig :19-37, ry fitting
:114-230, spline set-up :232-264, frame + curvature formulas :279-294).

The reference builds these quantities as CasADi expressions of 15 "param_terms"
(xc, xcs, xcss, ry, rys) and evaluates them one point at a time.  At solve time `s` is always a
fixed number (base_raceline.py:963-984), so here everything is plain vectorised numpy: the
raceline builders ask for the 13 frame constants per collocation point
(`frame_constants`: Rp row-major, ks, ky, kn, |xcs|) which is all the kernels need.
'''
from dataclasses import dataclass, field
from enum import Enum
from typing import Union

import numpy as np
import scipy.interpolate
from scipy.spatial.distance import cdist

from .pytypes import PythonMsg, RacerState, DroneState, RelativeOrientation


class GateShape(Enum):
    ''' gate shape options '''
    CIRCLE = 0
    SQUARE = 1


class SplineRyFitOptions(Enum):
    ''' how the lateral reference direction is chosen when not given '''
    TORSION_FREE = 0
    PLANAR = 1
    PRINCIPAL_CURVATURE = 2


@dataclass
class BaseCenterlineConfig(PythonMsg):
    ''' bounds, gates and periodicity of a centerline (base_centerline.py:26-54) '''
    s_min: float = 0
    s_max: float = 10
    y_min: float = -2
    y_max: float = 2
    n_min: float = -2
    n_max: float = 2
    gate_s: np.ndarray = None
    gate_shape: GateShape = GateShape.CIRCLE
    gate_ri: float = 1.25
    gate_ro: float = 1.35
    gate_w: float = 0.2
    gate_snap_fit: bool = True
    closed: bool = False
    gamma: float = 0.9
    N_grid: int = 10000


@dataclass
class SplineCenterlineConfig(BaseCenterlineConfig):
    ''' waypoints (3 x n) and optional lateral directions (spline_centerline.py:19-37) '''
    s: Union[np.ndarray, None] = None
    x: np.ndarray = None
    ry: Union[np.ndarray, None] = None
    ry_fit_method: SplineRyFitOptions = SplineRyFitOptions.PLANAR

    def __post_init__(self):
        if self.x is None:
            self.x = np.array([[0, 0, 0], [10, 0, 0], [10, 10, 0], [0, 10, 0], [0, 0, 0]]).T


class _PiecewiseCubic:
    '''
    scipy CubicSpline evaluated the way the reference's `scipy_spline_to_casadi`
    (drone3d/utils/interp.py:55-84, extrapolate='linear') evaluates it: right-continuous
    piece lookup, and a straight line through the end value/slope outside the knots
    (so the second derivative is 0 at and beyond the last knot).
    '''

    def __init__(self, x, y, bc_type):
        self.spline = scipy.interpolate.CubicSpline(x, y, bc_type=bc_type)
        self.kx = np.asarray(self.spline.x, dtype=float)
        self.c = np.asarray(self.spline.c, dtype=float)      # (4, n-1, dim)
        self.end_val = self.spline(self.kx[-1])
        self.end_slope = self.spline(self.kx[-1], 1)

    def __call__(self, s, nu=0):
        s = np.asarray(s, dtype=float)
        scalar = s.ndim == 0
        s = np.atleast_1d(s)
        n = len(self.kx)
        idx = np.searchsorted(self.kx, s, side='right') - 1     # -1 below, n-1 at/after the end
        inner = np.clip(idx, 0, n - 2)
        rel = s - self.kx[inner]
        c3, c2, c1, c0 = (self.c[k][inner] for k in range(4))
        rel = rel.reshape((-1,) + (1,) * (c0.ndim - 1))
        below = (idx < 0).reshape(rel.shape)
        above = (idx >= n - 1).reshape(rel.shape)
        # linear extrapolation pieces
        rel_hi = (s - self.kx[-1]).reshape(rel.shape)
        c0 = np.where(above, self.end_val, c0)
        c1 = np.where(above, self.end_slope, c1)
        rel = np.where(above, rel_hi, rel)
        dead = below | above
        c2 = np.where(dead, 0.0, c2)
        c3 = np.where(dead, 0.0, c3)
        if nu == 0:
            out = c0 + rel * (c1 + rel * (c2 + rel * c3))
        elif nu == 1:
            out = c1 + rel * (2 * c2 + 3 * rel * c3)
        elif nu == 2:
            out = 2 * c2 + 6 * rel * c3
        else:
            raise NotImplementedError('derivative order > 2')
        return out[0] if scalar else out


def _unit(v, axis=-1):
    return v / np.linalg.norm(v, axis=axis, keepdims=True)


class BaseCenterline:
    ''' geometry queries shared by all centerlines; subclasses provide `param_terms` '''
    config: BaseCenterlineConfig
    cleanly_closed = True
    xc_grid: np.ndarray

    def __init__(self, config: BaseCenterlineConfig):
        self.config = config
        self._setup_interp()

    # ---- to be provided ------------------------------------------------------------------
    def _setup_interp(self):
        raise NotImplementedError

    def param_terms(self, s) -> np.ndarray:
        ''' (..., 15) array [xc, xcs, xcss, ry, rys] at path length(s) s '''
        raise NotImplementedError

    # ---- bounds --------------------------------------------------------------------------
    def s_min(self):
        return self.config.s_min

    def s_max(self):
        return self.config.s_max

    def y_min(self, s: float = 0):
        return self.config.y_min

    def y_max(self, s: float = 0):
        return self.config.y_max

    def n_min(self, s: float = 0):
        return self.config.n_min

    def n_max(self, s: float = 0):
        return self.config.n_max

    # ---- frame ---------------------------------------------------------------------------
    def frame(self, s):
        '''
        Darboux frame and curvatures at s (scalar or (n,) array), following
        spline_centerline.py:279-294.  Returns a dict of arrays with leading shape of s.
        '''
        pt = self.param_terms(s)
        xc, xcs, xcss, ry, rys = (pt[..., 3 * k:3 * k + 3] for k in range(5))
        mag = np.linalg.norm(xcs, axis=-1)
        es = xcs / mag[..., None]
        ey = ry - es * np.sum(es * ry, axis=-1, keepdims=True)
        ey = _unit(ey)
        en = np.cross(es, ey)
        # [[xcs.es, xcs.ey], [ry.es, ry.ey]] [a; b] = [xcss.en; rys.en] / |xcs|,  ky = -a, ks = b
        m00 = np.sum(xcs * es, axis=-1)
        m01 = np.sum(xcs * ey, axis=-1)
        m10 = np.sum(ry * es, axis=-1)
        m11 = np.sum(ry * ey, axis=-1)
        r0 = np.sum(xcss * en, axis=-1)
        r1 = np.sum(rys * en, axis=-1)
        det = m00 * m11 - m01 * m10
        a = (m11 * r0 - m01 * r1) / det / mag
        b = (-m10 * r0 + m00 * r1) / det / mag
        kn = -np.sum(np.cross(xcss, xcs) * en, axis=-1) / mag ** 3
        return dict(xc=xc, es=es, ey=ey, en=en, ks=b, ky=-a, kn=kn, mag_xcs=mag)

    def frame_constants(self, s) -> np.ndarray:
        '''
        the 13 per-point numbers the kernels consume: Rp (row-major, columns es|ey|en),
        ks, ky, kn, |xcs|.  s: (n,) -> (n, 13)
        '''
        fr = self.frame(np.atleast_1d(np.asarray(s, dtype=float)))
        Rp = np.stack([fr['es'], fr['ey'], fr['en']], axis=-1)      # (n, 3, 3)
        return np.concatenate([Rp.reshape(-1, 9), fr['ks'][:, None], fr['ky'][:, None],
                               fr['kn'][:, None], fr['mag_xcs'][:, None]], axis=1)

    @staticmethod
    def _vec_out(v):
        ''' (3,) for scalar queries, (3, n) for array queries -- the reference's orientation '''
        return v if v.ndim == 1 else v.T

    def p2xc(self, s):
        return self._vec_out(self.frame(s)['xc'])

    def p2es(self, s):
        return self._vec_out(self.frame(s)['es'])

    def p2ey(self, s):
        return self._vec_out(self.frame(s)['ey'])

    def p2en(self, s):
        return self._vec_out(self.frame(s)['en'])

    def p2Rp(self, s):
        fr = self.frame(s)
        return np.stack([fr['es'], fr['ey'], fr['en']], axis=-1)

    def p2ks(self, s):
        return self.frame(s)['ks']

    def p2ky(self, s):
        return self.frame(s)['ky']

    def p2kn(self, s):
        return self.frame(s)['kn']

    def p2k(self, s):
        fr = self.frame(s)
        return np.stack([fr['ks'], fr['ky'], fr['kn']], axis=0)

    def p2mag_xcs(self, s):
        return self.frame(s)['mag_xcs']

    def p2x(self, s, y, n):
        fr = self.frame(s)
        y = np.asarray(y, dtype=float)[..., None]
        n = np.asarray(n, dtype=float)[..., None]
        return self._vec_out(fr['xc'] + y * fr['ey'] + n * fr['en'])

    fast_p2x = p2x
    fast_p2ey = p2ey
    fast_p2en = p2en

    # ---- gates ---------------------------------------------------------------------------
    def gate_position(self, s: float) -> np.ndarray:
        return self.p2xc(s)

    def gate_orientation(self, s: float) -> np.ndarray:
        ''' gate frame [es ey en] with the near-vertical snap of base_centerline.py:318-336 '''
        fr = self.frame(float(s))
        es, ey, en = fr['es'], fr['ey'], fr['en']
        if self.config.gate_snap_fit:
            up = np.array([0., 0., 1.])
            if abs(es[2]) > 0.9:
                es = up
                ey = _unit(ey - es * (ey @ es))
                en = np.cross(es, ey)
            elif abs(en[2]) > 0.9:
                en = up
                es = _unit(es - en * (es @ en))
                ey = np.cross(en, es)
        return np.array([es, ey, en]).T

    # ---- global <-> local ----------------------------------------------------------------
    def l2gx(self, state: RacerState):
        state.x.from_vec(self.p2x(*state.p.to_vec()))

    def l2gq(self, state: DroneState):
        if isinstance(state.r, RelativeOrientation):
            R = self.p2Rp(state.p.s) @ state.r.R()
        else:
            R = state.r.R()
        state.q.from_mat(R)

    def x2p(self, x_query: np.ndarray) -> np.ndarray:
        ''' nearest-grid projection of (n, 3) global positions to (s, y, n) '''
        x_query = np.atleast_2d(x_query)
        s = self.xc_grid[cdist(x_query, self.xc_grid[:, 1:], metric='sqeuclidean').argmin(axis=1), 0]
        fr = self.frame(s)
        d = x_query - fr['xc']
        return np.stack([s + np.sum(d * fr['es'], axis=1),
                         np.sum(d * fr['ey'], axis=1),
                         np.sum(d * fr['en'], axis=1)], axis=1)

    def g2lx(self, state: RacerState):
        state.p.from_vec(self.x2p(state.x.to_vec()[None])[0])


class SplineCenterline(BaseCenterline):
    ''' centerline through waypoints with a cubic spline for xc(s) and for ry(s) '''
    config: SplineCenterlineConfig

    def __init__(self, config: SplineCenterlineConfig):
        if not isinstance(config.gate_s, np.ndarray) and isinstance(config.s, np.ndarray):
            config.gate_s = config.s
        super().__init__(config)

    def _setup_interp(self):
        cfg = self.config
        cfg.x = np.asarray(cfg.x, dtype=float)
        if cfg.closed and not (cfg.x[:, 0] == cfg.x[:, -1]).all():
            cfg.x = np.hstack([cfg.x, cfg.x[:, 0:1]])
        if cfg.s is None:
            cfg.s = np.arange(cfg.x.shape[1]) * 1
            if cfg.gate_s is None:
                cfg.gate_s = cfg.s
        cfg.s_max = cfg.s.max()
        cfg.s_min = cfg.s.min()

        bc = 'periodic' if cfg.closed else 'not-a-knot'
        self._xc = _PiecewiseCubic(cfg.s, cfg.x.T, bc)
        self._center_spline = self._xc.spline

        s_grid = np.linspace(self.s_min(), self.s_max(), cfg.N_grid)
        self.xc_grid = np.concatenate([s_grid[:, None], self._center_spline(s_grid)], axis=1)

        if cfg.ry is not None:
            s_fit, ry_fit = cfg.s, np.asarray(cfg.ry, dtype=float)
        elif cfg.ry_fit_method == SplineRyFitOptions.PLANAR:
            s_fit, ry_fit = self._ry_planar()
        elif cfg.ry_fit_method == SplineRyFitOptions.PRINCIPAL_CURVATURE:
            s_fit, ry_fit = self._ry_principal_curvature()
        elif cfg.ry_fit_method == SplineRyFitOptions.TORSION_FREE:
            s_fit, ry_fit = self._ry_torsion_free()
        else:
            raise NotImplementedError(f'Unhandled ry fit option: {cfg.ry_fit_method}')

        ry_fit = np.array(ry_fit, dtype=float)
        self.cleanly_closed = bool(np.linalg.norm(ry_fit[0] - ry_fit[-1]) < 1e-3)
        if self.cleanly_closed:
            ry_fit[-1] = ry_fit[0]
        self._ry = _PiecewiseCubic(s_fit, ry_fit, 'periodic' if self.cleanly_closed else 'not-a-knot')
        self._lateral_spline = self._ry.spline

    def param_terms(self, s):
        s = np.asarray(s, dtype=float)
        return np.concatenate([self._xc(s), self._xc(s, 1), self._xc(s, 2),
                               self._ry(s), self._ry(s, 1)], axis=-1)

    # ---- lateral direction fits ----------------------------------------------------------
    def _ry_planar(self):
        ''' horizontal left-pointing direction from the unwrapped yaw (spline_centerline.py:151-176) '''
        s_fit = np.linspace(self.s_min(), self.s_max(), 100)
        tangent = self._center_spline(s_fit, 1)
        th = np.arctan2(tangent[:, 1], tangent[:, 0])
        for k in range(1, len(th)):
            while th[k] - th[k - 1] > np.pi:
                th[k] -= 2 * np.pi
            while th[k - 1] - th[k] > np.pi:
                th[k] += 2 * np.pi
        th = th + np.pi / 2
        fine = scipy.interpolate.CubicSpline(s_fit, th)
        coarse = scipy.interpolate.CubicSpline(self.config.s, fine(self.config.s))
        th_fit = coarse(s_fit)
        ry = np.stack([np.cos(th_fit), np.sin(th_fit), np.zeros_like(th_fit)], axis=1)
        if self.config.closed:
            ry[-1] = ry[0]
        return s_fit, ry

    def _ry_principal_curvature(self):
        s_fit = self.config.s
        es = _unit(self._center_spline(s_fit, 1))
        en = self._center_spline(s_fit, 2)
        en = _unit(en - es * np.sum(es * en, axis=1, keepdims=True))
        return s_fit, -np.cross(en, es)

    def _ry_torsion_free(self):
        ''' parallel transport of ey along the curve: ey' = -(es' . ey) es (spline_centerline.py:178-217) '''
        from scipy.integrate import solve_ivp
        s_grid = np.linspace(self.s_min(), self.s_max(), 100)

        def rhs(s, ey):
            d1 = self._center_spline(s, 1)
            d2 = self._center_spline(s, 2)
            nrm = np.linalg.norm(d1)
            es = d1 / nrm
            des = d2 / nrm - d1 * (d1 @ d2) / nrm ** 3
            return -(des @ ey) * es

        d0 = self._center_spline(self.s_min(), 1)
        ey0 = np.array([-d0[1], d0[0], 0.0])
        ey0 /= np.linalg.norm(ey0)
        sol = solve_ivp(rhs, (s_grid[0], s_grid[-1]), ey0, t_eval=s_grid, rtol=1e-11, atol=1e-12,
                        max_step=s_grid[1] - s_grid[0])
        return s_grid, sol.y.T
