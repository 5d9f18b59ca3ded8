'''
Oracle restatement of the reference centerline.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Follows drone3d/centerlines/spline_centerline.py (:106-176 s / ry fitting, :232-264 splines,
:266-322 symbolic frame) and drone3d/centerlines/base_centerline.py (:274-336 p2* helpers and
gate pose), with drone3d/utils/interp.py:55-84 for how a scipy spline becomes a piecewise
polynomial with linear extrapolation.

The reference keeps the frame as a CasADi expression of 15 symbolic `param_terms`
[xc, xcs, xcss, ry, rys]; calling it with a numeric s folds everything to numbers.  Here the
same expression is built on the SX-like engine with numeric leaves, so it constant-folds the
same way, and the resulting numbers are read back.
'''
import numpy as np
import scipy.interpolate

from aircraft_trajectory_optimization_b200 import symbolic as sx
from aircraft_trajectory_optimization_b200.centerlines import GateShape, SplineRyFitOptions, \
    SplineCenterlineConfig  # plain config dataclasses / enums only


class _CasadiSpline:
    ''' interp.py:55-84 with extrapolate='linear': coefficient lookup by pw_const, then a cubic '''

    def __init__(self, x_data, y_data, bc_type):
        sp = scipy.interpolate.CubicSpline(x_data, y_data, bc_type=bc_type)
        self.k_x = sp.x
        k_c = sp.c
        k_f = sp(self.k_x[-1])
        self.x_0 = [self.k_x[0], *self.k_x]
        self.c_0 = [k_c[3, 0], *k_c[3, :], k_f]
        self.c_1 = [k_c[2, 0], *k_c[2, :], sp(self.k_x[-1], 1)]
        self.c_2 = [0, *k_c[1, :], 0]
        self.c_3 = [0, *k_c[0, :], 0]

    def _piece(self, s):
        # ca.pw_const(t, tval, val): val[0] for t < tval[0]; val[i+1] for tval[i] <= t < tval[i+1]
        return int(np.searchsorted(self.k_x, s, side='right'))

    def __call__(self, s, nu=0):
        i = self._piece(s)
        r = s - self.x_0[i]
        c0, c1, c2, c3 = self.c_0[i], self.c_1[i], self.c_2[i], self.c_3[i]
        if nu == 0:
            return c0 + c1 * r + c2 * r ** 2 + c3 * r ** 3
        if nu == 1:
            return c1 + 2 * c2 * r + 3 * c3 * r ** 2
        return 2 * c2 + 6 * c3 * r


class RefSplineCenterline:
    ''' numeric-only restatement of SplineCenterline (closed or open, PLANAR / PRINCIPAL ry fit) '''

    def __init__(self, config: SplineCenterlineConfig):
        # spline_centerline.py:58-61
        if not isinstance(config.gate_s, np.ndarray) and isinstance(config.s, np.ndarray):
            config.gate_s = config.s
        self.config = config
        self.cleanly_closed = True
        self._setup_interp()

    def s_min(self):
        return self.config.s_min

    def s_max(self):
        return self.config.s_max

    def _setup_interp(self):
        # spline_centerline.py:232-264
        cfg = self.config
        if cfg.closed:
            if not (cfg.x[:, 0] == cfg.x[:, -1]).all():
                cfg.x = np.hstack([cfg.x, cfg.x[:, 0:1]])
        # _fill_in_s :106-112
        if cfg.s is None:
            cfg.s = np.arange(cfg.x.shape[1]) * 1
            if cfg.gate_s is None:
                cfg.gate_s = cfg.s
        cfg.s_max = cfg.s.max()
        cfg.s_min = cfg.s.min()

        bc_type = 'not-a-knot' if not cfg.closed else 'periodic'
        self._center_spline = scipy.interpolate.CubicSpline(cfg.s, cfg.x.T, bc_type=bc_type)
        self._xc = [_CasadiSpline(cfg.s, cfg.x[i], bc_type) for i in range(3)]
        self._fill_in_ry()

    def _fill_in_ry(self):
        # spline_centerline.py:114-149
        cfg = self.config
        if cfg.ry is not None:
            s_grid, ry_grid = cfg.s, np.array(cfg.ry, dtype=float)
        elif cfg.ry_fit_method == SplineRyFitOptions.PLANAR:
            s_grid, ry_grid = self._fill_in_ry_planar()
        elif cfg.ry_fit_method == SplineRyFitOptions.PRINCIPAL_CURVATURE:
            s_grid, ry_grid = self._fill_in_ry_principal_curvature()
        else:
            raise NotImplementedError('oracle: TORSION_FREE needs IDAS (third party); not restated')
        if np.linalg.norm(ry_grid[0] - ry_grid[-1]) < 1e-3:
            ry_grid[-1] = ry_grid[0]
            self.cleanly_closed = True
        else:
            self.cleanly_closed = False
        bc_type = 'periodic' if self.cleanly_closed else 'not-a-knot'
        self._ry = [_CasadiSpline(s_grid, ry_grid[:, i], bc_type) for i in range(3)]

    def _fill_in_ry_planar(self):
        # spline_centerline.py:151-176
        s_fit = np.linspace(self.s_min(), self.s_max(), 100)
        es = self._center_spline(s_fit, 1).T
        th = np.arctan2(es[1], es[0])
        for k in range(1, len(th)):
            while th[k] - th[k - 1] > np.pi:
                th[k] -= 2 * np.pi
            while th[k - 1] - th[k] > np.pi:
                th[k] += 2 * np.pi
        th = th + np.pi / 2
        thc = scipy.interpolate.CubicSpline(s_fit, th)
        s_waypoints = self.config.s
        thc = scipy.interpolate.CubicSpline(s_waypoints, thc(s_waypoints))
        th_fit = thc(s_fit)
        ry_fit = np.array([np.cos(th_fit), np.sin(th_fit), th_fit * 0])
        if self.config.closed:
            ry_fit[:, -1] = ry_fit[:, 0]
        return s_fit, ry_fit.T

    def _fill_in_ry_principal_curvature(self):
        # spline_centerline.py:219-230
        s_fit = self.config.s
        es = self._center_spline(s_fit, 1)
        en = self._center_spline(s_fit, 2)
        es = es / np.linalg.norm(es, axis=1)[:, np.newaxis]
        en = en - es * (es * en).sum(axis=1)[:, np.newaxis]
        en = en / np.linalg.norm(en, axis=1)[:, np.newaxis]
        return s_fit, -np.cross(en, es)

    # ---- param terms and the symbolic frame ------------------------------------------------
    def f_param_terms(self, s):
        ''' spline_centerline.py:297-307: [xc, xcs, xcss, ry, rys] at numeric s (15 floats) '''
        s = float(s)
        return np.array([f(s, nu) for nu in (0, 1, 2) for f in self._xc]
                        + [f(s, nu) for nu in (0, 1) for f in self._ry])

    @staticmethod
    def sym_rep(param_terms):
        '''
        spline_centerline.py:279-294 on the expression engine.  param_terms: 15 entries
        (floats fold to constants).  Returns dict of SX scalars / object arrays.
        '''
        pt = np.array([e if isinstance(e, sx.SX) else sx.SX.const(e) for e in param_terms],
                      dtype=object)
        xc, xcs, xcss, ry, rys = (pt[3 * k:3 * k + 3] for k in range(5))
        mag = sx.norm_2(xcs)
        es = xcs / mag
        ey = ry - es * (es @ ry)
        ey = ey / sx.norm_2(ey)
        en = np.array([es[1] * ey[2] - es[2] * ey[1],
                       es[2] * ey[0] - es[0] * ey[2],
                       es[0] * ey[1] - es[1] * ey[0]], dtype=object)
        # 2x2 inverse by adjugate / determinant
        a, b, c, d = xcs @ es, xcs @ ey, ry @ es, ry @ ey
        det = a * d - b * c
        r0, r1 = xcss @ en, rys @ en
        kyks0 = (d * r0 - b * r1) / det / mag
        kyks1 = (a * r1 - c * r0) / det / mag
        cr = np.array([xcss[1] * xcs[2] - xcss[2] * xcs[1],
                       xcss[2] * xcs[0] - xcss[0] * xcs[2],
                       xcss[0] * xcs[1] - xcss[1] * xcs[0]], dtype=object)
        kn = -(cr @ en) / mag ** 3
        Rp = np.empty((3, 3), dtype=object)
        Rp[:, 0], Rp[:, 1], Rp[:, 2] = es, ey, en
        return dict(xc=xc, es=es, ey=ey, en=en, Rp=Rp, ks=kyks1, ky=-kyks0, kn=kn, mag_xcs=mag,
                    k=np.array([kyks1, -kyks0, kn], dtype=object))

    def _num(self, s, key):
        v = self.sym_rep(self.f_param_terms(s))[key]
        if isinstance(v, sx.SX):
            return v.value()
        return np.vectorize(lambda e: e.value(), otypes=[float])(v)

    # base_centerline.py:300-312
    def p2xc(self, s):
        return self._num(s, 'xc')

    def p2es(self, s):
        return self._num(s, 'es')

    def p2ey(self, s):
        return self._num(s, 'ey')

    def p2en(self, s):
        return self._num(s, 'en')

    def p2Rp(self, s):
        return self._num(s, 'Rp')

    def p2ks(self, s):
        return self._num(s, 'ks')

    def p2ky(self, s):
        return self._num(s, 'ky')

    def p2kn(self, s):
        return self._num(s, 'kn')

    def p2mag_xcs(self, s):
        return self._num(s, 'mag_xcs')

    def p2x(self, s, y, n):
        return self.p2xc(s) + y * self.p2ey(s) + n * self.p2en(s)

    def gate_position(self, s):
        # base_centerline.py:314-316
        return self.p2xc(s)

    def gate_orientation(self, s):
        # base_centerline.py:318-336
        es = self.p2es(s)
        ey = self.p2ey(s)
        en = self.p2en(s)
        if self.config.gate_snap_fit:
            if abs(es[2]) > 0.9:
                es = np.array([0, 0, 1])
                ey = ey - es * (ey.T @ es)
                ey = ey / np.linalg.norm(ey)
                en = np.cross(es, ey)
            elif abs(en[2]) > 0.9:
                en = np.array([0, 0, 1])
                es = es - en * (es.T @ en)
                es = es / np.linalg.norm(es)
                ey = np.cross(en, es)
        return np.array([es, ey, en]).T

    def y_max(self, s=0):
        return self.config.y_max

    def y_min(self, s=0):
        return self.config.y_min

    def n_max(self, s=0):
        return self.config.n_max

    def n_min(self, s=0):
        return self.config.n_min
