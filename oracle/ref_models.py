'''
Oracle restatement of the reference's dynamics models.  TEST INFRASTRUCTURE ONLY.

Follows drone3d/dynamics/rotations.py:39-112 (R, M for ESP quaternion / YPR angles),
drone3d/dynamics/drone_models.py:47-123 (global drone) and :249-292 (parametric pose
evolution), drone3d/dynamics/point_model.py:28-75 and :149-213, and
drone3d/dynamics/dynamics_model.py:91-114 (RK4), :166-198 (helper functions), :351-365 (bounds).

Each model exposes `zdot(z, u, param_terms=None)` that *builds the expression* for the given
arguments (the reference calls an SX Function, which substitutes -- same result), matrix-style
like the reference so products are accumulated in the same order.
'''
import numpy as np

from aircraft_trajectory_optimization_b200 import symbolic as sx
from .ref_centerline import RefSplineCenterline


def _sxarr(A):
    ''' object array with every entry an SX (numbers become constants) '''
    A = np.asarray(A, dtype=object)
    out = np.empty(A.shape, dtype=object)
    for idx, x in np.ndenumerate(A):
        out[idx] = x if isinstance(x, sx.SX) else sx.SX.const(float(x))
    return out


def _vec(*e):
    return _sxarr(list(e))


def _hat(v):
    # drone_models.py:105-110
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=object)


def _mt(A, B):
    ''' dense matrix product with SX-style accumulation (0 + a*b + ...) '''
    A = _sxarr(A)
    B = _sxarr(B)
    vec = B.ndim == 1
    if vec:
        B = B[:, None]
    out = np.empty((A.shape[0], B.shape[1]), dtype=object)
    for i in range(A.shape[0]):
        for j in range(B.shape[1]):
            acc = sx.SX.const(0)
            for k in range(A.shape[1]):
                acc = acc + A[i, k] * B[k, j]
            out[i, j] = acc
    return out[:, 0] if vec else out


def rotation(use_quat, r):
    ''' rotations.py:44-102: returns (R 3x3, M) for orientation variables r '''
    if use_quat:
        qi, qj, qk, qr = r
        R = np.array([
            [1 - 2 * qj ** 2 - 2 * qk ** 2, 2 * (qi * qj - qk * qr), 2 * (qi * qk + qj * qr)],
            [2 * (qi * qj + qk * qr), 1 - 2 * qi ** 2 - 2 * qk ** 2, 2 * (qj * qk - qi * qr)],
            [2 * (qi * qk - qj * qr), 2 * (qj * qk + qi * qr), 1 - 2 * qi ** 2 - 2 * qj ** 2],
        ], dtype=object) / (qi ** 2 + qj ** 2 + qk ** 2 + qr ** 2)
        M = 0.5 * np.array([[qr, -qk, qj], [qk, qr, -qi], [-qj, qi, qr], [-qi, -qj, -qk]],
                           dtype=object)
    else:
        a, b, c = r
        ca_, sa, cb, sb, cc, sc = sx.cos(a), sx.sin(a), sx.cos(b), sx.sin(b), sx.cos(c), sx.sin(c)
        Ra = np.array([[ca_, -sa, 0], [sa, ca_, 0], [0, 0, 1]], dtype=object)
        Rb = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]], dtype=object)
        Rc = np.array([[1, 0, 0], [0, cc, -sc], [0, sc, cc]], dtype=object)
        R = _mt(_mt(Ra, Rb), Rc)
        M = np.array([[0, sc / cb, cc / cb], [0, cc, -sc], [1, sc * sx.tan(b), cc * sx.tan(b)]],
                     dtype=object)
    return R, M


class RefModel:
    ''' shared surface: dims, bounds, rk4, helper expressions '''
    nz: int
    nu: int
    parametric = False
    line: RefSplineCenterline = None

    def __init__(self, config, line=None):
        self.config = config
        self.line = line
        self.parametric = line is not None

    def terms(self, z, u, param_terms=None):
        ''' dict with z_dot, R (global), vg, Tg '''
        raise NotImplementedError

    def zdot(self, z, u, param_terms=None):
        return self.terms(z, u, param_terms)['z_dot']

    def rk4(self, z0, u0, dt, param_terms=None):
        # dynamics_model.py:104-109 ; base_raceline.py:1042-1047
        k1 = self.zdot(z0, u0, param_terms)
        k2 = self.zdot(z0 + dt / 2 * k1, u0, param_terms)
        k3 = self.zdot(z0 + dt / 2 * k2, u0, param_terms)
        k4 = self.zdot(z0 + dt * k3, u0, param_terms)
        return z0 + dt / 6 * (k1 + k2 * 2 + k3 * 2 + k4)

    def _frame(self, z, param_terms):
        ''' symbolic/numeric frame for the given param_terms (numbers fold to constants) '''
        return RefSplineCenterline.sym_rep(param_terms)

    # bounds: dynamics_model.py:351-365 on top of the model's own
    def zu(self, s=0):
        zu = self._zu()
        if self.parametric:
            zu[0] = self.line.s_max()
            zu[1] = self.line.y_max(s=s)
            zu[2] = self.line.n_max(s=s)
        return zu

    def zl(self, s=0):
        zl = self._zl()
        if self.parametric:
            zl[0] = self.line.s_min()
            zl[1] = self.line.y_min(s=s)
            zl[2] = self.line.n_min(s=s)
        return zl

    def uu(self):
        return [self.config.T_max] * self.nu

    def ul(self):
        return [self.config.T_min] * self.nu

    def duu(self):
        return [self.config.dT_max] * self.nu

    def dul(self):
        return [self.config.dT_min] * self.nu

    def add_model_stage_constraints(self, z, u, g, lbg, ubg):
        pass

    # numeric helpers (dynamics_model.py:166-198): evaluate the expression at numbers
    def _numeric(self, key, z, u):
        pt = self.line.f_param_terms(z[0]) if self.parametric else None
        v = self.terms(_vec(*z), _vec(*u), pt)[key]
        return np.vectorize(lambda e: e.value() if isinstance(e, sx.SX) else float(e),
                            otypes=[float])(v)

    def f_R(self, z, u):
        return self._numeric('R', z, u)

    def f_T(self, z, u):
        return self._numeric('Tg', z, u)

    def f_vg(self, z, u):
        return self._numeric('vg', z, u)


class RefDroneModel(RefModel):
    ''' drone_models.py:12-123 (global) and :236-292 (parametric) '''
    nu = 4

    def __init__(self, config, line=None):
        super().__init__(config, line)
        self.nz = 13 if config.use_quat else 12

    def terms(self, z, u, param_terms=None):
        cfg = self.config
        nr = 4 if cfg.use_quat else 3
        p, r, vb, wb = z[:3], z[3:3 + nr], z[3 + nr:6 + nr], z[6 + nr:9 + nr]
        Rq, M = rotation(cfg.use_quat, r)

        if not self.parametric:
            # drone_models.py:47-59
            R = Rq
            vg = _mt(R, vb)
            p_dot = vg
            r_dot = _mt(M, wb)
            R_rel = None
        else:
            # drone_models.py:249-292
            fr = self._frame(z, param_terms)
            Rp = fr['Rp']
            R_rel = _mt(Rp.T, Rq) if cfg.global_r else Rq
            vp = _mt(R_rel, vb)
            y, n = p[1], p[2]
            s_dot = vp[0] / fr['mag_xcs'] / (1 + fr['ky'] * n - fr['kn'] * y)
            y_dot = vp[1] + n * fr['ks'] * s_dot * fr['mag_xcs']
            n_dot = vp[2] - y * fr['ks'] * s_dot * fr['mag_xcs']
            p_dot = _vec(s_dot, y_dot, n_dot)
            wp = fr['k'] * s_dot * fr['mag_xcs']
            w_eff = wb if cfg.global_r else wb - _mt(Rq.T, wp)
            r_dot = _mt(M, w_eff)
            R = Rq if cfg.global_r else _mt(Rp, Rq)
            vg = _mt(R, vb)

        # drone_models.py:61-92
        Fgb = -cfg.m * cfg.g * _vec(R[2, 0], R[2, 1], R[2, 2])
        Fdb = _vec(-cfg.b1, -cfg.b2, -cfg.b3) * vb
        Kdb = _vec(-cfg.bw1, -cfg.bw2, -cfg.bw3) * wb
        Tb = _vec(0, 0, u[0] + u[1] + u[2] + u[3])
        TKb = _vec((u[0] + u[1] - u[2] - u[3]) * cfg.l,
                   (-u[0] + u[1] + u[2] - u[3]) * cfg.l,
                   (u[0] - u[1] + u[2] - u[3]) * cfg.k)
        Fb = Fdb + Fgb + Tb
        Kb = Kdb + TKb
        Tg = _mt(R, Tb)

        # drone_models.py:94-123
        Wb = _hat(wb)
        Ib = np.diag([cfg.I1, cfg.I2, cfg.I3])
        vb_dot = Fb / cfg.m - _mt(Wb, vb)
        wb_dot = _mt(np.linalg.inv(Ib), Kb - _mt(_mt(Wb, Ib), wb))
        z_dot = np.concatenate([p_dot, r_dot, vb_dot, wb_dot])
        return dict(z_dot=z_dot, R=R, vg=vg, Tg=Tg, R_rel=R_rel)

    def _zu(self):
        # drone_models.py:185-198 ; rotations.py:130-145
        c = self.config
        if c.use_quat:
            ubr = [np.inf] * 4
        elif c.global_r:
            ubr = [np.inf, np.pi / 2.1, np.pi / 2.1]
        else:
            ubr = [np.pi / 2, np.pi / 2.1, np.pi / 2.1]
        return [np.inf] * 3 + ubr + [np.inf] * 3 + [c.w_max] * 3

    def _zl(self):
        c = self.config
        if c.use_quat:
            lbr = [-np.inf] * 4
        elif c.global_r:
            lbr = [-np.inf, -np.pi / 2.1, -np.pi / 2.1]
        else:
            lbr = [-np.pi / 2, -np.pi / 2.1, -np.pi / 2.1]
        return [-np.inf] * 3 + lbr + [-np.inf] * 3 + [c.w_min] * 3


class RefPointModel(RefModel):
    ''' point_model.py:13-129 (global) and :131-213 (parametric) '''
    nz = 6
    nu = 3

    def terms(self, z, u, param_terms=None):
        cfg = self.config
        p, vb = z[:3], z[3:6]
        if not self.parametric:
            # point_model.py:28-40
            p_dot = vb
            R = np.eye(3)
            vg = vb
            wb = np.zeros(3)
        else:
            # point_model.py:149-188
            fr = self._frame(z, param_terms)
            Rp = fr['Rp']
            R_rel = Rp.T if cfg.global_r else np.eye(3)
            vp = _mt(R_rel, vb)
            y, n = p[1], p[2]
            s_dot = vp[0] / fr['mag_xcs'] / (1 + fr['ky'] * n - fr['kn'] * y)
            y_dot = vp[1] + n * fr['ks'] * s_dot * fr['mag_xcs']
            n_dot = vp[2] - y * fr['ks'] * s_dot * fr['mag_xcs']
            p_dot = _vec(s_dot, y_dot, n_dot)
            wp = fr['k'] * s_dot * fr['mag_xcs']
            wb = np.zeros(3) if cfg.global_r else wp
            R = np.eye(3) if cfg.global_r else Rp
            vg = _mt(R, vb)
        # point_model.py:42-59
        Tb = u
        Fgb = -cfg.m * cfg.g * _vec(R[2, 0], R[2, 1], R[2, 2])
        Fdb = _vec(-cfg.b1, -cfg.b2, -cfg.b3) * vb
        Fb = Tb + Fgb + Fdb
        Tg = _mt(R, Tb)
        # point_model.py:61-75 / :190-213
        vb_dot = Fb / cfg.m
        if self.parametric:
            vb_dot = vb_dot - _mt(_hat(wb), vb)
        z_dot = np.concatenate([p_dot, vb_dot])
        return dict(z_dot=z_dot, R=R, vg=vg, Tg=Tg)

    def _zu(self):
        return [np.inf] * 6

    def _zl(self):
        return [-np.inf] * 6

    def add_model_stage_constraints(self, z, u, g, lbg, ubg):
        # point_model.py:122-129
        u_mag = u @ u
        g += [np.array([u_mag / self.config.T_max / self.config.T_max], dtype=object)]
        ubg += [1]
        lbg += [-np.inf]
