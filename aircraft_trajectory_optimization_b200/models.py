'''
Vehicle models: symbolic right-hand sides for the build-time code generator, plus the numeric
host-side pieces of the reference's model API (bounds, helper functions, state packing).

Equations: reference drone3d/dynamics/drone_models.py:47-123 (global drone), :249-292 (curvilinear
pose), drone3d/dynamics/point_model.py:28-75 / :149-213, drone3d/dynamics/rotations.py:44-102;
summarised in SURVEY.md App. B.  Written here in scalar form over named parameters so that one
generated function serves every vehicle of a batch:

  x  = [z ; u]                       state then input
  fc = [Rp (row-major, columns es|ey|en), ks, ky, kn, |xcs|]      13 per-point frame constants
  vp = vehicle parameters (VP_DRONE / VP_POINT below), per problem

A *variant* fixes everything that changes the structure of the code:
  vehicle 'drone'|'point', orientation 'quat'|'ypr' (drone), frame 'global'|'param_gr'|'param_lr'
  (curvilinear frame with global / frame-relative orientation), drag False|True (linear drag
  coefficients all zero or not -- with b == 0 the reference's SX graph drops those terms and the
  sparsity pattern changes).
'''
from dataclasses import dataclass

import numpy as np

from . import symbolic as sx
from .pytypes import DroneConfig, PointConfig, RacerConfig, DroneState, PointState, \
    GlobalQuaternion, GlobalEulerAngles, RelativeQuaternion, RelativeEulerAngles, \
    quat_to_matrix, ypr_to_matrix

VP_DRONE = ['m', 'g', 'I1', 'I2', 'I3', 'l', 'k', 'bw1', 'bw2', 'bw3', 'b1', 'b2', 'b3']
VP_POINT = ['m', 'g', 'b1', 'b2', 'b3', 'T_max']
NFC = 13


@dataclass(frozen=True)
class Variant:
    vehicle: str = 'drone'
    orient: str = 'quat'
    frame: str = 'global'
    drag: bool = False

    @property
    def name(self):
        o = self.orient if self.vehicle == 'drone' else 'pm'
        return f'{self.vehicle}_{o}_{self.frame}' + ('_drag' if self.drag else '')

    @property
    def nz(self):
        if self.vehicle == 'point':
            return 6
        return 13 if self.orient == 'quat' else 12

    @property
    def nu(self):
        return 4 if self.vehicle == 'drone' else 3

    @property
    def nr(self):
        return 0 if self.vehicle == 'point' else (4 if self.orient == 'quat' else 3)

    @property
    def parametric(self):
        return self.frame != 'global'

    @property
    def vp_names(self):
        return VP_DRONE if self.vehicle == 'drone' else VP_POINT


def variant_of(config: RacerConfig, parametric: bool) -> Variant:
    ''' the code variant a vehicle config + frame choice maps to '''
    drag = not (config.b1 == 0 and config.b2 == 0 and config.b3 == 0)
    frame = 'global' if not parametric else ('param_gr' if config.global_r else 'param_lr')
    if isinstance(config, DroneConfig):
        return Variant('drone', 'quat' if config.use_quat else 'ypr', frame, drag)
    return Variant('point', 'quat', frame, drag)


def vehicle_params(config: RacerConfig) -> np.ndarray:
    ''' the vp vector of a config, in VP_DRONE / VP_POINT order '''
    names = VP_DRONE if isinstance(config, DroneConfig) else VP_POINT
    return np.array([float(getattr(config, k)) for k in names])


# ------------------------------------------------------------------------------------------
# symbolic right-hand side
# ------------------------------------------------------------------------------------------
def _rot_quat(q):
    qi, qj, qk, qr = q
    den = qi * qi + qj * qj + qk * qk + qr * qr
    R = [[1 - 2 * (qj * qj) - 2 * (qk * qk), 2 * (qi * qj - qk * qr), 2 * (qi * qk + qj * qr)],
         [2 * (qi * qj + qk * qr), 1 - 2 * (qi * qi) - 2 * (qk * qk), 2 * (qj * qk - qi * qr)],
         [2 * (qi * qk - qj * qr), 2 * (qj * qk + qi * qr), 1 - 2 * (qi * qi) - 2 * (qj * qj)]]
    return [[e / den for e in row] for row in R]


def _rot_ypr(r):
    a, b, c = r
    ca_, sa, cb, sb, cc, sc = sx.cos(a), sx.sin(a), sx.cos(b), sx.sin(b), sx.cos(c), sx.sin(c)
    # Ra(a) Rb(b) Rc(c) multiplied out
    return [[ca_ * cb, ca_ * sb * sc - sa * cc, ca_ * sb * cc + sa * sc],
            [sa * cb, sa * sb * sc + ca_ * cc, sa * sb * cc - ca_ * sc],
            [-sb, cb * sc, cb * cc]]


def _rdot(variant, r, w):
    if variant.orient == 'quat':
        qi, qj, qk, qr = r
        return [0.5 * (qr * w[0] - qk * w[1] + qj * w[2]),
                0.5 * (qk * w[0] + qr * w[1] - qi * w[2]),
                0.5 * (-qj * w[0] + qi * w[1] + qr * w[2]),
                0.5 * (-qi * w[0] - qj * w[1] - qk * w[2])]
    a, b, c = r
    cb, cc, sc, tb = sx.cos(b), sx.cos(c), sx.sin(c), sx.tan(b)
    return [sc / cb * w[1] + cc / cb * w[2],
            cc * w[1] - sc * w[2],
            w[0] + sc * tb * w[1] + cc * tb * w[2]]


def _matvec(M, v):
    return [M[i][0] * v[0] + M[i][1] * v[1] + M[i][2] * v[2] for i in range(3)]


def _matTvec(M, v):
    return [M[0][i] * v[0] + M[1][i] * v[1] + M[2][i] * v[2] for i in range(3)]


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def zdot(variant: Variant, z, u, fc, vp):
    '''
    state derivative as a list of SX.  z, u, fc, vp: sequences of SX (fc unused for frame
    'global'; vp indexed by VP_DRONE / VP_POINT).
    '''
    P = dict(zip(variant.vp_names, vp))
    nr = variant.nr
    p = z[:3]
    r = z[3:3 + nr]
    vb = z[3 + nr:6 + nr]
    if variant.parametric:
        Rp = [[fc[3 * i + j] for j in range(3)] for i in range(3)]
        ks, ky, kn, mag = fc[9], fc[10], fc[11], fc[12]

    if variant.vehicle == 'drone':
        wb = z[6 + nr:9 + nr]
        Rq = _rot_quat(r) if variant.orient == 'quat' else _rot_ypr(r)
    else:
        Rq = None

    # pose kinematics
    w_frame = None
    if not variant.parametric:
        if variant.vehicle == 'drone':
            p_dot = _matvec(Rq, vb)
            Rg = Rq
        else:
            p_dot = list(vb)
            Rg = None                      # identity
    else:
        if variant.vehicle == 'drone':
            if variant.frame == 'param_gr':
                vpar = _matTvec(Rp, _matvec(Rq, vb))
            else:
                vpar = _matvec(Rq, vb)
        else:
            vpar = _matTvec(Rp, vb) if variant.frame == 'param_gr' else list(vb)
        y, n = p[1], p[2]
        s_dot = vpar[0] / mag / (1 + ky * n - kn * y)
        adv = s_dot * mag
        p_dot = [s_dot, vpar[1] + n * ks * adv, vpar[2] - y * ks * adv]
        w_frame = [ks * adv, ky * adv, kn * adv]
        if variant.vehicle == 'drone':
            if variant.frame == 'param_gr':
                Rg = Rq
            else:
                Rg = [[Rp[i][0] * Rq[0][j] + Rp[i][1] * Rq[1][j] + Rp[i][2] * Rq[2][j]
                       for j in range(3)] for i in range(3)]
        else:
            Rg = None if variant.frame == 'param_gr' else Rp

    # gravity in the body frame: -m g (third row of the global rotation)
    if Rg is None:
        grav = [0, 0, -P['g']]
    else:
        grav = [-P['g'] * Rg[2][i] for i in range(3)]

    if variant.vehicle == 'drone':
        if variant.frame == 'param_lr':
            w_eff = [wb[i] - e for i, e in enumerate(_matTvec(Rq, w_frame))]
        else:
            w_eff = wb
        r_dot = _rdot(variant, r, w_eff)
        thrust = [0, 0, (u[0] + u[1] + u[2] + u[3]) / P['m']]
        wxv = _cross(wb, vb)
        vb_dot = [grav[i] + thrust[i] - wxv[i] for i in range(3)]
        if variant.drag:
            vb_dot = [vb_dot[i] - P[f'b{i + 1}'] * vb[i] / P['m'] for i in range(3)]
        Iw = [P['I1'] * wb[0], P['I2'] * wb[1], P['I3'] * wb[2]]
        wxIw = _cross(wb, Iw)
        tau = [(u[0] + u[1] - u[2] - u[3]) * P['l'] - P['bw1'] * wb[0],
               (-u[0] + u[1] + u[2] - u[3]) * P['l'] - P['bw2'] * wb[1],
               (u[0] - u[1] + u[2] - u[3]) * P['k'] - P['bw3'] * wb[2]]
        wb_dot = [(tau[i] - wxIw[i]) / P[f'I{i + 1}'] for i in range(3)]
        return [*p_dot, *r_dot, *vb_dot, *wb_dot]

    vb_dot = [grav[i] + u[i] / P['m'] for i in range(3)]
    if variant.drag:
        vb_dot = [vb_dot[i] - P[f'b{i + 1}'] * vb[i] / P['m'] for i in range(3)]
    if variant.frame == 'param_lr':
        wxv = _cross(w_frame, vb)
        vb_dot = [vb_dot[i] - wxv[i] for i in range(3)]
    return [*p_dot, *vb_dot]


def end_terms(variant: Variant, z, u, vp):
    '''
    the helper expressions the open-track end rows use (dynamics_model.py:166-198: ca_f_vg, ca_f_R, ca_f_T), global
    frame only: dict(vg = global-frame velocity, e3 = third column of the rotation matrix (None for the point mass),
    Tg = global-frame thrust (point mass: the input itself, point_model.py:28-58)).
    '''
    if variant.parametric:
        raise NotImplementedError('end-row helper expressions of parametric models need the centerline spline '
                                  'inside the graph (base_centerline.py:117-154); the reference cannot build them either')
    nr = variant.nr
    vb = z[3 + nr:6 + nr]
    if variant.vehicle == 'drone':
        r = z[3:3 + nr]
        Rq = _rot_quat(r) if variant.orient == 'quat' else _rot_ypr(r)
        return dict(vg=_matvec(Rq, vb), e3=[Rq[0][2], Rq[1][2], Rq[2][2]], Tg=None)
    return dict(vg=list(vb), e3=None, Tg=list(u))


# ------------------------------------------------------------------------------------------
# host-side model objects (reference API: dynamics_model.py:55-250, :253-365)
# ------------------------------------------------------------------------------------------
class DynamicsModel:
    '''
    numeric model object: bounds, helper functions f_R / f_T / f_Fg / f_vg, state packing and a
    `step` simulator.  The NLP never evaluates these on the host; the kernels run the generated
    code of `self.variant`.
    '''
    config: RacerConfig
    line = None

    def __init__(self, config: RacerConfig, line=None):
        self.config = config
        self.line = line
        self.variant = variant_of(config, line is not None)
        self.nz, self.nu = self.variant.nz, self.variant.nu
        self._rhs = None
        self._last_q = None

    # ---- numeric right-hand side through the expression engine ----------------------------
    def _build_rhs(self):
        g = sx.Graph()
        prev = sx.graph()
        sx.set_graph(g)
        try:
            z = list(sx.SX.sym('z', self.nz))
            u = list(sx.SX.sym('u', self.nu))
            fc = list(sx.SX.sym('fc', NFC))
            vp = list(sx.SX.sym('vp', len(self.variant.vp_names)))
            out = [sx._id(e) for e in zdot(self.variant, z, u, fc, vp)]
        finally:
            sx.set_graph(prev)
        self._rhs = (g, out, g.reachable(out))

    def f_zdot_full(self, z, u, fc):
        ''' z_dot at explicit frame constants fc (13,) '''
        if self._rhs is None:
            self._build_rhs()
        g, out, nodes = self._rhs
        vals = [*z, *u, *fc, *vehicle_params(self.config)]
        return np.array(g.evaluate(out, vals, nodes), dtype=float)

    def f_zdot(self, z, u):
        fc = self.line.frame_constants(z[0])[0] if self.line is not None else np.zeros(NFC)
        return self.f_zdot_full(z, u, fc)

    def step(self, state):
        ''' integrate one config.dt forward (reference: IDAS, dynamics_model.py:81-89; here LSODA) '''
        from scipy.integrate import solve_ivp
        z, u = self.state2zu(state)
        sol = solve_ivp(lambda t, zz: self.f_zdot(zz, u), (0, self.config.dt), np.array(z, dtype=float),
                        rtol=1e-10, atol=1e-12, method='LSODA')
        self.zu2state(state, sol.y[:, -1], u)

    # ---- helper functions -------------------------------------------------------------------
    def _orientation(self, z):
        ''' rotation matrix of the orientation variables alone '''
        if self.variant.vehicle == 'point':
            return np.eye(3)
        r = np.asarray(z[3:3 + self.variant.nr], dtype=float)
        if self.variant.orient == 'quat':
            return quat_to_matrix(r) / float(r @ r)
        return ypr_to_matrix(*r)

    def f_R(self, z, u=None):
        ''' body -> global rotation '''
        Rq = self._orientation(z)
        if self.variant.frame == 'param_lr':
            return self.line.p2Rp(z[0]) @ Rq
        return Rq

    def f_T(self, z, u):
        ''' global-frame thrust vector '''
        u = np.asarray(u, dtype=float)
        Tb = np.array([0., 0., u.sum()]) if self.variant.vehicle == 'drone' else u
        return self.f_R(z, u) @ Tb

    def f_Fg(self, z, u=None):
        ''' gravity force in the body frame '''
        return -self.config.m * self.config.g * self.f_R(z, u)[2, :]

    def f_vg(self, z, u=None):
        ''' global-frame velocity '''
        nr = self.variant.nr
        return self.f_R(z, u) @ np.asarray(z[3 + nr:6 + nr], dtype=float)

    # ---- bounds (drone_models.py:185-229, point_model.py:104-120, dynamics_model.py:351-365) -
    def _orient_bounds(self):
        if self.variant.vehicle == 'point':
            return []
        if self.variant.orient == 'quat':
            return [np.inf] * 4
        first = np.inf if self.config.global_r else np.pi / 2
        return [first, np.pi / 2.1, np.pi / 2.1]

    def zu(self, s=0):
        c = self.config
        zu = [np.inf] * 3 + self._orient_bounds() + [np.inf] * 3
        if self.variant.vehicle == 'drone':
            zu += [c.w_max] * 3
        if self.line is not None:
            zu[0:3] = [self.line.s_max(), self.line.y_max(s=s), self.line.n_max(s=s)]
        return zu

    def zl(self, s=0):
        c = self.config
        zl = [-np.inf] * 3 + [-b for b in self._orient_bounds()] + [-np.inf] * 3
        if self.variant.vehicle == 'drone':
            zl += [c.w_min] * 3
        if self.line is not None:
            zl[0:3] = [self.line.s_min(), self.line.y_min(s=s), self.line.n_min(s=s)]
        return zl

    def uu(self):
        return [self.config.T_max] * self.nu

    def ul(self):
        return [self.config.T_min] * self.nu

    def duu(self):
        return [self.config.dT_max] * self.nu

    def dul(self):
        return [self.config.dT_min] * self.nu

    # ---- state packing ------------------------------------------------------------------------
    def get_empty_state(self):
        if self.variant.vehicle == 'point':
            return PointState()
        c = self.config
        if c.global_r:
            r = GlobalQuaternion() if c.use_quat else GlobalEulerAngles()
        else:
            r = RelativeQuaternion() if c.use_quat else RelativeEulerAngles()
        return DroneState(r=r)

    def state2u(self, state):
        return state.u.to_vec()

    def state2zu(self, state):
        pos = state.p.to_vec() if self.line is not None else state.x.to_vec()
        z = [*pos]
        if self.variant.vehicle == 'drone':
            z += [*state.r.to_vec(), *state.v.to_vec(), *state.w.to_vec()]
        else:
            z += [*state.v.to_vec()]
        return z, self.state2u(state)

    def u2state(self, state, u):
        state.u.from_vec(u)

    def du2state(self, state, du):
        state.du.from_vec(du)

    def zu2state(self, state, z, u):
        ''' drone_models.py:162-183 / :306-328, point_model.py:98-102 / :239-252 '''
        self.u2state(state, u)
        nr = self.variant.nr
        if self.line is not None:
            state.p.from_vec(z[:3])
            state.x.from_vec(self.line.p2x(*z[:3]))
        else:
            state.x.from_vec(z[:3])
        state.v.from_vec(z[3 + nr:6 + nr])
        if self.variant.vehicle == 'drone':
            state.r.from_vec(z[3:3 + nr])
            state.w.from_vec(z[6 + nr:9 + nr])
            if self.line is None and self.variant.orient == 'quat':
                state.q.from_vec(z[3:7])
                return
            state.q.from_mat(self.f_R(z, u))
            if self._last_q is not None and np.linalg.norm(self._last_q - state.q.to_vec()) > 1.8:
                state.q.from_vec(-state.q.to_vec())
            self._last_q = state.q.to_vec()
        elif self.line is not None:
            state.q.from_mat(self.f_R(z, u))


class DroneModel(DynamicsModel):
    ''' inertial-frame quadrotor (drone_models.py:12) '''

    def __init__(self, config: DroneConfig):
        super().__init__(config, None)


class ParametricDroneModel(DynamicsModel):
    ''' quadrotor in the curvilinear frame of a centerline (drone_models.py:236) '''

    def __init__(self, config: DroneConfig, line):
        super().__init__(config, line)


class PointModel(DynamicsModel):
    ''' inertial-frame point mass (point_model.py:13) '''

    def __init__(self, config: PointConfig):
        super().__init__(config, None)


class ParametricPointModel(DynamicsModel):
    ''' point mass in the curvilinear frame (point_model.py:131) '''

    def __init__(self, config: PointConfig, line):
        super().__init__(config, line)
