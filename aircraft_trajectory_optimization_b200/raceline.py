'''
Raceline problem builders with the reference's Python API.

Class names, constructor signatures, `.solve()` / `.get_ws()`, `solver_w0/lbw/ubw/lbg/ubg`, `model`,
`ws_raceline / ws_model / ws_solver`, `tube`, `setup_time / solve_time / ipopt_time / feval_time`
follow drone3d/raceline/base_raceline.py (:27-98 configs and results, :100-199 BaseRaceline,
:867-937 global frame, :940-1251 parametric frame, :1254-1331 obstacle variant),
drone3d/raceline/drone_raceline.py (:24-427) and drone3d/raceline/point_raceline.py (:12-90).

Where the reference appends CasADi SX rows to nlp['g'] and calls `ca.nlpsol` (:799), these builders
append row *records* to a StructureBuilder in the same order and hand the
resulting NLPStructure to the CUDA library; the solver object mirrors `nlpsol`'s call signature
(`solver(x0=, lbx=, ubx=, lbg=, ubg=)` -> dict with 'x', 'f', 'g', 'lam_x', 'lam_g'; `.stats()`).
'''
from dataclasses import dataclass, field
from typing import List, Callable
import time

import numpy as np

from .pytypes import PythonMsg, RacerConfig, RacerState, DroneConfig, PointConfig, matrix_to_quat
from .centerlines import BaseCenterline, GateShape
from .collocation import get_collocation_coefficients, get_intermediate_collocation_coefficients
from .models import DynamicsModel, DroneModel, ParametricDroneModel, PointModel, \
    ParametricPointModel, VP_POINT
from .structure import StructureBuilder, NLPStructure, RK4, COLLOC
from . import codegen


@dataclass
class RacelineConfig(PythonMsg):
    ''' configuration of a raceline solver (base_raceline.py:27-57) '''
    verbose: bool = True
    plot_iterations: bool = False
    N: int = 30
    K: int = 7
    use_rk4: bool = False
    R: np.ndarray = 1e-7
    dR: np.ndarray = 1e-7
    h0: float = 1
    v0: float = 1
    closed: bool = False
    fix_gate_center: bool = False
    max_iter: int = 1000
    hsl_linear_solver: str = 'ma97'     # kept for API parity; the KKT solve is the library's own


@dataclass
class GlobalRacelineConfig(RacelineConfig):
    ''' global-frame raceline (base_raceline.py:61-66) '''
    gate_xi: np.ndarray = None
    gate_xj: np.ndarray = None
    gate_xk: np.ndarray = None


@dataclass
class ParametricRacelineConfig(RacelineConfig):
    ''' curvilinear-frame raceline (base_raceline.py:69-78) '''
    fixed_gates: List[float] = None
    force_regularity: bool = True


@dataclass
class RacelineResults(PythonMsg):
    ''' results of a raceline solve (base_raceline.py:82-98) '''
    solve_time: float = None
    ipopt_time: float = None
    feval_time: float = None
    feasible: bool = None
    states: List[RacerState] = None
    step_sizes: List[float] = None
    time: float = None
    periodic: bool = False
    label: str = None
    color: List[float] = None
    z_interp: Callable[[float], np.ndarray] = None
    u_interp: Callable[[float], np.ndarray] = None
    du_interp: Callable[[float], np.ndarray] = None
    global_frame: bool = None


@dataclass
class ObstacleFreeTube:
    '''
    obstacle-free tube samples (drone3d/obstacles/mesh_obstacle.py:198-237): parametric centre
    offsets ball_p[:, (s, dy, dn)], radii ball_r and the vehicle collision radius.
    '''
    ball_p: np.ndarray
    ball_r: np.ndarray
    collision_r: float
    line: BaseCenterline = None
    ball_center: np.ndarray = None
    ball_tangent_pts: np.ndarray = None

    def disc(self, s):
        ''' (delta_y, delta_n, available radius) of the sample nearest in s '''
        idx = int(np.argmin(np.abs(self.ball_p[:, 0] - s)))
        avail = max(self.ball_r[idx] - self.collision_r, 0.01)
        return self.ball_p[idx, 1], self.ball_p[idx, 2], avail


class BaseRaceline:
    ''' shared builder; subclasses set the frame / vehicle flags and labels '''
    parametric = False
    drone = False
    global_frame: bool = None
    label = ''
    color = [1, 0, 0, 1]

    setup_time: float = -1
    solve_time: float = -1
    ipopt_time: float = -1
    feval_time: float = -1

    def __init__(self, line: BaseCenterline, config: RacelineConfig, vehicle_config: RacerConfig,
                 ws_raceline: RacelineResults = None, ws_model: DynamicsModel = None):
        self.line = line
        self.config = config
        self.vehicle_config = vehicle_config
        self.ws_raceline = ws_raceline
        self.ws_model = ws_model
        self._first_ws_r = None
        self._last_ws_r = None
        self.solver = None
        self.sol = None
        self._setup()

    # ---- public API -------------------------------------------------------------------------
    def solve(self) -> RacelineResults:
        ''' solve the raceline NLP on the GPU and unpack (base_raceline.py:157-191) '''
        t0 = time.time()
        sol = self.solver(x0=self.solver_w0, ubx=self.solver_ubw, lbx=self.solver_lbw,
                          ubg=self.solver_ubg, lbg=self.solver_lbg)
        self.solve_time = time.time() - t0
        self.sol = sol
        if self.config.verbose:
            H = np.asarray(sol['x'])[:self.config.N]
            print(f'reached target in    {np.sum(H):0.3f} seconds')
            print(f'with a total cost of {float(sol["f"]):0.3f}')
            print(f'it took              {self.setup_time:0.3f} seconds to set up the problem')
            print(f'    and              {self.solve_time:0.3f} seconds to solve it')
        stats = self.solver.stats()
        self.feval_time = sum(stats[k] for k in ('t_wall_nlp_f', 't_wall_nlp_g', 't_wall_nlp_grad_f',
                                                 't_wall_nlp_hess_l', 't_wall_nlp_jac_g'))
        self.ipopt_time = self.solve_time - self.feval_time
        return self._unpack_soln(sol)

    def get_ws(self) -> RacelineResults:
        ''' the initial guess packaged as results (base_raceline.py:193-198) '''
        self.solve_time = self.ipopt_time = self.feval_time = -1
        return self._unpack_soln({'x': self.solver_w0})

    # ---- set-up -------------------------------------------------------------------------------
    def _get_model(self, config) -> DynamicsModel:
        raise NotImplementedError

    def _ws_available(self):
        return self.ws_model is not None and self.ws_raceline is not None

    def _setup(self):
        t0 = time.time()
        self.model = self._get_model(self.vehicle_config)
        self._setup_checks()
        self._create_nlp()
        self._create_solver()
        self.setup_time = time.time() - t0

    def _setup_checks(self):
        cfg = self.config
        if self.parametric:
            if not self.line.cleanly_closed and not self.model.config.global_r:
                raise NotImplementedError('Global orientation must be used for skewly closed centerlines')
        if cfg.use_rk4:
            # base_raceline.py:226-230 (mutates the config exactly like the reference)
            cfg.h0 /= cfg.K
            cfg.N *= cfg.K
            cfg.K = 0
        if not self.parametric:
            x = np.array([cfg.gate_xi, cfg.gate_xj, cfg.gate_xk])
            if cfg.closed and not (x[:, 0] == x[:, -1]).all():
                x = np.hstack([x, x[:, 0:1]])
            num_phases = x.shape[1] - 1
            cfg.N = int(num_phases * np.ceil(cfg.N / num_phases))
            self.gate_n_interval = int(cfg.N / num_phases)

    def _get_s(self, n, k):
        ''' fixed path length of collocation point (n, k) (base_raceline.py:972-984) '''
        ds = (self.line.s_max() - self.line.s_min()) / self.config.N
        if self.config.use_rk4:
            return self.line.s_min() + ds * n
        return self.line.s_min() + ds * (n + self.tau[k])

    def _create_nlp(self):
        cfg = self.config
        nu = self.model.nu
        if isinstance(cfg.R, (float, int)):
            cfg.R = np.eye(nu) * cfg.R
        if isinstance(cfg.dR, (float, int)):
            cfg.dR = np.eye(nu) * cfg.dR
        for M in (cfg.R, cfg.dR):
            if np.count_nonzero(M - np.diag(np.diagonal(M))):
                raise NotImplementedError('only diagonal R / dR cost matrices are supported')
        N, K = cfg.N, cfg.K
        if not cfg.use_rk4:
            self.tau, self.B, self.C, self.D = get_collocation_coefficients(K)
        fc = None
        if self.parametric:
            s_all = np.array([self._get_s(n, k) for n in range(N) for k in range(K + 1)])
            fc = self.line.frame_constants(s_all)
        meta = codegen.load_meta(self.model.variant)
        self.sb = StructureBuilder(RK4 if cfg.use_rk4 else COLLOC, meta, N, K,
                                   np.diagonal(cfg.R), np.diagonal(cfg.dR), fc)
        if not cfg.use_rk4:
            self.sb.set_collocation(self.tau, self.B, self.C, self.D)
        self._enforce_model()
        self._add_gate_constraints()
        self._create_problem()

    # ---- model rows -----------------------------------------------------------------------------
    def _enforce_model(self):
        cfg, sb = self.config, self.sb
        if not self.parametric:
            # equal step sizes inside a gate-to-gate phase (base_raceline.py:891-905)
            for n in range(0, cfg.N, self.gate_n_interval):
                for n2 in range(n + 1, n + self.gate_n_interval):
                    sb.add_affine([sb.iH(n2), sb.iH(n)], [1.0, -1.0], 0.0, 0., 0.)
        for n in range(cfg.N):
            if cfg.use_rk4:
                self._enforce_rk4_interval(n)
            else:
                self._enforce_collocation_interval(n)
        if cfg.closed:
            if not self.drone:
                self._enforce_loop_closure()
        else:
            self._enforce_initial_constraints()
            self._enforce_terminal_constraints()

    # ---- open tracks: expression rows at both ends (tail.py) -----------------------------------
    def _tail(self):
        if self.sb.tail is None:
            from .tail import TailRows
            self.sb.tail = TailRows(self.sb, self.model.variant)
        return self.sb.tail

    def _zF_uF(self):
        ''' end state and input of the last interval (base_raceline.py:322-348) '''
        t = self._tail()
        D = None if self.config.use_rk4 else self.D
        return t.zF(D), t.uF(D)

    def _end_rows(self, z, u):
        ''' base_raceline.py:516-543, drone_raceline.py:110-148, point_raceline.py:15-45 '''
        t = self._tail()
        tm = t.terms(z, u)
        vg = tm['vg']
        t.add_rows([vg[0] * vg[0] + vg[1] * vg[1] + vg[2] * vg[2]], -np.inf, 0.)
        if self.drone:
            t.add_rows(tm['e3'], [0., 0., 1.], [0., 0., 1.])
            t.add_rows(z[-3:], 0., 0.)
        else:
            t.add_rows(tm['Tg'][:2], 0., 0.)

    def _enforce_initial_constraints(self):
        if self.parametric:
            # the reference's helper functions of parametric models evaluate the centerline spline at the symbolic
            # path length (base_centerline.py:117-154), which its SX graph cannot hold either
            raise NotImplementedError('open racelines are built for the global frame only')
        z, u, _ = self._tail().point(0, 0)
        self._end_rows(z, u)

    def _enforce_terminal_constraints(self):
        self._end_rows(*self._zF_uF())

    def _fix_gate_expr(self, x, s, include_axial_fix):
        ''' gate rows for a position given as expressions (the end state zF): base_raceline.py:545-595 '''
        t = self._tail()
        t.use()
        gate_x = self.line.gate_position(s)
        rc = self.model.config.collision_radius
        lcfg = self.line.config
        if self.config.fix_gate_center:
            t.add_rows([x[i] - float(gate_x[i]) for i in range(3)], 0., 0.)
            return
        R = self.line.gate_orientation(s)
        d = [x[i] - float(gate_x[i]) for i in range(3)]

        def dot(vec, col):
            acc = 0
            for i in range(3):
                acc = acc + vec[i] * float(R[i, col])
            return acc
        if lcfg.gate_shape == GateShape.CIRCLE:
            a, b = dot(d, 1), dot(d, 2)
            t.add_rows([a * a + b * b], -np.inf, (lcfg.gate_ri - rc) ** 2)
            if include_axial_fix:
                t.add_rows([dot(x, 0) - float(gate_x @ R[:, 0])], 0., 0.)
        elif lcfg.gate_shape == GateShape.SQUARE:
            d_max = lcfg.gate_ri - rc
            delta = [dot(d, i) for i in range(3)]
            if include_axial_fix:
                t.add_rows(delta, [0., -d_max, -d_max], [0., d_max, d_max])
            else:
                t.add_rows(delta[1:], [-d_max, -d_max], [d_max, d_max])
        else:
            raise NotImplementedError('Unhandled Gate Shape')

    def _stage_constraint(self, n, k):
        ''' point-mass thrust sphere u'u / T_max^2 <= 1 (point_model.py:122-129) '''
        if self.drone:
            return
        sb = self.sb
        nu = self.model.nu
        sb.add_squares([sb.iU(n, k, j) for j in range(nu)], np.eye(nu), np.zeros(nu), -np.inf, 1,
                       scale_vp=VP_POINT.index('T_max'))

    def _regularity(self, n, k):
        ''' kn*y - ky*n <= gamma where the frame curves strongly (base_raceline.py:1105-1129) '''
        if not (self.parametric and self.config.force_regularity):
            return
        sb = self.sb
        s = self._get_s(n, k)
        ky, kn = float(self.line.p2ky(s)), float(self.line.p2kn(s))
        if ky ** 2 + kn ** 2 > 0.1:
            sb.add_affine([sb.iZ(n, k, 1), sb.iZ(n, k, 2)], [kn, -ky], 0.0, -np.inf, self.line.config.gamma)

    def _enforce_rk4_interval(self, n):
        ''' base_raceline.py:363-391 (global) / :1052-1112 (parametric) '''
        cfg, sb = self.config, self.sb
        nz, nu = self.model.nz, self.model.nu
        if self.parametric:
            sb.add_affine([sb.iZ(n, 0, 0)], [1.0], -self._get_s(n, 0), 0., 0.)
        if n == cfg.N - 1:
            return
        first = 1 if self.parametric else 0
        rows = sb.alloc_rows(nz - first, 0., 0.)
        for c, r in zip(range(first, nz), rows):
            self._cell_out_row(n, c, r, -1.0, sb.iZ(n + 1, 0, c), 1.0, 0.0)
        rows = sb.alloc_rows(nu, 0., 0.)
        for j, r in enumerate(rows):
            self._cell_out_row(n, nz + j, r, -1.0, sb.iU(n + 1, 0, j), 1.0, 0.0)
        sb.cell_par[n, 0] = 0.5          # un = U + dU*h/2  
        if self.parametric:
            r = sb.alloc_rows(1, 0., 0.)[0]
            self._cell_out_row(n, 0, r, 1.0, -1, 0.0, -self._get_s(n + 1, 0))
        self._stage_constraint(n, 0)
        self._regularity(n, 0)

    def _cell_out_row(self, n, c, row, coef, partner, pcoef, off):
        sb = self.sb
        sb.cell_row[n, c] = row
        sb.cell_coef[n, c] = coef
        sb.cell_partner[n, c] = partner
        sb.cell_pcoef[n, c] = pcoef
        sb.cell_off[n, c] = off

    def _enforce_collocation_interval(self, n):
        ''' base_raceline.py:398-490 and the parametric overrides :1114-1181 '''
        cfg, sb = self.config, self.sb
        K = cfg.K
        nz, nu = self.model.nz, self.model.nu
        # ode rows
        for k in range(K + 1):
            if self.parametric:
                sb.colloc_sdot_row(n, k, sb.alloc_rows(1, 0, np.inf)[0])
            if k > 0:
                sb.colloc_defect_rows(n, k, sb.alloc_rows(nz, 0., 0.))
            sb.colloc_du_rows(n, k, sb.alloc_rows(nu, 0., 0.))
        # constraints
        for k in range(K + 1):
            self._regularity(n, k)
        for k in range(K + 1):
            self._stage_constraint(n, k)
        # continuity from the previous interval: those rows are produced by cell n-1
        if n >= 1:
            first = 1 if self.parametric else 0
            rows = sb.alloc_rows(nz - first, 0., 0.)
            for c, r in zip(range(first, nz), rows):
                sb.colloc_end_row(n - 1, c, r, -1.0, sb.iZ(n, 0, c), 1.0, 0.0)
            rows = sb.alloc_rows(nu, 0., 0.)
            for j, r in enumerate(rows):
                sb.colloc_end_row(n - 1, nz + j, r, -1.0, sb.iU(n, 0, j), 1.0, 0.0)
        if self.parametric:
            sb.add_affine([sb.iZ(n, 0, 0)], [1.0], -self._get_s(n, 0), 0., 0.)
            sb.add_affine([sb.iZ(n, k, 0) for k in range(K + 1)], self.D, -self._get_s(n + 1, 0), 0., 0.)

    def _end_row(self, c, row, coef, partner, pcoef, off):
        ''' a row built from the end state of the LAST interval (closure rows) '''
        n = self.config.N - 1
        if self.config.use_rk4:
            self._cell_out_row(n, c, row, coef, partner, pcoef, off)
        else:
            self.sb.colloc_end_row(n, c, row, coef, partner, pcoef, off)

    def _enforce_loop_closure(self):
        ''' point-mass closure: [uF - u0], [zF - z0] (base_raceline.py:492-514, :1183-1227) '''
        sb = self.sb
        nz, nu = self.model.nz, self.model.nu
        rows = sb.alloc_rows(nu, 0., 0.)
        for j, r in enumerate(rows):
            self._end_row(nz + j, r, 1.0, sb.iU(0, 0, j), -1.0, 0.0)
        if self.config.use_rk4:
            sb.cell_par[self.config.N - 1, 0] = 1.0      # uF = U + dU*H  
        if self.parametric and not self.line.cleanly_closed:
            # skewly closed centerline (base_raceline.py:1208-1227): the frame at s_max is rotated about the tangent
            # against the frame at s_min, so the lateral offsets close through a 2x2 rotation.  Those two rows have two
            # partners each and go through the expression rows (tail.py); the rest are ordinary end rows, sign flipped.
            line = self.line
            ey1, en1 = line.p2ey(line.s_min()), line.p2en(line.s_min())
            ey2, en2 = line.p2ey(line.s_max() - 0.001), line.p2en(line.s_max() - 0.001)
            A = np.array([[ey1 @ ey2, en1 @ ey2], [ey1 @ en2, en1 @ en2]], dtype=float)
            t = self._tail()
            zF, _ = self._zF_uF()
            z0, _, _ = t.point(0, 0)
            t.use()
            t.add_rows([(0 + z0[1] * float(A[i, 0]) + z0[2] * float(A[i, 1])) - zF[1 + i] for i in range(2)], 0., 0.)
            rows = sb.alloc_rows(nz - 3, 0., 0.)
            for c, r in zip(range(3, nz), rows):
                self._end_row(c, r, -1.0, sb.iZ(0, 0, c), 1.0, 0.0)
            return
        first = 1 if self.parametric else 0
        rows = sb.alloc_rows(nz - first, 0., 0.)
        for c, r in zip(range(first, nz), rows):
            self._end_row(c, r, 1.0, sb.iZ(0, 0, c), -1.0, 0.0)

    def _enforce_modified_loop_closure(self):
        ''' drone closure with quaternion sign / yaw wrap taken from the warm start (drone_raceline.py:47-104) '''
        sb = self.sb
        nz, nu = self.model.nz, self.model.nu
        quat = self.model.config.use_quat
        rows = sb.alloc_rows(nu, 0., 0.)
        for j, r in enumerate(rows):
            self._end_row(nz + j, r, 1.0, sb.iU(0, 0, j), -1.0, 0.0)
        if self.config.use_rk4:
            sb.cell_par[self.config.N - 1, 0] = 1.0
        plain = [1, 2] + list(range(7 if quat else 4, nz))
        rows = sb.alloc_rows(len(plain), 0., 0.)
        for c, r in zip(plain, rows):
            self._end_row(c, r, 1.0, sb.iZ(0, 0, c), -1.0, 0.0)
        ws_known = self._first_ws_r is not None and self._last_ws_r is not None
        if quat:
            flipped = ws_known and np.linalg.norm(self._first_ws_r - self._last_ws_r) > 1
            rows = sb.alloc_rows(4, 0., 0.)
            for c, r in zip(range(3, 7), rows):
                self._end_row(c, r, 1.0, sb.iZ(0, 0, c), 1.0 if flipped else -1.0, 0.0)
        else:
            wraps = np.round((self._last_ws_r - self._first_ws_r)[0] / 2 / np.pi) if ws_known else 0.0
            r = sb.alloc_rows(1, 0., 0.)[0]
            self._end_row(3, r, 1.0, sb.iZ(0, 0, 3), -1.0, -2 * np.pi * wraps if ws_known else 0.0)
        if not self.parametric:
            r = sb.alloc_rows(1, 0., 0.)[0]
            self._end_row(0, r, 1.0, sb.iZ(0, 0, 0), -1.0, 0.0)

    # ---- gates ------------------------------------------------------------------------------------
    def _fix_gate(self, vars_, X, Xs, xconst, s, include_axial_fix):
        '''
        gate rows for the affine position x = X @ w[vars] + xconst (base_raceline.py:545-595).
        Xs: structural mask of X (which entries the reference's expression keeps).
        '''
        sb = self.sb
        gate_x = self.line.gate_position(s)
        rc = self.model.config.collision_radius
        lcfg = self.line.config

        def form(direction):
            ''' coefficients, structural mask and constant of (x - gate_x) . direction '''
            struct = ((direction != 0)[:, None] & Xs).any(axis=0)
            return direction @ X, struct, float(np.sum((xconst - gate_x) * direction))

        def affine(direction, lb, ub):
            coefs, struct, const = form(direction)
            sb.add_affine(vars_, coefs, const, lb, ub, mask=struct)

        if self.config.fix_gate_center:
            for i in range(3):
                affine(np.eye(3)[i], 0, 0)
            return
        R = self.line.gate_orientation(s)
        if lcfg.gate_shape == GateShape.CIRCLE:
            forms = [form(R[:, 1]), form(R[:, 2])]
            sb.add_squares(vars_, np.array([f[0] for f in forms]), [f[2] for f in forms],
                           -np.inf, (lcfg.gate_ri - rc) ** 2, mask=np.array([f[1] for f in forms]))
            if include_axial_fix:
                affine(R[:, 0], 0., 0.)
        elif lcfg.gate_shape == GateShape.SQUARE:
            d_max = lcfg.gate_ri - rc
            if include_axial_fix:
                affine(R[:, 0], 0., 0.)
            affine(R[:, 1], -d_max, d_max)
            affine(R[:, 2], -d_max, d_max)
        else:
            raise NotImplementedError('Unhandled Gate Shape')

    def _add_gate_constraints(self):
        raise NotImplementedError

    # ---- decision vector, bounds, initial guess ------------------------------------------------------
    def _create_problem(self):
        cfg, sb = self.config, self.sb
        N, K = cfg.N, cfg.K
        w0, lbw, ubw = [], [], []
        for n in range(N):
            h0 = self._guess_h(n)
            ubw.append(h0 * 10)
            lbw.append(h0 / 100)
            w0.append(h0)
        for n in range(N):
            for k in range(K + 1):
                s = self._get_s(n, k) if self.parametric else 0
                lbw += [*self.model.zl(s), *self.model.ul(), *self.model.dul()]
                ubw += [*self.model.zu(s), *self.model.uu(), *self.model.duu()]
                w0 += [*self._guess_z(n, k), *self._guess_u(n, k), *self._guess_du(n, k)]
        if self.drone and cfg.closed:
            self._enforce_modified_loop_closure()       # appended last (drone_raceline.py:150-156)
        self.structure: NLPStructure = sb.finalize()
        st = self.structure
        st.w0 = np.array(w0, dtype=float)
        st.lbw = np.array(lbw, dtype=float)
        st.ubw = np.array(ubw, dtype=float)
        self.solver_w0, self.solver_lbw, self.solver_ubw = st.w0, st.lbw, st.ubw
        self.solver_lbg, self.solver_ubg = st.lbg, st.ubg

    def _guess_h(self, n):
        if self._ws_available():
            return self.ws_raceline.step_sizes[n]
        if self.config.h0:
            return self.config.h0
        if self.parametric:
            ds = (self.line.s_max() - self.line.s_min()) / self.config.N
            return ds / self.config.v0 * float(self.line.p2mag_xcs(ds * n))
        return 1

    def _guess_pos_vel(self, n, k):
        ''' position (or s, y, n) and linear velocity guess (base_raceline.py:920-937, :1240-1251) '''
        cfg = self.config
        z = [0.] * 6
        if not self.parametric:
            gate_no = n / self.gate_n_interval if cfg.use_rk4 \
                else (n + k / cfg.K) / self.gate_n_interval
            v = self.line.p2es(gate_no)
            z[:3] = self.line.p2xc(gate_no)
            z[3:6] = v / np.linalg.norm(v) * cfg.v0
        else:
            z[0] = self._get_s(n, k)
            if not self.model.config.global_r:
                z[3] = cfg.v0
            else:
                z[3:6] = cfg.v0 * self.line.p2es(self._get_s(n, k))
        return [float(e) for e in z]

    def _guess_z(self, n, k):
        return self._guess_pos_vel(n, k)

    def _guess_u(self, n, k):
        return [0.] * self.model.nu

    def _guess_du(self, n, k):
        return [0.] * self.model.nu

    # ---- solver + unpacking ------------------------------------------------------------------------
    def _create_solver(self):
        ''' where the reference calls ca.nlpsol('solver', 'ipopt', prob, opts) (base_raceline.py:752-799) '''
        from .solver import InteriorPointSolver
        self.solver = InteriorPointSolver(lambda: self.functions, max_iter=self.config.max_iter,
                                          verbose=self.config.verbose)

    @property
    def functions(self):
        ''' the five NLP functions on the GPU (created on first use; raises without a CUDA device) '''
        if getattr(self, '_functions', None) is None:
            from .functions import NlpFunctions
            self._functions = NlpFunctions(self.structure, self.vehicle_config)
        return self._functions

    def unpack_w(self, x):
        ''' (H, Z, U, dU) views of a decision vector '''
        cfg = self.config
        N, P = cfg.N, cfg.K + 1
        nz, nu = self.model.nz, self.model.nu
        x = np.asarray(x, dtype=float).ravel()
        body = x[N:].reshape(N, P, nz + 2 * nu)
        return x[:N], body[..., :nz], body[..., nz:nz + nu], body[..., nz + nu:]

    def _unpack_soln(self, sol) -> RacelineResults:
        ''' states at every collocation point and interpolants (base_raceline.py:801-864) '''
        from .interpolation import make_interpolants
        cfg = self.config
        H, Z, U, dU = self.unpack_w(sol['x'])
        N, K = cfg.N, cfg.K
        t0 = np.concatenate([[0.0], np.cumsum(H)[:-1]])
        tau = self.tau if not cfg.use_rk4 else np.zeros(1)
        T = t0[:, None] + tau[None, :] * H[:, None]
        states = []
        for n in range(N):
            for k in range(K + 1):
                state = self.model.get_empty_state()
                state.t = float(T[n, k])
                self.model.zu2state(state, Z[n, k], U[n, k])
                self.model.du2state(state, dU[n, k])
                states.append(state)
        z_i, u_i, du_i = make_interpolants(H, Z, U, dU, None if cfg.use_rk4 else (self.tau, self.D))
        feasible = bool(self.solver.stats().get('success', False)) if self.solver is not None else False
        return RacelineResults(
            states=states, step_sizes=H.copy(), time=float(np.sum(H)), periodic=cfg.closed,
            label=self.label, color=list(self.color), z_interp=z_i, u_interp=u_i, du_interp=du_i,
            solve_time=self.solve_time, feval_time=self.feval_time, ipopt_time=self.ipopt_time,
            feasible=feasible, global_frame=self.global_frame)


class BaseGlobalRaceline(BaseRaceline):
    ''' global-frame raceline (base_raceline.py:867-937) '''
    parametric = False
    global_frame = True

    def _add_gate_constraints(self):
        cfg, sb = self.config, self.sb
        for gate_no, n in enumerate(range(0, cfg.N, self.gate_n_interval)):
            vars_ = [sb.iZ(n, 0, i) for i in range(3)]
            self._fix_gate(vars_, np.eye(3), np.eye(3, dtype=bool), np.zeros(3), gate_no, True)
        if not cfg.closed:
            # the gate at the very end of an open track sits on the end state (base_raceline.py:914-918)
            zF, _ = self._zF_uF()
            self._fix_gate_expr(zF[:3], len(cfg.gate_xi) - 1, True)


class BaseParametricRaceline(BaseRaceline):
    ''' curvilinear-frame raceline (base_raceline.py:940-1251) '''
    parametric = True
    global_frame = False

    def _add_gate_constraints(self):
        ''' gates at fixed path lengths, interpolated inside their interval (base_raceline.py:986-1032) '''
        cfg, line, sb = self.config, self.line, self.sb
        fixed_gates = cfg.fixed_gates
        if fixed_gates is None:
            if line.config.gate_s is None:
                return
            fixed_gates = line.config.gate_s
            if line.s_min() in fixed_gates and line.config.closed and cfg.closed:
                fixed_gates = np.array([k for k in fixed_gates if k != line.s_max()])
        for s in fixed_gates:
            s0 = self._get_s(0, 0)
            if s < s0:
                raise TypeError('Gate is before start')
            n = 0
            while not self._get_s(n + 1, 0) > s:
                n += 1
                s0 = self._get_s(n, 0)
                if n == cfg.N and s > s0 + 0.1:
                    raise TypeError('Gate is after end')
            if n == cfg.N:
                raise NotImplementedError('gate at the very end of an open track')
            sf = self._get_s(n + 1, 0)
            d = (s - s0) / (sf - s0)
            if cfg.use_rk4:
                # z_gate = Z[n] + d (Z[n+1] - Z[n]);  d == 0 folds the second term away
                pts = [(n, 0, 1.0 - d, True)] + ([(n + 1, 0, d, True)] if d != 0 else [])
            else:
                Dg = get_intermediate_collocation_coefficients(cfg.K, d)
                pts = [(n, k, Dg[k], Dg[k] != 0) for k in range(cfg.K + 1)]
            ey, en = line.p2ey(s), line.p2en(s)
            vars_, cols, struct = [], [], []
            for (nn, kk, wgt, present) in pts:
                if not present:
                    continue
                for comp, direction in ((1, ey), (2, en)):
                    vars_.append(sb.iZ(nn, kk, comp))
                    cols.append(wgt * direction)
                    struct.append(direction != 0)
            self._fix_gate(vars_, np.array(cols).T, np.array(struct).T, line.p2xc(s), s, False)


class BaseParametricObstacleRaceline(BaseParametricRaceline):
    ''' parametric raceline inside an obstacle-free tube (base_raceline.py:1254-1331) '''
    tube: ObstacleFreeTube = None
    calc_tube_time: float = -1

    def __init__(self, line, config, vehicle_config, mesh_obstacle=None, tube=None,
                 ws_raceline=None, ws_model=None):
        self.mesh_obstacle = mesh_obstacle
        self.tube = tube
        super().__init__(line, config, vehicle_config, ws_raceline, ws_model)

    def _add_gate_constraints(self):
        super()._add_gate_constraints()
        self._add_mesh_constraints()

    def _add_mesh_constraints(self):
        cfg, sb = self.config, self.sb
        if self.tube is None:
            if self.mesh_obstacle is None:
                raise ValueError('either a tube or a mesh obstacle is needed')
            s = np.array([self._get_s(n, k) for n in range(cfg.N) for k in range(cfg.K + 1)])
            t0 = time.time()
            self.tube = self.mesh_obstacle.compute_plannning_tube(
                self.line, s, self.model.config.collision_radius)
            self.calc_tube_time = time.time() - t0
        for n in range(cfg.N):
            for k in range(cfg.K + 1):
                dy, dn, avail = self.tube.disc(self._get_s(n, k))
                sb.add_squares([sb.iZ(n, k, 1), sb.iZ(n, k, 2)], np.eye(2), [-dy, -dn],
                               -np.inf, avail ** 2)


# ---------------------------------------------------------------------------------------------------
class PointRaceline(BaseRaceline):
    ''' point-mass mixin (point_raceline.py:12-45) '''
    drone = False


class DroneRaceline(BaseRaceline):
    ''' drone mixin: warm-start generation and orientation guesses (drone_raceline.py:24-277) '''
    drone = True
    ws_solver: BaseRaceline = None

    def _gemerate_ws(self, solver_class, solver_args):
        print('Generating Warmstart... ')
        t0 = time.time()
        self.ws_solver = solver_class(*solver_args)
        ws_raceline = self.ws_solver.solve()
        print(f'Done (Lap Time: {ws_raceline.time:0.2f}s) (Setup + Solve: {time.time() - t0:0.2f}s)')
        return ws_raceline, self.ws_solver

    def _ws_point(self, n, k):
        ''' warm-start state, input, input rate at collocation point (n, k) '''
        idx = n * (self.config.K + 1) + k
        t = self.ws_raceline.states[idx].t
        return self.ws_raceline.z_interp(t), self.ws_raceline.u_interp(t), self.ws_raceline.du_interp(t)

    def _guess_Rp(self, n, k):
        return self.line.p2Rp(self._get_s(n, k))

    def _guess_z(self, n, k):
        z6 = self._guess_pos_vel(n, k)
        quat = self.model.config.use_quat
        z = [*z6[:3], *([1, 0, 0, 0] if quat else [0, 0, 0]), *z6[3:6], 0, 0, 0]
        if self.parametric:
            z[0] = self._get_s(n, k)
        if not self._ws_available():
            return [float(e) for e in z]
        z_ws, u_ws, du_ws = self._ws_point(n, k)
        z[0:3] = z_ws[0:3]
        T = self.ws_model.f_T(z_ws, u_ws)
        vg = self.ws_model.f_vg(z_ws, u_ws)
        if self.config.closed:
            e1 = vg / np.linalg.norm(vg)
            e3 = T / np.linalg.norm(T)
            e1 = e1 - e3 * (e1 @ e3)
            e1 = e1 / np.linalg.norm(e1)
            R = np.array([e1, np.cross(e3, e1), e3]).T
        else:
            hat = lambda v: np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
            b = np.array([0, 0, 1])
            tn = T / np.linalg.norm(T)
            v = -np.cross(tn, b)
            s2 = float(v @ v)
            # exactly vertical thrust (the end rows of an open point-mass raceline pin T_x = T_y = 0): 0 / 0 in the
            # reference's expression (drone_raceline.py:201-204); its limit (1 - cos) / sin^2 -> 1/2 gives R = I
            R = np.eye(3) + hat(v) + hat(v) @ hat(v) * ((1 - tn @ b) / s2 if s2 > 0 else 0.5)
        if not self.model.config.global_r:
            R = self._guess_Rp(n, k).T @ R
        if quat:
            r = matrix_to_quat(R)
            if self._last_ws_r is not None and np.linalg.norm(r - self._last_ws_r) >= 1:
                r = -r
            z[3:7] = r
        else:
            from scipy.spatial.transform import Rotation
            r = np.flip(Rotation.from_matrix(R).as_euler('xyz', degrees=False))
            if self._last_ws_r is not None and np.linalg.norm(r - self._last_ws_r) > 1:
                if r[0] - self._last_ws_r[0] > np.pi:
                    r[0] -= 2 * np.pi
                elif r[0] - self._last_ws_r[0] <= -np.pi:
                    r[0] += 2 * np.pi
                if np.linalg.norm(r - self._last_ws_r) > 1:
                    raise NotImplementedError('Warmstart continuity failed for euler angles, try quaternion')
            z[3:6] = r
        if self._first_ws_r is None:
            self._first_ws_r = r
        self._last_ws_r = r
        z[-6:-3] = R.T @ vg
        dT = self.ws_model.f_T(z_ws, du_ws)
        z[-3:] = R.T @ np.cross(T, dT) / np.linalg.norm(T) ** 2
        return [float(e) for e in z]

    def _guess_u(self, n, k):
        if self._ws_available():
            z_ws, u_ws, _ = self._ws_point(n, k)
            return [float(np.linalg.norm(self.ws_model.f_T(z_ws, u_ws)) / 4)] * 4
        return super()._guess_u(n, k)

    # ---- batched warm-start chain on the device --------------------------------
    def guess_batch(self, w_pm):
        '''
        drone initial guesses of B point-mass solutions (rows of w_pm, decision-vector layout of the warm-start
        solver) -- the batched form of _guess_z / _guess_u (drone_raceline.py:158-277), csrc/warm_start.cuh.
        Returns (w0 (B, nw) CUDA tensor clipped to the variable bounds like the NLP start point, info (B, 4)).
        '''
        import torch
        from .warm_start import drone_guess_batch
        cfg = self.config
        fc = None
        if self.parametric and not self.model.config.global_r:
            s_all = np.array([self._get_s(n, k) for n in range(cfg.N) for k in range(cfg.K + 1)])
            fc = self.line.frame_constants(s_all)
        w0, info = drone_guess_batch(w_pm, cfg.N, cfg.K, quat=self.model.config.use_quat, closed=cfg.closed,
                                     global_r=self.model.config.global_r, fc=fc)
        return w0, info

    def solve_batch(self, vp_drone, vp_point=None):
        '''
        the scripts' chain (point-mass solve -> drone guess -> drone solve, drone_raceline.py:294-312) for B vehicle
        variants at once, every instance warm-started from its OWN point-mass solution.  vp_drone: (B, 13) vehicle
        parameters (models.VP_DRONE order); vp_point: (B, 6) (models.VP_POINT), default: the warm-start solver's
        parameters with the mass of every drone variant.  Instances whose warm start closes the lap with the other
        quaternion sign / yaw wrap than the one the closure rows were built for (info[:, 0] / info[:, 1]) are reported
        in `closure_mismatch`; they need a structure of their own.
        '''
        from .models import vehicle_params, VP_POINT, VP_DRONE
        if self.ws_solver is None:
            raise RuntimeError('solve_batch needs the warm-start solver (generate_ws=True)')
        vp_drone = np.atleast_2d(np.asarray(vp_drone, dtype=float))
        B = vp_drone.shape[0]
        ws = self.ws_solver
        if vp_point is None:
            vp_point = np.tile(vehicle_params(ws.vehicle_config), (B, 1))
            vp_point[:, VP_POINT.index('m')] = vp_drone[:, VP_DRONE.index('m')]
        wst = ws.structure
        sol_pm = ws.solver(x0=np.tile(wst.w0, (B, 1)), lbx=wst.lbw, ubx=wst.ubw, lbg=wst.lbg, ubg=wst.ubg, p=vp_point)
        pm_ok = np.asarray(ws.solver.stats()['success_each'])
        w0, info = self.guess_batch(sol_pm['x'])
        info = info.cpu().numpy()
        st = self.structure
        X0 = np.clip(w0.cpu().numpy(), st.lbw, st.ubw)
        sol = self.solver(x0=X0, lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg, p=vp_drone)
        quat = self.model.config.use_quat
        ws_known = self._first_ws_r is not None and self._last_ws_r is not None
        if quat:
            built_flip = bool(ws_known and np.linalg.norm(self._first_ws_r - self._last_ws_r) > 1)
            mismatch = info[:, 0].astype(bool) != built_flip
        else:
            built_wraps = int(np.round((self._last_ws_r - self._first_ws_r)[0] / 2 / np.pi)) if ws_known else 0
            mismatch = info[:, 1] != built_wraps
        sol['lap_time'] = sol['x'][:, :self.config.N].sum(1)
        sol['lap_time_ws'] = sol_pm['x'][:, :ws.config.N].sum(1)
        sol['ws_success'] = pm_ok
        sol['success'] = np.asarray(self.solver.stats()['success_each'])
        sol['closure_mismatch'] = mismatch | (info[:, 2] > 0)
        sol['x0'] = X0
        return sol


class GlobalPointRaceline(PointRaceline, BaseGlobalRaceline):
    label, color = 'Global PM', [1, 1, 1, 1]

    def _get_model(self, config: PointConfig):
        return PointModel(config)


class ParametricPointRaceline(PointRaceline, BaseParametricRaceline):
    label, color = 'Parametric PM', [.5, .5, .5, 1]

    def _get_model(self, config: PointConfig):
        return ParametricPointModel(config, self.line)


class ParametricObstaclePointRaceline(PointRaceline, BaseParametricObstacleRaceline):
    label, color = 'PM w/ obstacles', [.3, .3, .3, 1]

    def _get_model(self, config: PointConfig):
        return ParametricPointModel(config, self.line)


def _ws_config(config):
    ws = config.copy()
    ws.verbose = False
    ws.plot_iterations = False
    return ws


class GlobalDroneRaceline(DroneRaceline, BaseGlobalRaceline):
    ''' drone_raceline.py:280-322 '''
    label, color = 'Global Drone', [0, 1, 0, 1]

    def __init__(self, line, config: GlobalRacelineConfig, vehicle_config: DroneConfig,
                 ws_raceline=None, ws_model=None, generate_ws: bool = True):
        if generate_ws:
            point_config = PointConfig(global_r=vehicle_config.global_r,
                                       collision_radius=vehicle_config.collision_radius)
            ws_raceline, ws_solver = self._gemerate_ws(GlobalPointRaceline,
                                                       (line, _ws_config(config), point_config))
            ws_model = ws_solver.model
        super().__init__(line, config, vehicle_config, ws_raceline, ws_model)

    def _get_model(self, config: DroneConfig):
        config.global_r = True
        return DroneModel(config)


class ParametricDroneRaceline(DroneRaceline, BaseParametricRaceline):
    ''' drone_raceline.py:325-368 '''
    label, color = 'Parametric Drone', [1, 0, 0, 1]

    def __init__(self, line, config: ParametricRacelineConfig, vehicle_config: DroneConfig,
                 ws_raceline=None, ws_model=None, generate_ws: bool = True):
        if generate_ws:
            point_config = PointConfig(global_r=vehicle_config.global_r)
            ws_raceline, ws_solver = self._gemerate_ws(ParametricPointRaceline,
                                                       (line, _ws_config(config), point_config))
            ws_model = ws_solver.model
        super().__init__(line, config, vehicle_config, ws_raceline, ws_model)

    def _get_model(self, config: DroneConfig):
        return ParametricDroneModel(config, self.line)


class ParametricObstacleDroneRaceline(DroneRaceline, BaseParametricObstacleRaceline):
    ''' drone_raceline.py:371-427 '''
    label, color = 'Drone w/ obstacles', [0, .3, 1, 1]

    def __init__(self, line, config: ParametricRacelineConfig, vehicle_config: DroneConfig,
                 mesh_obstacle=None, tube: ObstacleFreeTube = None, ws_raceline=None, ws_model=None,
                 generate_ws: bool = True):
        if generate_ws:
            point_config = PointConfig(global_r=vehicle_config.global_r,
                                       collision_radius=vehicle_config.collision_radius)
            ws_raceline, ws_solver = self._gemerate_ws(
                ParametricObstaclePointRaceline,
                (line, _ws_config(config), point_config, mesh_obstacle, tube))
            ws_model = ws_solver.model
            tube = ws_solver.tube
        BaseParametricObstacleRaceline.__init__(self, line, config, vehicle_config, mesh_obstacle,
                                                tube, ws_raceline, ws_model)

    def _get_model(self, config: DroneConfig):
        return ParametricDroneModel(config, self.line)
