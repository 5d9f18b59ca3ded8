'''
Oracle restatement of the reference's raceline NLP construction.  TEST INFRASTRUCTURE ONLY.

Follows drone3d/raceline/base_raceline.py (:226-239 build order, :279-320 variables,
:363-391 / :1052-1112 RK4 intervals, :398-490 / :1114-1181 collocation intervals, :492-543 closure /
end constraints, :545-595 gates, :601-623 cost, :670-750 decision vector, :867-937 global frame,
:940-1251 parametric frame, :1293-1310 tube rows), drone3d/raceline/drone_raceline.py
(:42-45 quaternion renormalisation, :47-104 modified closure, :110-148 end constraints,
:150-156 closure appended last, :158-274 initial guess) and drone3d/raceline/point_raceline.py
(:15-45), plus drone3d/obstacles/mesh_obstacle.py:219-237 for the tube disc rows.

One flag-driven class instead of the reference's mixin hierarchy; the *order* in which rows are
appended to g is the reference's, which is what fixes every index of
jac_g / hess_l.
'''
from dataclasses import dataclass

import numpy as np
from scipy.spatial.transform import Rotation as _SciRot

from aircraft_trajectory_optimization_b200 import symbolic as sx
from aircraft_trajectory_optimization_b200.centerlines import GateShape
from .ref_discretization import get_collocation_coefficients, \
    get_intermediate_collocation_coefficients
from .ref_models import RefDroneModel, RefPointModel, _vec
from .ref_centerline import RefSplineCenterline


@dataclass
class RefTube:
    ''' what ObstacleFreeTube keeps (mesh_obstacle.py:202-217): samples (s, dy, dn) and radii '''
    ball_p: np.ndarray       # (n, 3) columns s, delta_y, delta_n
    ball_r: np.ndarray       # (n,)
    collision_r: float

    def row(self, s, z):
        # mesh_obstacle.py:219-237 (nearest sample in s; KD-tree on a 1-D set == argmin |ds|)
        idx = int(np.argmin(np.abs(self.ball_p[:, 0] - s)))
        dy, dn = self.ball_p[idx, 1], self.ball_p[idx, 2]
        avail = np.maximum(self.ball_r[idx] - self.collision_r, 0.01)
        return (z[1] - dy) ** 2 + (z[2] - dn) ** 2, avail ** 2


class RefWarmstart:
    ''' the pieces of a solved RacelineResults + model that _guess_* read (drone_raceline.py:158-274) '''

    def __init__(self, model, H, Z, U, dU):
        self.model = model
        self.step_sizes = np.asarray(H, dtype=float)
        self.Z, self.U, self.dU = (np.asarray(a, dtype=float) for a in (Z, U, dU))


class RefRaceline:
    '''
    frame:   'global' | 'parametric'
    vehicle: 'drone' | 'point'
    config:  a (Global|Parametric)RacelineConfig-like object (fields N, K, use_rk4, R, dR, h0, v0,
             closed, fix_gate_center, [gate_xi/xj/xk] or [fixed_gates, force_regularity]);
             it is mutated the same way the reference mutates it
    '''

    def __init__(self, line: RefSplineCenterline, config, vehicle_config, frame, vehicle,
                 tube: RefTube = None, ws: RefWarmstart = None, graph=None):
        self.line, self.config, self.vehicle_config = line, config, vehicle_config
        self.frame, self.vehicle, self.tube, self.ws = frame, vehicle, tube, ws
        self.parametric = frame == 'parametric'
        self.drone = vehicle == 'drone'
        self._first_ws_r = None
        self._last_ws_r = None
        self.graph = graph if graph is not None else sx.new_graph()
        sx.set_graph(self.graph)

        # _get_model (drone_raceline.py:314-316, :353-357 ; point_raceline.py:51-52, :64-68)
        if self.drone and not self.parametric:
            vehicle_config.global_r = True
        cls = RefDroneModel if self.drone else RefPointModel
        self.model = cls(vehicle_config, line if self.parametric else None)
        self._setup_checks()
        self._create_nlp()

    # ---------------------------------------------------------------------------------------
    def _setup_checks(self):
        cfg = self.config
        if self.parametric:
            # base_raceline.py:946-954
            if not self.line.cleanly_closed and not self.model.config.global_r:
                raise NotImplementedError('Global orientation must be used for skewly closed centerlines')
        # base_raceline.py:226-230
        if cfg.use_rk4:
            cfg.h0 /= cfg.K
            cfg.N *= cfg.K
            cfg.K = 0
        if not self.parametric:
            # base_raceline.py:873-885
            x = np.array([cfg.gate_xi, cfg.gate_xj, cfg.gate_xk])
            if cfg.closed and not (x[:, 0] == x[:, -1]).all():
                x = np.hstack([x, x[:, 0:1]])
            num_phases = x.shape[1] - 1
            cfg.N = int(num_phases * np.ceil(cfg.N / num_phases))
            self.gate_n_interval = int(cfg.N / num_phases)

    def _create_nlp(self):
        # base_raceline.py:232-270
        self.g, self.ubg, self.lbg = [], [], []
        self.J = sx.SX.const(0)
        self.nz, self.nu = self.model.nz, self.model.nu
        cfg = self.config
        if isinstance(cfg.R, (float, int)):
            cfg.R = np.eye(self.nu) * cfg.R
        if isinstance(cfg.dR, (float, int)):
            cfg.dR = np.eye(self.nu) * cfg.dR
        self._create_nlp_vars()
        self._enforce_model()
        self._add_gate_constraints()
        self._add_costs()
        self._create_problem()

    def _create_nlp_vars(self):
        # base_raceline.py:279-320 ; symbols are created in decision-vector order
        N, K = self.config.N, self.config.K
        if not self.config.use_rk4:
            self.tau, self.B, self.C, self.D = get_collocation_coefficients(K)
        self.H = np.array([sx.SX.sym(f'h_{n}') for n in range(N)], dtype=object)
        self.Z = np.empty((N, K + 1), dtype=object)
        self.U = np.empty((N, K + 1), dtype=object)
        self.dU = np.empty((N, K + 1), dtype=object)
        for n in range(N):
            for k in range(K + 1):
                self.Z[n, k] = sx.SX.sym(f'z_{n}_{k}', self.nz)
                self.U[n, k] = sx.SX.sym(f'u_{n}_{k}', self.nu)
                self.dU[n, k] = sx.SX.sym(f'du_{n}_{k}', self.nu)

    # ---- helpers ----------------------------------------------------------------------------
    def _row(self, expr, lb, ub):
        e = np.atleast_1d(np.asarray(expr, dtype=object))
        self.g.append(e)
        n = len(e)
        self.lbg += list(lb) if np.ndim(lb) else [lb] * n
        self.ubg += list(ub) if np.ndim(ub) else [ub] * n

    def _get_s(self, n, k):
        # base_raceline.py:972-984
        ds = (self.line.s_max() - self.line.s_min()) / self.config.N
        if self.config.use_rk4:
            return self.line.s_min() + ds * n
        return self.line.s_min() + ds * (n + self.tau[k])

    def _eval_ode(self, n, k=0):
        # base_raceline.py:272-277 / :963-970
        if self.parametric:
            pt = self.line.f_param_terms(self._get_s(n, k))
            return self.model.zdot(self.Z[n, k], self.U[n, k], pt)
        return self.model.zdot(self.Z[n, k], self.U[n, k])

    def _cont(self, z):
        # drone_raceline.py:42-45 (in place on a copy; the reference mutates a fresh expression)
        z = np.array(z, dtype=object)
        if self.drone and self.model.config.use_quat:
            z[3:7] = z[3:7] / sx.norm_2(z[3:7])
        return z

    def _rk4_step(self, n):
        h = self.H[n]
        z, u = self.Z[n, 0], self.U[n, 0]
        pt = self.line.f_param_terms(self._get_s(n, 0)) if self.parametric else None
        return self.model.rk4(z, u, h, pt)

    def _zF(self):
        # base_raceline.py:322-336 / :1034-1050
        if self.config.use_rk4:
            return self._cont(self._rk4_step(self.config.N - 1))
        zF = 0
        for k in range(self.config.K + 1):
            zF = zF + self.Z[-1, k] * self.D[k]
        return self._cont(zF)

    def _uF(self):
        # base_raceline.py:338-348
        if self.config.use_rk4:
            return self.U[-1, 0] + self.dU[-1, 0] * self.H[-1]
        uF = 0
        for k in range(self.config.K + 1):
            uF = uF + self.U[-1, k] * self.D[k]
        return uF

    # ---- model rows -------------------------------------------------------------------------
    def _enforce_model(self):
        cfg = self.config
        if not self.parametric:
            # base_raceline.py:891-905: equal step sizes inside a gate-to-gate phase
            for n in range(0, cfg.N, self.gate_n_interval):
                for n2 in range(n + 1, n + self.gate_n_interval):
                    self._row(self.H[n2] - self.H[n], 0., 0.)
        # base_raceline.py:350-361
        for n in range(cfg.N):
            if not cfg.use_rk4:
                self._collocation_ode(n)
                self._collocation_constraints(n)
                self._collocation_continuity(n)
            elif self.parametric:
                self._rk4_interval_parametric(n)
            else:
                self._rk4_interval_global(n)
        if cfg.closed:
            if not self.drone:
                self._loop_closure_point()
            # drone: deferred (drone_raceline.py:106-108)
        else:
            self._initial_constraints()
            self._terminal_constraints()

    def _rk4_interval_global(self, n):
        # base_raceline.py:363-391
        if n == self.config.N - 1:
            return
        h = self.H[n]
        zn = self._cont(self._rk4_step(n))
        un = self.U[n, 0] + self.dU[n, 0] * h / 2
        self._row(self.Z[n + 1, 0] - zn, 0., 0.)
        self._row(self.U[n + 1, 0] - un, 0., 0.)
        self.model.add_model_stage_constraints(self.Z[n, 0], self.U[n, 0], self.g, self.lbg, self.ubg)

    def _rk4_interval_parametric(self, n):
        # base_raceline.py:1052-1112
        self._row(self.Z[n, 0][0] - self._get_s(n, 0), 0., 0.)
        if n == self.config.N - 1:
            return
        h = self.H[n]
        zn = self._cont(self._rk4_step(n))
        un = self.U[n, 0] + self.dU[n, 0] * h / 2
        self._row(self.Z[n + 1, 0][1:] - zn[1:], 0., 0.)
        self._row(self.U[n + 1, 0] - un, 0., 0.)
        self._row(zn[0] - self._get_s(n + 1, 0), 0., 0.)
        self.model.add_model_stage_constraints(self.Z[n, 0], self.U[n, 0], self.g, self.lbg, self.ubg)
        if self.config.force_regularity:
            ky = self.line.p2ky(self._get_s(n, 0))
            kn = self.line.p2kn(self._get_s(n, 0))
            if ky ** 2 + kn ** 2 > 0.1:
                self._row(kn * self.Z[n, 0][1] - ky * self.Z[n, 0][2], -np.inf, self.line.config.gamma)

    def _collocation_ode(self, n):
        # base_raceline.py:398-434
        K, C, H = self.config.K, self.C, self.H
        for k in range(K + 1):
            poly_ode = 0
            poly_du = 0
            for k2 in range(K + 1):
                poly_ode = poly_ode + C[k2][k] * self.Z[n, k2] / H[n]
                poly_du = poly_du + C[k2][k] * self.U[n, k2] / H[n]
            func_ode = self._eval_ode(n, k)
            if self.parametric:
                self._row(poly_ode[0], 0, np.inf)
            if k > 0:
                self._row(func_ode - poly_ode, 0., 0.)
            self._row(self.dU[n, k] - poly_du, 0., 0.)

    def _collocation_constraints(self, n):
        K = self.config.K
        if self.parametric and self.config.force_regularity:
            # base_raceline.py:1121-1129
            for k in range(K + 1):
                ky = self.line.p2ky(self._get_s(n, k))
                kn = self.line.p2kn(self._get_s(n, k))
                if ky ** 2 + kn ** 2 > 0.1:
                    self._row(kn * self.Z[n, k][1] - ky * self.Z[n, k][2], -np.inf,
                              self.line.config.gamma)
        # base_raceline.py:445-451
        for k in range(K + 1):
            self.model.add_model_stage_constraints(self.Z[n, k], self.U[n, k],
                                                   self.g, self.lbg, self.ubg)

    def _collocation_continuity(self, n):
        K, D = self.config.K, self.D
        if n >= 1:
            prev_z, prev_u = 0, 0
            for k in range(K + 1):
                prev_z = prev_z + self.Z[n - 1, k] * D[k]
                prev_u = prev_u + self.U[n - 1, k] * D[k]
            prev_z = self._cont(prev_z)
            if self.parametric:
                # base_raceline.py:1156-1163
                self._row(self.Z[n, 0][1:] - prev_z[1:], 0., 0.)
            else:
                # base_raceline.py:484-486
                self._row(self.Z[n, 0] - prev_z, 0., 0.)
            self._row(self.U[n, 0] - prev_u, 0., 0.)
        if self.parametric:
            # base_raceline.py:1165-1181: pin s at both ends of the interval
            zN = 0
            for k in range(K + 1):
                zN = zN + self.Z[n, k] * D[k]
            self._row(self.Z[n, 0][0] - self._get_s(n, 0), 0., 0.)
            self._row(zN[0] - self._get_s(n + 1, 0), 0., 0.)

    def _loop_closure_point(self):
        z0, u0 = self.Z[0, 0], self.U[0, 0]
        zF = self._cont(self._zF())
        uF = self._uF()
        self._row(uF - u0, 0., 0.)
        if not self.parametric:
            # base_raceline.py:509-514
            self._row(zF - z0, 0., 0.)
        elif self.line.cleanly_closed:
            # base_raceline.py:1203-1206
            self._row(zF[1:] - z0[1:], 0., 0.)
        else:
            # base_raceline.py:1208-1227
            ey1 = self.line.p2ey(self.line.s_min())
            en1 = self.line.p2en(self.line.s_min())
            ey2 = self.line.p2ey(self.line.s_max() - 0.001)
            en2 = self.line.p2en(self.line.s_max() - 0.001)
            A = np.array([[ey1 @ ey2, en1 @ ey2], [ey1 @ en2, en1 @ en2]])
            self._row(A @ z0[1:3] - zF[1:3], 0., 0.)
            self._row(z0[3:] - zF[3:], 0., 0.)

    def _vg(self, z, u):
        if self.parametric:
            raise NotImplementedError('oracle: open parametric tracks need f(s) inside the graph')
        return self.model.terms(z, u)['vg']

    def _end_rows(self, z, u):
        # base_raceline.py:516-543 + drone_raceline.py:110-148 / point_raceline.py:15-45
        t = self.model.terms(z, u)
        vg = t['vg']
        self._row(vg @ vg, -np.inf, 0.)
        if self.drone:
            self._row(t['R'][:, 2], [0, 0, 1], [0, 0, 1])
            self._row(z[-3:], 0., 0.)
        else:
            T = t['Tg']
            self._row(_vec(T[0], T[1]), 0., 0.)

    def _initial_constraints(self):
        if self.parametric:
            raise NotImplementedError('oracle: open parametric tracks not restated')
        self._end_rows(self.Z[0, 0], self.U[0, 0])

    def _terminal_constraints(self):
        self._end_rows(self._zF(), self._uF())

    def _modified_loop_closure(self):
        # drone_raceline.py:47-104
        z0, u0 = self.Z[0, 0], self.U[0, 0]
        zF, uF = self._zF(), self._uF()
        dz = zF - z0
        quat = self.model.config.use_quat
        rows = [uF - u0, dz[1:3], dz[7:] if quat else dz[4:]]
        ws_known = self._first_ws_r is not None and self._last_ws_r is not None
        if quat:
            if ws_known and np.linalg.norm(self._first_ws_r - self._last_ws_r) > 1:
                rows.append(zF[3:7] + z0[3:7])
            else:
                rows.append(zF[3:7] - z0[3:7])
        else:
            if ws_known:
                wraps = np.round((self._last_ws_r - self._first_ws_r)[0] / 2 / np.pi)
                rows.append(_vec(dz[3] - 2 * np.pi * wraps))
            else:
                rows.append(_vec(dz[3]))
        if not self.parametric:
            rows.append(_vec(dz[0]))
        for r in rows:
            self._row(r, 0., 0.)

    # ---- gates ------------------------------------------------------------------------------
    def _fix_gate(self, x_var, s, include_axial_fix):
        # base_raceline.py:545-595
        gate_x = self.line.gate_position(s)
        cfg = self.config
        if cfg.fix_gate_center:
            self._row(x_var - gate_x, 0, 0)
            return
        R = self.line.gate_orientation(s)
        rc = self.model.config.collision_radius
        shape = self.line.config.gate_shape
        if shape == GateShape.CIRCLE:
            e1, e2, e3 = R[:, 0], R[:, 1], R[:, 2]
            r_sq = ((x_var - gate_x) @ e2) ** 2 + ((x_var - gate_x) @ e3) ** 2
            self._row(r_sq, -np.inf, (self.line.config.gate_ri - rc) ** 2)
            if include_axial_fix:
                self._row(x_var @ e1 - gate_x @ e1, 0., 0.)
        elif shape == GateShape.SQUARE:
            d = x_var - gate_x
            delta = np.array([R[:, i] @ d for i in range(3)], dtype=object)
            d_max = self.line.config.gate_ri - rc
            if include_axial_fix:
                self._row(delta, [0., -d_max, -d_max], [0., d_max, d_max])
            else:
                self._row(delta[1:], [-d_max, -d_max], [d_max, d_max])
        else:
            raise NotImplementedError('Unhandled Gate Shape')

    def _add_gate_constraints(self):
        cfg = self.config
        if not self.parametric:
            # base_raceline.py:907-918
            for gate_no, n in enumerate(range(0, cfg.N, self.gate_n_interval)):
                self._fix_gate(self.Z[n, 0][:3], gate_no, True)
            if not cfg.closed:
                self._fix_gate(self._zF()[:3], len(cfg.gate_xi) - 1, True)
        else:
            self._add_gate_constraints_parametric()
        if self.tube is not None:
            # base_raceline.py:1303-1310
            for n in range(cfg.N):
                for k in range(cfg.K + 1):
                    r_sq, ub = self.tube.row(self._get_s(n, k), self.Z[n, k])
                    self._row(r_sq, -np.inf, ub)

    def _add_gate_constraints_parametric(self):
        # base_raceline.py:986-1032
        cfg, line = self.config, self.line
        fixed_gates = cfg.fixed_gates
        if fixed_gates is None:
            if line.config.gate_s is not None:
                fixed_gates = line.config.gate_s
                if line.s_min() in fixed_gates:
                    if line.config.closed and cfg.closed:
                        fixed_gates = np.array([k for k in fixed_gates if k != line.s_max()])
            else:
                return
        for s in fixed_gates:
            s0 = self._get_s(0, 0)
            if s < s0:
                raise TypeError('Gate is before start')
            n = 0
            while not self._get_s(n + 1, 0) > s:
                n += 1
                s0 = self._get_s(n, 0)
                if n == cfg.N:
                    if s > s0 + 0.1:
                        raise TypeError('Gate is after end')
            if n == cfg.N:
                z_gate = self._zF()
            else:
                sf = self._get_s(n + 1, 0)
                d = (s - s0) / (sf - s0)
                if cfg.use_rk4:
                    z_gate = self.Z[n, 0] + d * (self.Z[n + 1, 0] - self.Z[n, 0])
                else:
                    Dg = get_intermediate_collocation_coefficients(cfg.K, d)
                    z_gate = 0
                    for k in range(cfg.K + 1):
                        z_gate = z_gate + self.Z[n, k] * Dg[k]
            x_gate = line.p2xc(s) + z_gate[1] * line.p2ey(s) + z_gate[2] * line.p2en(s)
            self._fix_gate(x_gate, s, False)

    # ---- cost -------------------------------------------------------------------------------
    def _add_costs(self):
        # base_raceline.py:601-623
        cfg = self.config
        for n in range(cfg.N):
            for k in range(cfg.K + 1):
                u, du = self.U[n, k], self.dU[n, k]
                stage = u @ (cfg.R @ u) + du @ (cfg.dR @ du) + 1
                if cfg.use_rk4:
                    self.J = self.J + stage * self.H[n]
                else:
                    self.J = self.J + stage * self.H[n] * self.B[k]

    # ---- decision vector ----------------------------------------------------------------------
    def _create_problem(self):
        # base_raceline.py:625-646, :670-717 (+ drone_raceline.py:150-156)
        cfg = self.config
        w, w0, ubw, lbw = [], [], [], []
        for n in range(cfg.N):
            w.append(self.H[n])
            h0 = self._guess_h(n)
            ubw.append(h0 * 10)
            lbw.append(h0 / 100)
            w0.append(h0)
        for n in range(cfg.N):
            for k in range(cfg.K + 1):
                s = self._get_s(n, k) if self.parametric else 0
                w += [*self.Z[n, k], *self.U[n, k], *self.dU[n, k]]
                lbw += [*self.model.zl(s), *self.model.ul(), *self.model.dul()]
                ubw += [*self.model.zu(s), *self.model.uu(), *self.model.duu()]
                w0 += [*self._guess_z(n, k), *self._guess_u(n, k), *([0.] * self.nu)]
        if self.drone and cfg.closed:
            self._modified_loop_closure()
        self.w = w
        self.w0 = np.array(w0, dtype=float)
        self.lbw = np.array(lbw, dtype=float)
        self.ubw = np.array(ubw, dtype=float)
        self.g_flat = [e for blk in self.g for e in blk]
        self.lbg = np.array(self.lbg, dtype=float)
        self.ubg = np.array(self.ubg, dtype=float)
        assert len(self.g_flat) == len(self.lbg) == len(self.ubg)

    def _guess_h(self, n):
        # base_raceline.py:731-736 / :1232-1238
        if self.ws is not None:
            return self.ws.step_sizes[n]
        if self.config.h0:
            return self.config.h0
        if self.parametric:
            ds = (self.line.s_max() - self.line.s_min()) / self.config.N
            return ds / self.config.v0 * self.line.p2mag_xcs(ds * n)
        return 1

    def _guess_z_base(self, n, k):
        cfg = self.config
        z = [0.] * 6      # position / s-y-n, then linear velocity; the drone mixin re-packs it
        if not self.parametric:
            # base_raceline.py:920-937
            if cfg.use_rk4:
                gate_no = n / self.gate_n_interval
            else:
                gate_no = (n + k / cfg.K) / self.gate_n_interval
            x = self.line.p2xc(gate_no)
            v = self.line.p2es(gate_no)
            v = v / np.linalg.norm(v) * cfg.v0
            z[:3] = x
            z[3:6] = v
        else:
            # base_raceline.py:1240-1251
            z[0] = self._get_s(n, k)
            if not self.model.config.global_r:
                z[3] = cfg.v0
            else:
                z[3:6] = cfg.v0 * self.line.p2es(self._get_s(n, k))
        return [float(e) for e in z]

    def _guess_z(self, n, k):
        z = self._guess_z_base(n, k)
        if not self.drone:
            return z
        # drone_raceline.py:158-262
        quat = self.model.config.use_quat
        z = [*z[:3], *([1, 0, 0, 0] if quat else [0, 0, 0]), *z[3:6], 0, 0, 0]
        if self.parametric:
            z[0] = self._get_s(n, k)
        if self.ws is None:
            return [float(e) for e in z]
        ws = self.ws
        z_ws, u_ws, du_ws = ws.Z[n, k], ws.U[n, k], ws.dU[n, k]   # == interp at the node time
        z[0:3] = z_ws[0:3]
        T = ws.model.f_T(z_ws, u_ws)
        vgw = ws.model.f_vg(z_ws, u_ws)
        if self.config.closed:
            e1 = vgw / np.linalg.norm(vgw)
            e3 = T / np.linalg.norm(T)
            e1 = e1 - e3 * (e1.T @ e3)
            e1 = e1 / np.linalg.norm(e1)
            e2 = np.cross(e3, e1)
            R = np.array([e1, e2, e3]).T
        else:
            def _hat(v):
                return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
            b = np.array([0, 0, 1])
            v = -np.cross(T / np.linalg.norm(T), b)
            s = np.linalg.norm(v)
            c = np.dot(T / np.linalg.norm(T), b)
            R = np.eye(3) + _hat(v) + _hat(v) @ _hat(v) * (1 - c) / s ** 2
        if not self.model.config.global_r:
            R = self.line.p2Rp(self._get_s(n, k)).T @ R
        if quat:
            r = _SciRot.from_matrix(R).as_quat()
            if self._last_ws_r is not None and np.linalg.norm(r - self._last_ws_r) >= 1:
                r = -r
            z[3:7] = r
        else:
            r = np.flip(_SciRot.from_matrix(R).as_euler('xyz', degrees=False))
            if self._last_ws_r is not None and np.linalg.norm(r - self._last_ws_r) > 1:
                if r[0] - self._last_ws_r[0] > np.pi:
                    r[0] -= 2 * np.pi
                elif r[0] - self._last_ws_r[0] <= -np.pi:
                    r[0] += 2 * np.pi
                if np.linalg.norm(r - self._last_ws_r) > 1:
                    raise NotImplementedError('Warmstart continuity failed for euler angles')
            z[3:6] = r
        if self._first_ws_r is None:
            self._first_ws_r = r
        self._last_ws_r = r
        vb = R.T @ vgw
        z[-6:-3] = vb
        dT = ws.model.f_T(z_ws, du_ws)
        wb = R.T @ np.cross(T, dT) / np.linalg.norm(T) ** 2
        z[-3:] = wb
        return [float(e) for e in z]

    def _guess_u(self, n, k):
        # drone_raceline.py:264-274 / base_raceline.py:742-745
        if self.drone and self.ws is not None:
            T = self.ws.model.f_T(self.ws.Z[n, k], self.ws.U[n, k])
            return [float(np.linalg.norm(T) / 4)] * 4
        return [0.] * self.nu

    # ---- unpacking (base_raceline.py:664-668) ------------------------------------------------
    def unpack(self, x):
        cfg = self.config
        N, P, S = cfg.N, cfg.K + 1, self.nz + 2 * self.nu
        x = np.asarray(x, dtype=float)
        H = x[:N]
        body = x[N:].reshape(N, P, S)
        return H, body[..., :self.nz], body[..., self.nz:self.nz + self.nu], body[..., self.nz + self.nu:]

    def warmstart_from(self, x):
        return RefWarmstart(self.model, *self.unpack(x))
