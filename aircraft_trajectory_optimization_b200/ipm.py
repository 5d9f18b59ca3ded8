'''
Batched primal-dual interior-point driver (the `nlpsol('solver', 'ipopt', ...)` replacement).

The reference hands its NLP to IPOPT through CasADi (drone3d/raceline/base_raceline.py:752-799)
and calls `self.solver(x0=, lbx=, ubx=, lbg=, ubg=)` (:160-165).  This module is the same kind of
solver -- a filter line-search barrier method following Waechter & Biegler (Math. Prog. 106, 2006),
the algorithm IPOPT implements [third party, not in the reference tree] -- written so that B
independent problem instances of one structure advance in lock step:

    min f(x)   s.t.  lbg <= g(x) <= ubg,   lbx <= x <= ubx

  * equality rows (lbg == ubg) stay equalities, inequality rows get slacks s with bounds;
  * every iteration evaluates f, grad f, g, jac_g, hess_l for the whole batch (one launch of the
    interval-cell kernel), condenses the slacks, and solves the KKT system with the batched
    block-tridiagonal + border kernel (csrc/kkt_blocks.cuh) followed by iterative refinement;
  * the block solver pivots symmetrically (Bunch-Parlett), so it reports the inertia of the KKT
    matrix; a wrong inertia (not exactly ng negative eigenvalues) or a vanishing pivot triggers
    IPOPT's delta_w / delta_c escalation schedule (Algorithm IC of the paper);
  * monotone barrier update, fraction-to-the-boundary rule, filter with switching / Armijo
    conditions;
  * feasibility restoration (IPOPT switches to it when the line search cannot make progress): the
    instance leaves the filter method and minimises the constraint violation from the reference
    point x_R where it got stuck,
        min  rho/2 ||c(x, s)||^2 + zeta/2 ||D_R (x - x_R)||^2 - mu sum log(distance to the bounds),
    by regularised Gauss-Newton steps -- the same KKT kernel with W = 0, diagonal -1/rho on every row
    (quasi-definite: always the right inertia, also where the Jacobian is rank deficient) and an
    Armijo search on that merit function -- until theta has dropped by the factor resto_kappa and
    the point is acceptable to the original filter; multipliers are reset on return.  (IPOPT's
    own restoration problem uses the l1 norm with extra slack variables; the l2 form needs no new
    unknowns and no new kernel.)

All state lives in torch tensors on the device of the inputs; the two heavy operations come from a
*backend* object (`eval`, `kkt_solve`, `kkt_matvec`).  The product backend is `CudaBackend`
(functions.NlpFunctions + kkt.KktSolver); there is no CPU backend in this package.
'''
from dataclasses import dataclass, field
import time

import numpy as np
import torch


@dataclass
class IpmOptions:
    tol: float = 1e-8
    acceptable_tol: float = 1e-6
    acceptable_iter: int = 15
    max_iter: int = 1000
    mu_init: float = 0.1
    mu_min_factor: float = 0.1         # mu >= tol * mu_min_factor
    kappa_eps: float = 10.0
    kappa_mu: float = 0.2
    theta_mu: float = 1.5
    tau_min: float = 0.99
    s_max: float = 100.0
    kappa_sigma: float = 1e10
    bound_push: float = 1e-2
    bound_frac: float = 1e-2
    bound_relax: float = 1e-8
    kappa_d: float = 1e-4
    gamma_theta: float = 1e-5
    gamma_phi: float = 1e-8
    delta_ls: float = 1.0
    s_theta: float = 1.1
    s_phi: float = 2.3
    eta_phi: float = 1e-8
    max_ls: int = 30
    ls_width: int = 4                  # step lengths tried per evaluation call in the line search
    filter_size: int = 64
    delta_w_first: float = 1e-4
    delta_w_min: float = 1e-20
    delta_w_max: float = 1e40
    kappa_w_minus: float = 1.0 / 3.0
    kappa_w_plus: float = 8.0
    kappa_w_plus_first: float = 100.0
    delta_c_bar: float = 1e-8
    # a Newton step this much larger than its right-hand side means the KKT matrix is numerically singular although every
    # pivot passed (a constraint whose gradient vanishes on the feasible set, e.g. the third thrust-axis row of an open
    # drone raceline, drone_raceline.py:110-148): treated like a vanishing pivot -> delta_c perturbation
    # Off by default (inf): with 1e10 the hard multi-start instances of the regular problems, whose steps are legitimately
    # huge for a few iterations, take other paths and some run out of iterations (tests/test_gpu_ipm.py).
    singular_step_ratio: float = float('inf')
    kappa_c: float = 0.25
    refine_steps: int = 4               # at most this many refinement steps per solve; a step is taken only while some
                                       # instance's scaled residual is above refine_tol (IPOPT: residual_ratio_max 1e-10)
    compact: bool = True               # copy the surviving instances into smaller tensors once few are left
    compact_frac: float = 0.5
    compact_min: int = 16
    speculate: int = 2                 # spare slots of a factorisation wave try this many further delta_w candidates
    speculate_max: int = 4             # ... up to this many when the wave is mostly empty
    use_glue: bool = True              # fused CUDA kernels for the vector work when the backend provides them
    max_soc: int = 0                   # second-order correction steps when the full step is rejected and theta grows (IPOPT:
    kappa_soc: float = 0.99            # max_soc = 4, kappa_soc); needs a backend that re-solves single instances.  Off by
                                       # default: on the C5 batch it is taken ~0.7 times per instance and changes neither the
                                       # iteration counts (median 79 -> 80) nor the converged set, at +19 % wall time
    restoration: bool = True           # feasibility restoration when the line search fails (see below)
    resto_kappa: float = 0.9           # leave restoration once theta <= resto_kappa * theta at entry
    resto_rho: float = 1e4             # weight of the constraint violation in the restoration merit (IPOPT: rho = 1000 on the l1 norm)
    resto_max_iter: int = 200          # restoration iterations per visit before the instance is given up
    window: int = 0                    # > 0: at most this many instances iterate at a time; finished ones are
                                       # replaced from the pending queue (keeps the batched kernels full)
    verbose: bool = False


@dataclass
class IpmResult:
    x: torch.Tensor
    f: torch.Tensor
    g: torch.Tensor
    lam_g: torch.Tensor
    lam_x: torch.Tensor
    success: torch.Tensor          # bool [B]
    status: torch.Tensor           # int [B]: 0 optimal, 1 acceptable, 2 max_iter, 3 failed
    iterations: torch.Tensor       # int [B]
    kkt_error: torch.Tensor
    n_iter: int = 0
    n_eval: int = 0
    n_factor: int = 0
    n_speculated: int = 0              # extra delta_w candidates factorised in spare wave slots
    n_restorations: int = 0            # visits of the feasibility restoration, summed over the instances
    n_soc: int = 0                     # steps accepted after a second-order correction
    t_eval: float = 0.0
    t_kkt: float = 0.0
    t_total: float = 0.0
    history: list = field(default_factory=list)
    factorisations_each: np.ndarray = None     # KKT factorisations every instance took part in


def _inf_norm(t):
    return t.abs().amax(dim=1) if t.shape[1] else torch.zeros(t.shape[0], dtype=t.dtype, device=t.device)


class InteriorPoint:
    def __init__(self, backend, options: IpmOptions = None):
        self.be = backend
        self.opt = options or IpmOptions()

    # ------------------------------------------------------------------------------------------------
    def solve(self, x0, lbx, ubx, lbg, ubg, lam_g0=None) -> IpmResult:
        o, be = self.opt, self.be
        t_start = time.perf_counter()
        dev, dt = x0.device, torch.float64
        B, n = x0.shape
        m = be.ng
        exp = lambda a: torch.as_tensor(a, dtype=dt, device=dev).expand(B, -1)
        lbx, ubx, lbg, ubg = exp(lbx), exp(ubx), exp(lbg), exp(ubg)
        eq = (lbg == ubg)
        ineq = ~eq
        # relaxed bounds (IPOPT bound_relax_factor)
        relax = lambda b, sgn: b + sgn * o.bound_relax * torch.clamp(b.abs(), min=1.0)
        hasL, hasU = torch.isfinite(lbx), torch.isfinite(ubx)
        xL = torch.where(hasL, relax(lbx, -1.0), torch.zeros_like(lbx))
        xU = torch.where(hasU, relax(ubx, +1.0), torch.zeros_like(ubx))
        shasL, shasU = torch.isfinite(lbg) & ineq, torch.isfinite(ubg) & ineq
        sL = torch.where(shasL, relax(lbg, -1.0), torch.zeros_like(lbg))
        sU = torch.where(shasU, relax(ubg, +1.0), torch.zeros_like(ubg))
        ceq = torch.where(eq, lbg, torch.zeros_like(lbg))
        # damping only for one-sided bounds
        dampL, dampU = (hasL & ~hasU).to(dt), (hasU & ~hasL).to(dt)
        sdampL, sdampU = (shasL & ~shasU).to(dt), (shasU & ~shasL).to(dt)
        fL, fU, sfL, sfU = hasL.to(dt), hasU.to(dt), shasL.to(dt), shasU.to(dt)
        n_bounds = (fL + fU).sum(1) + (sfL + sfU).sum(1)
        # fused CUDA kernels for the vector work of a sweep (backend.glue; the torch arithmetic below is the same maths)
        glue = getattr(be, 'glue', None) if o.use_glue else None
        xflag = sflag = None
        if glue is not None:
            xflag = (hasL.to(torch.uint8) | (hasU.to(torch.uint8) << 1)).contiguous()
            sflag = (shasL.to(torch.uint8) | (shasU.to(torch.uint8) << 1) | (eq.to(torch.uint8) << 2)).contiguous()
            xL, xU, sL, sU, ceq = (t_.contiguous() for t_ in (xL, xU, sL, sU, ceq))

        def push(v, L, U, hL, hU):
            ''' move v strictly inside [L, U] (IPOPT bound_push / bound_frac) '''
            span = torch.where(hL & hU, U - L, torch.full_like(v, float('inf')))
            pL = torch.minimum(o.bound_push * torch.clamp(L.abs(), min=1.0), o.bound_frac * span)
            pU = torch.minimum(o.bound_push * torch.clamp(U.abs(), min=1.0), o.bound_frac * span)
            v = torch.where(hL, torch.maximum(v, L + pL), v)
            v = torch.where(hU, torch.minimum(v, U - pU), v)
            return v

        x = push(x0.to(dt).clone(), xL, xU, hasL, hasU)
        res = IpmResult(None, None, None, None, None, None, None, None, None)
        ones_B = torch.ones(B, dtype=dt, device=dev)

        def evaluate(xx, yy, derivs, mask=None, prev=None):
            ''' mask: instances whose values are needed; the other rows keep the values of `prev` (the driver's
            current evaluation, updated in place -- it survives compaction because the driver owns it) '''
            t0 = time.perf_counter()
            idx = None
            if mask is not None and not bool(mask.all()):
                idx = torch.nonzero(mask).squeeze(1)
            out = be.eval(xx, yy, ones_B, derivs, idx, prev if idx is not None else None)
            if dev.type == 'cuda':
                torch.cuda.synchronize(dev)
            res.t_eval += time.perf_counter() - t0
            res.n_eval += 1
            return out

        y = torch.zeros(B, m, dtype=dt, device=dev) if lam_g0 is None else lam_g0.to(dt).clone()
        ev = evaluate(x, y, True)
        s = push(torch.where(ineq, ev['g'], ceq), sL, sU, shasL, shasU)
        s = torch.where(ineq, s, ceq)
        zL, zU = fL.clone(), fU.clone()
        vL, vU = sfL.clone(), sfU.clone()
        mu = torch.full((B,), o.mu_init, dtype=dt, device=dev)
        zero_h = torch.zeros_like(ev['hess'])

        def kkt(hess, jac, dxd, negd, rhs, mask=None):
            t0 = time.perf_counter()
            idx = None
            if mask is not None and not bool(mask.all()):
                idx = torch.nonzero(mask).squeeze(1)
            sol, st = be.kkt_solve(hess, jac, dxd, negd, rhs, o.refine_steps, idx)
            if dev.type == 'cuda':
                torch.cuda.synchronize(dev)
            res.t_kkt += time.perf_counter() - t0
            res.n_factor += 1
            return sol, st

        def jt_y(jac, yy):
            vec = torch.cat([torch.zeros(B, n, dtype=dt, device=dev), yy], dim=1)
            z = torch.zeros(B, n, dtype=dt, device=dev)
            return be.kkt_matvec(zero_h, jac, z, torch.zeros(B, m, dtype=dt, device=dev), vec)[:, :n]

        # ---- least-squares multiplier estimate (IPOPT 3.6 in the paper) ------------------------------
        if lam_g0 is None:
            rhs = torch.cat([-(ev['grad_f'] - zL + zU), torch.where(ineq, vL - vU, torch.zeros_like(s))], dim=1)
            negd = torch.where(ineq, -torch.ones_like(s), torch.zeros_like(s))
            sol, st = kkt(zero_h, ev['jac'], torch.ones(B, n, dtype=dt, device=dev), negd, rhs)
            y0 = sol[:, n:]
            ok = torch.isfinite(y0).all(1) & (_inf_norm(y0) <= 1e3) & (st[:, 0] == 0)
            y = torch.where(ok[:, None], y0, torch.zeros_like(y0))
            ev = evaluate(x, y, True)

        def slacks(xx, ss):
            return (xx - xL, xU - xx, ss - sL, sU - ss)

        def barrier(fv, xx, ss, mu_):
            dL, dU, eL, eU = slacks(xx, ss)
            lg = lambda d_, f_: (torch.log(torch.where(f_ > 0, d_, torch.ones_like(d_))) * f_).sum(1)
            phi = fv - mu_ * (lg(dL, fL) + lg(dU, fU) + lg(eL, sfL) + lg(eU, sfU))
            phi = phi + o.kappa_d * mu_ * ((dL * dampL).sum(1) + (dU * dampU).sum(1)
                                          + (eL * sdampL).sum(1) + (eU * sdampU).sum(1))
            return phi

        def infeas(gv, ss):
            return torch.where(eq, gv - ceq, gv - ss)

        def error_parts(ev_, xx, ss, yy, zL_, zU_, vL_, vU_):
            ''' the mu-independent pieces of IPOPT's scaled optimality error (one J'y product) '''
            gradL_x = ev_['grad_f'] + jt_y(ev_['jac'], yy) - zL_ + zU_
            gradL_s = torch.where(ineq, -yy - vL_ + vU_, torch.zeros_like(yy))
            zsum = (zL_ * fL + zU_ * fU).sum(1) + (vL_ * sfL + vU_ * sfU).sum(1)
            n_tot = n + int(ineq[0].sum())
            s_d = torch.clamp((yy.abs().sum(1) + zsum) / max(1, m + n_tot), min=o.s_max) / o.s_max
            s_c = torch.clamp(zsum / torch.clamp(n_bounds, min=1.0), min=o.s_max) / o.s_max
            dual = torch.maximum(_inf_norm(gradL_x), _inf_norm(gradL_s))
            prim = _inf_norm(infeas(ev_['g'], ss))
            dL, dU, eL, eU = slacks(xx, ss)
            prods = (dL * zL_ * fL, dU * zU_ * fU, eL * vL_ * sfL, eU * vU_ * sfU)
            return dual, prim, s_d, s_c, prods

        def error_at(parts, mu_):
            if parts[0] == 'glue':
                _, dual, prim, s_d, s_c, pmin, pmax = parts
                comp = torch.clamp(torch.maximum(pmax - mu_, mu_ - pmin), min=0.0)
                return torch.stack([dual / s_d, prim, comp / s_c]).amax(0), dual, prim, comp
            dual, prim, s_d, s_c, prods = parts
            mc = mu_[:, None]
            comp = torch.stack([_inf_norm((pr - mc * f_)) for pr, f_ in zip(prods, (fL, fU, sfL, sfU))]).amax(0)
            return torch.stack([dual / s_d, prim, comp / s_c]).amax(0), dual, prim, comp

        def errors(ev_, xx, ss, yy, zL_, zU_, vL_, vU_, mu_):
            return error_at(error_parts(ev_, xx, ss, yy, zL_, zU_, vL_, vU_), mu_)

        # ---- state -------------------------------------------------------------------------------------
        F = o.filter_size
        filt_theta = torch.full((B, F), float('inf'), dtype=dt, device=dev)
        filt_phi = torch.full((B, F), float('inf'), dtype=dt, device=dev)
        filt_n = torch.zeros(B, dtype=torch.long, device=dev)
        theta0 = infeas(ev['g'], s).abs().sum(1)
        theta_max = 1e4 * torch.clamp(theta0, min=1.0)
        theta_min = 1e-4 * torch.clamp(theta0, min=1.0)
        filt_theta[:, 0] = theta_max          # the initial filter {theta >= theta_max}
        filt_phi[:, 0] = -float('inf')
        filt_n[:] = 1
        delta_w_last = torch.zeros(B, dtype=dt, device=dev)
        delta_w = torch.zeros(B, dtype=dt, device=dev)
        delta_c = torch.zeros(B, dtype=dt, device=dev)
        first_try = torch.ones(B, dtype=torch.bool, device=dev)
        attempts = torch.zeros(B, dtype=torch.long, device=dev)
        n_fact = torch.zeros(B, dtype=torch.long, device=dev)
        status = torch.full((B,), -1, dtype=torch.long, device=dev)      # -1 iterating, -2 pending, >= 0 finished
        if 0 < o.window < B:
            status[o.window:] = -2
        iters = torch.zeros(B, dtype=torch.long, device=dev)
        acc_count = torch.zeros(B, dtype=torch.long, device=dev)
        ls_fail = torch.zeros(B, dtype=torch.long, device=dev)
        moved_prev = torch.ones(B, dtype=torch.bool, device=dev)
        resto = torch.zeros(B, dtype=torch.bool, device=dev)             # in feasibility restoration
        resto_it = torch.zeros(B, dtype=torch.long, device=dev)
        theta_R = torch.zeros(B, dtype=dt, device=dev)                   # theta at entry
        x_R = x.clone()                                                  # reference point of the restoration
        DR2 = torch.ones_like(x)                                         # D_R^2 = 1 / max(1, |x_R|)^2
        n_resto = torch.zeros(B, dtype=torch.long, device=dev)           # restoration visits per instance
        mu_floor = o.tol * o.mu_min_factor

        def reset_filter(mask):
            filt_theta[mask] = float('inf')
            filt_phi[mask] = float('inf')
            filt_theta[mask, 0] = theta_max[mask]
            filt_phi[mask, 0] = -float('inf')
            filt_n[mask] = 1

        def GS():
            ''' the current state tensors for the fused kernels (the names are re-bound as the sweep goes on) '''
            return dict(x=x, s=s, y=y, zL=zL, zU=zU, vL=vL, vU=vU, xL=xL, xU=xU, sL=sL, sU=sU, ceq=ceq, xflag=xflag,
                        sflag=sflag, mu=mu, delta_w=delta_w, delta_c=delta_c,
                        resto=resto.to(torch.uint8) if bool(resto.any()) else None, x_R=x_R, DR2=DR2)

        # ---- physical compaction: once most instances have finished, the survivors are copied into smaller tensors so
        # that the elementwise work of a sweep scales with what is still iterating, not with the original batch
        B0 = B
        orig = torch.arange(B, device=dev)
        OUT = {}

        def stash(rows_local):
            ''' park the final state of finished instances (local row indices) in the full-size result arrays '''
            if not OUT:
                OUT.update(x=torch.empty(B0, n, dtype=dt, device=dev), f=torch.empty(B0, dtype=dt, device=dev),
                           g=torch.empty(B0, m, dtype=dt, device=dev), lam_g=torch.empty(B0, m, dtype=dt, device=dev),
                           lam_x=torch.empty(B0, n, dtype=dt, device=dev), err=torch.empty(B0, dtype=dt, device=dev),
                           status=torch.empty(B0, dtype=torch.long, device=dev),
                           iters=torch.empty(B0, dtype=torch.long, device=dev),
                           n_fact=torch.empty(B0, dtype=torch.long, device=dev))
            if rows_local.numel() == 0:
                return
            dst = orig[rows_local]
            e_fin = error_at(error_parts(ev, x, s, y, zL, zU, vL, vU), torch.zeros_like(mu))[0]
            for key, src in (('x', torch.minimum(torch.maximum(x, lbx), ubx)), ('f', ev['f']), ('g', ev['g']), ('lam_g', y), ('lam_x', zU - zL), ('err', e_fin),
                             ('status', status), ('iters', iters), ('n_fact', n_fact)):
                OUT[key][dst] = src[rows_local]

        it = 0
        while True:
            active = status == -1
            n_live = int(((status == -1) | (status == -2)).sum())
            if o.compact and B > o.compact_min and n_live <= o.compact_frac * B and n_live > 0:
                live = (status == -1) | (status == -2)
                stash(torch.nonzero(~live).squeeze(1))
                keep = torch.nonzero(live).squeeze(1)
                sel = lambda t_: t_[keep].contiguous()
                (x, s, y, zL, zU, vL, vU, mu, filt_theta, filt_phi, filt_n, theta_max, theta_min, delta_w_last, delta_w,
                 delta_c, first_try, attempts, n_fact, status, iters, acc_count, ls_fail, orig, ones_B, zero_h,
                 resto, resto_it, theta_R, x_R, DR2, n_resto) = map(sel, (
                     x, s, y, zL, zU, vL, vU, mu, filt_theta, filt_phi, filt_n, theta_max, theta_min, delta_w_last, delta_w,
                     delta_c, first_try, attempts, n_fact, status, iters, acc_count, ls_fail, orig, ones_B, zero_h,
                     resto, resto_it, theta_R, x_R, DR2, n_resto))
                (xL, xU, sL, sU, ceq, eq, ineq, fL, fU, sfL, sfU, dampL, dampU, sdampL, sdampU, n_bounds, lbx, ubx,
                 moved_prev) = map(sel, (xL, xU, sL, sU, ceq, eq, ineq, fL, fU, sfL, sfU, dampL, dampU, sdampL, sdampU,
                                         n_bounds, lbx, ubx, moved_prev))
                ev = {k_: sel(v_) for k_, v_ in ev.items()}
                if glue is not None:
                    xflag, sflag = sel(xflag), sel(sflag)
                be.select(keep)
                B = keep.numel()
                active = status == -1
            # ---- leave the restoration: theta has dropped by resto_kappa and the filter accepts the point --------
            if bool((resto & active).any()):
                th0 = infeas(ev['g'], s).abs().sum(1)
                ph0 = barrier(ev['f'], x, s, mu)
                in_f = ((th0[:, None] >= (1 - o.gamma_theta) * filt_theta)
                        & (ph0[:, None] >= filt_phi - o.gamma_phi * filt_theta)).any(1)
                leave = resto & active & (th0 <= o.resto_kappa * theta_R) & ~in_f
                if bool(leave.any()):
                    lv = leave[:, None]
                    resto = resto & ~leave
                    y = torch.where(lv, torch.zeros_like(y), y)            # multipliers restart; the kappa_sigma
                    zL, zU = torch.where(lv, fL, zL), torch.where(lv, fU, zU)      # safeguard pulls z towards mu / d
                    vL, vU = torch.where(lv, sfL, vL), torch.where(lv, sfU, vU)
                    ev = evaluate(x, y, True, leave, ev)                   # hess_l with the new multipliers
            # ---- convergence and barrier update --------------------------------------------------------
            if glue is not None:
                jty_cur = jt_y(ev['jac'], y).contiguous()
                E8 = glue.error(GS(), ev, jty_cur)
                n_tot = n + int(ineq[0].sum())
                zsum_, ysum_ = E8[:, 2], E8[:, 3]
                parts = ('glue', E8[:, 0], E8[:, 1],
                         torch.clamp((ysum_ + zsum_) / max(1, m + n_tot), min=o.s_max) / o.s_max,
                         torch.clamp(zsum_ / torch.clamp(n_bounds, min=1.0), min=o.s_max) / o.s_max, E8[:, 4], E8[:, 5])
            else:
                parts = error_parts(ev, x, s, y, zL, zU, vL, vU)
            E0, dual0, prim0, comp0 = error_at(parts, torch.zeros_like(mu))
            done = active & ~resto & (E0 <= o.tol)
            status[done] = 0
            acc = active & ~resto & ~done & (E0 <= o.acceptable_tol)
            # counted per iteration (like IPOPT), not per sweep: an instance stalled on inertia correction has not moved
            acc_count = torch.where(acc, acc_count + moved_prev.long(), torch.zeros_like(acc_count))
            done_acc = acc & (acc_count >= o.acceptable_iter)
            status[done_acc] = 1
            active = status == -1
            if o.verbose:
                k = 0
                print(f'{it:4d} f={float(ev["f"][k]):.8e} inf_pr={float(prim0[k]):.2e} inf_du={float(dual0[k]):.2e} '
                      f'compl={float(comp0[k]):.2e} lg(mu)={np.log10(float(mu[k])):.1f} dw={float(delta_w_last[k]):.1e} '
                      f'active={int(active.sum())}')
            # refill the window from the pending queue
            n_act = int(active.sum())
            if o.window > 0 and n_act < o.window:
                pend = torch.nonzero(status == -2).squeeze(1)
                if pend.numel():
                    status[pend[:o.window - n_act]] = -1
                    active = status == -1
                    n_act = int(active.sum())
            res.history.append((it, n_act))
            if n_act == 0:
                break
            for _ in range(4):
                Emu = error_at(parts, mu)[0]
                upd = active & ~resto & (Emu <= o.kappa_eps * mu) & (mu > mu_floor)
                if not bool(upd.any()):
                    break
                mu_new = torch.clamp(torch.minimum(o.kappa_mu * mu, mu ** o.theta_mu), min=mu_floor)
                mu = torch.where(upd, mu_new, mu)
                reset_filter(upd)
            tau = torch.clamp(1.0 - mu, min=o.tau_min)

            # ---- the Newton system ---------------------------------------------------------------------
            mu_c = mu[:, None]
            if glue is not None:
                x, s, y, zL, zU, vL, vU, mu, delta_w, delta_c = (t_.contiguous() for t_ in (x, s, y, zL, zU, vL, vU, mu,
                                                                                            delta_w, delta_c))
                NW = glue.newton(GS(), ev, jty_cur)
                c, r_s, gphi_x, gphi_s = NW['c'], NW['r_s'], NW['gphi_x'], NW['gphi_s']
                Sx = Ss = r_x = None
            else:
                dL, dU, eL, eU = slacks(x, s)
                inv = lambda d_, f_: f_ / torch.where(f_ > 0, d_, torch.ones_like(d_))
                iL, iU, jL, jU = inv(dL, fL), inv(dU, fU), inv(eL, sfL), inv(eU, sfU)
                Sx = zL * iL + zU * iU
                Ss = vL * jL + vU * jU
                gphi_x = ev['grad_f'] - mu_c * iL + mu_c * iU + o.kappa_d * mu_c * (dampL - dampU)
                gphi_s = torch.where(ineq, -mu_c * jL + mu_c * jU + o.kappa_d * mu_c * (sdampL - sdampU),
                                     torch.zeros_like(s))
                r_x = gphi_x + jt_y(ev['jac'], y)
                r_s = torch.where(ineq, gphi_s - y, torch.zeros_like(s))
                c = infeas(ev['g'], s)

            # One factorisation call per sweep (IPOPT's Algorithm IC, de-serialised over the batch): an instance
            # whose KKT matrix has the wrong inertia does not step in this sweep; it keeps its iterate, escalates
            # its own (delta_w, delta_c) and is factorised again together with everybody else in the next sweep.
            def newton(rows, dw, dc):
                ''' (dx_diag, neg_d, rhs, Ss + dw) of the Newton system with regularisation (dw, dc); rows: index
                tensor into the batch or None for every instance '''
                pick = (lambda t: t) if rows is None else (lambda t: t[rows])
                if glue is not None:
                    # from the fused kernel's outputs (computed with the instance's current delta_w)
                    if rows is None:
                        return NW['dxd'], NW['negd'], NW['rhs'], NW['Ssr']
                    iq, c_, dw0 = pick(ineq), pick(c), delta_w[rows][:, None]
                    Ssr = torch.where(iq, pick(NW['Ssr']) - dw0 + dw[:, None], torch.ones_like(c_))
                    negd_ = torch.where(iq, -1.0 / Ssr, torch.zeros_like(c_)) - dc[:, None]
                    rhs_ = torch.cat([pick(NW['rhs'])[:, :n], torch.where(iq, -c_ - pick(r_s) / Ssr, -c_)], dim=1)
                    return pick(NW['dxd']) - dw0 + dw[:, None], negd_, rhs_, Ssr
                iq, Ss_, c_ = pick(ineq), pick(Ss), pick(c)
                Ssr = torch.where(iq, Ss_ + dw[:, None], torch.ones_like(Ss_))
                negd_ = torch.where(iq, -1.0 / Ssr, torch.zeros_like(Ss_)) - dc[:, None]
                rhs_ = torch.cat([-pick(r_x), torch.where(iq, -c_ - pick(r_s) / Ssr, -c_)], dim=1)
                return pick(Sx) + dw[:, None], negd_, rhs_, Ssr

            def escalate(dw, ft, dwl):
                ''' the next delta_w of Algorithm IC after a wrong-inertia factorisation with delta_w = dw '''
                start = torch.where(dwl == 0, torch.full_like(dw, o.delta_w_first),
                                    torch.clamp(o.kappa_w_minus * dwl, min=o.delta_w_min))
                grow = torch.where(dwl == 0, o.kappa_w_plus_first * dw, o.kappa_w_plus * dw)
                return torch.where(ft, start, grow)

            dxd, negd, rhs, Ss_reg = newton(None, delta_w, delta_c)
            any_resto = bool((resto & active).any())
            hess_in = ev['hess']
            if any_resto and glue is not None:
                hess_in = torch.where(resto[:, None], torch.zeros_like(hess_in), hess_in)      # W = 0 in the restoration
            if any_resto and glue is None:
                # regularised Gauss-Newton step on the constraint violation (module docstring): primal barrier terms
                rr = resto[:, None]
                zeta = torch.sqrt(mu)[:, None]
                SxR = mu_c * (iL * iL + iU * iU)
                SsR = torch.where(ineq, mu_c * (jL * jL + jU * jU), torch.ones_like(s))
                gbx = -mu_c * iL + mu_c * iU + o.kappa_d * mu_c * (dampL - dampU)
                gbs = torch.where(ineq, -mu_c * jL + mu_c * jU + o.kappa_d * mu_c * (sdampL - sdampU), torch.zeros_like(s))
                dxd = torch.where(rr, zeta * DR2 + SxR, dxd)
                negd = torch.where(rr, torch.where(ineq, -(1.0 / o.resto_rho + 1.0 / SsR), torch.full_like(s, -1.0 / o.resto_rho)), negd)
                rhs = torch.where(rr, torch.cat([-(zeta * DR2 * (x - x_R) + gbx),
                                                 torch.where(ineq, -c - gbs / SsR, -c)], dim=1), rhs)
                hess_in = torch.where(rr, torch.zeros_like(hess_in), hess_in)
            # Speculative candidates: a launch of the factorisation kernel takes as long for one instance as for a
            # full wave of them, so the spare slots of the last wave factorise the next `speculate` entries of
            # some instances' delta_w escalation sequences alongside; an instance whose first matrix has the wrong
            # inertia then picks up the first later candidate with the right one in the same sweep.  The iterates
            # are exactly those of the one-candidate-per-sweep schedule.
            spec_rows, spec_dw, spec_parts = None, [], []
            wave = int(getattr(be, 'kkt_wave', 0)) if o.speculate > 0 else 0
            if wave > 0:
                free = (-n_act) % wave
                # few instances in flight: every one of them gets a longer run of candidates
                depth_s = max(o.speculate, min(o.speculate_max, free // n_act))
                n_spec = min(free // depth_s, n_act)
                if n_spec > 0:
                    score = active.to(dt) * (1.0 + 2.0 * (delta_w > 0).to(dt) + (delta_w_last > 0).to(dt))
                    spec_rows = torch.argsort(-score, stable=True)[:n_spec]
                    dwc, ft = delta_w[spec_rows], first_try[spec_rows]
                    for _ in range(depth_s):
                        dwc = escalate(dwc, ft, delta_w_last[spec_rows])
                        ft = torch.zeros_like(ft)
                        spec_dw.append(dwc)
                        spec_parts.append(newton(spec_rows, dwc, delta_c[spec_rows]))
            fac_slot = torch.full((B,), -1, dtype=torch.long, device=dev)     # factor slot of every instance's accepted system
            if spec_rows is None:
                sol, st = kkt(hess_in, ev['jac'], dxd, negd, rhs, active)
                ia_ = torch.nonzero(active).squeeze(1)
                fac_slot[ia_] = torch.arange(ia_.numel(), device=dev)
            else:
                t0 = time.perf_counter()
                idx_act = torch.nonzero(active).squeeze(1)
                idx_all = torch.cat([idx_act] + [spec_rows] * len(spec_parts))
                cat = lambda k, full: torch.cat([full[idx_act]] + [p_[k] for p_ in spec_parts])
                sol_all, st_all = be.kkt_solve_rows(hess_in, ev['jac'], idx_all, cat(0, dxd), cat(1, negd), cat(2, rhs),
                                                    o.refine_steps)
                if dev.type == 'cuda':
                    torch.cuda.synchronize(dev)
                res.t_kkt += time.perf_counter() - t0
                res.n_factor += 1
                res.n_speculated += int(spec_rows.numel()) * len(spec_parts)
                na = int(idx_act.numel())
                sol = torch.zeros(B, n + m, dtype=dt, device=dev)
                st = torch.zeros(B, 2, dtype=st_all.dtype, device=dev)
                st[:, 1] = m
                sol.index_copy_(0, idx_act, sol_all[:na])
                st.index_copy_(0, idx_act, st_all[:na])
                fac_slot[idx_act] = torch.arange(na, device=dev)
                spec_pos = torch.full((B,), -1, dtype=torch.long, device=dev)
                spec_pos[spec_rows] = torch.arange(spec_rows.numel(), device=dev)
            n_fact = n_fact + active.long()
            if o.verbose:
                print(f'        dw={float(delta_w[0]):.2e} dc={float(delta_c[0]):.2e} bad_piv={int(st[0, 0])} '
                      f'neg={int(st[0, 1])} (want {m}) attempt {int(attempts[0])}')
            pending = active.clone()                 # instances whose current candidate is still to be judged
            bad = torch.zeros_like(active)
            moved = torch.zeros_like(active)
            for depth in range(len(spec_parts) + 1):
                finite = torch.isfinite(sol).all(1)
                singular = (st[:, 0] != 0) | ~finite
                if o.singular_step_ratio < float('inf') and rhs.shape == sol.shape:
                    singular = singular | (sol.abs().amax(1) > o.singular_step_ratio * torch.clamp(rhs.abs().amax(1), min=1.0))
                wrong_inertia = st[:, 1] != m
                bad_j = pending & (singular | wrong_inertia)
                ok_j = pending & ~bad_j
                moved = moved | ok_j
                delta_w_last = torch.where(ok_j & (delta_w > 0), delta_w, delta_w_last)
                # too few negative eigenvalues or a vanishing pivot: the constraint Jacobian is (numerically) rank
                # deficient -> perturb the constraint block first and retry with the same delta_w
                degenerate = bad_j & (singular | (st[:, 1] < m)) & (delta_c == 0)
                esc = bad_j & ~degenerate
                delta_w_next = torch.where(esc, escalate(delta_w, first_try, delta_w_last), delta_w)
                delta_c_next = torch.where(degenerate, o.delta_c_bar * mu ** o.kappa_c, delta_c)
                attempts = torch.where(bad_j, attempts + 1, torch.where(ok_j, torch.zeros_like(attempts), attempts))
                status[bad_j & ((delta_w_next > o.delta_w_max) | (attempts > 60))] = 3
                # instances that step start their next iteration from delta_w = delta_c = 0 again
                delta_w = torch.where(bad_j, delta_w_next, torch.where(ok_j, torch.zeros_like(delta_w), delta_w))
                delta_c = torch.where(bad_j, delta_c_next, torch.where(ok_j, torch.zeros_like(delta_c), delta_c))
                first_try = torch.where(bad_j, first_try & ~esc, torch.where(ok_j, torch.ones_like(first_try), first_try))
                take = torch.zeros_like(active)
                if depth < len(spec_parts):
                    # the candidate factorised alongside is this instance's next attempt
                    cand_dw = torch.zeros_like(delta_w)
                    cand_dw[spec_rows] = spec_dw[depth]
                    take = esc & (status == -1) & (spec_pos >= 0) & (cand_dw == delta_w)
                bad = bad | (bad_j & ~take)
                if not bool(take.any()):
                    break
                rows_t = torch.nonzero(take).squeeze(1)
                src = na + depth * int(spec_rows.numel()) + spec_pos[rows_t]
                sol.index_copy_(0, rows_t, sol_all[src])
                st.index_copy_(0, rows_t, st_all[src])
                fac_slot[rows_t] = src
                Ss_reg.index_copy_(0, rows_t, spec_parts[depth][3][spec_pos[rows_t]])
                n_fact = n_fact + take.long()
                pending = take
            stalled = bad & (status == -1)
            if glue is not None:
                NW['Ssr'] = Ss_reg if Ss_reg.is_contiguous() else Ss_reg.contiguous()      # (rows of accepted speculative candidates differ)
                DIR = glue.direction(GS(), sol.contiguous(), moved, tau.contiguous(), NW)
                dx, dy, ds, dzL, dzU, dvL, dvU = (DIR[k_] for k_ in ('dx', 'dy', 'ds', 'dzL', 'dzU', 'dvL', 'dvU'))
                a_pr, a_du = DIR['sc'][:, 0].clone(), DIR['sc'][:, 1].clone()
            else:
                mv = moved[:, None]
                dx = torch.where(mv, sol[:, :n], torch.zeros_like(x))
                dy = torch.where(mv, sol[:, n:], torch.zeros_like(y))
                ds = torch.where(mv & ineq, (dy - r_s) / Ss_reg, torch.zeros_like(s))
                if any_resto:
                    rr = resto[:, None]
                    ds = torch.where(rr, torch.where(mv & ineq, (sol[:, n:] - gbs) / SsR, torch.zeros_like(s)), ds)
                    dy = torch.where(rr, torch.zeros_like(dy), dy)          # multipliers rest during restoration

                dzL = (mu_c * iL - zL - zL * iL * dx) * fL
                dzU = (mu_c * iU - zU + zU * iU * dx) * fU
                dvL = (mu_c * jL - vL - vL * jL * ds) * sfL
                dvU = (mu_c * jU - vU + vU * jU * ds) * sfU
                if any_resto:
                    keep_ = (~resto)[:, None].to(dt)
                    dzL, dzU, dvL, dvU = dzL * keep_, dzU * keep_, dvL * keep_, dvU * keep_

                def max_step(d_, step, f_, t_):
                    ''' largest alpha in (0, 1] with d_ + alpha * step >= (1 - tau) d_ '''
                    ratio = torch.where((step < 0) & (f_ > 0), -t_[:, None] * d_ / step, torch.full_like(d_, float('inf')))
                    return torch.clamp(ratio.amin(1), max=1.0) if ratio.shape[1] else torch.ones_like(t_)

                a_pr = torch.stack([max_step(dL, dx, fL, tau), max_step(dU, -dx, fU, tau),
                                    max_step(eL, ds, sfL, tau), max_step(eU, -ds, sfU, tau)]).amin(0)
                a_du = torch.stack([max_step(zL, dzL, fL, tau), max_step(zU, dzU, fU, tau),
                                    max_step(vL, dvL, sfL, tau), max_step(vU, dvU, sfU, tau)]).amin(0)

            # ---- filter line search ---------------------------------------------------------------------
            if glue is not None:
                theta, phi, phiR = NW['sc'][:, 0].clone(), NW['sc'][:, 1].clone(), NW['sc'][:, 2].clone()
                dphi, dphiR = DIR['sc'][:, 2].clone(), DIR['sc'][:, 3].clone()
            else:
                theta = c.abs().sum(1)
                phi = barrier(ev['f'], x, s, mu)
                dphi = (gphi_x * dx).sum(1) + (gphi_s * ds).sum(1)
            if any_resto and glue is None:
                # merit of the restoration problem and its slope along the step
                zeta1 = torch.sqrt(mu)
                phiR = 0.5 * o.resto_rho * (c * c).sum(1) + 0.5 * zeta1 * (DR2 * (x - x_R) ** 2).sum(1) + barrier(torch.zeros_like(mu), x, s, mu)
                cJd = sol[:, n:] - o.resto_rho * c       # rho (J dx - ds) on inequality rows, rho J dx on equality rows: w = rho (c + Jc d)
                dphiR = (c * cJd).sum(1) + (zeta1[:, None] * DR2 * (x - x_R) * dx).sum(1) + (gbx * dx).sum(1) + (gbs * ds).sum(1)
            alpha = a_pr.clone()
            searching = moved & (status == -1)
            accepted_alpha = torch.zeros_like(alpha)
            augment = torch.zeros(B, dtype=torch.bool, device=dev)
            switch_base = (dphi < 0) & (theta <= theta_min)
            last_alpha = alpha.clone()
            # Every round tries `ls_width` step lengths alpha, alpha/2, ... of every instance still searching in ONE
            # evaluation call (an evaluation costs microseconds per instance; a round trip through the driver does
            # not) and accepts the longest admissible one -- the same step a sequential backtracking search takes.
            Kw = max(1, int(o.ls_width))
            halves = 0.5 ** torch.arange(Kw, dtype=dt, device=dev)
            soc_on = o.max_soc > 0 and glue is not None and bool(getattr(be, 'can_soc', False))
            ls = 0
            while ls < o.max_ls:
                rows = torch.nonzero(searching).squeeze(1)
                if rows.numel() == 0:
                    break
                ns = rows.numel()
                al = alpha[rows][:, None] * halves[None, :]                                    # [ns, Kw]
                if glue is not None:
                    al = al.contiguous()
                    xt2, rows32 = glue.trial(x, dx, rows, al)                                  # [ns * Kw, n]
                    t0 = time.perf_counter()
                    evt = be.eval_points(xt2, rows.repeat_interleave(Kw))
                    if dev.type == 'cuda':
                        torch.cuda.synchronize(dev)
                    res.t_eval += time.perf_counter() - t0
                    res.n_eval += 1
                    TM = glue.trial_merit(GS(), rows32, al, xt2, ds, evt['f'].contiguous(), evt['g'].contiguous())
                    th_t, ph_t, phR_t, okfin = TM[:, :, 0], TM[:, :, 1], TM[:, :, 2], TM[:, :, 3] > 0.5
                else:
                    xt = x[rows][:, None, :] + al[:, :, None] * dx[rows][:, None, :]               # [ns, Kw, n]
                    st_ = torch.where(ineq[rows][:, None, :], s[rows][:, None, :] + al[:, :, None] * ds[rows][:, None, :],
                                      ceq[rows][:, None, :])
                    t0 = time.perf_counter()
                    evt = be.eval_points(xt.reshape(ns * Kw, n), rows.repeat_interleave(Kw))
                    if dev.type == 'cuda':
                        torch.cuda.synchronize(dev)
                    res.t_eval += time.perf_counter() - t0
                    res.n_eval += 1
                    g_t, f_t = evt['g'].reshape(ns, Kw, m), evt['f'].reshape(ns, Kw)
                    th_t = torch.where(eq[rows][:, None, :], g_t - ceq[rows][:, None, :], g_t - st_).abs().sum(2)
                    # barrier function at the trial points
                    lg = lambda d_, f_: (torch.log(torch.where(f_ > 0, d_, torch.ones_like(d_))) * f_).sum(2)
                    r_ = lambda t: t[rows][:, None, :]
                    dLt, dUt, eLt, eUt = xt - r_(xL), r_(xU) - xt, st_ - r_(sL), r_(sU) - st_
                    mu_r = mu[rows][:, None]
                    ph_t = f_t - mu_r * (lg(dLt, r_(fL)) + lg(dUt, r_(fU)) + lg(eLt, r_(sfL)) + lg(eUt, r_(sfU)))
                    ph_t = ph_t + o.kappa_d * mu_r * ((dLt * r_(dampL)).sum(2) + (dUt * r_(dampU)).sum(2)
                                                      + (eLt * r_(sdampL)).sum(2) + (eUt * r_(sdampU)).sum(2))
                    okfin = torch.isfinite(th_t) & torch.isfinite(ph_t)
                ft, fp = filt_theta[rows][:, None, :], filt_phi[rows][:, None, :]
                in_filter = ((th_t[:, :, None] >= (1 - o.gamma_theta) * ft)
                             & (ph_t[:, :, None] >= fp - o.gamma_phi * ft)).any(2)
                th0, ph0, dph = theta[rows][:, None], phi[rows][:, None], dphi[rows][:, None]
                sw = switch_base[rows][:, None] & (al * (-dph).clamp(min=0) ** o.s_phi > o.delta_ls * th0 ** o.s_theta)
                armijo = ph_t <= ph0 + o.eta_phi * al * dph
                suff = (th_t <= (1 - o.gamma_theta) * th0) | (ph_t <= ph0 - o.gamma_phi * th0)
                ok = okfin & ~in_filter & (th_t <= theta_max[rows][:, None]) & torch.where(sw, armijo, suff)
                if any_resto:
                    if glue is None:
                        c_t = torch.where(eq[rows][:, None, :], g_t - ceq[rows][:, None, :], g_t - st_)
                        bar_t = ph_t - f_t                                                        # barrier terms only
                        phR_t = 0.5 * o.resto_rho * (c_t * c_t).sum(2) + 0.5 * zeta1[rows][:, None] * (
                            r_(DR2) * (xt - r_(x_R)) ** 2).sum(2) + bar_t
                    okR = torch.isfinite(phR_t) & (phR_t <= phiR[rows][:, None] + 1e-4 * al * dphiR[rows][:, None])
                    ok = torch.where(resto[rows][:, None], okR, ok)
                # never accept a candidate beyond the max_ls-th halving
                ok = ok & ((ls + torch.arange(Kw, device=dev)) < o.max_ls)[None, :]
                if ls == 0 and soc_on:
                    # ---- second-order correction (Waechter & Biegler 2.4; IPOPT max_soc = 4): the full step was rejected
                    # and the constraint violation did not drop -> correct the step with the constraint values at the trial
                    # point, re-solving with the factors of this sweep (rb_kkt_resolve_rows), before any backtracking
                    soc_m = (~ok[:, 0]) & okfin[:, 0] & (th_t[:, 0] >= th0[:, 0]) & ~resto[rows] & (fac_slot[rows] >= 0)
                    if bool(soc_m.any()):
                        r_soc = rows[soc_m]
                        a0 = al[soc_m, 0]
                        iq_s, eq_s = ineq[r_soc], eq[r_soc]
                        g_t0 = evt['g'].reshape(ns, Kw, m)[soc_m, 0]
                        c_old = c[r_soc]
                        c_t0 = torch.where(eq_s, g_t0 - ceq[r_soc], g_t0 - (s[r_soc] + a0[:, None] * ds[r_soc]))
                        c_soc = a0[:, None] * c_old + c_t0
                        th_old = th_t[soc_m, 0].clone()
                        sw0 = sw[soc_m, 0]
                        alive = torch.ones(r_soc.numel(), dtype=torch.bool, device=dev)
                        got = torch.zeros_like(alive)
                        dx_keep = torch.zeros(r_soc.numel(), n, dtype=dt, device=dev)
                        ds_keep = torch.zeros(r_soc.numel(), m, dtype=dt, device=dev)
                        a_keep = torch.zeros(r_soc.numel(), dtype=dt, device=dev)
                        for _soc in range(o.max_soc):
                            la = torch.nonzero(alive).squeeze(1)
                            if la.numel() == 0:
                                break
                            rr_ = r_soc[la]
                            Ssr_ = NW['Ssr'][rr_]
                            rs_ = r_s[rr_]
                            rhs_soc = torch.cat([rhs[rr_, :n], torch.where(ineq[rr_], -c_soc[la] - rs_ / Ssr_, -c_soc[la])], dim=1)
                            t0 = time.perf_counter()
                            sol_s = be.kkt_resolve_slots(fac_slot[rr_].to(torch.int32).contiguous(), rhs_soc.contiguous())
                            res.t_kkt += time.perf_counter() - t0
                            dx_s = sol_s[:, :n]
                            ds_s = torch.where(ineq[rr_], (sol_s[:, n:] - rs_) / Ssr_, torch.zeros_like(rs_))
                            # fraction-to-the-boundary rule for the corrected step
                            def ftb(d_, st_, f_):
                                ratio = torch.where((st_ < 0) & (f_ > 0), -tau[rr_][:, None] * d_ / st_, torch.full_like(d_, float('inf')))
                                return torch.clamp(ratio.amin(1), max=1.0) if ratio.shape[1] else torch.ones(rr_.numel(), dtype=dt, device=dev)
                            xr, sr = x[rr_], s[rr_]
                            a_s = torch.stack([ftb(xr - xL[rr_], dx_s, fL[rr_]), ftb(xU[rr_] - xr, -dx_s, fU[rr_]),
                                               ftb(sr - sL[rr_], ds_s, sfL[rr_]), ftb(sU[rr_] - sr, -ds_s, sfU[rr_])]).amin(0)
                            dx_f = torch.zeros_like(x)
                            ds_f = torch.zeros_like(s)
                            dx_f[rr_] = dx_s
                            ds_f[rr_] = ds_s
                            al1 = a_s[:, None].contiguous()
                            xt1, rows32s = glue.trial(x, dx_f, rr_, al1)
                            t0 = time.perf_counter()
                            evs = be.eval_points(xt1, rr_)
                            if dev.type == 'cuda':
                                torch.cuda.synchronize(dev)
                            res.t_eval += time.perf_counter() - t0
                            res.n_eval += 1
                            TMs = glue.trial_merit(GS(), rows32s, al1, xt1, ds_f, evs['f'].contiguous(), evs['g'].contiguous())
                            th_s, ph_s, fin_s = TMs[:, 0, 0], TMs[:, 0, 1], TMs[:, 0, 3] > 0.5
                            ft_, fp_ = filt_theta[rr_], filt_phi[rr_]
                            inf_s = ((th_s[:, None] >= (1 - o.gamma_theta) * ft_) & (ph_s[:, None] >= fp_ - o.gamma_phi * ft_)).any(1)
                            th0s, ph0s, dphs = theta[rr_], phi[rr_], dphi[rr_]
                            arm_s = ph_s <= ph0s + o.eta_phi * a0[la] * dphs
                            suf_s = (th_s <= (1 - o.gamma_theta) * th0s) | (ph_s <= ph0s - o.gamma_phi * th0s)
                            ok_s = fin_s & ~inf_s & (th_s <= theta_max[rr_]) & torch.where(sw0[la], arm_s, suf_s)
                            acc = la[ok_s]
                            got[acc] = True
                            dx_keep[acc] = dx_s[ok_s]
                            ds_keep[acc] = ds_s[ok_s]
                            a_keep[acc] = a_s[ok_s]
                            alive[acc] = False
                            # not accepted: give up unless the violation keeps dropping, else correct again
                            rej = ~ok_s
                            stop = rej & (~fin_s | (th_s > o.kappa_soc * th_old[la]))
                            alive[la[stop]] = False
                            cont = rej & ~stop
                            if bool(cont.any()):
                                lc_ = la[cont]
                                st1 = torch.where(ineq[rr_], sr + a_s[:, None] * ds_s, ceq[rr_])
                                c_t1 = torch.where(eq[rr_], evs['g'] - ceq[rr_], evs['g'] - st1)
                                c_soc[lc_] = a_s[cont][:, None] * c_soc[lc_] + c_t1[cont]
                                th_old[lc_] = th_s[cont]
                            res.n_soc += int(ok_s.sum())
                        if bool(got.any()):
                            rg = r_soc[got]
                            # the corrected step replaces the direction of these instances (primal part; the duals keep theirs)
                            DIR['dx'][rg] = dx_keep[got]
                            DIR['ds'][rg] = ds_keep[got]
                            dx, ds = DIR['dx'], DIR['ds']
                            accepted_alpha[rg] = a_keep[got]
                            augment[rg] = ~sw0[got]       # (an Armijo-accepted switching step does not augment the filter)
                            searching[rg] = False
                            keep_rows = ~torch.isin(rows, rg)
                            rows, al, ok = rows[keep_rows], al[keep_rows], ok[keep_rows]
                            sw, armijo = sw[keep_rows], armijo[keep_rows]
                any_ok = ok.any(1)
                first = torch.argmax(ok.to(torch.int8), dim=1)                                   # first admissible halving
                pick = first[:, None]
                acc_rows = rows[any_ok]
                accepted_alpha[acc_rows] = al.gather(1, pick).squeeze(1)[any_ok]
                augment[acc_rows] = ~(sw & armijo).gather(1, pick).squeeze(1)[any_ok]
                searching[acc_rows] = False
                rej_rows = rows[~any_ok]
                last_alpha[rej_rows] = al[~any_ok, Kw - 1]
                alpha[rej_rows] = al[~any_ok, Kw - 1] * 0.5
                ls += Kw
            failed = searching
            if bool(failed.any()):
                # A failed search of the filter method far from feasibility starts the restoration at this point (no
                # step now); a failed search inside the restoration ends the instance.  Close to feasibility (theta below
                # theta_min, where the filter only asks for descent) the search usually fails on rounding noise: the
                # filter is reset and the shortest trial step taken, three times in a row at most.
                enter = failed & ~resto & (theta > theta_min) if o.restoration else torch.zeros_like(failed)
                status[failed & resto] = 3
                if bool(enter.any()):
                    # the current point joins the filter, so that the method does not return here
                    idx_f = torch.clamp(filt_n, max=F - 1)
                    rows_e = torch.nonzero(enter).squeeze(1)
                    filt_theta[rows_e, idx_f[rows_e]] = ((1 - o.gamma_theta) * theta)[rows_e]
                    filt_phi[rows_e, idx_f[rows_e]] = (phi - o.gamma_phi * theta)[rows_e]
                    filt_n[rows_e] = torch.clamp(filt_n[rows_e] + 1, max=F - 1)
                    resto = resto | enter
                    resto_it = torch.where(enter, torch.zeros_like(resto_it), resto_it)
                    theta_R = torch.where(enter, theta, theta_R)
                    x_R = torch.where(enter[:, None], x, x_R)
                    DR2 = torch.where(enter[:, None], 1.0 / torch.clamp(x.abs(), min=1.0) ** 2, DR2)
                    n_resto = n_resto + enter.long()
                    res.n_restorations += int(enter.sum())
                soft = failed & ~resto
                ls_fail = torch.where(soft, ls_fail + 1, ls_fail)
                reset_filter(soft)
                accepted_alpha = torch.where(soft, last_alpha, accepted_alpha)
                augment = augment & ~failed
                status[soft & (ls_fail >= 3)] = 3
            ls_fail = torch.where(moved & ~failed, torch.zeros_like(ls_fail), ls_fail)
            was_resto = resto & moved & ~failed                   # took a restoration step in this sweep
            augment = augment & ~resto

            # ---- filter augmentation, step, multiplier reset ----------------------------------------------
            if bool(augment.any()):
                idx = torch.clamp(filt_n, max=F - 1)
                rows = torch.nonzero(augment).squeeze(1)
                filt_theta[rows, idx[rows]] = ((1 - o.gamma_theta) * theta)[rows]
                filt_phi[rows, idx[rows]] = (phi - o.gamma_phi * theta)[rows]
                filt_n[rows] = torch.clamp(filt_n[rows] + 1, max=F - 1)
            if glue is not None:
                # (x_R is a separate tensor: the in-place update of x does not touch the restoration reference)
                glue.update(GS(), accepted_alpha.contiguous(), (a_du * moved).contiguous(), DIR, o.kappa_sigma)
            else:
                a = accepted_alpha[:, None]
                x = x + a * dx
                s = torch.where(ineq, s + a * ds, ceq)
                y = y + a * dy
                ad = (a_du * moved)[:, None]
                zL, zU, vL, vU = zL + ad * dzL, zU + ad * dzU, vL + ad * dvL, vU + ad * dvU
                dL, dU, eL, eU = slacks(x, s)

                def reset(z_, d_, f_):
                    d_ = torch.where(f_ > 0, d_, torch.ones_like(d_))
                    lo, hi = mu_c / (o.kappa_sigma * d_), o.kappa_sigma * mu_c / d_
                    return torch.where(f_ > 0, torch.maximum(torch.minimum(z_, hi), lo), z_)

                zL, zU, vL, vU = reset(zL, dL, fL), reset(zU, dU, fU), reset(vL, eL, sfL), reset(vU, eU, sfU)
            resto_it = resto_it + was_resto.long()
            status[(resto_it >= o.resto_max_iter) & resto & (status == -1)] = 3
            iters = iters + moved.long()
            it += 1
            status[(status == -1) & (iters >= o.max_iter)] = 2
            active = status == -1
            if o.verbose and any_resto and bool(resto[0] | was_resto[0]):
                print(f'      [restoration] phi_R={float(phiR[0]):.6e} slope={float(dphiR[0]):.3e} theta_R={float(theta_R[0]):.3e} '
                      f'it={int(resto_it[0])}')
            if o.verbose:
                print(f'      alpha_pr={float(accepted_alpha[0]):.3e} a_max={float(a_pr[0]):.3e} alpha_du={float(a_du[0]):.3e} '
                      f'ls<={ls} failed={bool(failed[0])} theta={float(theta[0]):.3e} dphi={float(dphi[0]):.3e} '
                      f'|dx|={float(dx[0].abs().max()):.2e} |dy|={float(dy[0].abs().max()):.2e} stalled={bool(stalled[0])}')
            moved_prev = moved
            # instances that just ended with "maximum iterations" / "line search failed" took the step too: their
            # returned f, g and kkt_error belong to the returned x
            if bool(moved.any()):
                ev = evaluate(x, y, True, moved, ev)

        status[status < 0] = 2
        if OUT:
            stash(torch.arange(B, device=dev))
            res.x, res.f, res.g, res.lam_g, res.lam_x = OUT['x'], OUT['f'], OUT['g'], OUT['lam_g'], OUT['lam_x']
            res.status, res.iterations, res.kkt_error = OUT['status'], OUT['iters'], OUT['err']
            res.factorisations_each = OUT['n_fact'].cpu().numpy()
            be.select(None)
        else:
            E0 = errors(ev, x, s, y, zL, zU, vL, vU, torch.zeros_like(mu))[0]
            # ipopt.honor_original_bounds = 'yes' (drone3d/raceline/base_raceline.py:788): the returned point lies inside
            # the original bounds, not the relaxed ones the iterations work with
            res.x, res.f, res.g, res.lam_g = torch.minimum(torch.maximum(x, lbx), ubx), ev['f'], ev['g'], y
            res.lam_x = zU - zL
            res.status, res.iterations, res.kkt_error = status, iters, E0
            res.factorisations_each = n_fact.cpu().numpy()
        res.success = res.status <= 1
        res.n_iter = it
        res.t_total = time.perf_counter() - t_start
        return res


class CudaBackend:
    ''' eval through rb_eval_batch, KKT through rb_kkt_* (device tensors only) '''

    def __init__(self, functions, vp, kkt_solver=None):
        from .kkt import KktSolver
        self.F = functions
        self.st = functions.st
        self.ng = self.st.ng
        self.vp = vp                                   # device tensor (nvp,) or (B, nvp)
        self._vp_full = vp
        self.K = kkt_solver or KktSolver(self.st)
        self.refine_tol = 1e-10
        self._buf = {}
        self._glue = None

    @property
    def glue(self):
        ''' fused kernels for the vector work of the interior-point sweep (glue.py) '''
        if self._glue is None:
            from .glue import IpmGlue
            d = IpmOptions()
            self._glue = IpmGlue(d.kappa_d, d.resto_rho)
        return self._glue

    def _out(self, name, shape, dev):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.device != dev:
            t = torch.empty(shape, dtype=torch.float64, device=dev)
            self._buf[name] = t
        return t

    def select(self, keep):
        ''' the driver compacted its state to the instances `keep` (None: back to the full batch) '''
        if keep is None:
            self.vp = self._vp_full
        elif self.vp.dim() == 2:
            if getattr(self, '_vp_full', None) is None:
                self._vp_full = self.vp
            self.vp = self.vp[keep].contiguous()

    def _vp_rows(self, idx):
        return self.vp if (idx is None or self.vp.dim() == 1) else self.vp[idx].contiguous()

    def eval(self, x, lam_g, lam_f, derivs, idx=None, prev=None):
        """ idx: evaluate only these instances; the other rows keep the values of `prev` (a dict returned by an earlier
        call, possibly row-compacted by the caller since), which is updated in place and returned """
        st, dev, B = self.st, x.device, x.shape[0]
        f64 = dict(dtype=torch.float64, device=dev)
        names = ('f', 'grad_f', 'g', 'jac', 'hess') if derivs else ('f', 'g')
        shapes = dict(f=(B,), grad_f=(B, st.nw), g=(B, st.ng), jac=(B, st.nnz_jac), hess=(B, st.nnz_hess))
        if idx is not None and prev is None:
            raise ValueError('a masked evaluation needs the previous evaluation to update')
        if prev is not None:
            full = {k: prev[k] for k in names}
            assert all(tuple(full[k].shape) == shapes[k] for k in names)
        else:
            full = {k: torch.empty(shapes[k], **f64) for k in names}
        if idx is None:
            xc, lc, out = x.contiguous(), (lam_g.contiguous() if derivs else None), full
        else:
            Bc = idx.numel()
            xc, lc = x[idx].contiguous(), (lam_g[idx].contiguous() if derivs else None)
            out = {k: torch.empty((Bc,) + shapes[k][1:], **f64) for k in names}
        lf = None if not derivs else (lam_f if idx is None else lam_f[idx].contiguous())
        if xc.shape[0] > 0:
            self._scratch = self.F.eval_device(xc, lc, lf, self._vp_rows(idx), None, out['f'], out.get('grad_f'),
                                               out['g'],
                                               out.get('jac'), out.get('hess'), None)
        if idx is not None:
            for k in names:
                full[k].index_copy_(0, idx, out[k])
        return dict(full)

    def eval_points(self, x_rows, inst):
        """ f and g at arbitrary trial points; inst[r] = instance (vehicle parameters) of row r """
        st, dev, R = self.st, x_rows.device, x_rows.shape[0]
        f = torch.empty(R, dtype=torch.float64, device=dev)
        g = torch.empty(R, st.ng, dtype=torch.float64, device=dev)
        vp = self.vp if self.vp.dim() == 1 else self.vp[inst].contiguous()
        if R > 0:
            self.F.eval_device(x_rows.contiguous(), None, None, vp, None, f, None, g, None, None, None)
        return dict(f=f, g=g)

    def kkt_matvec(self, hess, jac, dx_diag, neg_d, vec):
        return self.K.matvec(hess, jac, dx_diag.contiguous(), neg_d.contiguous(), vec.contiguous())

    @property
    def kkt_wave(self):
        ''' instances one launch of the factorisation kernel holds at once (CTAs per SM x SMs): a launch with fewer
        instances takes as long, so the driver fills the spare slots with speculative regularisation candidates '''
        sms = torch.cuda.get_device_properties(self.vp.device).multi_processor_count
        if getattr(self.K, 'cs', None) is not None:
            # condensed collocation intervals: the interior kernel runs one CTA per (interval, instance), one CTA per SM
            return max(1, sms // self.K.cs.NI)
        return sms * (3 if self.K.ks.bmax <= 64 else 1)

    @property
    def can_soc(self):
        ''' further right-hand sides for single instances of the last factorisation (second-order correction) '''
        return bool(getattr(self.K, 'can_resolve_rows', False))

    def kkt_resolve_slots(self, slots, rhs):
        ''' solve with the factors of the last kkt_solve / kkt_solve_rows call: rhs row p <-> factor slot slots[p] '''
        return self.K.resolve_rows(slots, rhs)

    def kkt_solve_rows(self, hess, jac, idx, dx_diag, neg_d, rhs, refine_steps):
        ''' rows `idx` of hess / jac (repeats allowed) with dx_diag / neg_d / rhs given row by row '''
        return self._kkt_core(hess[idx], jac[idx], dx_diag, neg_d, rhs, refine_steps)

    def _kkt_core(self, hess, jac, dx_diag, neg_d, rhs, refine_steps):
        dx_diag, neg_d, rhs = dx_diag.contiguous(), neg_d.contiguous(), rhs.contiguous()
        sol, status = self.K.factor_solve(hess, jac, dx_diag, neg_d, rhs)
        sol = sol.clone()
        scale = torch.clamp(rhs.abs().amax(1), min=1.0)
        for _ in range(refine_steps):
            r = rhs - self.K.matvec(hess, jac, dx_diag, neg_d, sol)
            # iterative refinement only while some instance is above the fp64 noise floor of its right-hand side;
            # the correction is applied instance by instance, so a result does not depend on the rest of the batch
            need = (r.abs().amax(1) / scale) > self.refine_tol
            n_need = int(need.sum())
            if n_need == 0:
                break
            if n_need <= rhs.shape[0] // 2 and getattr(self.K, 'can_resolve_rows', False):
                # the re-solve streams every instance's factors from HBM: run it for the instances that need it only
                idx = torch.nonzero(need).squeeze(1)
                dsol = self.K.resolve_rows(idx.to(torch.int32), r.index_select(0, idx).contiguous())
                sol.index_add_(0, idx, dsol)
            else:
                sol = torch.where(need[:, None], sol + self.K.resolve(hess, jac, dx_diag, neg_d, r), sol)
        return sol, status

    def kkt_solve(self, hess, jac, dx_diag, neg_d, rhs, refine_steps, idx=None):
        B = rhs.shape[0]
        if idx is not None:
            hess, jac, dx_diag, neg_d, rhs = hess[idx], jac[idx], dx_diag[idx], neg_d[idx], rhs[idx]
        sol, status = self._kkt_core(hess, jac, dx_diag, neg_d, rhs, refine_steps)
        if idx is None:
            return sol, status
        sol_f = torch.zeros(B, sol.shape[1], dtype=sol.dtype, device=sol.device)
        st_f = torch.zeros(B, 2, dtype=status.dtype, device=sol.device)
        st_f[:, 1] = self.ng
        sol_f.index_copy_(0, idx, sol)
        st_f.index_copy_(0, idx, status)
        return sol_f, st_f
