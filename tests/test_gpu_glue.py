'''
GPU unit tests of the fused interior-point kernels (csrc/ipm_glue.cuh through the C ABI rb_ipm_*) against a plain torch
statement of the same formulas (those of ipm.py) on random states: bounds of every kind (none, lower, upper, both),
equality and inequality rows, instances in and out of the feasibility restoration.  fp64, tolerance 1e-12 relative.
'''
import numpy as np
import pytest
import torch


def _state(B=5, n=61, m=47, seed=0, resto=(False, True, False, False, True)):
    g_ = torch.Generator().manual_seed(seed)
    R = lambda *s: torch.rand(*s, generator=g_, dtype=torch.float64)
    N = lambda *s: torch.randn(*s, generator=g_, dtype=torch.float64)
    kx = torch.randint(0, 4, (B, n), generator=g_)          # 0 none, 1 lower, 2 upper, 3 both
    ks = torch.randint(0, 4, (B, m), generator=g_)          # 0 equality, 1 lower, 2 upper, 3 both
    xL, xU = -1 - R(B, n), 1 + R(B, n)
    x = xL + (xU - xL) * (0.05 + 0.9 * R(B, n))
    sL, sU = -1 - R(B, m), 1 + R(B, m)
    s = sL + (sU - sL) * (0.05 + 0.9 * R(B, m))
    hasL, hasU = (kx == 1) | (kx == 3), (kx == 2) | (kx == 3)
    eq = ks == 0
    shasL, shasU = (ks == 1) | (ks == 3), (ks == 2) | (ks == 3)
    S = dict(x=x, s=torch.where(eq, R(B, m), s), y=N(B, m), zL=R(B, n) * hasL, zU=R(B, n) * hasU, vL=R(B, m) * shasL,
             vU=R(B, m) * shasU, xL=xL * hasL, xU=xU * hasU, sL=sL * shasL, sU=sU * shasU, ceq=torch.where(eq, N(B, m), torch.zeros(B, m, dtype=torch.float64)),
             xflag=(hasL.to(torch.uint8) | (hasU.to(torch.uint8) << 1)),
             sflag=(shasL.to(torch.uint8) | (shasU.to(torch.uint8) << 1) | (eq.to(torch.uint8) << 2)),
             mu=0.01 + R(B), delta_w=R(B) * 0.1, delta_c=R(B) * 1e-3, resto=torch.tensor(resto[:B]).to(torch.uint8),
             x_R=x + 0.1 * N(B, n), DR2=0.5 + R(B, n))
    S['s'] = torch.where(eq, S['ceq'], S['s'])
    ev = dict(f=N(B), grad_f=N(B, n), g=N(B, m))
    jty = N(B, n)
    flags = dict(hasL=hasL, hasU=hasU, eq=eq, shasL=shasL, shasU=shasU)
    return S, ev, jty, flags


def _cuda(d):
    return {k: (v.cuda().contiguous() if torch.is_tensor(v) else v) for k, v in d.items()}


def _ref(S, ev, jty, F, kd, rho):
    ''' the formulas of ipm.py in plain torch (fp64) '''
    dt = torch.float64
    fL, fU, sfL, sfU = (F[k].to(dt) for k in ('hasL', 'hasU', 'shasL', 'shasU'))
    eq, ineq = F['eq'], ~F['eq']
    dampL, dampU = (F['hasL'] & ~F['hasU']).to(dt), (F['hasU'] & ~F['hasL']).to(dt)
    sdL, sdU = (F['shasL'] & ~F['shasU']).to(dt), (F['shasU'] & ~F['shasL']).to(dt)
    x, s, y, mu = S['x'], S['s'], S['y'], S['mu'][:, None]
    dL, dU, eL, eU = x - S['xL'], S['xU'] - x, s - S['sL'], S['sU'] - s
    one = torch.ones_like
    inv = lambda d_, f_: f_ / torch.where(f_ > 0, d_, one(d_))
    iL, iU, jL, jU = inv(dL, fL), inv(dU, fU), inv(eL, sfL), inv(eU, sfU)
    c = torch.where(eq, ev['g'] - S['ceq'], ev['g'] - s)
    gradL_x = ev['grad_f'] + jty - S['zL'] + S['zU']
    gradL_s = torch.where(ineq, -y - S['vL'] + S['vU'], torch.zeros_like(y))
    prods = torch.cat([(dL * S['zL'])[F['hasL']], (dU * S['zU'])[F['hasU']]]) if False else None
    out = {}
    P = [torch.where(F['hasL'], dL * S['zL'], torch.full_like(x, float('nan'))), torch.where(F['hasU'], dU * S['zU'], torch.full_like(x, float('nan'))),
         torch.where(F['shasL'], eL * S['vL'], torch.full_like(s, float('nan'))), torch.where(F['shasU'], eU * S['vU'], torch.full_like(s, float('nan')))]
    Pcat = torch.cat(P, dim=1)
    out['err'] = torch.stack([torch.maximum(gradL_x.abs().amax(1), gradL_s.abs().amax(1)), c.abs().amax(1),
                              (S['zL'] * fL + S['zU'] * fU).sum(1) + (S['vL'] * sfL + S['vU'] * sfU).sum(1), y.abs().sum(1),
                              torch.where(torch.isnan(Pcat), torch.full_like(Pcat, float('inf')), Pcat).amin(1),
                              torch.where(torch.isnan(Pcat), torch.full_like(Pcat, -float('inf')), Pcat).amax(1),
                              c.abs().sum(1)], dim=1)
    R = S['resto'].bool()[:, None]
    zeta = torch.sqrt(mu)
    Sx = S['zL'] * iL + S['zU'] * iU
    Ss = S['vL'] * jL + S['vU'] * jU
    gbx = -mu * iL + mu * iU + kd * mu * (dampL - dampU)
    gbs = torch.where(ineq, -mu * jL + mu * jU + kd * mu * (sdL - sdU), torch.zeros_like(s))
    gphi_x = ev['grad_f'] + gbx
    r_x = gphi_x + jty
    r_s = torch.where(ineq, gbs - y, torch.zeros_like(s))
    dw, dc = S['delta_w'][:, None], S['delta_c'][:, None]
    Ssr = torch.where(ineq, Ss + dw, one(s))
    SsR = torch.where(ineq, mu * (jL * jL + jU * jU), one(s))
    out['dxd'] = torch.where(R, zeta * S['DR2'] + mu * (iL * iL + iU * iU), Sx + dw)
    out['negd'] = torch.where(R, torch.where(ineq, -(1 / rho + 1 / SsR), torch.full_like(s, -1 / rho)),
                              torch.where(ineq, -1 / Ssr, torch.zeros_like(s)) - dc)
    out['rhs'] = torch.cat([torch.where(R, -(zeta * S['DR2'] * (x - S['x_R']) + gbx), -r_x),
                            torch.where(R, torch.where(ineq, -c - gbs / SsR, -c), torch.where(ineq, -c - r_s / Ssr, -c))], dim=1)
    lg = lambda d_, f_: (torch.log(torch.where(f_ > 0, d_, one(d_))) * f_).sum(1)
    bar = -S['mu'] * (lg(dL, fL) + lg(dU, fU) + lg(eL, sfL) + lg(eU, sfU)) + kd * S['mu'] * (
        (dL * dampL).sum(1) + (dU * dampU).sum(1) + (eL * sdL).sum(1) + (eU * sdU).sum(1))
    out['theta'], out['phi'] = c.abs().sum(1), ev['f'] + bar
    out['phiR'] = 0.5 * rho * (c * c).sum(1) + 0.5 * zeta[:, 0] * (S['DR2'] * (x - S['x_R']) ** 2).sum(1) + bar
    out.update(c=c, gbx=gbx, gbs=gbs, gphi_x=gphi_x, r_s=r_s, Ssr=Ssr, SsR=SsR, iL=iL, iU=iU, jL=jL, jU=jU, dL=dL, dU=dU, eL=eL,
               eU=eU, fL=fL, fU=fU, sfL=sfL, sfU=sfU, ineq=ineq, R=R, zeta=zeta)
    return out


def _close(a, b, tol=1e-12):
    a, b = a.detach().cpu(), b.detach().cpu()
    scale = torch.clamp(b.abs(), min=1.0)
    bad = ~(torch.isfinite(a) == torch.isfinite(b))
    return (not bool(bad.any())) and float((torch.where(torch.isfinite(b), (a - b).abs() / scale, torch.zeros_like(b))).max()) <= tol


@pytest.mark.gpu
def test_fused_kernels_against_torch_formulas(built_library):
    from aircraft_trajectory_optimization_b200.glue import IpmGlue
    kd, rho, ks = 1e-4, 1e4, 1e10
    G = IpmGlue(kd, rho)
    S, ev, jty, F = _state()
    ref = _ref(S, ev, jty, F, kd, rho)
    Sc, evc, jtyc = _cuda(S), _cuda(ev), jty.cuda()
    E = G.error(Sc, evc, jtyc)
    assert _close(E[:, :7], ref['err'])
    NW = G.newton(Sc, evc, jtyc)
    for k in ('dxd', 'negd', 'rhs'):
        assert _close(NW[k], ref[k]), k
    assert _close(NW['sc'][:, 0], ref['theta']) and _close(NW['sc'][:, 1], ref['phi'])
    R1 = S['resto'].bool()
    assert _close(NW['sc'][R1, 2], ref['phiR'][R1])
    # direction
    B, n = S['x'].shape
    m = S['s'].shape[1]
    gg = torch.Generator().manual_seed(5)
    sol = torch.randn(B, n + m, generator=gg, dtype=torch.float64)
    moved = torch.tensor([True, True, False, True, True])
    tau = 0.99 + 0.005 * torch.rand(B, generator=gg, dtype=torch.float64)
    D = G.direction(Sc, sol.cuda(), moved.cuda(), tau.cuda(), NW)
    mv, R, ineq, mu = moved[:, None], ref['R'], ref['ineq'], S['mu'][:, None]
    dx = torch.where(mv, sol[:, :n], torch.zeros(B, n, dtype=torch.float64))
    w = torch.where(mv, sol[:, n:], torch.zeros(B, m, dtype=torch.float64))
    ds = torch.where(R, torch.where(mv & ineq, (w - ref['gbs']) / ref['SsR'], torch.zeros_like(w)),
                     torch.where(mv & ineq, (w - ref['r_s']) / ref['Ssr'], torch.zeros_like(w)))
    dy = torch.where(R, torch.zeros_like(w), w)
    keep = (~R).to(torch.float64)
    dzL = (mu * ref['iL'] - S['zL'] - S['zL'] * ref['iL'] * dx) * ref['fL'] * keep
    dzU = (mu * ref['iU'] - S['zU'] + S['zU'] * ref['iU'] * dx) * ref['fU'] * keep
    dvL = (mu * ref['jL'] - S['vL'] - S['vL'] * ref['jL'] * ds) * ref['sfL'] * keep
    dvU = (mu * ref['jU'] - S['vU'] + S['vU'] * ref['jU'] * ds) * ref['sfU'] * keep
    for k, v in (('dx', dx), ('dy', dy), ('ds', ds), ('dzL', dzL), ('dzU', dzU), ('dvL', dvL), ('dvU', dvU)):
        assert _close(D[k], v), k

    def max_step(d_, step, f_):
        ratio = torch.where((step < 0) & (f_ > 0), -tau[:, None] * d_ / step, torch.full_like(d_, float('inf')))
        return torch.clamp(ratio.amin(1), max=1.0)

    a_pr = torch.stack([max_step(ref['dL'], dx, ref['fL']), max_step(ref['dU'], -dx, ref['fU']),
                        max_step(ref['eL'], ds, ref['sfL']), max_step(ref['eU'], -ds, ref['sfU'])]).amin(0)
    a_du = torch.stack([max_step(S['zL'], dzL, ref['fL']), max_step(S['zU'], dzU, ref['fU']),
                        max_step(S['vL'], dvL, ref['sfL']), max_step(S['vU'], dvU, ref['sfU'])]).amin(0)
    assert _close(D['sc'][:, 0], a_pr) and _close(D['sc'][:, 1], a_du)
    nr = ~R1
    dphi = (ref['gphi_x'] * dx).sum(1) + (ref['gbs'] * ds).sum(1)
    assert _close(D['sc'][nr, 2], dphi[nr], 1e-11)
    dphiR = (ref['c'] * (w - rho * ref['c'])).sum(1) + (ref['zeta'] * S['DR2'] * (S['x'] - S['x_R']) * dx).sum(1) \
        + (ref['gbx'] * dx).sum(1) + (ref['gbs'] * ds).sum(1)
    assert _close(D['sc'][R1, 3], dphiR[R1], 1e-11)
    # trial points and their merits
    rows = torch.tensor([0, 1, 3, 4])
    Kw = 3
    al = (a_pr[rows][:, None] * 0.5 ** torch.arange(Kw, dtype=torch.float64)[None, :]).contiguous()
    xt, rows32 = G.trial(Sc['x'], D['dx'], rows.cuda(), al.cuda())
    xt_ref = S['x'][rows][:, None, :] + al[:, :, None] * dx[rows][:, None, :]
    assert _close(xt.view(len(rows), Kw, n), xt_ref)
    f_t = torch.randn(len(rows) * Kw, generator=gg, dtype=torch.float64)
    g_t = torch.randn(len(rows) * Kw, m, generator=gg, dtype=torch.float64)
    TM = G.trial_merit(Sc, rows32, al.cuda(), xt, D['ds'], f_t.cuda(), g_t.cuda())
    for r_i, b in enumerate(rows.tolist()):
        for k in range(Kw):
            St = {kk: (vv[b:b + 1] if torch.is_tensor(vv) and vv.dim() >= 1 and vv.shape[0] == B else vv) for kk, vv in S.items()}
            St['x'] = xt_ref[r_i, k][None]
            St['s'] = torch.where(F['eq'][b:b + 1], S['ceq'][b:b + 1], S['s'][b:b + 1] + al[r_i, k] * ds[b:b + 1])
            evt = dict(f=f_t[r_i * Kw + k][None], grad_f=ev['grad_f'][b:b + 1], g=g_t[r_i * Kw + k][None])
            rt = _ref(St, evt, jty[b:b + 1], {kk: vv[b:b + 1] for kk, vv in F.items()}, kd, rho)
            assert _close(TM[r_i, k, 0][None], rt['theta'], 1e-11) and _close(TM[r_i, k, 1][None], rt['phi'], 1e-11)
            if bool(S['resto'][b]):
                assert _close(TM[r_i, k, 2][None], rt['phiR'], 1e-11)
    # update
    alpha = torch.where(moved, a_pr * 0.5, torch.zeros_like(a_pr))
    ad = a_du * moved
    mu_c = S['mu'][:, None]
    x2 = S['x'] + alpha[:, None] * dx
    s2 = torch.where(ineq, S['s'] + alpha[:, None] * ds, S['ceq'])
    y2 = S['y'] + alpha[:, None] * dy

    def reset(z_, d_, f_):
        d_ = torch.where(f_ > 0, d_, torch.ones_like(d_))
        lo, hi = mu_c / (ks * d_), ks * mu_c / d_
        return torch.where(f_ > 0, torch.maximum(torch.minimum(z_, hi), lo), z_)

    zL2 = reset(S['zL'] + ad[:, None] * dzL, x2 - S['xL'], ref['fL'])
    zU2 = reset(S['zU'] + ad[:, None] * dzU, S['xU'] - x2, ref['fU'])
    vL2 = reset(S['vL'] + ad[:, None] * dvL, s2 - S['sL'], ref['sfL'])
    vU2 = reset(S['vU'] + ad[:, None] * dvU, S['sU'] - s2, ref['sfU'])
    G.update(Sc, alpha.cuda(), ad.cuda().contiguous(), D, ks)
    for k, v in (('x', x2), ('s', s2), ('y', y2), ('zL', zL2), ('zU', zU2), ('vL', vL2), ('vU', vU2)):
        assert _close(Sc[k], v), k
