'''
ctypes binding of the fused interior-point kernels (csrc/ipm_glue.cuh, C ABI `rb_ipm_*` in include/raceline_b200.h).

The batched interior-point driver (ipm.py) keeps its state in torch CUDA tensors; with this object as `backend.glue`
the element-wise and reduction work of a sweep runs in six fused kernels instead of a few hundred torch launches.
The torch arithmetic in ipm.py stays as the CPU-testable statement of the same formulas (tests/test_gpu_ipm.py checks
that both paths give the same iterates).  Device tensors only; no CPU fallback in here.
'''
import ctypes

import torch

_vp = ctypes.c_void_p


class _IpmArgs(ctypes.Structure):
    _fields_ = ([('B', ctypes.c_int), ('n', ctypes.c_int), ('m', ctypes.c_int), ('sx', ctypes.c_longlong),
                 ('ss', ctypes.c_longlong)]
                + [(k, _vp) for k in ('xL', 'xU', 'sL', 'sU', 'ceq', 'xflag', 'sflag', 'x', 's', 'y', 'zL', 'zU', 'vL', 'vU',
                                      'grad_f', 'g', 'jty', 'mu', 'delta_w', 'delta_c', 'resto', 'x_R', 'DR2')]
                + [('kappa_d', ctypes.c_double), ('rho', ctypes.c_double)])


def _p(t):
    return None if t is None else _vp(t.data_ptr())


def _chk(t, dtype=torch.float64):
    assert t.is_cuda and t.is_contiguous() and t.dtype == dtype, (t.dtype, t.is_contiguous())
    return t


class IpmGlue:
    def __init__(self, kappa_d, rho):
        from .functions import load_library, _check
        self.lib, self._check = load_library(), _check
        self.kappa_d, self.rho = float(kappa_d), float(rho)
        lib = self.lib
        A = ctypes.POINTER(_IpmArgs)
        lib.rb_ipm_error.argtypes = [A, _vp, _vp]
        lib.rb_ipm_newton.argtypes = [A] + [_vp] * 11
        lib.rb_ipm_direction.argtypes = [A] + [_vp] * 17
        lib.rb_ipm_trial.argtypes = [ctypes.c_int] * 3 + [_vp] * 6
        lib.rb_ipm_trial_merit.argtypes = [A, ctypes.c_int, ctypes.c_int] + [_vp] * 8
        lib.rb_ipm_update.argtypes = [A] + [_vp] * 9 + [ctypes.c_double, _vp]
        self._buf = {}

    # ---- helpers -------------------------------------------------------------------------------------------
    def buf(self, name, shape, dev, dtype=torch.float64):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.device != dev:
            t = torch.empty(shape, dtype=dtype, device=dev)
            self._buf[name] = t
        return t

    def args(self, S, ev=None, jty=None):
        ''' S: dict of the driver's current state tensors '''
        a = _IpmArgs()
        x = _chk(S['x'])
        a.B, a.n = x.shape
        a.m = S['s'].shape[1]
        a.sx, a.ss = a.n, a.m
        for k in ('xL', 'xU', 'sL', 'sU', 'ceq', 'x', 's', 'y', 'zL', 'zU', 'vL', 'vU', 'mu', 'delta_w', 'delta_c'):
            setattr(a, k, _p(_chk(S[k])))
        a.xflag, a.sflag = _p(_chk(S['xflag'], torch.uint8)), _p(_chk(S['sflag'], torch.uint8))
        if ev is not None:
            a.grad_f, a.g = _p(_chk(ev['grad_f'])), _p(_chk(ev['g']))
        if jty is not None:
            a.jty = _p(_chk(jty))
        if S.get('resto') is not None:
            a.resto, a.x_R, a.DR2 = _p(_chk(S['resto'], torch.uint8)), _p(_chk(S['x_R'])), _p(_chk(S['DR2']))
        a.kappa_d, a.rho = self.kappa_d, self.rho
        self._keep = (S, ev, jty)
        return a

    @staticmethod
    def _stream(t):
        return _vp(torch.cuda.current_stream(t.device).cuda_stream)

    # ---- kernels -------------------------------------------------------------------------------------------
    def error(self, S, ev, jty):
        a = self.args(S, ev, jty)
        out = self.buf('err', (a.B, 8), S['x'].device)
        self._check(self.lib.rb_ipm_error(ctypes.byref(a), _p(out), self._stream(out)), 'rb_ipm_error')
        return out

    def newton(self, S, ev, jty):
        a = self.args(S, ev, jty)
        dev, B, n, m = S['x'].device, a.B, a.n, a.m
        o = dict(dxd=self.buf('dxd', (B, n), dev), negd=self.buf('negd', (B, m), dev), rhs=self.buf('rhs', (B, n + m), dev),
                 gphi_x=self.buf('gphi_x', (B, n), dev), gphi_s=self.buf('gphi_s', (B, m), dev), c=self.buf('c', (B, m), dev),
                 r_s=self.buf('r_s', (B, m), dev), Ssr=self.buf('Ssr', (B, m), dev), sc=self.buf('nsc', (B, 4), dev))
        self._check(self.lib.rb_ipm_newton(ctypes.byref(a), _p(_chk(ev['f'])), _p(o['dxd']), _p(o['negd']), _p(o['rhs']),
                                           _p(o['gphi_x']), _p(o['gphi_s']), _p(o['c']), _p(o['r_s']), _p(o['Ssr']),
                                           _p(o['sc']), self._stream(o['sc'])), 'rb_ipm_newton')
        return o

    def direction(self, S, sol, moved, tau, nw):
        a = self.args(S)
        dev, B, n, m = S['x'].device, a.B, a.n, a.m
        d = {k: self.buf(k, (B, n if k in ('dx', 'dzL', 'dzU') else m), dev) for k in ('dx', 'dy', 'ds', 'dzL', 'dzU', 'dvL', 'dvU')}
        sc = self.buf('dsc', (B, 4), dev)
        mv = moved.to(torch.uint8).contiguous()
        self._check(self.lib.rb_ipm_direction(ctypes.byref(a), _p(_chk(sol)), _p(mv), _p(_chk(tau)), _p(nw['gphi_x']),
                                              _p(nw['gphi_s']), _p(nw['c']), _p(nw['r_s']), _p(nw['Ssr']), _p(d['dx']),
                                              _p(d['dy']), _p(d['ds']), _p(d['dzL']), _p(d['dzU']), _p(d['dvL']), _p(d['dvU']),
                                              _p(sc), self._stream(sc)), 'rb_ipm_direction')
        d['sc'] = sc
        return d

    def trial(self, x, dx, rows, al):
        ns, Kw = al.shape
        n = x.shape[1]
        xt = torch.empty(ns * Kw, n, dtype=torch.float64, device=x.device)
        rows32 = rows.to(torch.int32).contiguous()
        self._check(self.lib.rb_ipm_trial(n, Kw, ns, _p(rows32), _p(_chk(al)), _p(_chk(x)), _p(_chk(dx)), _p(xt),
                                          self._stream(xt)), 'rb_ipm_trial')
        return xt, rows32

    def trial_merit(self, S, rows32, al, xt, ds, f_t, g_t):
        a = self.args(S)
        ns, Kw = al.shape
        out = torch.empty(ns * Kw, 4, dtype=torch.float64, device=xt.device)
        self._check(self.lib.rb_ipm_trial_merit(ctypes.byref(a), Kw, ns, _p(rows32), _p(_chk(al)), _p(_chk(xt)), _p(_chk(ds)),
                                                _p(_chk(f_t)), _p(_chk(g_t)), _p(out), self._stream(out)), 'rb_ipm_trial_merit')
        return out.view(ns, Kw, 4)

    def update(self, S, alpha, alpha_du, d, kappa_sigma):
        a = self.args(S)
        self._check(self.lib.rb_ipm_update(ctypes.byref(a), _p(_chk(alpha)), _p(_chk(alpha_du)), _p(d['dx']), _p(d['dy']),
                                           _p(d['ds']), _p(d['dzL']), _p(d['dzU']), _p(d['dvL']), _p(d['dvU']),
                                           float(kappa_sigma), self._stream(alpha)), 'rb_ipm_update')
