'''
ctypes binding of libraceline_b200.so and CasADi-Function-like wrappers.

The reference obtains nlp_f / nlp_g / nlp_grad_f / nlp_jac_g / nlp_hess_l implicitly from
`ca.nlpsol` (drone3d/raceline/base_raceline.py:799).  `NlpFunctions` exposes the same five
functions over the CUDA library, with CasADi's calling convention and sparsity queries:

    F = nlp.nlp_jac_g ; g, J = F(x, p) ; F.sparsity_out(1).colind() / .row() / .nnz()

There is no CPU fallback: if the shared library or a CUDA device is missing the constructor raises.
'''
import ctypes
import os

import numpy as np

from .models import vehicle_params
from .structure import NLPStructure

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RACELINE_B200_LIB') or os.path.join(_HERE, 'libraceline_b200.so')   # override: kernel experiments
_LIB = None

_c_int_p = ctypes.POINTER(ctypes.c_int32)
_c_i64_p = ctypes.POINTER(ctypes.c_int64)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)


class _Desc(ctypes.Structure):
    _fields_ = [
        ('transcription', ctypes.c_int), ('variant', ctypes.c_char_p), ('N', ctypes.c_int), ('K', ctypes.c_int),
        ('nw', ctypes.c_int), ('ng', ctypes.c_int), ('nnz_jac', ctypes.c_int), ('nnz_hess', ctypes.c_int),
        ('R', _c_dbl_p), ('dR', _c_dbl_p), ('fc', _c_dbl_p),
        ('cell_row', _c_int_p), ('cell_coef', _c_dbl_p), ('cell_partner', _c_int_p), ('cell_pcoef', _c_dbl_p),
        ('cell_off', _c_dbl_p), ('cell_par', _c_dbl_p), ('cell_jslot', _c_int_p), ('cell_hslot', _c_int_p),
        ('cell_nj', ctypes.c_int), ('cell_nh', ctypes.c_int), ('cell_ncp', ctypes.c_int),
        ('tmpl_j', _c_int_p), ('tmpl_h', _c_int_p), ('n_tmpl_j', ctypes.c_int), ('n_tmpl_h', ctypes.c_int),
        ('colloc_C', _c_dbl_p), ('colloc_D', _c_dbl_p), ('colloc_B', _c_dbl_p),
        ('n_srow', ctypes.c_int), ('srow_row', _c_int_p), ('srow_kind', _c_int_p), ('srow_scale', _c_int_p),
        ('srow_var_ptr', _c_int_p), ('srow_var', _c_int_p), ('srow_jslot', _c_int_p),
        ('srow_form_ptr', _c_int_p), ('srow_coef_ptr', _c_int_p), ('srow_A', _c_dbl_p), ('srow_c', _c_dbl_p),
        ('n_shess', ctypes.c_int), ('shess_slot', _c_int_p), ('shess_add', _c_int_p), ('shess_ptr', _c_int_p),
        ('shess_row', _c_int_p), ('shess_coef', _c_dbl_p), ('shess_scale', _c_int_p),
        ('jac_colind', _c_i64_p), ('jac_row', _c_i64_p), ('hess_colind', _c_i64_p), ('hess_row', _c_i64_p),
    ]


class _TailDesc(ctypes.Structure):
    ''' rb_tail_desc (include/raceline_b200.h): tape of the end rows of an open track '''
    _fields_ = [('n_ins', ctypes.c_int), ('ins', _c_int_p), ('n_levels', ctypes.c_int * 3), ('lvl_ptr', _c_int_p),
                ('n_const', ctypes.c_int), ('cval', _c_dbl_p), ('n_slots', ctypes.c_int)]


EXPORTS = ['rb_last_error', 'rb_device_count', 'rb_set_device', 'rb_problem_create', 'rb_problem_destroy',
           'rb_problem_nvp', 'rb_problem_set_tail', 'rb_sparsity_size', 'rb_sparsity_get', 'rb_eval_scratch_bytes', 'rb_eval_batch',
           'rb_nlp_f', 'rb_nlp_g', 'rb_nlp_grad_f', 'rb_nlp_jac_g', 'rb_nlp_hess_l', 'rb_nlp_eval_all',
           'rb_launch_count', 'rb_profile_enable', 'rb_profile_cell_ms', 'rb_fp64_peak',
           'rb_kkt_create', 'rb_kkt_set_interiors', 'rb_kkt_destroy', 'rb_kkt_factor_bytes', 'rb_kkt_factor_solve', 'rb_kkt_resolve', 'rb_kkt_resolve_rows',
           'rb_kkt_matvec', 'rb_casadi_bind', 'rb_mesh_sdf', 'rb_ws_drone_guess', 'rb_centerline_frames', 'rb_traj_interp', 'rb_ipm_error', 'rb_ipm_newton',
           'rb_ipm_direction', 'rb_ipm_trial', 'rb_ipm_trial_merit', 'rb_ipm_update']
# CasADi generated-code shaped symbols (include/raceline_b200.h, RB_CASADI_DECLARE)
CASADI_FUNCTIONS = ['nlp_f', 'nlp_g', 'nlp_grad_f', 'nlp_jac_g', 'nlp_hess_l']
CASADI_SUFFIXES = ['', '_n_in', '_n_out', '_name_in', '_name_out', '_default_in', '_sparsity_in', '_sparsity_out',
                   '_work', '_alloc_mem', '_init_mem', '_free_mem', '_checkout', '_release', '_incref', '_decref']
EXPORTS += [f + sfx for f in CASADI_FUNCTIONS for sfx in CASADI_SUFFIXES]


def load_library():
    ''' load libraceline_b200.so (raises if it has not been built: run __graft_entry__.build()) '''
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: build it with __graft_entry__.build(); '
                           'there is no CPU fallback for the raceline NLP functions')
    lib = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    lib.rb_last_error.restype = ctypes.c_char_p
    lib.rb_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    lib.rb_set_device.argtypes = [ctypes.c_int]
    lib.rb_problem_create.argtypes = [ctypes.POINTER(_Desc), ctypes.POINTER(vp)]
    lib.rb_problem_destroy.argtypes = [vp]
    lib.rb_problem_destroy.restype = None
    lib.rb_problem_nvp.argtypes = [vp]
    lib.rb_problem_set_tail.argtypes = [vp, ctypes.POINTER(_TailDesc)]
    lib.rb_sparsity_size.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]
    lib.rb_sparsity_get.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    lib.rb_eval_scratch_bytes.argtypes = [vp, ctypes.c_int]
    lib.rb_eval_scratch_bytes.restype = ctypes.c_size_t
    lib.rb_eval_batch.argtypes = [vp, ctypes.c_int] + [vp] * 4 + [ctypes.c_int] + [vp] * 8
    for name in ('rb_nlp_f', 'rb_nlp_g'):
        getattr(lib, name).argtypes = [vp, ctypes.c_int, vp, vp, vp]
    for name in ('rb_nlp_grad_f', 'rb_nlp_jac_g'):
        getattr(lib, name).argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
    lib.rb_nlp_hess_l.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp]
    lib.rb_nlp_eval_all.argtypes = [vp, ctypes.c_int] + [vp] * 9
    lib.rb_launch_count.restype = ctypes.c_longlong
    lib.rb_profile_enable.argtypes = [ctypes.c_int]
    lib.rb_profile_cell_ms.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    lib.rb_fp64_peak.argtypes = [ctypes.POINTER(ctypes.c_double)]
    _LIB = lib
    return lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f'{what} failed: {load_library().rb_last_error().decode()}')


class Sparsity:
    ''' CasADi-Sparsity-like view of a CCS pattern '''

    def __init__(self, nrow, ncol, colind, row):
        self._shape = (int(nrow), int(ncol))
        self._colind = np.asarray(colind, dtype=np.int64)
        self._row = np.asarray(row, dtype=np.int64)

    def size(self):
        return self._shape

    def size1(self):
        return self._shape[0]

    def size2(self):
        return self._shape[1]

    def nnz(self):
        return len(self._row)

    def colind(self):
        return self._colind.tolist()

    def row(self):
        return self._row.tolist()

    def compressed(self):
        ''' [nrow, ncol, colind..., row...] -- CasADi's compressed form '''
        return [*self._shape, *self._colind.tolist(), *self._row.tolist()]

    @staticmethod
    def dense(nrow, ncol=1):
        return Sparsity(nrow, ncol, np.arange(ncol + 1) * nrow, np.tile(np.arange(nrow), ncol))


class Function:
    ''' minimal CasADi-Function-like callable over one of the five entry points '''

    def __init__(self, name, names_in, names_out, sp_in, sp_out, call):
        self._name, self._in, self._out = name, names_in, names_out
        self._sp_in, self._sp_out, self._call = sp_in, sp_out, call

    def name(self):
        return self._name

    def n_in(self):
        return len(self._in)

    def n_out(self):
        return len(self._out)

    def name_in(self, i=None):
        return list(self._in) if i is None else self._in[i]

    def name_out(self, i=None):
        return list(self._out) if i is None else self._out[i]

    def sparsity_in(self, i):
        return self._sp_in[i]

    def sparsity_out(self, i):
        return self._sp_out[i]

    def size_in(self, i):
        return self._sp_in[i].size()

    def size_out(self, i):
        return self._sp_out[i].size()

    def nnz_out(self, i):
        return self._sp_out[i].nnz()

    def __call__(self, *args, **kwargs):
        if kwargs:
            args = [kwargs.get(k) for k in self._in]
        out = self._call(*args)
        return out if len(self._out) > 1 else out[0]


class NlpFunctions:
    ''' the five NLP functions of one structured raceline problem, evaluated on the GPU '''

    def __init__(self, structure: NLPStructure, vehicle_config=None, device=None):
        self.lib = load_library()
        cnt = ctypes.c_int(0)
        if self.lib.rb_device_count(ctypes.byref(cnt)) != 0 or cnt.value == 0:
            raise RuntimeError('no CUDA device: the raceline NLP functions run only on the GPU '
                               f'({self.lib.rb_last_error().decode()})')
        if device is not None:
            _check(self.lib.rb_set_device(int(device)), 'rb_set_device')
        self.st = structure
        self._keep = []
        self.handle = ctypes.c_void_p()
        _check(self.lib.rb_problem_create(ctypes.byref(self._make_desc()), ctypes.byref(self.handle)),
               'rb_problem_create')
        self.nvp = self.lib.rb_problem_nvp(self.handle)
        if structure.tail is not None:
            tp = structure.tail
            td = _TailDesc()
            td.n_ins, td.ins = len(tp['ins']), self._arr(tp['ins'], np.int32, _c_int_p)
            td.n_levels = (ctypes.c_int * 3)(*tp['n_levels'])
            td.lvl_ptr = self._arr(tp['lvl_ptr'], np.int32, _c_int_p)
            td.n_const, td.cval = len(tp['cval']), self._arr(tp['cval'], np.float64, _c_dbl_p)
            td.n_slots = int(tp['n_slots'])
            _check(self.lib.rb_problem_set_tail(self.handle, ctypes.byref(td)), 'rb_problem_set_tail')
        self.vp = vehicle_params(vehicle_config) if vehicle_config is not None else None
        st = structure
        self.sp_jac = Sparsity(st.ng, st.nw, st.jac_colind, st.jac_row)
        self.sp_hess = Sparsity(st.nw, st.nw, st.hess_colind, st.hess_row)
        dx, dg, d1 = Sparsity.dense(st.nw), Sparsity.dense(st.ng), Sparsity.dense(1)
        dp = Sparsity.dense(self.nvp)
        self.nlp_f = Function('nlp_f', ['x', 'p'], ['f'], [dx, dp], [d1],
                              lambda x, p=None: (self.eval(x, p, want='f')['f'],))
        self.nlp_g = Function('nlp_g', ['x', 'p'], ['g'], [dx, dp], [dg],
                              lambda x, p=None: (self.eval(x, p, want='g')['g'],))
        self.nlp_grad_f = Function('nlp_grad_f', ['x', 'p'], ['f', 'grad_f_x'], [dx, dp], [d1, dx],
                                   lambda x, p=None: tuple(self.eval(x, p, want=('f', 'grad_f'))[k]
                                                           for k in ('f', 'grad_f')))
        self.nlp_jac_g = Function('nlp_jac_g', ['x', 'p'], ['g', 'jac_g_x'], [dx, dp], [dg, self.sp_jac],
                                  lambda x, p=None: tuple(self.eval(x, p, want=('g', 'jac'))[k]
                                                          for k in ('g', 'jac')))
        self.nlp_hess_l = Function('nlp_hess_l', ['x', 'p', 'lam_f', 'lam_g'], ['triu_hess_gamma_x_x'],
                                   [dx, dp, d1, dg], [self.sp_hess],
                                   lambda x, p=None, lam_f=1.0, lam_g=None:
                                   (self.eval(x, p, lam_f, lam_g, want='hess')['hess'],))

    def bind_casadi_symbols(self, p=None):
        ''' make this problem the one the library's CasADi-style symbols (nlp_f, nlp_g, nlp_grad_f, nlp_jac_g,
        nlp_hess_l and their _n_in / _sparsity_out / ... companions, include/raceline_b200.h) evaluate;
        p: vehicle parameters used when a caller passes arg[1] == NULL '''
        vp = self.vp if p is None else p
        arr = None if vp is None else np.ascontiguousarray(vp, dtype=np.float64)
        if arr is not None and arr.size != self.nvp:
            raise ValueError(f'p has {arr.size} entries, expected {self.nvp}')
        self.lib.rb_casadi_bind.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _check(self.lib.rb_casadi_bind(self.handle, None if arr is None else arr.ctypes.data_as(ctypes.c_void_p)),
               'rb_casadi_bind')

    def __del__(self):
        try:
            if getattr(self, 'handle', None) is not None and self.handle.value:
                self.lib.rb_problem_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    def _arr(self, a, dtype, ptr):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype)
        self._keep.append(a)
        return a.ctypes.data_as(ptr)

    def _make_desc(self):
        st = self.st
        i32 = lambda a: self._arr(a, np.int32, _c_int_p)
        i64 = lambda a: self._arr(a, np.int64, _c_i64_p)
        f64 = lambda a: self._arr(a, np.float64, _c_dbl_p)
        c, s, h = st.cell, st.srow, st.shess
        d = _Desc()
        d.transcription, d.variant, d.N, d.K = st.transcription, st.variant.encode(), st.N, st.K
        d.nw, d.ng, d.nnz_jac, d.nnz_hess = st.nw, st.ng, st.nnz_jac, st.nnz_hess
        d.R, d.dR, d.fc = f64(st.R), f64(st.dR), f64(st.fc)
        d.cell_row, d.cell_coef, d.cell_partner = i32(c['row']), f64(c['coef']), i32(c['partner'])
        d.cell_pcoef, d.cell_off, d.cell_par = f64(c['pcoef']), f64(c['off']), f64(c['par'])
        d.cell_jslot, d.cell_hslot = i32(c['jslot']), i32(c['hslot'])
        d.cell_nj, d.cell_nh, d.cell_ncp = c['nj'], c['nh'], c['ncp']
        if c.get('tmpl_j') is not None:
            d.tmpl_j, d.tmpl_h = i32(c['tmpl_j']), i32(c['tmpl_h'])
            d.n_tmpl_j, d.n_tmpl_h = len(c['tmpl_j']), len(c['tmpl_h'])
            d.colloc_C, d.colloc_D, d.colloc_B = f64(c['C']), f64(c['D']), f64(c['B'])
        d.n_srow = s['n']
        d.srow_row, d.srow_kind, d.srow_scale = i32(s['row']), i32(s['kind']), i32(s['scale'])
        d.srow_var_ptr, d.srow_var, d.srow_jslot = i32(s['var_ptr']), i32(s['var']), i32(s['jslot'])
        d.srow_form_ptr, d.srow_coef_ptr = i32(s['form_ptr']), i32(s['coef_ptr'])
        d.srow_A, d.srow_c = f64(s['A']), f64(s['c'])
        d.n_shess = h['n']
        d.shess_slot, d.shess_add, d.shess_ptr = i32(h['slot']), i32(h['add']), i32(h['ptr'])
        d.shess_row, d.shess_coef, d.shess_scale = i32(h['row']), f64(h['coef']), i32(h['scale'])
        d.jac_colind, d.jac_row = i64(st.jac_colind), i64(st.jac_row)
        d.hess_colind, d.hess_row = i64(st.hess_colind), i64(st.hess_row)
        return d

    # ---- host-buffer evaluation (the CasADi-shaped path) ------------------------------------
    def eval(self, x, p=None, lam_f=None, lam_g=None, want=('f', 'grad_f', 'g', 'jac', 'hess')):
        '''
        evaluate at x (nw,) or (B, nw) through the host entry point rb_nlp_eval_all.
        p: vehicle parameters (nvp,) or (B, nvp); defaults to the builder's vehicle config.
        Returns a dict of numpy arrays for the requested outputs.
        '''
        if isinstance(want, str):
            want = (want,)
        st = self.st
        x = np.asarray(x, dtype=np.float64)
        single = x.ndim == 1
        X = np.ascontiguousarray(np.atleast_2d(x))
        B = X.shape[0]
        if X.shape[1] != st.nw:
            raise ValueError(f'x has {X.shape[1]} entries, expected {st.nw}')
        if p is None:
            if self.vp is None:
                raise ValueError('vehicle parameters p are required')
            p = self.vp
        P = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(p, dtype=np.float64)), (B, self.nvp)))
        out = {}
        ptr = {}
        shapes = dict(f=(B,), grad_f=(B, st.nw), g=(B, st.ng), jac=(B, st.nnz_jac), hess=(B, st.nnz_hess))
        for k in ('f', 'grad_f', 'g', 'jac', 'hess'):
            if k in want:
                out[k] = np.empty(shapes[k])
                ptr[k] = out[k].ctypes.data_as(ctypes.c_void_p)
            else:
                ptr[k] = None
        LF = LG = None
        if 'hess' in want:
            LF = np.ascontiguousarray(np.broadcast_to(np.asarray(1.0 if lam_f is None else lam_f,
                                                                 dtype=np.float64).reshape(-1), (B,)))
            if st.ng:
                if lam_g is None:
                    raise ValueError('nlp_hess_l needs lam_g')
                LG = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(lam_g, dtype=np.float64)),
                                                          (B, st.ng)))
        vp_ = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        _check(self.lib.rb_nlp_eval_all(self.handle, B, vp_(X), vp_(P), vp_(LF), vp_(LG),
                                        ptr['f'], ptr['grad_f'], ptr['g'], ptr['jac'], ptr['hess']),
               'rb_nlp_eval_all')
        if single:
            out = {k: (float(v[0]) if k == 'f' else v[0]) for k, v in out.items()}
        return out

    # ---- device-pointer evaluation (torch tensors) -------------------------------------------
    def eval_device(self, x, lam_g=None, lam_f=None, vp=None, fc=None, f=None, grad_f=None, g=None,
                    jac=None, hess=None, scratch=None, stream=None):
        '''
        asynchronous batched evaluation on torch CUDA tensors (fp64, contiguous).
        x (B, nw); vp (B, nvp) or (nvp,); outputs are written in place into the given tensors.
        '''
        import torch
        B = x.shape[0]
        if vp is None:
            raise ValueError('vp (device tensor) is required')
        vp_stride = 0 if vp.dim() == 1 else vp.shape[1]
        if scratch is None:
            need = self.lib.rb_eval_scratch_bytes(self.handle, B)
            cached = getattr(self, '_scratch', None)
            if cached is None or cached.numel() < need or cached.device != x.device:
                cached = torch.empty(need, dtype=torch.uint8, device=x.device)
                self._scratch = cached
            scratch = cached
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        dp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        for t in (x, lam_g, lam_f, vp, fc, f, grad_f, g, jac, hess):
            if t is not None:
                assert t.is_cuda and t.is_contiguous() and (t.dtype == torch.float64)
        _check(self.lib.rb_eval_batch(self.handle, B, dp(x), dp(lam_g), dp(lam_f), dp(vp), vp_stride,
                                      dp(fc), dp(f), dp(grad_f), dp(g), dp(jac), dp(hess), dp(scratch),
                                      ctypes.c_void_p(stream)), 'rb_eval_batch')
        return scratch

    def launch_count(self):
        return int(self.lib.rb_launch_count())
