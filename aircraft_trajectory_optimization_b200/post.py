'''
Device-side steps around the solve (host side of csrc/post.cuh).

  frame_constants_batch   the per-point constants (Rp, ks, ky, kn, |xc'|) of M path lengths on each of T spline
                          centerlines -- what SplineCenterline.frame_constants computes on the host
                          (drone3d/centerlines/spline_centerline.py:232-322) -- in one launch; the result is the `fc_b`
                          argument of the evaluation kernels, so a track sweep never leaves the device.
  interp_batch            z(t), u(t), du(t) of B solutions at M times each -- the interpolants of
                          RacelineResults (drone3d/raceline/base_raceline.py:801-864,
                          drone3d/utils/discretization_utils.py:53-137).
'''
import ctypes

import numpy as np

from .functions import load_library, _check


class _SplineTab(ctypes.Structure):
    _fields_ = [('nkx', ctypes.c_int), ('nkr', ctypes.c_int)] + [(k, ctypes.c_void_p) for k in
                                                                 ('kx', 'cx', 'ex', 'kr', 'cr', 'er')]


def _cuda(device):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('this step runs only on a CUDA device (no CPU fallback)')
    return torch, torch.device('cuda', device)


def spline_tables(lines, device=0):
    ''' device copy of the cubic-spline tables of T SplineCenterline objects; returns (tabs tensor, keep-alive list) '''
    torch, dev = _cuda(device)
    keep, recs = [], []
    for line in lines:
        ptrs = []
        for sp in (line._xc, line._ry):
            kx = torch.as_tensor(np.ascontiguousarray(sp.kx), dtype=torch.float64, device=dev)
            c = torch.as_tensor(np.ascontiguousarray(sp.c), dtype=torch.float64, device=dev)        # (4, n-1, 3)
            e = torch.as_tensor(np.ascontiguousarray(np.stack([sp.end_val, sp.end_slope])), dtype=torch.float64, device=dev)
            keep += [kx, c, e]
            ptrs += [kx.data_ptr(), c.data_ptr(), e.data_ptr()]
        recs.append(_SplineTab(len(line._xc.kx), len(line._ry.kx), *ptrs))
    arr = (_SplineTab * len(recs))(*recs)
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    tabs = torch.as_tensor(raw, device=dev)
    keep.append(tabs)
    return tabs, keep


def frame_constants_batch(lines, s, yn=None, device=0):
    '''
    lines: T SplineCenterline objects; s: (M,) shared or (T, M) path lengths; yn: optional (T, M, 2) lateral offsets.
    Returns fc (T, M, 13), xc (T, M, 3) and, with yn, the global positions xg (T, M, 3) -- CUDA tensors.
    '''
    torch, dev = _cuda(device)
    lib = load_library()
    vp = ctypes.c_void_p
    lib.rb_centerline_frames.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]
    T = len(lines)
    tabs, keep = spline_tables(lines, device)
    s_t = torch.as_tensor(np.ascontiguousarray(s) if isinstance(s, np.ndarray) else s, dtype=torch.float64, device=dev).contiguous()
    shared = s_t.dim() == 1
    M = s_t.shape[-1]
    assert shared or s_t.shape[0] == T
    fc = torch.empty(T, M, 13, dtype=torch.float64, device=dev)
    xc = torch.empty(T, M, 3, dtype=torch.float64, device=dev)
    yn_t = xg = None
    if yn is not None:
        yn_t = torch.as_tensor(np.ascontiguousarray(yn) if isinstance(yn, np.ndarray) else yn, dtype=torch.float64, device=dev).contiguous()
        assert yn_t.shape == (T, M, 2)
        xg = torch.empty(T, M, 3, dtype=torch.float64, device=dev)
    p = lambda t: None if t is None else vp(t.data_ptr())
    _check(lib.rb_centerline_frames(p(tabs), T, p(s_t), M, 0 if shared else M, p(fc), p(xc), p(yn_t), p(xg),
                                    vp(torch.cuda.current_stream(dev).cuda_stream)), 'rb_centerline_frames')
    return (fc, xc, xg) if yn is not None else (fc, xc)


def interp_batch(W, N, K, S, tq, tau=None, D=None, device=0):
    '''
    W: (B, N + N (K+1) S) solutions; tq: (M,) shared or (B, M) query times; tau, D: collocation nodes and end weights
    (K > 0).  Returns (B, M, S) CUDA tensor: (z, u, du) at every query time.
    '''
    torch, dev = _cuda(device)
    lib = load_library()
    vp = ctypes.c_void_p
    ci = ctypes.c_int
    lib.rb_traj_interp.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, ci, ci, vp, vp, vp]
    T_ = lambda a: torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a, dtype=torch.float64, device=dev).contiguous()
    W = T_(W)
    if W.dim() == 1:
        W = W[None]
    B, P = W.shape[0], K + 1
    assert W.shape[1] == N + N * P * S
    tq = T_(tq)
    shared = tq.dim() == 1
    M = tq.shape[-1]
    tau_t = T_(tau) if K > 0 else None
    D_t = T_(D) if K > 0 else None
    tp = torch.empty(B, N + 1, dtype=torch.float64, device=dev)
    out = torch.empty(B, M, S, dtype=torch.float64, device=dev)
    p = lambda t: None if t is None else vp(t.data_ptr())
    _check(lib.rb_traj_interp(p(W), B, N, P, S, p(tau_t), p(D_t), p(tq), M, 0 if shared else M, p(tp), p(out),
                              vp(torch.cuda.current_stream(dev).cuda_stream)), 'rb_traj_interp')
    return out
