'''
Batched warm-start chain on the device (host side of csrc/warm_start.cuh).

The reference maps ONE point-mass raceline to ONE drone initial guess point by point in Python
(drone3d/raceline/drone_raceline.py:158-277, with scipy Rotation calls per point).  `drone_guess_batch` does it for B
point-mass solutions at once -- the multi-start / vehicle-parameter sweeps warm-start every instance from its own
point-mass solve -- and reports, per instance, what the builder needs for the modified loop closure
(drone_raceline.py:82-95): whether the quaternion closes with a sign flip, or how often the yaw angle wraps.
'''
import ctypes

import numpy as np

from .functions import load_library, _check


class _WsArgs(ctypes.Structure):
    _fields_ = ([(k, ctypes.c_int) for k in ('B', 'N', 'P', 'quat', 'closed', 'global_r', 'reserved', 'nw_pm', 'nw_dr')]
                + [(k, ctypes.c_void_p) for k in ('w_pm', 'fc', 'w_dr', 'info')])


def drone_guess_batch(w_pm, N, K, quat=True, closed=True, global_r=True, fc=None, device=0):
    '''
    w_pm: (B, N + N (K+1) 12) point-mass decision vectors (numpy or CUDA tensor).  fc: (N (K+1), 13) frame constants
    of the centerline (SplineCenterline.frame_constants), needed when global_r is False.
    Returns (w0 (B, N + N (K+1) (nz + 8)) CUDA tensor, info (B, 4) int32 CUDA tensor: flipped closure, yaw wraps,
    continuity failures, 0).
    '''
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('drone_guess_batch needs a CUDA device (no CPU fallback)')
    lib = load_library()
    lib.rb_ws_drone_guess.argtypes = [ctypes.POINTER(_WsArgs), ctypes.c_void_p]
    dev = torch.device('cuda', device)
    w_pm = torch.as_tensor(np.ascontiguousarray(w_pm) if isinstance(w_pm, np.ndarray) else w_pm,
                           dtype=torch.float64, device=dev).contiguous()
    if w_pm.dim() == 1:
        w_pm = w_pm[None, :]
    B, P = w_pm.shape[0], K + 1
    nz = 13 if quat else 12
    nw_dr = N + N * P * (nz + 8)
    assert w_pm.shape[1] == N + N * P * 12, 'point-mass decision vector: N step sizes + N (K+1) x (6 states, 3 inputs, 3 rates)'
    w_dr = torch.empty(B, nw_dr, dtype=torch.float64, device=dev)
    info = torch.zeros(B, 4, dtype=torch.int32, device=dev)
    fc_t = None
    if fc is not None:
        fc_t = torch.as_tensor(np.ascontiguousarray(fc), dtype=torch.float64, device=dev).contiguous()
        assert fc_t.shape == (N * P, 13)
    a = _WsArgs(B=B, N=N, P=P, quat=int(quat), closed=int(closed), global_r=int(global_r), reserved=0,
                nw_pm=w_pm.shape[1], nw_dr=nw_dr, w_pm=w_pm.data_ptr(), fc=fc_t.data_ptr() if fc_t is not None else None,
                w_dr=w_dr.data_ptr(), info=info.data_ptr())
    _check(lib.rb_ws_drone_guess(ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
           'rb_ws_drone_guess')
    return w_dr, info
