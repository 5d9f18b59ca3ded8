'''
GPU edge cases of the C ABI: empty batches, outputs not wanted, per-instance frame constants, error paths,
and the maximum-size structures of BASELINE.json (C2 at N = 490, C3 at N = 100 x K = 7).
'''
import ctypes

import numpy as np
import pytest

from cases import build_product, eval_point


@pytest.mark.gpu
def test_empty_batch_and_unwanted_outputs(built_library):
    prod = build_product('race_param_rk4_point', small=True)
    st, F = prod.structure, prod.functions
    lib = F.lib
    # B = 0 is a no-op that succeeds
    assert lib.rb_nlp_eval_all(F.handle, 0, None, None, None, None, None, None, None, None, None) == 0
    x, lam = eval_point(st, 0)
    full = F.eval(x, lam_f=1.0, lam_g=lam)
    only_g = F.eval(x, want='g')
    assert set(only_g) == {'g'} and np.array_equal(only_g['g'], full['g'])
    only_h = F.eval(x, lam_f=1.0, lam_g=lam, want='hess')
    assert np.array_equal(only_h['hess'], full['hess'])
    # hess_l without multipliers is an error, reported through rb_last_error
    hbuf = np.empty(st.nnz_hess)
    vp = np.ascontiguousarray(F.vp)
    rc = lib.rb_nlp_hess_l(F.handle, 1, x.ctypes.data_as(ctypes.c_void_p), vp.ctypes.data_as(ctypes.c_void_p), None, None,
                           hbuf.ctypes.data_as(ctypes.c_void_p))
    assert rc != 0 and b'lam_g' in lib.rb_last_error()


@pytest.mark.gpu
def test_unknown_variant_is_refused(built_library):
    from aircraft_trajectory_optimization_b200.functions import NlpFunctions
    prod = build_product('race_param_rk4_point', small=True)
    st = prod.structure
    bad = type(st)(**{**st.__dict__, 'variant': 'drone_quat_param_lr_drag'})
    with pytest.raises(RuntimeError, match='not compiled'):
        NlpFunctions(bad, prod.vehicle_config)


@pytest.mark.gpu
def test_malformed_tail_tape_is_refused(built_library):
    ''' rb_problem_set_tail checks every instruction against the problem's sizes before anything reaches the device '''
    from aircraft_trajectory_optimization_b200.functions import NlpFunctions
    prod = build_product('race_global_rk4_point_open', small=True)
    st = prod.structure
    for col, bad_value, what in ((3, st.tail['n_slots'], 'out of range'),       # destination slot beyond the work array
                                 (0, 99, 'out of range')):                      # unknown opcode
        tape = {**st.tail, 'ins': st.tail['ins'].copy()}
        tape['ins'][5, col] = bad_value
        bad = type(st)(**{**st.__dict__, 'tail': tape})
        with pytest.raises(RuntimeError, match=what):
            NlpFunctions(bad, prod.vehicle_config)
    # a store beyond the Jacobian values
    tape = {**st.tail, 'ins': st.tail['ins'].copy()}
    k = int(np.nonzero(tape['ins'][:, 0] == 15)[0][0])
    tape['ins'][k, 2] = st.nnz_jac
    with pytest.raises(RuntimeError, match='out of range'):
        NlpFunctions(type(st)(**{**st.__dict__, 'tail': tape}), prod.vehicle_config)
    # level table that does not end at the last instruction
    tape = {**st.tail, 'lvl_ptr': st.tail['lvl_ptr'].copy()}
    tape['lvl_ptr'][-1] -= 1
    with pytest.raises(RuntimeError, match='level table'):
        NlpFunctions(type(st)(**{**st.__dict__, 'tail': tape}), prod.vehicle_config)


@pytest.mark.gpu
def test_per_instance_frame_constants_match_shared(built_library):
    ''' track sweeps pass frame constants per instance (fc_b); equal tables must reproduce the shared result '''
    import torch
    prod = build_product('race_param_rk4_drone', small=True)
    st, F = prod.structure, prod.functions
    B = 3
    dev = torch.device('cuda', 0)
    X = torch.from_numpy(np.stack([eval_point(st, s)[0] for s in range(B)])).to(dev)
    L = torch.from_numpy(np.stack([eval_point(st, s)[1] for s in range(B)])).to(dev)
    vp = torch.from_numpy(np.asarray(F.vp)).to(dev)
    sig = torch.ones(B, dtype=torch.float64, device=dev)
    fcb = torch.from_numpy(np.tile(st.fc[None], (B, 1, 1))).to(dev).contiguous()
    outs = []
    for fc in (None, fcb):
        f = torch.empty(B, dtype=torch.float64, device=dev)
        gf = torch.empty(B, st.nw, dtype=torch.float64, device=dev)
        g = torch.empty(B, st.ng, dtype=torch.float64, device=dev)
        j = torch.empty(B, st.nnz_jac, dtype=torch.float64, device=dev)
        h = torch.empty(B, st.nnz_hess, dtype=torch.float64, device=dev)
        F.eval_device(X, L, sig, vp, fc, f, gf, g, j, h)
        torch.cuda.synchronize()
        outs.append([t.cpu().numpy() for t in (f, gf, g, j, h)])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    # and a perturbed table changes the result of that instance only
    fcb2 = fcb.clone()
    fcb2[1, :, 9:] *= 1.01
    g2 = torch.empty(B, st.ng, dtype=torch.float64, device=dev)
    F.eval_device(X, None, None, vp, fcb2, None, None, g2, None, None)
    torch.cuda.synchronize()
    g2 = g2.cpu().numpy()
    assert np.array_equal(g2[0], outs[0][2][0]) and np.array_equal(g2[2], outs[0][2][2])
    assert not np.array_equal(g2[1], outs[0][2][1])


@pytest.mark.gpu
def test_maximum_size_structures(built_library):
    ''' C2 (nw = 10780) and C3 (nw = 16900): finite outputs, chunked and unchunked batches agree '''
    for name, B in (('race_param_rk4_drone', 130), ('obs_param_colloc_drone', 3)):
        prod = build_product(name)
        st, F = prod.structure, prod.functions
        X = np.stack([eval_point(st, s)[0] for s in range(B)])
        L = np.stack([eval_point(st, s)[1] for s in range(B)])
        out = F.eval(X, lam_f=np.ones(B), lam_g=L)
        assert all(np.isfinite(v).all() for v in out.values())
        one = F.eval(X[B - 1], lam_f=1.0, lam_g=L[B - 1])       # last instance: second chunk of the batch call
        for k in ('g', 'jac', 'hess', 'grad_f'):
            assert np.array_equal(out[k][B - 1], one[k]), (name, k)


@pytest.mark.gpu
def test_kkt_accepts_collocation_structures(built_library):
    from aircraft_trajectory_optimization_b200.kkt import KktSolver
    st = build_product('fig8_global_colloc_drone', small=True).structure
    # default: interiors condensed, the reduced system goes through the shared-memory chain kernels
    K = KktSolver(st)
    assert K.cs is not None and K.cs.amax > 250 and K.ks.bmax <= 64 and K.ks.nb < 64
    # the uncondensed chain of large triples (csrc/kkt_big.cuh) stays available
    K = KktSolver(st, condensed=False)
    assert K.cs is None and K.ks.bmax > 300 and K.ks.nb < 64
