'''
GPU parity of the device-side steps around the solve (csrc/post.cuh through the C ABI) against the host numpy
implementations that follow the reference (aircraft_trajectory_optimization_b200/centerlines.py::frame_constants after
drone3d/centerlines/spline_centerline.py:232-322; interpolation.py::make_interpolants after
drone3d/utils/discretization_utils.py:53-137) and against the oracle's centerline (oracle/ref_centerline.py).
'''
import numpy as np
import pytest

from cases import make_line, build_product


@pytest.mark.gpu
def test_frame_constants_of_three_tracks_in_one_launch(built_library):
    from aircraft_trajectory_optimization_b200.post import frame_constants_batch
    from oracle.ref_centerline import RefSplineCenterline
    lines = [make_line(t) for t in ('race', 'fig8', 'obs')]
    rng = np.random.default_rng(0)
    M = 4001
    # per-track path lengths: a grid across the whole lap, the knots themselves, and points beyond both ends
    S = np.stack([np.concatenate([np.linspace(l.s_min() - 0.5, l.s_max() + 0.5, M - len(l.config.s)), l.config.s])
                  for l in lines])
    yn = rng.uniform(-1, 1, (3, M, 2))
    fc, xc, xg = frame_constants_batch(lines, S, yn=yn)
    fc, xc, xg = fc.cpu().numpy(), xc.cpu().numpy(), xg.cpu().numpy()
    for t, (name, line) in enumerate(zip(('race', 'fig8', 'obs'), lines)):
        ref = line.frame_constants(S[t])
        assert np.abs(fc[t] - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), name
        fr = line.frame(S[t])
        assert np.abs(xc[t] - fr['xc']).max() <= 1e-12 * max(1.0, np.abs(fr['xc']).max())
        xg_ref = fr['xc'] + fr['ey'] * yn[t, :, :1] + fr['en'] * yn[t, :, 1:]
        assert np.abs(xg[t] - xg_ref).max() <= 1e-12 * max(1.0, np.abs(xg_ref).max())
        # the oracle's restatement of the reference centerline, inside the lap
        rl = make_line(name, RefSplineCenterline)
        inside = (S[t] >= line.s_min()) & (S[t] < line.s_max())
        for i in np.nonzero(inside)[0][::400]:
            Rp = np.asarray(rl.p2Rp(S[t, i]), dtype=float)
            assert np.abs(fc[t, i, :9].reshape(3, 3) - Rp).max() <= 1e-10
            assert abs(fc[t, i, 10] - float(rl.p2ky(S[t, i]))) <= 1e-9 * max(1.0, abs(fc[t, i, 10]))
    # shared path lengths (stride 0) give the same numbers as per-track rows
    fc2, _ = frame_constants_batch(lines[:1], S[0])
    assert np.array_equal(fc2.cpu().numpy()[0], fc[0])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['fig8_global_colloc_drone', 'race_param_rk4_drone', 'obs_param_colloc_point'])
def test_interpolants_match_the_host_closures(name, built_library):
    from aircraft_trajectory_optimization_b200.post import interp_batch
    from aircraft_trajectory_optimization_b200.interpolation import make_interpolants
    prod = build_product(name, small=True)
    st, cfg = prod.structure, prod.config
    N, K = cfg.N, cfg.K
    S = prod.model.nz + 2 * prod.model.nu
    rng = np.random.default_rng(1)
    B, M = 3, 257
    W = np.stack([st.w0 + 0.05 * rng.standard_normal(st.nw) for _ in range(B)])
    W[:, :N] = np.abs(W[:, :N]) + 1e-3
    lap = W[:, :N].sum(1)
    # query times: across the lap, before the start and after the end, and exactly at interval boundaries
    tq = np.stack([np.concatenate([np.linspace(-0.1 * l, 1.1 * l, M - N - 1), np.concatenate([[0.0], np.cumsum(W[b, :N])])])
                   for b, l in enumerate(lap)])
    colloc = None if cfg.use_rk4 else (prod.tau, prod.D)
    out = interp_batch(W, N, K, S, tq, *(colloc or (None, None))).cpu().numpy()
    nz, nu = prod.model.nz, prod.model.nu
    for b in range(B):
        H, Z, U, dU = prod.unpack_w(W[b])
        zi, ui, dui = make_interpolants(H, Z, U, dU, colloc)
        ref = np.stack([np.concatenate([zi(t), ui(t), dui(t)]) for t in tq[b]])
        scale = max(1.0, np.abs(ref).max())
        # (exactly at an interval boundary the host closure and the kernel may pick neighbouring intervals when the
        # cumulative sums differ in the last bit; both are continuous there for z, so compare z only at those times)
        inner = slice(0, M - N - 1)
        assert np.abs(out[b, inner] - ref[inner]).max() <= 1e-10 * scale, (name, b)
    # a shared row of query times
    out1 = interp_batch(W[:1], N, K, S, tq[0], *(colloc or (None, None))).cpu().numpy()
    assert np.array_equal(out1[0], out[0])


@pytest.mark.gpu
def test_track_sweep_tables_built_on_the_device_feed_the_evaluation(built_library):
    ''' s8(f)-3 end to end: frame tables of several tracks from rb_centerline_frames go straight into rb_eval_batch as
    per-instance constants; the nominal track reproduces the host-built table's results to 1e-10 '''
    import copy
    import torch
    from cases import eval_point
    from aircraft_trajectory_optimization_b200.post import frame_constants_batch
    from aircraft_trajectory_optimization_b200.centerlines import SplineCenterline
    prod = build_product('race_param_rk4_drone', small=True)
    st, F, cfg = prod.structure, prod.functions, prod.config
    s_all = np.array([prod._get_s(n, k) for n in range(cfg.N) for k in range(cfg.K + 1)])
    base = prod.line
    # same waypoints, stretched vertically: a different track with the same path-length grid
    c2 = copy.deepcopy(base.config)
    c2.x = c2.x * np.array([[1.0], [1.0], [1.3]])
    lines = [base, SplineCenterline(c2), base]
    fcb, _ = frame_constants_batch(lines, s_all)
    assert fcb.shape == (3, len(s_all), 13)
    dev = torch.device('cuda', 0)
    B = 3
    x, lam = eval_point(st, 0)
    X = torch.from_numpy(np.tile(x, (B, 1))).to(dev)
    L = torch.from_numpy(np.tile(lam, (B, 1))).to(dev)
    vp = torch.from_numpy(np.asarray(F.vp)).to(dev)
    sig = torch.ones(B, dtype=torch.float64, device=dev)
    g = torch.empty(B, st.ng, dtype=torch.float64, device=dev)
    j = torch.empty(B, st.nnz_jac, dtype=torch.float64, device=dev)
    h = torch.empty(B, st.nnz_hess, dtype=torch.float64, device=dev)
    F.eval_device(X, L, sig, vp, fcb.contiguous(), None, None, g, j, h)
    torch.cuda.synchronize()
    ref = F.eval(x, lam_f=1.0, lam_g=lam)
    for name, t in (('g', g), ('jac', j), ('hess', h)):
        a = t.cpu().numpy()
        scale = max(1.0, np.abs(ref[name]).max())
        assert np.abs(a[0] - ref[name]).max() <= 1e-10 * scale, name
        assert np.array_equal(a[0], a[2])
        assert np.abs(a[1] - ref[name]).max() > 1e-6 * scale, name      # the stretched track is a different problem
