#!/usr/bin/env python
'''
bench.py -- jac_g + hess_lag evaluations per second on the BASELINE.json workload.

Workload (configs[1] / C2 / C5 of SURVEY.md s8d): scripts/race.py's racetrack, parametric
(non-Euclidean) pose, quaternion drone, RK4 shooting with N = 70*7 = 490 intervals
(nw = 10780).  One *step* evaluates a batch of independent problem instances -- seeded eval points
x = clip(w0 + 1e-2*N(0,1)), lam_g ~ N(0,1), vehicle parameters x U[0.9, 1.1] -- and produces
f, grad_f, g, jac_g values and hess_l values for every instance (one *eval* each).

  value   evals/s with inputs and outputs resident in HBM (CUDA events, max over ranks)
  e2e     evals/s through the C-ABI host entry point rb_nlp_eval_all with pinned HOST buffers,
          host->device and device->host copies inside the timed region
  --impl reference   the CPU path: the oracle's flat-tape interpreter (the execution model of
          CasADi's SX virtual machine) on all host cores, one instance per thread

Launch: python bench.py --gpus N --steps K --warmup W   (N > 1 under torch.distributed.run)
'''
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = 'jac_g+hess_lag evals/sec'
UNIT = 'evals/s'
CASE = 'race_param_rk4_drone'
WORKLOAD = ('scripts/race.py racetrack, parametric (non-Euclidean) pose, quaternion drone, RK4 shooting '
            'N=490 (C2), batch of seeded eval points x vehicle-parameter variants (C5 shape)')


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    ''' nvidia-smi clocks / throttle reasons sampled while the timed region runs '''
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            p = [q.strip() for q in ln.split(',')]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
            except ValueError:
                continue
            for nme, val in zip(names, p[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(nme)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=float(np.max(smax)) if smax else None,
                    samples=len(sm), reasons=sorted(reasons))


def make_inputs(st, vp0, B, seed0):
    ''' eval points and vehicle variants, seeded per instance '''
    X = np.empty((B, st.nw))
    L = np.empty((B, st.ng))
    VP = np.empty((B, len(vp0)))
    for b in range(B):
        rng = np.random.default_rng(seed0 + b)
        X[b] = np.clip(st.w0 + 1e-2 * rng.standard_normal(st.nw), st.lbw, st.ubw)
        L[b] = rng.standard_normal(st.ng)
        scale = np.ones(len(vp0))
        scale[[0, 2, 3, 4, 5, 6]] = np.random.default_rng(10_000 + seed0 + b).uniform(0.9, 1.1, 6)  # m, I1..3, l, k
        VP[b] = vp0 * scale
    return X, L, VP


def algorithmic_bytes(st, n_consts=13):
    ''' bytes one eval must move '''
    n_pts = st.N * (st.K + 1)
    return 8 * (st.nw + st.ng + n_pts * n_consts + st.ng + st.nw + st.nnz_jac + st.nnz_hess) + 8


def cpu_port_rate(seconds_budget=12.0, nthreads=0, sample_intervals=49):
    '''
    the CPU path on a bounded sample: the oracle's interpreter on the first `sample_intervals`
    of the 490 intervals (same track, same model; cost is linear in the interval count), rate
    scaled by sample_intervals / 490.  Returns (evals/s for the full workload, cores, sample text).
    '''
    from cases import build_oracle
    from oracle.nlp_functions import OracleNLP, max_threads
    ref = build_oracle(CASE, N=sample_intervals // 7)
    nlp = OracleNLP(ref)
    cores = nthreads or len(os.sched_getaffinity(0))
    cores = min(cores, max_threads())
    rng = np.random.default_rng(0)
    B = max(cores * 2, 8)
    IN = np.zeros((B, nlp.n_in))
    for b in range(B):
        IN[b, :nlp.nw] = np.clip(ref.w0 + 1e-2 * rng.standard_normal(nlp.nw), ref.lbw, ref.ubw)
        IN[b, nlp.nw:nlp.nw + nlp.ng] = rng.standard_normal(nlp.ng)
        IN[b, -1] = 1.0
    nlp.t_jac_g.batch(IN[:cores], cores)
    t0 = time.time()
    done = 0
    while time.time() - t0 < seconds_budget:
        nlp.t_grad_f.batch(IN, cores)
        nlp.t_jac_g.batch(IN, cores)
        nlp.t_hess_l.batch(IN, cores)
        done += B
    dt = time.time() - t0
    full_N = 490
    rate = done / dt * (ref.config.N / full_N)
    sample = (f'oracle flat-tape interpreter (CasADi SX-VM execution model), {cores} threads, one instance '
              f'per thread; {done} evals of the first {ref.config.N} of {full_N} intervals in {dt:.1f} s '
              f'(tapes: jac {nlp.t_jac_g.n}, hess {nlp.t_hess_l.n} instructions), rate scaled by '
              f'{ref.config.N}/{full_N}')
    return rate, cores, sample



def cpu_compiled_rate(st, vehicle_config, variant):
    ''' BASELINE.md plan B: the compiled straight-line port at FULL size (oracle/compiled_cell.py) '''
    from oracle.compiled_cell import compiled_rate
    from aircraft_trajectory_optimization_b200.models import vehicle_params
    r, cores, sample = compiled_rate(st, vehicle_params(vehicle_config), variant)
    return dict(value=r, unit=UNIT, cores=cores, kind='port', sample=sample)


def full_config(st, B, world):
    ''' the `config` object of the bench line; the reference arm prints the same one '''
    ab = algorithmic_bytes(st)
    return dict(workload=WORKLOAD, instances_per_gpu_per_step=B, nw=st.nw, ng=st.ng, nnz_jac=st.nnz_jac,
                nnz_hess=st.nnz_hess,
                l2_policy=f'inputs+outputs per step {ab * B / 1e9:.2f} GB per GPU, larger than the 126 MB L2',
                parallelism=f'{world} x independent instance shards, no collective on the data path')


def build_c2_with_warm_start():
    ''' scripts/race.py second solve: parametric quaternion drone, RK4, warm-started from the point-mass solve '''
    from cases import make_line
    from aircraft_trajectory_optimization_b200 import raceline as RL
    from aircraft_trajectory_optimization_b200.pytypes import DroneConfig
    line = make_line('race')
    cfg = RL.ParametricRacelineConfig(N=70, use_rk4=True, closed=True, verbose=False)
    cfg.fixed_gates = line.config.s[:-1]
    return RL.ParametricDroneRaceline(line, cfg, DroneConfig(global_r=True, use_quat=True))


def multistart_inputs(st, vp0, B, seed0, nz=13, nu=4):
    ''' w0_b = clip(w0_ws + 0.05 * scale * N(0,1)), vehicle parameters x U[0.9, 1.1] '''
    S = nz + 2 * nu
    scale = np.zeros(st.nw)
    scale[:st.N] = st.w0[:st.N]                                   # h: h0
    per = np.array([0, .5, .5, .1, .1, .1, .1, 1, 1, 1, .5, .5, .5, 1, 1, 1, 1, 0, 0, 0, 0])   # s pinned, du: 0
    scale[st.N:] = np.tile(per, (st.nw - st.N) // S)
    X0 = np.empty((B, st.nw))
    VP = np.empty((B, len(vp0)))
    for b in range(B):
        rng = np.random.default_rng(seed0 + b)
        X0[b] = np.clip(st.w0 + (0.05 * scale * rng.standard_normal(st.nw) if seed0 + b > 0 else 0.0), st.lbw, st.ubw)
        sc = np.ones(len(vp0))
        if seed0 + b > 0:
            sc[[0, 2, 3, 4, 5, 6]] = np.random.default_rng(10_000 + seed0 + b).uniform(0.9, 1.1, 6)
        VP[b] = vp0 * sc
    return X0, VP


def cpu_iteration_seconds(prod, n_rep=2):
    '''
    one interior-point iteration of the C2 drone NLP on ONE host core, the way the reference's CPU path
    spends it: interpreted evaluation of grad_f, jac_g, hess_l (flat tapes, CasADi's SX-VM execution model)
    on a 49-interval sample scaled to 490, plus a sparse direct factorisation + solve of the full-size KKT
    matrix (scipy SuperLU standing in for MUMPS / MA97).
    '''
    from cases import build_oracle
    from oracle.nlp_functions import OracleNLP
    from oracle.kkt_blocks_ref import kkt_matrix
    import scipy.sparse.linalg as spla
    ref = build_oracle(CASE, N=7)
    nlp = OracleNLP(ref)
    IN = np.zeros((1, nlp.n_in))
    IN[0, :nlp.nw] = ref.w0
    IN[0, -1] = 1.0
    t0 = time.perf_counter()
    for _ in range(n_rep):
        nlp.t_grad_f.batch(IN, 1), nlp.t_jac_g.batch(IN, 1), nlp.t_hess_l.batch(IN, 1)
    t_eval = (time.perf_counter() - t0) / n_rep * (490 / ref.config.N)
    st = prod.structure
    x = st.w0
    out = prod.functions.eval(x, lam_f=1.0, lam_g=np.ones(st.ng), want=('jac', 'hess'))
    K = kkt_matrix(st, out['hess'], out['jac'], np.ones(st.nw), np.where(st.lbg == st.ubg, 0.0, 1.0)).tocsc()
    t0 = time.perf_counter()
    for _ in range(n_rep):
        spla.splu(K).solve(np.ones(st.nw + st.ng))
    t_kkt = (time.perf_counter() - t0) / n_rep
    # the same evaluation through compiled straight-line code, full size, one core (oracle/compiled_cell.py)
    t_eval_c = None
    try:
        from oracle.compiled_cell import compiled_rate
        from aircraft_trajectory_optimization_b200.models import vehicle_params
        r, _, _ = compiled_rate(st, vehicle_params(prod.vehicle_config), prod.model.variant, seconds_budget=3.0, nthreads=1)
        t_eval_c = 1.0 / r
    except Exception:
        pass
    return t_eval, t_kkt, t_eval_c


def kkt_kernel_rate(prod, sol, VP, dev):
    ''' one full wave of the KKT factor/solve kernel on converged C5 iterates, timed with CUDA events, next to its
    algorithmic flop count (Gauss-Jordan block inverses + the dense products of the sweep) and the FP64 peak '''
    import torch
    be = prod.solver._backend
    K, ks, st = be.K, be.K.ks, prod.structure
    W = int(be.kkt_wave)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    reps = -(-W // sol['x'].shape[0])
    x = T(np.tile(sol['x'], (reps, 1))[:W])
    lam = T(np.tile(sol['lam_g'], (reps, 1))[:W])
    vp = T(np.tile(VP, (reps, 1))[:W])
    be.vp = vp
    ev = be.eval(x, lam, torch.ones(W, dtype=torch.float64, device=dev), True)
    dxd = torch.full((W, st.nw), 1e-2, dtype=torch.float64, device=dev)
    negd = torch.where(torch.as_tensor(st.lbg == st.ubg, device=dev)[None, :], torch.zeros(W, st.ng, dtype=torch.float64, device=dev),
                       torch.full((W, st.ng), -1.0, dtype=torch.float64, device=dev))
    rhs = torch.ones(W, st.nw + st.ng, dtype=torch.float64, device=dev)
    hess, jac = ev['hess'].contiguous(), ev['jac'].contiguous()
    K.factor_solve(hess, jac, dxd, negd, rhs)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        K.factor_solve(hess, jac, dxd, negd, rhs)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / 3
    # algorithmic flops of kkt_factor_kernel + kkt_solve_kernel (csrc/kkt_chain.cuh): Gauss-Jordan inverse of the upper
    # triangle (b^3), Q_n, YL_n, carry, P_{n+1}, border Schur complement; one solve = forward + backward products
    flop = 0.0
    for n in range(ks.N):
        b = int(ks.blk_ptr[n + 1] - ks.blk_ptr[n])
        sn, an = int(ks.sup_ptr[n + 1] - ks.sup_ptr[n]), int(ks.act[n])
        flop += 1.0 * b ** 3 + 2.0 * b * sn * an + 2.0 * an * an * sn
        flop += 2.0 * b * b + 2.0 * sn * an + 2.0 * b * an                       # solve: z, r_b, Q x_b
        if n < ks.N - 1:
            m, q = int(ks.cr_ptr[n + 1] - ks.cr_ptr[n]), int(ks.cc_ptr[n + 1] - ks.cc_ptr[n])
            flop += 2.0 * b * q * m + 2.0 * m * m * q + 2.0 * m * sn * an
            flop += 4.0 * b * m                                                  # solve: carry, back-substitution
    flop += 1.0 * ks.nb ** 3
    fp = ctypes.c_double(0)
    prod.functions.lib.rb_fp64_peak(ctypes.byref(fp))
    return dict(kernel='kkt_factor_kernel + kkt_solve_kernel', instances_per_launch=W, ms_per_launch=ms,
                factorisations_per_s=W / ms * 1e3, flop_per_factorisation=flop,
                achieved_tflops=flop * W / (ms * 1e-3) / 1e12, fp64_peak_tflops=fp.value,
                fp64_peak_source='rb_fp64_peak micro-kernel in this run (MEASURED_PEAKS.json has no fp64 entry)',
                frac=flop * W / (ms * 1e-3) / 1e12 / fp.value,
                note='latency bound: ~41 dependent Bunch-Kaufman pivot steps per stage block (one pivot warp, ~190 '
                     'instructions per step), 489 blocks in sequence per instance; dependent DFMA latency is 23 cycles')


def run_solves(args, dev, rank, world, dist):
    ''' converged raceline solves/s on a multi-start x vehicle-parameter batch of C2 (the C5 shape) '''
    import torch
    from aircraft_trajectory_optimization_b200.models import vehicle_params
    t_build = time.perf_counter()
    prod = build_c2_with_warm_start()                 # includes the point-mass warm-start solve on the GPU
    t_build = time.perf_counter() - t_build
    st = prod.structure
    vp0 = vehicle_params(prod.vehicle_config)
    B = args.solves_batch
    X0, VP = multistart_inputs(st, vp0, B, seed0=rank * B)
    prod.solver.verbose = False
    prod.solver.max_iter = args.solves_max_iter
    from aircraft_trajectory_optimization_b200.ipm import IpmOptions
    prod.solver.options = IpmOptions(window=args.solves_window, refine_steps=args.solves_refine)
    lib_launch0 = prod.functions.launch_count()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sol = prod.solver(x0=X0, lbx=st.lbw, ubx=st.ubw, lbg=st.lbg, ubg=st.ubg, p=VP)
    e1.record()
    torch.cuda.synchronize(dev)
    s = prod.solver.stats()
    ok = s['success_each']
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3, float(ok.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t[0] = tmax[0]
    # the only exchange of the multi-GPU path: a final gather of per-instance results on rank 0
    from aircraft_trajectory_optimization_b200.sharding import gather_results
    full = gather_results(dict(lap_time=sol['x'][:, :st.N].sum(1), success=np.asarray(ok), iterations=s['iterations_each']),
                          world * B, dist if world > 1 else None)
    if rank != 0:
        return None
    laps, okf, its = full['lap_time'], full['success'].astype(bool), full['iterations']
    out = dict(value=float(t[1]) / float(t[0]), unit='converged solves/s', instances=world * B, window=args.solves_window,
               converged=int(okf.sum()), seconds=float(t[0]), sweeps=int(prod.solver.result.n_iter),
               iterations_median=float(np.median(its[okf])) if okf.any() else None,
               iterations_p90=float(np.percentile(its[okf], 90)) if okf.any() else None,
               iterations_p99=float(np.percentile(its[okf], 99)) if okf.any() else None,
               iterations_max=int(its.max()), max_iter=args.solves_max_iter, kkt_factorisations=int(s['n_factor']), evaluations=int(s['n_eval']),
               t_eval_s=s['t_wall_nlp_hess_l'], t_kkt_s=s['t_wall_linear_solver'],
               lap_time_nominal=float(laps[0]), lap_time_min=float(laps[okf].min()) if okf.any() else None,
               lap_time_max=float(laps[okf].max()) if okf.any() else None,
               e2e=dict(value=float(t[1]) / float(t[0]), unit='converged solves/s',
                        api='solver(x0=, lbx=, ubx=, lbg=, ubg=, p=) with HOST arrays in and out (integration level A: the '
                            'interior-point driver runs on the device, only x0 / vehicle parameters go in and '
                            'x, f, g, lam_g, lam_x, status come back)',
                        h2d_bytes_per_step=int(8 * B * (st.nw + VP.shape[1]) + 8 * 2 * (st.nw + st.ng)),
                        d2h_bytes_per_step=int(8 * B * (2 * st.nw + 2 * st.ng + 3))),
               warm_start_setup_s=t_build, gpu_launches=int(prod.functions.launch_count() - lib_launch0),
               speculative_factorisations=int(prod.solver.result.n_speculated),
               restoration_visits=int(prod.solver.result.n_restorations),
               return_status_rank0={str(k): int(v) for k, v in zip(*np.unique(np.asarray(s['return_status']), return_counts=True))},
               workload='C5 shape: C2 x multi-start (w0_ws + 0.05*scale*N(0,1)) x vehicle parameters U[0.9,1.1]; '
                        'instance 0 of rank 0 is the nominal race.py problem')
    if rank == 0:
        try:
            out['kkt_kernel'] = kkt_kernel_rate(prod, sol, VP, dev)
        except Exception as exc:        # a side measurement must not take the bench line down
            out['kkt_kernel'] = dict(error=repr(exc))
    if rank == 0 and not args.no_cpu:
        t_eval, t_kkt, t_eval_c = cpu_iteration_seconds(prod)
        cores = len(os.sched_getaffinity(0))
        its = out['iterations_median'] or out['iterations_max']
        # IPOPT on this problem class: ~1.5 function evaluations and (measured here, inertia retries included)
        # `fac_per_it` factorisations per iteration of ONE instance
        fac_per_it = float(np.mean(prod.solver.result.factorisations_each / np.maximum(1, s['iterations_each'])))
        per_solve = its * (t_eval * 1.5 + t_kkt * fac_per_it)
        out['cpu_baseline'] = dict(value=cores / per_solve, unit='converged solves/s', cores=cores, kind='port',
                                   modelled=True,
                                   sample=f'MODELLED from measured parts (no CPU interior-point run at full size: the full-size '
                                          f'tape takes minutes to build): one IP iteration on one core = tape evaluation '
                                          f'{t_eval * 1e3:.1f} ms (49-interval sample x10) + SuperLU factor/solve of the full '
                                          f'{st.nw + st.ng}-dim KKT matrix {t_kkt * 1e3:.1f} ms x {fac_per_it:.2f} '
                                          f'factorisations/iteration; x {its:.0f} iterations (median of the GPU run); one solve '
                                          f'per core on {cores} cores')
        if t_eval_c is not None:
            per_solve_c = its * (t_eval_c * 1.5 + t_kkt * fac_per_it)
            out['cpu_baseline']['compiled'] = dict(
                value=cores / per_solve_c, unit='converged solves/s', modelled=True,
                sample=f'the same with the compiled straight-line evaluation, full size: {t_eval_c * 1e3:.2f} ms per evaluation')
    return out


def run_colloc_eval(dev, local_rank, steps, B=512):
    ''' C1 (scripts/fig_8.py: global-frame quaternion drone, Legendre collocation N=56, K=7): device-resident evals/s '''
    import torch
    from cases import build_product
    from aircraft_trajectory_optimization_b200.functions import NlpFunctions, load_library
    from aircraft_trajectory_optimization_b200.models import vehicle_params
    prod = build_product('fig8_global_colloc_drone')
    st = prod.structure
    F = NlpFunctions(st, prod.vehicle_config, device=local_rank)
    lib = load_library()
    X, L, VP = make_inputs(st, vehicle_params(prod.vehicle_config), B, seed0=0)
    f64 = dict(dtype=torch.float64, device=dev)
    x_d, l_d, vp_d = (torch.from_numpy(a).to(dev) for a in (X, L, VP))
    sig_d = torch.ones(B, **f64)
    outs = [torch.empty(B, **f64), torch.empty(B, st.nw, **f64), torch.empty(B, st.ng, **f64),
            torch.empty(B, st.nnz_jac, **f64), torch.empty(B, st.nnz_hess, **f64)]
    scratch = torch.empty(max(1, lib.rb_eval_scratch_bytes(F.handle, B)), dtype=torch.uint8, device=dev)
    step = lambda: F.eval_device(x_d, l_d, sig_d, vp_d, None, *outs, scratch)
    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    ab = algorithmic_bytes(st)
    peak, _ = _peaks()
    # KKT factor + solve of C1 (condensed interiors + chain kernels), one instance and one wave, CUDA events
    kkt = None
    try:
        from aircraft_trajectory_optimization_b200.kkt import KktSolver
        K = KktSolver(st)
        kkt = dict(path='kktc_interior_factor_kernel (one CTA per interval and instance) + kktc_gather_kernel + '
                        'kkt_factor_kernel + kkt_solve_kernel on the reduced system + kktc_interior_back_kernel',
                   interiors=int(K.cs.NI), interior_unknowns=int(K.cs.amax), separator_unknowns=int(K.cs.smax),
                   reduced_block=int(K.ks.bmax), border=int(K.ks.nb))
        for Bk in (1, 8):
            dxd = torch.full((Bk, st.nw), 1e-2, **f64)
            negd = torch.where(torch.as_tensor(st.lbg == st.ubg, device=dev)[None, :], torch.zeros(Bk, st.ng, **f64),
                               torch.full((Bk, st.ng), -1.0, **f64))
            rhs = torch.ones(Bk, st.nw + st.ng, **f64)
            hj = (outs[4][:Bk].contiguous(), outs[3][:Bk].contiguous())
            K.factor_solve(*hj, dxd, negd, rhs)
            torch.cuda.synchronize(dev)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for _ in range(5):
                K.factor_solve(*hj, dxd, negd, rhs)
            k1.record()
            torch.cuda.synchronize(dev)
            kkt[f'ms_per_factor_solve_B{Bk}'] = k0.elapsed_time(k1) / 5
    except Exception as exc:
        kkt = dict(error=repr(exc))
    return dict(workload='C1: scripts/fig_8.py, global-frame quaternion drone, Legendre collocation N=56 K=7', kkt=kkt,
                value=B / (ms * 1e-3), unit=UNIT, instances_per_step=B, ms_per_step=ms, nw=st.nw, ng=st.ng,
                nnz_jac=st.nnz_jac, nnz_hess=st.nnz_hess, algorithmic_bytes_per_eval=ab,
                roofline=dict(bound='hbm', achieved=ab * B / (ms * 1e-3) / 1e9, peak=peak, unit='GB/s',
                              frac=ab * B / (ms * 1e-3) / 1e9 / peak,
                              kernel='colloc_point_kernel<PF_drone_quat_global> + colloc_gather_kernel (all chunks of a step)'))


def run_single_solves(dev):
    '''
    the reference scripts as ONE problem each (configs 0-3 of BASELINE.json): warm-started drone solve of
    scripts/fig_8.py (C1, global frame, collocation N=56), scripts/race.py (C2, parametric RK4 N=490),
    scripts/obstacles.py (C3, collocation N=100, synthetic tube) and the N=200 narrow-gate stand-in (C4).  Every problem
    is solved twice; the second solve is reported (the first pays CUDA module loading and allocations).
    '''
    import contextlib
    import torch
    from cases import make_line, synthetic_tube_arrays
    from aircraft_trajectory_optimization_b200 import raceline as RL
    from aircraft_trajectory_optimization_b200.collocation import get_collocation_coefficients
    from aircraft_trajectory_optimization_b200.pytypes import DroneConfig
    from aircraft_trajectory_optimization_b200.solve_util import solve_util

    def obstacles():
        line = make_line('obs')
        cfg = RL.ParametricRacelineConfig(verbose=False, N=100)
        cfg.closed = True
        dc = DroneConfig(global_r=True, use_quat=True, collision_radius=0.4)
        tau = get_collocation_coefficients(cfg.K)[0]
        ds = (line.s_max() - line.s_min()) / cfg.N
        sgrid = np.array([line.s_min() + ds * (n + tau[k]) for n in range(cfg.N) for k in range(cfg.K + 1)])
        tube = RL.ObstacleFreeTube(*synthetic_tube_arrays(sgrid), dc.collision_radius)
        return RL.ParametricObstacleDroneRaceline(line, cfg, dc, None, tube, generate_ws=True)

    def cpc():
        line = make_line('fig8cpc')
        return solve_util(line, global_frame=False, drone=True, use_quaternion=True, global_r=True, use_ws=True,
                          solve=False, verbose=False, N=200)[0]

    cases = [
        ('C1 scripts/fig_8.py: global-frame quaternion drone, collocation N=56 K=7, warm-started',
         lambda: solve_util(make_line('fig8'), global_frame=True, drone=True, use_quaternion=True, global_r=True,
                            use_ws=True, solve=False, verbose=False, N=50)[0]),
        ('C2 scripts/race.py: parametric quaternion drone, RK4 N=490, warm-started', build_c2_with_warm_start),
        ('C3 scripts/obstacles.py: parametric quaternion drone in a tube, collocation N=100 K=7, warm-started', obstacles),
        ('C4 stand-in: narrow-gate fig-8, parametric collocation N=200 K=7, warm-started', cpc),
    ]
    out = []
    for label, make in cases:
        try:
            with contextlib.redirect_stdout(sys.stderr):
                solver = make()
                solver.solve()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                res = solver.solve()
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
            st = solver.solver.stats()
            out.append(dict(problem=label, lap_time=float(res.time), wall_s=dt, solver_s=float(solver.solve_time),
                            iterations=int(st['iter_count']), return_status=str(st['return_status']),
                            kkt_factorisations=int(st['n_factor']), t_kkt_s=float(st['t_wall_linear_solver']),
                            t_eval_s=float(st['t_wall_nlp_hess_l']), nw=int(solver.structure.nw), ng=int(solver.structure.ng)))
        except Exception as exc:        # a side measurement must not take the bench line down
            out.append(dict(problem=label, error=repr(exc)))
    return out


def run_measured_cpu_solve(dev):
    '''
    a MEASURED converged-solve comparison at full size with the same interior-point driver on both sides: the warm-start
    NLP of scripts/race.py (parametric point mass, RK4 N=490, 6370 variables, cold start).  CPU: oracle tapes (interpreted
    SX evaluation, the reference's execution model) + SuperLU on one core (oracle/cpu_backend.py); GPU: the product path.
    The drone NLP of C2 is not run on the CPU: its full-size tape takes minutes to build.
    '''
    import contextlib
    import torch
    from cases import build_oracle, build_product
    from oracle.nlp_functions import OracleNLP
    from oracle.cpu_backend import OracleBackend
    from aircraft_trajectory_optimization_b200.ipm import InteriorPoint, IpmOptions
    name = 'race_param_rk4_point'
    out = dict(problem='warm-start NLP of scripts/race.py: parametric point mass, RK4 N=490 (cold start)')
    with contextlib.redirect_stdout(sys.stderr):
        prod = build_product(name)
        prod.solve()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res = prod.solve()
        torch.cuda.synchronize(dev)
        out['gpu'] = dict(wall_s=time.perf_counter() - t0, lap_time=float(res.time), iterations=int(prod.solver.stats()['iter_count']),
                          return_status=str(prod.solver.stats()['return_status']))
        t0 = time.perf_counter()
        ref = build_oracle(name)
        nlp = OracleNLP(ref)
        t_build = time.perf_counter() - t0
        T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
        ip = InteriorPoint(OracleBackend(nlp, nlp), IpmOptions(max_iter=1000, verbose=False))
        t0 = time.perf_counter()
        r = ip.solve(T(ref.w0[None]), T(ref.lbw), T(ref.ubw), T(ref.lbg), T(ref.ubg))
        dt = time.perf_counter() - t0
    out['cpu'] = dict(wall_s=dt, lap_time=float(r.x[0, :ref.config.N].sum()), iterations=int(r.iterations[0]), cores=1,
                      t_eval_s=float(r.t_eval), t_kkt_s=float(r.t_kkt), tape_build_s=t_build,
                      kind='port: oracle tape interpreter + scipy SuperLU, same interior-point driver; SuperLU reports no inertia, so the '
                           'CPU run never regularises (it still ends at the same minimiser, in more iterations)')
    out['speedup_single_instance'] = dt / out['gpu']['wall_s']
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rates = []
    t0 = time.time()
    sample = ''
    cores = 1
    for i in range(args.warmup + args.steps):
        r, cores, sample = cpu_port_rate(seconds_budget=max(2.0, min(10.0, 60.0 / max(1, args.steps + args.warmup))))
        if i >= args.warmup:
            rates.append(r)
    v = float(np.mean(rates))
    from cases import build_product
    prod = build_product(CASE)                      # host-side structure only: sizes for the shared config object
    cpu = dict(value=v, unit=UNIT, cores=cores, kind='port', sample=sample)
    try:
        cpu['compiled'] = cpu_compiled_rate(prod.structure, prod.vehicle_config, prod.model.variant)
    except Exception as exc:
        cpu['compiled'] = dict(error=repr(exc))
    line = dict(impl='reference', metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * (time.time() - t0) / max(1, args.steps + args.warmup),
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=full_config(prod.structure, args.batch, args.gpus),
                cpu_baseline=cpu,
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=2048, help='problem instances per GPU per step')
    ap.add_argument('--e2e-batch', type=int, default=1024)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-solves', action='store_true', help='skip the converged-solves leg')
    ap.add_argument('--no-colloc', action='store_true', help='skip the C1 (collocation) evaluation leg')
    ap.add_argument('--no-single', action='store_true', help='skip the single-problem solves of the reference scripts')
    ap.add_argument('--solves-batch', type=int, default=2048, help='multi-start instances per GPU in the solves leg')
    ap.add_argument('--solves-window', type=int, default=2048, help='instances iterating at a time (continuous batching)')
    ap.add_argument('--solves-refine', type=int, default=4, help='iterative-refinement steps per KKT solve')
    ap.add_argument('--solves-max-iter', type=int, default=300,
                    help='iteration cap per instance in the multi-start sweep (p99 of converged instances is ~280; '
                         'the reference sets 1000 for its single solves)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    from cases import build_product
    from aircraft_trajectory_optimization_b200.functions import NlpFunctions, load_library
    from aircraft_trajectory_optimization_b200.models import vehicle_params
    prod = build_product(CASE)
    st = prod.structure
    F = NlpFunctions(st, prod.vehicle_config, device=local_rank)
    lib = load_library()
    vp0 = vehicle_params(prod.vehicle_config)
    B = args.batch
    X, L, VP = make_inputs(st, vp0, B, seed0=rank * B)

    f64 = dict(dtype=torch.float64, device=dev)
    x_d = torch.from_numpy(X).to(dev)
    l_d = torch.from_numpy(L).to(dev)
    vp_d = torch.from_numpy(VP).to(dev)
    sig_d = torch.ones(B, **f64)
    f_d = torch.empty(B, **f64)
    gf_d = torch.empty(B, st.nw, **f64)
    g_d = torch.empty(B, st.ng, **f64)
    j_d = torch.empty(B, st.nnz_jac, **f64)
    h_d = torch.empty(B, st.nnz_hess, **f64)
    scratch = torch.empty(lib.rb_eval_scratch_bytes(F.handle, B), dtype=torch.uint8, device=dev)

    def step():
        F.eval_device(x_d, l_d, sig_d, vp_d, None, f_d, gf_d, g_d, j_d, h_d, scratch)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()         # samples cover warm-up, the timed region and the end-to-end leg (all under load)
    for _ in range(args.warmup):
        step()
    barrier()
    lib.rb_profile_enable(1)
    l0 = lib.rb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = int(lib.rb_launch_count() - l0)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    cell_ms, cell_n = ctypes.c_double(0), ctypes.c_int(0)
    lib.rb_profile_cell_ms(ctypes.byref(cell_ms), ctypes.byref(cell_n))
    lib.rb_profile_enable(0)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host entry point (pinned host buffers) ----------------------------
    Be = min(args.e2e_batch, B)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hx, hl, hvp = pin(X[:Be]), pin(L[:Be]), pin(VP[:Be])
    hsig = torch.ones(Be, dtype=torch.float64).pin_memory()
    hf = torch.empty(Be, dtype=torch.float64).pin_memory()
    hgf = torch.empty(Be, st.nw, dtype=torch.float64).pin_memory()
    hg = torch.empty(Be, st.ng, dtype=torch.float64).pin_memory()
    hj = torch.empty(Be, st.nnz_jac, dtype=torch.float64).pin_memory()
    hh = torch.empty(Be, st.nnz_hess, dtype=torch.float64).pin_memory()
    vptr = lambda t: ctypes.c_void_p(t.data_ptr())

    def e2e_step():
        rc = lib.rb_nlp_eval_all(F.handle, Be, vptr(hx), vptr(hvp), vptr(hsig), vptr(hl), vptr(hf), vptr(hgf),
                                 vptr(hg), vptr(hj), vptr(hh))
        if rc:
            raise RuntimeError(lib.rb_last_error().decode())

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * e2e_steps / float(dt.item())
    clocks = sampler.stop() if rank == 0 else None
    h2d = 8 * Be * (st.nw + st.ng + F.nvp + 1)
    d2h = 8 * Be * (1 + st.nw + st.ng + st.nnz_jac + st.nnz_hess)
    # spot check: the host path and the device path agree bit for bit on instance 0
    assert np.array_equal(hh[0].numpy(), h_d[0].cpu().numpy()) and np.array_equal(hj[0].numpy(), j_d[0].cpu().numpy())

    colloc = None
    if rank == 0 and not args.no_colloc:
        try:
            colloc = run_colloc_eval(dev, local_rank, max(3, args.steps // 2))
        except Exception as exc:        # a side measurement must not take the bench line down
            colloc = dict(error=repr(exc))
    single = None
    if rank == 0 and not args.no_single:
        single = run_single_solves(dev)
        if not args.no_cpu:
            try:
                single.append(run_measured_cpu_solve(dev))
            except Exception as exc:        # a side measurement must not take the bench line down
                single.append(dict(problem='measured CPU solve', error=repr(exc)))
    solves = None
    if not args.no_solves:
        del x_d, l_d, j_d, h_d, gf_d, g_d
        torch.cuda.empty_cache()
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):      # the builders print progress like the reference does
            solves = run_solves(args, dev, rank, world, dist)

    if rank == 0:
        peak, peak_src = _peaks()
        ab = algorithmic_bytes(st)
        cell_avg_ms = cell_ms.value / max(1, cell_n.value)
        achieved = ab * B / (cell_avg_ms * 1e-3) / 1e9
        fp = ctypes.c_double(0)
        lib.rb_fp64_peak(ctypes.byref(fp))
        # fp64 instructions per eval of the two shooting kernels: per interval 4 x (fJ_s + vjpW_s) straight-line
        # instructions in the point kernel, and per (interval, direction) 4 x (J dX + W dX + J' dkb) multiply-adds
        meta = __import__('aircraft_trajectory_optimization_b200.codegen', fromlist=['load_meta']).load_meta(prod.model.variant)
        nv = st.nz + st.nu + 1
        nj, nwz = len(meta['J']), len(meta['W'])
        flop_eval = st.N * (4 * (meta['ops']['fJ_s'] + meta['ops']['vjpW_s']) + nv * 4 * (2 * nj + 2 * nwz))
        # DRAM bytes of the two shooting kernels per 128-instance chunk, from the ncu --set full capture of this
        # command (profiles/r02_ncu_rk4.txt: point 31.7 + 442.6 MB, direction 453.1 + 177.1 MB), scaled to a step
        traffic_bytes = (31.672576e6 + 442.574848e6 + 453.142016e6 + 177.075968e6) * (B / 128.0)
        roofline = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                        traffic=traffic_bytes, traffic_unit='bytes per step (ncu dram__bytes_read+write of one --set full capture of this command, profiles/r02_ncu_rk4.txt; not re-measured in this run)',
                        algorithmic_bytes_per_step=ab * B, peak_source=peak_src, kernel='rk4_point_kernel + rk4_dir_kernel <PF_drone_quat_param_gr> (all chunks of a step)',
                        algorithmic_bytes_per_eval=ab, evals_per_launch=B, kernel_ms=cell_avg_ms,
                        kernel_share_of_step=cell_ms.value / total_ms,
                        fp64=dict(note='the shooting kernels are FP64 / latency bound, not HBM bound: fp64 arithmetic '
                                       'instructions per eval x evals / kernel time vs a measured FMA-rate peak',
                                  arith_instr_per_eval=flop_eval, achieved_ginstr_s=flop_eval * B / (cell_avg_ms * 1e-3) / 1e9,
                                  peak_gfma_s=fp.value * 1e3 / 2, frac=flop_eval * B / (cell_avg_ms * 1e-3) / (fp.value * 1e12 / 2)))
        cpu = None
        if not args.no_cpu:
            r, cores, sample = cpu_port_rate()
            cpu = dict(value=r, unit=UNIT, cores=cores, kind='port', sample=sample,
                       note='value = the execution model the reference runs (interpreted SX tape); `compiled` = what '
                            'C code generation of the same interval function gives, full size, no scaling')
            try:
                cpu['compiled'] = cpu_compiled_rate(st, prod.vehicle_config, prod.model.variant)
            except Exception as exc:
                cpu['compiled'] = dict(error=repr(exc))
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f64', data='synthetic',
                    config=full_config(st, B, world),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             instances_per_step=Be, api='rb_nlp_eval_all (host buffers, pinned)'),
                    gpu_launches=launches, clocks=clocks, roofline=roofline, cpu_baseline=cpu, solves=solves,
                    collocation=colloc, single_solves=single)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
