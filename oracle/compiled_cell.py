'''
Compiled straight-line CPU baseline for the shooting interval.  TEST / BENCH INFRASTRUCTURE ONLY (oracle/__init__.py).

BASELINE.md plan B: next to the flat-tape interpreter (oracle/sxvm.c, the execution model of CasADi's SX virtual
machine that the reference actually runs -- ca.nlpsol at drone3d/raceline/base_raceline.py:799 is created without
jit / code-generation options), this module times what CasADi's C code generation + an optimising compiler would
give: the whole RK4 shooting interval  out = cont(rk4(f; z, u, h))  (base_raceline.py:1052-1112,
dynamics_model.py:91-114, drone_raceline.py:42-45) differentiated as ONE expression graph -- Jacobian of out and
Hessian of mu' out with respect to (z, u, h), common sub-expressions shared -- emitted as one straight-line C function
and compiled with `gcc -O1 -march=x86-64-v3` (47 s for the 34 k statements of the
quaternion drone; -O2 takes twice as long to compile and runs no faster).  One NLP evaluation = N calls (one per interval, each with its own frame
constants) whose results are stored contiguously; the CCS scatter and the few hundred simple rows are left out, so
the number is a lower bound of the CPU time per evaluation (it favours the CPU).
'''
import ctypes
import hashlib
import os
import subprocess
import time

import numpy as np

from aircraft_trajectory_optimization_b200 import symbolic as sx
from aircraft_trajectory_optimization_b200.models import Variant, zdot, NFC

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = os.path.join(_HERE, '_cache')

_DRIVER = r'''
#include <pthread.h>
#include <stdlib.h>
typedef struct { int N, nin, nout, B; const double* in; double* out; int* next; pthread_mutex_t* mu; } job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (;;) {
    pthread_mutex_lock(j->mu);
    int p = (*j->next)++;
    pthread_mutex_unlock(j->mu);
    if (p >= j->B) break;
    const double* in = j->in + (size_t)p * j->N * j->nin;
    double* out = j->out + (size_t)p * j->N * j->nout;
    for (int n = 0; n < j->N; ++n) cell(in + (size_t)n * j->nin, out + (size_t)n * j->nout);
  }
  return 0;
}
/* B instances x N intervals; in [B][N][nin], out [B][N][nout]; one instance per thread at a time */
void eval_batch(int B, int N, int nin, int nout, const double* in, double* out, int nthreads) {
  int next = 0;
  pthread_mutex_t mu;
  pthread_mutex_init(&mu, 0);
  job_t job = {N, nin, nout, B, in, out, &next, &mu};
  pthread_t th[256];
  if (nthreads > 256) nthreads = 256;
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], 0, worker, &job);
  worker(&job);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], 0);
  pthread_mutex_destroy(&mu);
}
'''


def build_cell_graph(variant: Variant):
    ''' (graph, input ids, output ids): inputs [z, u, h, fc, vp, mu], outputs [out (nz), nnz(J), nnz(triu H)] '''
    nz, nu = variant.nz, variant.nu
    g = sx.new_graph()
    S = lambda name, n: [sx.SX(g.input(f'{name}{k}')) for k in range(n)]
    z, u = S('z', nz), S('u', nu)
    h = sx.SX(g.input('h'))
    fc, vp, mu = S('fc', NFC), S('vp', len(variant.vp_names)), S('mu', nz)
    f = lambda zz: zdot(variant, zz, u, fc, vp)
    k1 = f(z)
    k2 = f([z[i] + h / 2 * k1[i] for i in range(nz)])
    k3 = f([z[i] + h / 2 * k2[i] for i in range(nz)])
    k4 = f([z[i] + h * k3[i] for i in range(nz)])
    zn = [z[i] + h / 6 * (k1[i] + k2[i] * 2 + k3[i] * 2 + k4[i]) for i in range(nz)]
    if variant.vehicle == 'drone' and variant.orient == 'quat':
        nrm = sx.norm_2(zn[3:7])
        zn[3:7] = [e / nrm for e in zn[3:7]]
    out = [sx._id(e) for e in zn]
    v = [e.i for e in (*z, *u, h)]
    wrt = {n: c for c, n in enumerate(v)}
    rows = g.forward_sparse(out, wrt)
    J = [n for d in rows for _, n in sorted(d.items())]
    L = g.sum([g.mul(m.i, o) for m, o in zip(mu, out)])
    adj = g.reverse([L], [g.one])
    hrows = g.forward_sparse([adj.get(n, g.zero) for n in v], wrt)
    W = {}
    for i, d in enumerate(hrows):
        for j, n in d.items():
            W.setdefault((min(i, j), max(i, j)), n)
    H = [W[k] for k in sorted(W)]
    return g, out + J + H, (nz, len(J), len(H))


class CompiledCell:
    def __init__(self, variant: Variant, opt='-O1'):
        g, outs, (nz, nj, nh) = build_cell_graph(variant)
        self.nin, self.nout = len(g.input_names), len(outs)
        self.n_ops = len(g.reachable(outs))
        code, _ = g.emit_c(outs, [f'in[{k}]' for k in range(self.nin)], [f'out[{k}]' for k in range(self.nout)])
        src = ('#include <math.h>\nstatic void cell(const double* __restrict__ in, double* __restrict__ out) {\n'
               + code + '\n}\n' + _DRIVER)
        os.makedirs(_CACHE, exist_ok=True)
        tag = hashlib.sha256((src + opt).encode()).hexdigest()[:16]
        so = os.path.join(_CACHE, f'cell_{variant.name}_{tag}.so')
        self.compile_s = 0.0
        if not os.path.exists(so):
            c = os.path.join(_CACHE, f'cell_{variant.name}_{tag}.c')
            with open(c, 'w') as fh:
                fh.write(src)
            t0 = time.time()
            subprocess.check_call(['/usr/bin/gcc', opt, '-march=x86-64-v3', '-fPIC', '-shared', '-o', so, c, '-lm', '-lpthread'])
            self.compile_s = time.time() - t0
        self.lib = ctypes.CDLL(so)
        dp = ctypes.POINTER(ctypes.c_double)
        self.lib.eval_batch.argtypes = [ctypes.c_int] * 4 + [dp, dp, ctypes.c_int]
        self.lib.eval_batch.restype = None
        self.sizes = (nz, nj, nh)

    def eval_batch(self, IN, nthreads):
        ''' IN [B][N][nin] -> [B][N][nout] '''
        B, N, _ = IN.shape
        IN = np.ascontiguousarray(IN, dtype=np.float64)
        out = np.empty((B, N, self.nout))
        dp = ctypes.POINTER(ctypes.c_double)
        self.lib.eval_batch(B, N, self.nin, self.nout, IN.ctypes.data_as(dp), out.ctypes.data_as(dp), nthreads)
        return out


def cell_inputs(st, vp, B, seed0=0):
    ''' per-interval inputs [z, u, h, fc, vp, mu] of B seeded eval points of an RK4 problem structure (host side) '''
    N, nz, nu = st.N, st.nz, st.nu
    S = nz + 2 * nu
    fc = np.ones((N, NFC)) if st.fc is None else np.asarray(st.fc).reshape(N, -1)[:, :NFC]
    IN = np.empty((B, N, nz + nu + 1 + NFC + len(vp) + nz))
    for b in range(B):
        rng = np.random.default_rng(seed0 + b)
        x = np.clip(st.w0 + 1e-2 * rng.standard_normal(st.nw), st.lbw, st.ubw)
        zu = x[N:].reshape(N, S)
        IN[b, :, :nz + nu] = zu[:, :nz + nu]
        IN[b, :, nz + nu] = x[:N]
        IN[b, :, nz + nu + 1:nz + nu + 1 + NFC] = fc
        IN[b, :, nz + nu + 1 + NFC:nz + nu + 1 + NFC + len(vp)] = vp
        IN[b, :, -nz:] = rng.standard_normal((N, nz))
    return IN


def compiled_rate(st, vp, variant: Variant, seconds_budget=8.0, nthreads=0):
    ''' (evals/s, cores, sample text): full-size evaluations (all N intervals) of the compiled cell on all cores '''
    cores = nthreads or len(os.sched_getaffinity(0))
    C = CompiledCell(variant)
    B = max(2 * cores, 8)
    IN = cell_inputs(st, vp, B)
    C.eval_batch(IN[:cores], cores)
    t0 = time.time()
    done = 0
    while time.time() - t0 < seconds_budget:
        C.eval_batch(IN, cores)
        done += B
    dt = time.time() - t0
    sample = (f'compiled straight-line C (gcc -O1 -march=x86-64-v3) of the whole shooting interval: value, Jacobian and '
              f'Hessian of mu\' out in one function of {C.n_ops} operations; {done} full evaluations ({st.N} intervals each, '
              f'no scaling) in {dt:.1f} s on {cores} threads, one instance per thread; CCS scatter and simple rows left out')
    return done / dt, cores, sample
