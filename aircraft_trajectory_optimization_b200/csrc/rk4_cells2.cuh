// RK4 shooting cells, materialised form: two kernels per batch chunk.
//
// Reference definition of what is computed: drone3d/raceline/base_raceline.py:363-391 (global) and
// :1052-1112 (parametric) -- zn = cont(rk4(f; z, u, h)), rows  Z[n+1]-zn, U[n+1]-(U+dU*h/2), zn[0]-s;
// drone3d/raceline/drone_raceline.py:42-45 (cont = quaternion renormalisation), :47-104 (closure rows
// reuse the last interval's step), drone3d/raceline/base_raceline.py:601-623 (stage cost * h).
//
//   rk4_point_kernel : one thread per (instance, interval).  Runs the primal RK4 chain and the primal
//                      adjoint chain of mu' out once and writes, per stage i = 1..4, the non-zeros of
//                      J_i = df/dx(X_i) and W_i = sum_c kb_i[c] d2f_c/dx2(X_i), the slopes k_i, k_{i-1} and the
//                      adjoints xb_i = J_i' kb_i, xb_{i+1} into scratch.
//   rk4_dir_kernel   : one thread per (instance, interval, direction v_col of (z, u, h)), CPB cells per CTA.
//                      Pushes e_col through the four stages with sparse products against J_i (Jacobian column)
//                      and pulls the tangent of the adjoint sweep back with W_i dX_i + J_i' dkb_i (Hessian
//                      column); then rows, objective pieces and the CCS slot writes.
//
// Scratch layout: the CPB cells of one direction-kernel CTA form a group; per group 4 stage blocks of
// SB * CPB doubles, entry-major with the cells of the group adjacent.  A stage block is one contiguous,
// 16-byte aligned span, so the direction kernel fetches the eight blocks it needs (stages 1..4 forward, 4..1
// backward) with one TMA bulk copy each (cp.async.bulk + mbarrier), double-buffered in shared memory one step
// ahead of the arithmetic; all ~1.3 k operand reads of a thread are then short-latency shared-memory reads.
#pragma once
#include "common.cuh"

template <class PF>
struct Rk4Scratch {
  static constexpr int NZ = PF::NZ, NJ = PF::NJ, NW = PF::NW, CPB = PF::CPB;
  // entries of a stage block: J, W, k_i, xb_i (z part), k_{i-1}, xb_{i+1} (z part); padded to an even count so
  // that SB * CPB * 8 bytes is a multiple of 16 (TMA bulk copies)
  // J and W are stored in pairs (entries 2i, 2i+1 of a cell adjacent; pair i of local cell lc at double
  // offset (i * CPB + lc) * 2 of the region), everything else one double per (entry, cell)
  static constexpr int NJP = (NJ + 1) / 2 * 2, NWP = (NW + 1) / 2 * 2;
  static constexpr int oJ = 0, oW = NJP, oK = NJP + NWP, oX = oK + NZ, oKp = oX + NZ, oXn = oKp + NZ;
  static constexpr int SB = (oXn + NZ + 1) / 2 * 2;
  static constexpr int NS = 4 * SB;                  // doubles per cell
  static constexpr int STAGE_DOUBLES = SB * CPB;     // doubles per (group, stage) block
  // direction kernel: per-cell staging area (doubles) [z u du | h | row multipliers (z, u rows) | row coefficients]
  static constexpr int NU_ = PF::NU, NR_ = NZ + NU_;
  static constexpr int cZU = 0, cH = NZ + 2 * NU_, cMU = cH + 1, cCOEF = cMU + NR_, CB = cCOEF + NR_;
  static constexpr int NIDX = NZ + NU_ + 1;          // thread-private slot indices (Jacobian column, then Hessian column)
  static constexpr size_t dir_smem_bytes(int threads) {
    return (2 * (size_t)STAGE_DOUBLES + 3 * (size_t)NZ * threads + (size_t)CPB * CB) * sizeof(double) +
           (size_t)NIDX * threads * sizeof(int);
  }
  __host__ __device__ static size_t doubles(long long cells) { return (size_t)((cells + CPB - 1) / CPB) * NS * CPB; }
};

// multipliers of the cell's state rows pulled through cont(): muz = d(mu' cont(zn)) / d zn
template <class PF>
__device__ __forceinline__ void rk4_row_multipliers(const RbDev& d, const double* __restrict__ lam, size_t cr,
                                                    const double* zn, double* mu, double* muz) {
  constexpr int NZ = PF::NZ;
#pragma unroll
  for (int c = 0; c < NZ; ++c) {
    const int r = d.cell_row[cr + c];
    mu[c] = (r >= 0 && lam) ? lam[r] * d.cell_coef[cr + c] : 0.0;
    muz[c] = mu[c];
  }
  if (PF::QUAT) {
    const double r2 = zn[3] * zn[3] + zn[4] * zn[4] + zn[5] * zn[5] + zn[6] * zn[6];
    const double ri = 1.0 / sqrt(r2);
    double phi = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) phi += zn[3 + a] * ri * mu[3 + a];
#pragma unroll
    for (int a = 0; a < 4; ++a) muz[3 + a] = (mu[3 + a] - phi * zn[3 + a] * ri) * ri;
  }
}

#ifndef RB_POINT_MINBLOCKS
#define RB_POINT_MINBLOCKS 1
#endif
template <class PF>
__global__ void __launch_bounds__(128, RB_POINT_MINBLOCKS)
rk4_point_kernel(const RbDev d, const RbBatch b, double* __restrict__ scr) {
  using SC = Rk4Scratch<PF>;
  constexpr int NZ = PF::NZ, NU = PF::NU, NX = PF::NX, NR = NZ + NU, CPB = SC::CPB;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= (long long)b.B * d.N) return;
  // k_1..k_3 of this thread, kept for the adjoint chain (thread-private shared-memory slots)
  __shared__ double k_keep[3 * NZ * 128];
  double* __restrict__ kk = k_keep + threadIdx.x;
  const int p = (int)(cell / d.N);
  const int n = (int)(cell - (long long)p * d.N);
  const double* __restrict__ w = b.x + (size_t)p * d.nw;
  const double h = w[n];
  const double* __restrict__ zu = w + d.N + (size_t)n * (NZ + 2 * NU);
  const double* __restrict__ fcp =
      PF::USES_FC ? (b.fc_b ? b.fc_b + (size_t)p * d.N * PF::NFC : d.fc) + (size_t)n * PF::NFC : d.fc;
  const double* __restrict__ vpp = b.vp + (size_t)p * b.vp_stride;
  // entry e of stage st of this cell: S[(st * SB + e) * CPB]; pair i of J / W: (S - lc)[(st * SB + o) * CPB + 2 * (i * CPB + lc)]
  const int lc = (int)(cell % CPB);
  double* __restrict__ S = scr + (size_t)(cell / CPB) * SC::NS * CPB + lc;

  double x1[NX], xs[NX], k[NZ], Ks[NZ];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x1[i] = zu[i];
    xs[i] = x1[i];
  }
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    Ks[i] = 0.0;
    k[i] = 0.0;
  }
  // ---- primal chain: k_i = f(X_i), J_i;  X_{i+1} = x + a_{i+1} h k_i,  a = (0, 1/2, 1/2, 1),  b = (1, 2, 2, 1)
#pragma unroll 1
  for (int st = 0; st < 4; ++st) {
    double* __restrict__ Sst = S + (size_t)st * SC::SB * CPB;
#pragma unroll
    for (int i = 0; i < NZ; ++i) Sst[(SC::oKp + i) * CPB] = k[i];          // k_{st-1} (zero for the first stage)
    PF::fJ_s(xs, fcp, vpp, k, Sst + SC::oJ * CPB + lc);
    const double bw = (st == 0 || st == 3) ? 1.0 : 2.0;
    const double an = (st == 2) ? h : 0.5 * h;        // a_{st+2} h
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
      Sst[(SC::oK + i) * CPB] = k[i];
      Ks[i] += bw * k[i];
      if (st < 3) {
        kk[(st * NZ + i) * 128] = k[i];
        xs[i] = x1[i] + an * k[i];
      }
    }
  }
  if (!b.hess) return;

  // ---- primal adjoint chain of mu' cont(zn): kb_i = (b_i h / 6) muz + a_{i+1} h xb_{i+1},  xb_i = J_i' kb_i,  W_i
  const double h6 = h / 6.0;
  double zn[NZ], mu[NZ], muz[NZ], kb[NZ], xb[NX];
#pragma unroll
  for (int i = 0; i < NZ; ++i) zn[i] = x1[i] + h6 * Ks[i];
  const double* __restrict__ lam = b.lam_g ? b.lam_g + (size_t)p * d.ng : nullptr;
  rk4_row_multipliers<PF>(d, lam, (size_t)n * NR, zn, mu, muz);
#pragma unroll
  for (int i = 0; i < NX; ++i) xb[i] = 0.0;
#pragma unroll 1
  for (int st = 3; st >= 0; --st) {
    double* __restrict__ Sst = S + (size_t)st * SC::SB * CPB;
    const double bw = (st == 0 || st == 3) ? h6 : 2.0 * h6;
    const double an = (st == 3) ? 0.0 : ((st == 2) ? h : 0.5 * h);     // a_{st+2} h (coefficient of xb_{st+1})
    const double ax = (st == 0) ? 0.0 : ((st == 3) ? h : 0.5 * h);     // a_{st+1} h (X_st = x + ax k_{st-1})
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
      Sst[(SC::oXn + i) * CPB] = xb[i];                                  // xb_{st+1} (zero for the last stage)
      kb[i] = bw * muz[i] + an * xb[i];
      xs[i] = x1[i] + ax * (st > 0 ? kk[((st - 1) * NZ + i) * 128] : 0.0);
    }
    PF::vjpW_s(xs, kb, fcp, vpp, xb, Sst + SC::oW * CPB + lc);
#pragma unroll
    for (int i = 0; i < NZ; ++i) Sst[(SC::oX + i) * CPB] = xb[i];
  }
}

#ifndef RB_DIR_MINBLOCKS
#define RB_DIR_MINBLOCKS 3
#endif
template <class PF>
__global__ void __launch_bounds__(RB_CELL_THREADS, RB_DIR_MINBLOCKS)
rk4_dir_kernel(const RbDev d, const RbBatch b, const double* __restrict__ scr) {
  using SC = Rk4Scratch<PF>;
  constexpr int NZ = PF::NZ, NU = PF::NU, NX = PF::NX, NV = NX + 1, NL = NV + NU, NR = NZ + NU;
  constexpr int CPB = SC::CPB;
  static_assert(CPB == RB_CELL_THREADS / NV, "cells per CTA");
  constexpr unsigned STAGE_BYTES = SC::STAGE_DOUBLES * sizeof(double);
  static_assert(STAGE_BYTES % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
  // shared memory: two stage buffers (TMA destinations), then [3 NZ][RB_CELL_THREADS] thread-private tangents
  extern __shared__ __align__(128) double rk4_smem[];
  __shared__ __align__(8) unsigned long long rk4_bar[2];
  double* __restrict__ dks = rk4_smem + 2 * SC::STAGE_DOUBLES + (size_t)threadIdx.x;
  const double* __restrict__ grp = scr + (size_t)blockIdx.x * SC::NS * CPB;     // this CTA's group
  // step s = 0..7 reads stage (s < 4 ? s : 7 - s) from buffer s & 1
  auto issue = [&](int step) {
    const int stg = step < 4 ? step : 7 - step;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&rk4_bar[step & 1]);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(rk4_smem + (size_t)(step & 1) * SC::STAGE_DOUBLES);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STAGE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(grp + (size_t)stg * SC::STAGE_DOUBLES), "r"(STAGE_BYTES), "r"(bar)
                 : "memory");
  };
  auto wait = [&](int step) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&rk4_bar[step & 1]);
    const unsigned parity = (step >> 1) & 1;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RK4_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RK4_DONE_%=;\n"
        "bra RK4_WAIT_%=;\n"
        "RK4_DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
  };
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&rk4_bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&rk4_bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool want_h = b.hess != nullptr;
  const int nsteps = want_h ? 8 : 4;
  if (threadIdx.x == 0) {
    issue(0);
    issue(1);
  }
  const int lc = threadIdx.x / NV;
  const int col = threadIdx.x - lc * NV;
  const long long cell = (long long)blockIdx.x * CPB + lc;
  // threads without a cell still take part in the barriers of the pipeline
  const bool live = lc < CPB && cell < (long long)b.B * d.N;
  const long long cell_c = live ? cell : 0;
  const int p = (int)(cell_c / d.N);
  const int n = (int)(cell_c - (long long)p * d.N);

  const double* __restrict__ w = b.x + (size_t)p * d.nw;
  const size_t cr = (size_t)n * NR;
  const double* __restrict__ lam = b.lam_g ? b.lam_g + (size_t)p * d.ng : nullptr;
  const int lcs = live ? lc : 0;
  // Per-cell operands every direction thread of the cell needs (z, u, du, h, row multipliers, row coefficients)
  // are fetched once per cell by its threads together and read back from shared memory; the slot indices of
  // this thread's Jacobian / Hessian column arrive by cp.async in thread-private shared-memory slots.  All of
  // these global-memory latencies overlap the first TMA stage instead of stalling the sweeps one by one.
  double* __restrict__ cellbuf = rk4_smem + 2 * SC::STAGE_DOUBLES + 3 * NZ * RB_CELL_THREADS;
  int* __restrict__ idx = reinterpret_cast<int*>(cellbuf + CPB * SC::CB) + threadIdx.x;
  double* __restrict__ cb = cellbuf + lcs * SC::CB;
  if (live) {
    const double* __restrict__ zu_g = w + d.N + (size_t)n * (NZ + 2 * NU);
    for (int it = col; it < SC::CB; it += NV) {
      double v;
      if (it < SC::cH) {
        v = zu_g[it];
      } else if (it == SC::cH) {
        v = w[n];
      } else if (it < SC::cCOEF) {
        const int c = it - SC::cMU;
        const int r = d.cell_row[cr + c];
        v = (r >= 0 && lam) ? lam[r] * d.cell_coef[cr + c] : 0.0;
      } else {
        v = d.cell_coef[cr + it - SC::cCOEF];
      }
      cb[it] = v;
    }
  }
  auto fetch_idx = [&](int k, const int32_t* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(idx + k * RB_CELL_THREADS)),
                 "l"(src)
                 : "memory");
  };
  {
    const int32_t* __restrict__ js0 = d.cell_jslot + (size_t)n * d.cell_nj;
#pragma unroll
    for (int c = 0; c < NZ; ++c) {
      if (b.jac) fetch_idx(c, js0 + c * NV + col);
      else idx[c * RB_CELL_THREADS] = -1;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __syncthreads();
  const double h = cb[SC::cH];
  const double* __restrict__ zu = cb;                 // z, u, du of this cell
  double du[NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) du[j] = zu[NX + j];

  // direction of this thread
  const double dh = (col == NX) ? 1.0 : 0.0;
  double e[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) e[i] = (i == col) ? 1.0 : 0.0;

  // ---------------------------------------------------------------- forward sweep: dk_i = J_i dX_i
  // dX_i = e + a_i (h dk_{i-1} + dh k_{i-1}),  a = (0, 1/2, 1/2, 1);  dk_1..dk_3 are kept in this thread's
  // shared-memory slots for the reverse sweep
  const double h6 = h / 6.0, dh6 = dh / 6.0;
  double dxs[NX], dk[NZ], Ks[NZ], dKs[NZ];
#pragma unroll
  for (int i = 0; i < NX; ++i) dxs[i] = e[i];
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    Ks[i] = 0.0;
    dKs[i] = 0.0;
  }
#pragma unroll 1
  for (int st = 0; st < 4; ++st) {
    wait(st);
    const double* __restrict__ Sst = rk4_smem + (size_t)(st & 1) * SC::STAGE_DOUBLES + lcs;
    PF::jmul(Sst + SC::oJ * CPB + lcs, dxs, dk);
    const double bw = (st == 0 || st == 3) ? 1.0 : 2.0;
    const double an = (st == 2) ? 1.0 : 0.5;
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
      const double ki = Sst[(SC::oK + i) * CPB];
      Ks[i] += bw * ki;
      dKs[i] += bw * dk[i];
      if (st < 3) {
        dks[(st * NZ + i) * RB_CELL_THREADS] = dk[i];
        dxs[i] = e[i] + an * (h * dk[i] + dh * ki);
      }
    }
    __syncthreads();                                  // everyone is done with this buffer
    if (threadIdx.x == 0 && st + 2 < nsteps) issue(st + 2);
  }
  double zn[NZ], dzn[NZ];
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    zn[i] = zu[i] + h6 * Ks[i];
    dzn[i] = e[i] + h6 * dKs[i] + dh6 * Ks[i];
  }

  // ---------------------------------------------------------------- rows, multipliers
  double mu[NZ];   // multiplier on out_c (already times the row coefficient)
#pragma unroll
  for (int c = 0; c < NZ; ++c) mu[c] = cb[SC::cMU + c];

  // cont(): quaternion renormalisation of zn[3:7]
  double out[NZ], dout[NZ], muz[NZ], dmuz[NZ];
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    out[i] = zn[i];
    dout[i] = dzn[i];
    muz[i] = mu[i];
    dmuz[i] = 0.0;
  }
  if (PF::QUAT) {
    const double r2 = zn[3] * zn[3] + zn[4] * zn[4] + zn[5] * zn[5] + zn[6] * zn[6];
    const double ri = 1.0 / sqrt(r2);
    double nq[4];
    double nd = 0.0, phi = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      nq[a] = zn[3 + a] * ri;
      out[3 + a] = nq[a];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      nd += nq[a] * dzn[3 + a];
      phi += nq[a] * mu[3 + a];
    }
    double mud = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) mud += mu[3 + a] * dzn[3 + a];
    const double ri2 = ri * ri;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      dout[3 + a] = (dzn[3 + a] - nq[a] * nd) * ri;
      muz[3 + a] = (mu[3 + a] - phi * nq[a]) * ri;
      // [-(mu n' + n mu') - phi I + 3 phi n n'] dq / r^2
      dmuz[3 + a] = (-(mu[3 + a] * nd + nq[a] * mud) - phi * dzn[3 + a] + 3.0 * phi * nq[a] * nd) * ri2;
    }
  }

  // ---------------------------------------------------------------- g, jac_g
  if (live && (b.g || b.jac)) {
    double* __restrict__ g = b.g ? b.g + (size_t)p * d.ng : nullptr;
    double* __restrict__ jac = b.jac ? b.jac + (size_t)p * d.nnzj : nullptr;
    const int32_t* __restrict__ js = d.cell_jslot + (size_t)n * d.cell_nj;
    if (jac) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int c = 0; c < NZ; ++c) {
        const int s = idx[c * RB_CELL_THREADS];
        if (s >= 0) jac[s] = cb[SC::cCOEF + c] * dout[c];
      }
    }
    if (col < NZ) {
      // thread c also owns row c: value and the partner entry
      const int c = col;
      const int r = d.cell_row[cr + c];
      if (r >= 0) {
        const int pv = d.cell_partner[cr + c];
        double val = 0.0;
#pragma unroll
        for (int i = 0; i < NZ; ++i) val = (i == c) ? out[i] : val;
        val = cb[SC::cCOEF + c] * val + d.cell_off[cr + c];
        if (pv >= 0) val += d.cell_pcoef[cr + c] * w[pv];
        if (g) g[r] = val;
        if (jac) {
          const int s = js[NZ * NV + c];
          if (s >= 0) jac[s] = d.cell_pcoef[cr + c];
        }
      }
    } else if (col < NX) {
      // thread NZ+j owns input row j:  su * (u_j + du_j * h * ducoef) + pcoef * w[partner]
      const int j = col - NZ;
      const int r = d.cell_row[cr + NZ + j];
      if (r >= 0) {
        const double su = cb[SC::cCOEF + NZ + j];
        const double dc = d.cell_par[(size_t)n * d.cell_ncp];
        double uj = 0.0, duj = 0.0;
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          uj = (i == j) ? zu[NZ + i] : uj;
          duj = (i == j) ? du[i] : duj;
        }
        const int pv = d.cell_partner[cr + NZ + j];
        double val = su * (uj + duj * h * dc) + d.cell_off[cr + NZ + j];
        if (pv >= 0) val += d.cell_pcoef[cr + NZ + j] * w[pv];
        if (g) g[r] = val;
        if (jac) {
          const int32_t* ju = js + NZ * NV + NZ + 4 * j;
          if (ju[0] >= 0) jac[ju[0]] = su;
          if (ju[1] >= 0) jac[ju[1]] = su * h * dc;
          if (ju[2] >= 0) jac[ju[2]] = su * duj * dc;
          if (ju[3] >= 0) jac[ju[3]] = d.cell_pcoef[cr + NZ + j];
        }
      }
    }
  }

  // ---------------------------------------------------------------- objective pieces
  const double sig = b.lam_f ? b.lam_f[p] : 1.0;
  if (!live) {
  } else if (col >= NZ && col < NX) {
    const int j = col - NZ;
    double uj = 0.0, duj = 0.0;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      uj = (i == j) ? zu[NZ + i] : uj;
      duj = (i == j) ? du[i] : duj;
    }
    if (b.grad_f) {
      double* __restrict__ gf = b.grad_f + (size_t)p * d.nw + d.N + (size_t)n * (NZ + 2 * NU);
      gf[NZ + j] = 2.0 * d.R[j] * uj * h;
      gf[NX + j] = 2.0 * d.dR[j] * duj * h;
    }
  } else if (col < NZ) {
    if (b.grad_f) b.grad_f[(size_t)p * d.nw + d.N + (size_t)n * (NZ + 2 * NU) + col] = 0.0;
  } else {
    double stage = 1.0;
#pragma unroll
    for (int j = 0; j < NU; ++j) stage += d.R[j] * zu[NZ + j] * zu[NZ + j] + d.dR[j] * du[j] * du[j];
    if (b.grad_f) b.grad_f[(size_t)p * d.nw + n] = stage;
    if (b.fpart) b.fpart[(size_t)p * d.N + n] = stage * h;
  }

  if (!want_h) return;

  // slot indices of this thread's Hessian column: issued before the reverse sweep (into the slots the Jacobian
  // column's indices occupied), consumed after it
  const int32_t* __restrict__ hs = d.cell_hslot + (size_t)n * d.cell_nh;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int r = 0; r < NX; ++r) {
    if (r <= col) fetch_idx(r, hs + r * NL + col);
    else idx[r * RB_CELL_THREADS] = -1;
  }
  if (col == NX) fetch_idx(NX, hs + NX * NL + NX);
  else idx[NX * RB_CELL_THREADS] = -1;
  asm volatile("cp.async.commit_group;" ::: "memory");
  // ---------------------------------------------------------------- reverse sweep: dxb_i = W_i dX_i + J_i' dkb_i
  // kb_i = (b_i h / 6) muz + a_{i+1} h xb_{i+1};  dkb_i is its tangent;  hbar gathers a_i (dxb_i . k_{i-1} + xb_i . dk_{i-1})
  double dkb[NZ], dxb[NX];
  double gz[NZ], gu[NU];   // this thread's Hessian column: d/dv_col of (zbar, ubar); hbar below
  double dhb = 0.0;
#pragma unroll
  for (int i = 0; i < NZ; ++i) gz[i] = dmuz[i];
#pragma unroll
  for (int j = 0; j < NU; ++j) gu[j] = 0.0;
  // h-bar terms that involve only forward quantities: (1/6)(dmu.K + mu.dK)
#pragma unroll
  for (int i = 0; i < NZ; ++i) dhb += (dmuz[i] * Ks[i] + muz[i] * dKs[i]) * (1.0 / 6.0);
#pragma unroll
  for (int i = 0; i < NX; ++i) dxb[i] = 0.0;
#pragma unroll 1
  for (int st = 3; st >= 0; --st) {
    const int step = 7 - st;
    wait(step);
    const double* __restrict__ Sst = rk4_smem + (size_t)(step & 1) * SC::STAGE_DOUBLES + lcs;
    const double bw = (st == 0 || st == 3) ? 1.0 / 6.0 : 1.0 / 3.0;
    const double an = (st == 3) ? 0.0 : ((st == 2) ? 1.0 : 0.5);       // a_{st+2}: coefficient of (xb, dxb) of stage st+1
    const double ax = (st == 0) ? 0.0 : ((st == 3) ? 1.0 : 0.5);       // a_{st+1}: X_st = x + ax h k_{st-1}
#pragma unroll
    for (int i = 0; i < NZ; ++i)
      dkb[i] = bw * (dh * muz[i] + h * dmuz[i]) + an * (dh * Sst[(SC::oXn + i) * CPB] + h * dxb[i]);
#pragma unroll
    for (int i = 0; i < NX; ++i) dxs[i] = e[i];
    if (st > 0) {
#pragma unroll
      for (int i = 0; i < NZ; ++i)
        dxs[i] = e[i] + ax * (h * dks[((st - 1) * NZ + i) * RB_CELL_THREADS] + dh * Sst[(SC::oKp + i) * CPB]);
    }
    PF::wmul(Sst + SC::oW * CPB + lcs, dxs, dxb);
    PF::jtmul(Sst + SC::oJ * CPB + lcs, dkb, dxb);
#pragma unroll
    for (int i = 0; i < NZ; ++i) gz[i] += dxb[i];
#pragma unroll
    for (int j = 0; j < NU; ++j) gu[j] += dxb[NZ + j];
    if (st > 0) {
#pragma unroll
      for (int i = 0; i < NZ; ++i)
        dhb += ax * (dxb[i] * Sst[(SC::oKp + i) * CPB] + Sst[(SC::oX + i) * CPB] * dks[((st - 1) * NZ + i) * RB_CELL_THREADS]);
    }
    __syncthreads();
    if (threadIdx.x == 0 && step + 2 < 8) issue(step + 2);
  }
  if (!live) return;

  // ---------------------------------------------------------------- hess_l column `col`
  double* __restrict__ H = b.hess + (size_t)p * d.nnzh;
  // objective and input-row bilinear terms that land in this column
  if (col >= NZ && col < NX) {
    const int j = col - NZ;
#pragma unroll
    for (int i = 0; i < NU; ++i) gu[i] += (i == j) ? sig * 2.0 * d.R[j] * h : 0.0;
  } else if (col == NX) {
#pragma unroll
    for (int j = 0; j < NU; ++j) gu[j] += sig * 2.0 * d.R[j] * zu[NZ + j];
  }
  // pairs (r, col) with r <= col in local order (z, u, h)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int r = 0; r < NZ; ++r) {
    const int s = idx[r * RB_CELL_THREADS];
    if (s >= 0) H[s] = gz[r];
  }
#pragma unroll
  for (int j = 0; j < NU; ++j) {
    const int s = idx[(NZ + j) * RB_CELL_THREADS];
    if (s >= 0) H[s] = gu[j];
  }
  {
    const int s = idx[NX * RB_CELL_THREADS];
    if (s >= 0) H[s] = dhb;
  }
  // du entries are owned by the matching u thread: (du_j, du_j) and (h, du_j)
  if (col >= NZ && col < NX) {
    const int j = col - NZ;
    double duj = 0.0;
#pragma unroll
    for (int i = 0; i < NU; ++i) duj = (i == j) ? du[i] : duj;
    const double lu = cb[SC::cMU + NZ + j] * d.cell_par[(size_t)n * d.cell_ncp];
    int s = hs[(NV + j) * NL + (NV + j)];
    if (s >= 0) H[s] = sig * 2.0 * d.dR[j] * h;
    s = hs[NX * NL + (NV + j)];
    if (s >= 0) H[s] = sig * 2.0 * d.dR[j] * duj + lu;
  }
}
