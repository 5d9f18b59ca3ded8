// RK4 shooting cells, materialised form: two kernels per batch chunk.
//
// Reference definition of what is computed: drone3d/raceline/base_raceline.py:363-391 (global) and
// :1052-1112 (parametric) -- zn = cont(rk4(f; z, u, h)), rows  Z[n+1]-zn, U[n+1]-(U+dU*h/2), zn[0]-s;
// drone3d/raceline/drone_raceline.py:42-45 (cont = quaternion renormalisation), :47-104 (closure rows
// reuse the last interval's step), drone3d/raceline/base_raceline.py:601-623 (stage cost * h).
//
//   rk4_point_kernel : one thread per (instance, interval).  Runs the primal RK4 chain and the primal
//                      adjoint chain of mu' out once, and writes per stage i = 1..4 the non-zeros of
//                      J_i = df/dx(X_i), of W_i = sum_c kb_i[c] d2f_c/dx2(X_i), the slope k_i and the
//                      adjoint xb_i = J_i' kb_i into a scratch column (entry e of cell c at
//                      scr[(c / 32 * NS + e) * 32 + c % 32]: coalesced for this kernel, and the 32 cells a
//                      warp of the second kernel touches share cache lines).
//   rk4_dir_kernel   : one thread per (instance, interval, direction v_col of (z, u, h)).  Pushes e_col
//                      through the four stages with sparse products against the stored J_i (Jacobian column)
//                      and pulls the tangent of the adjoint sweep back with W_i dX_i + J_i' dkb_i (Hessian
//                      column); then the same epilogue as before: rows, objective pieces, CCS slot writes.
//
// The straight-line point functions (fJ_s, vjpW_s) therefore run once per cell instead of once per direction,
// and the per-direction work is ~1.3 k multiply-adds against L1-resident operands.
#pragma once
#include "common.cuh"

template <class PF>
struct Rk4Scratch {
  static constexpr int NZ = PF::NZ, NJ = PF::NJ, NW = PF::NW;
  static constexpr int SS = NJ + NW + 2 * NZ;     // per stage: J, W, k, xb(z part)
  static constexpr int NS = 4 * SS;
  __host__ __device__ static constexpr int oJ(int i) { return i * SS; }
  __host__ __device__ static constexpr int oW(int i) { return i * SS + NJ; }
  __host__ __device__ static constexpr int oK(int i) { return i * SS + NJ + NW; }
  __host__ __device__ static constexpr int oX(int i) { return i * SS + NJ + NW + NZ; }
  __host__ __device__ static size_t doubles(long long cells) { return (size_t)((cells + 31) / 32) * NS * 32; }
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

template <class PF>
__global__ void __launch_bounds__(128)
rk4_point_kernel(const RbDev d, const RbBatch b, double* __restrict__ scr) {
  using SC = Rk4Scratch<PF>;
  constexpr int NZ = PF::NZ, NU = PF::NU, NX = PF::NX, NR = NZ + NU;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= (long long)b.B * d.N) return;
  const int p = (int)(cell / d.N);
  const int n = (int)(cell - (long long)p * d.N);
  const double* __restrict__ w = b.x + (size_t)p * d.nw;
  const double h = w[n];
  const double* __restrict__ zu = w + d.N + (size_t)n * (NZ + 2 * NU);
  const double* __restrict__ fcp =
      PF::USES_FC ? (b.fc_b ? b.fc_b + (size_t)p * d.N * PF::NFC : d.fc) + (size_t)n * PF::NFC : d.fc;
  const double* __restrict__ vpp = b.vp + (size_t)p * b.vp_stride;
  double* __restrict__ S = scr + ((size_t)(cell >> 5) * SC::NS) * 32 + (cell & 31);

  double x1[NX], xs[NX], k[NZ], Ks[NZ];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x1[i] = zu[i];
    xs[i] = x1[i];
  }
#pragma unroll
  for (int i = 0; i < NZ; ++i) Ks[i] = 0.0;
  // ---- primal chain: k_i = f(X_i), J_i;  X_{i+1} = x + a_{i+1} h k_i,  a = (0, 1/2, 1/2, 1),  b = (1, 2, 2, 1)
#pragma unroll 1
  for (int st = 0; st < 4; ++st) {
    double* __restrict__ Sst = S + (size_t)st * SC::SS * 32;
    PF::fJ_s(xs, fcp, vpp, k, Sst);
    const double bw = (st == 0 || st == 3) ? 1.0 : 2.0;
    const double an = (st == 2) ? h : 0.5 * h;        // a_{st+1} h
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
      Sst[(SC::NJ + SC::NW + i) * 32] = k[i];
      Ks[i] += bw * k[i];
      if (st < 3) xs[i] = x1[i] + an * k[i];
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
    double* __restrict__ Sst = S + (size_t)st * SC::SS * 32;
    const double bw = (st == 0 || st == 3) ? h6 : 2.0 * h6;
    const double an = (st == 3) ? 0.0 : ((st == 2) ? h : 0.5 * h);     // a_{st+2} h (coefficient of xb_{st+1})
    const double ax = (st == 0) ? 0.0 : ((st == 3) ? h : 0.5 * h);     // a_{st+1} h (X_st = x + ax k_{st-1})
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
      kb[i] = bw * muz[i] + an * xb[i];
      xs[i] = x1[i] + (st > 0 ? ax * Sst[((long long)(SC::NJ + SC::NW + i) - SC::SS) * 32] : 0.0);
    }
    PF::vjpW_s(xs, kb, fcp, vpp, xb, Sst + SC::NJ * 32);
#pragma unroll
    for (int i = 0; i < NZ; ++i) Sst[(SC::NJ + SC::NW + NZ + i) * 32] = xb[i];
  }
}

template <class PF>
__global__ void __launch_bounds__(RB_CELL_THREADS)
rk4_dir_kernel(const RbDev d, const RbBatch b, const double* __restrict__ scr) {
  using SC = Rk4Scratch<PF>;
  constexpr int NZ = PF::NZ, NU = PF::NU, NX = PF::NX, NV = NX + 1, NL = NV + NU, NR = NZ + NU;
  constexpr int CPB = RB_CELL_THREADS / NV;
  const int lc = threadIdx.x / NV;
  const int col = threadIdx.x - lc * NV;
  if (lc >= CPB) return;
  const long long cell = (long long)blockIdx.x * CPB + lc;
  if (cell >= (long long)b.B * d.N) return;
  const int p = (int)(cell / d.N);
  const int n = (int)(cell - (long long)p * d.N);
  const double* __restrict__ S = scr + ((size_t)(cell >> 5) * SC::NS) * 32 + (cell & 31);

  const double* __restrict__ w = b.x + (size_t)p * d.nw;
  const double h = w[n];
  const double* __restrict__ zu = w + d.N + (size_t)n * (NZ + 2 * NU);
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
  extern __shared__ double rk4_smem[];
  double* __restrict__ dks = rk4_smem + (size_t)threadIdx.x;     // dks[(st * NZ + i) * RB_CELL_THREADS]
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
    const double* __restrict__ Sst = S + (size_t)st * SC::SS * 32;
    PF::jmul(Sst, dxs, dk);
    const double bw = (st == 0 || st == 3) ? 1.0 : 2.0;
    const double an = (st == 2) ? 1.0 : 0.5;
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
      const double ki = Sst[(SC::NJ + SC::NW + i) * 32];
      Ks[i] += bw * ki;
      dKs[i] += bw * dk[i];
      if (st < 3) {
        dks[(st * NZ + i) * RB_CELL_THREADS] = dk[i];
        dxs[i] = e[i] + an * (h * dk[i] + dh * ki);
      }
    }
  }
  double zn[NZ], dzn[NZ];
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    zn[i] = zu[i] + h6 * Ks[i];
    dzn[i] = e[i] + h6 * dKs[i] + dh6 * Ks[i];
  }

  // ---------------------------------------------------------------- rows, multipliers
  const size_t cr = (size_t)n * NR;
  const double* __restrict__ lam = b.lam_g ? b.lam_g + (size_t)p * d.ng : nullptr;
  double mu[NZ];   // multiplier on out_c (already times the row coefficient)
#pragma unroll
  for (int c = 0; c < NZ; ++c) {
    const int r = d.cell_row[cr + c];
    mu[c] = (r >= 0 && lam) ? lam[r] * d.cell_coef[cr + c] : 0.0;
  }

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
  if (b.g || b.jac) {
    double* __restrict__ g = b.g ? b.g + (size_t)p * d.ng : nullptr;
    double* __restrict__ jac = b.jac ? b.jac + (size_t)p * d.nnzj : nullptr;
    const int32_t* __restrict__ js = d.cell_jslot + (size_t)n * d.cell_nj;
    if (jac) {
#pragma unroll
      for (int c = 0; c < NZ; ++c) {
        const int s = js[c * NV + col];
        if (s >= 0) jac[s] = d.cell_coef[cr + c] * dout[c];
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
        val = d.cell_coef[cr + c] * val + d.cell_off[cr + c];
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
        const double su = d.cell_coef[cr + NZ + j];
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
  if (col >= NZ && col < NX) {
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

  if (!b.hess) return;

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
    const double* __restrict__ Sst = S + (size_t)st * SC::SS * 32;
    const double bw = (st == 0 || st == 3) ? 1.0 / 6.0 : 1.0 / 3.0;
    const double an = (st == 3) ? 0.0 : ((st == 2) ? 1.0 : 0.5);       // a_{st+2}: coefficient of (xb, dxb) of stage st+1
    const double ax = (st == 0) ? 0.0 : ((st == 3) ? 1.0 : 0.5);       // a_{st+1}: X_st = x + ax h k_{st-1}
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
      const double xbn = (st < 3) ? Sst[((long long)(SC::NJ + SC::NW + NZ + i) + SC::SS) * 32] : 0.0;   // xb_{st+1}
      dkb[i] = bw * (dh * muz[i] + h * dmuz[i]) + an * (dh * xbn + h * dxb[i]);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) dxs[i] = e[i];
    if (st > 0) {
#pragma unroll
      for (int i = 0; i < NZ; ++i) {
        const double kp = Sst[((long long)(SC::NJ + SC::NW + i) - SC::SS) * 32];                          // k_{st-1}
        dxs[i] = e[i] + ax * (h * dks[((st - 1) * NZ + i) * RB_CELL_THREADS] + dh * kp);
      }
    }
    PF::wmul(Sst + SC::NJ * 32, dxs, dxb);
    PF::jtmul(Sst, dkb, dxb);
#pragma unroll
    for (int i = 0; i < NZ; ++i) gz[i] += dxb[i];
#pragma unroll
    for (int j = 0; j < NU; ++j) gu[j] += dxb[NZ + j];
    if (st > 0) {
#pragma unroll
      for (int i = 0; i < NZ; ++i) {
        const double kp = Sst[((long long)(SC::NJ + SC::NW + i) - SC::SS) * 32];
        const double xbi = Sst[(SC::NJ + SC::NW + NZ + i) * 32];
        dhb += ax * (dxb[i] * kp + xbi * dks[((st - 1) * NZ + i) * RB_CELL_THREADS]);
      }
    }
  }

  // ---------------------------------------------------------------- hess_l column `col`
  double* __restrict__ H = b.hess + (size_t)p * d.nnzh;
  const int32_t* __restrict__ hs = d.cell_hslot + (size_t)n * d.cell_nh;
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
#pragma unroll
  for (int r = 0; r < NZ; ++r) {
    if (r <= col) {
      const int s = hs[r * NL + col];
      if (s >= 0) H[s] = gz[r];
    }
  }
#pragma unroll
  for (int j = 0; j < NU; ++j) {
    if (NZ + j <= col) {
      const int s = hs[(NZ + j) * NL + col];
      if (s >= 0) H[s] = gu[j];
    }
  }
  if (col == NX) {
    const int s = hs[NX * NL + NX];
    if (s >= 0) H[s] = dhb;
  }
  // du entries are owned by the matching u thread: (du_j, du_j) and (h, du_j)
  if (col >= NZ && col < NX) {
    const int j = col - NZ;
    double duj = 0.0;
#pragma unroll
    for (int i = 0; i < NU; ++i) duj = (i == j) ? du[i] : duj;
    const int r = d.cell_row[cr + NZ + j];
    const double lu = (r >= 0 && lam) ? lam[r] * d.cell_coef[cr + NZ + j] * d.cell_par[(size_t)n * d.cell_ncp] : 0.0;
    int s = hs[(NV + j) * NL + (NV + j)];
    if (s >= 0) H[s] = sig * 2.0 * d.dR[j] * h;
    s = hs[NX * NL + (NV + j)];
    if (s >= 0) H[s] = sig * 2.0 * d.dR[j] * duj + lu;
  }
}
