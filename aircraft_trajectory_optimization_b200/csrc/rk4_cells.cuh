// RK4 shooting-cell kernel: one thread per (problem, interval, direction).
//
// Reference definition of what is computed: drone3d/raceline/base_raceline.py:363-391 (global) and
// :1052-1112 (parametric) -- zn = cont(rk4(f; z, u, h)), rows  Z[n+1]-zn, U[n+1]-(U+dU*h/2), zn[0]-s;
// drone3d/raceline/drone_raceline.py:42-45 (cont = quaternion renormalisation), :47-104 (closure rows
// reuse the last interval's step), drone3d/raceline/base_raceline.py:601-623 (stage cost * h).
//
// Mapping.  A cell owns the local variables v = (z[NZ], u[NU], h) of interval n (+ du[NU], which only
// enters bilinearly with h).  Thread `col` in [0, NV) carries direction e_col through the four stages:
//   forward : generated jvp  -> k_i and dk_i = d k_i / d v_col                 (Jacobian column)
//   reverse : generated hvp  -> tangent of the adjoint sweep = d/dv_col grad(mu' out)  (Hessian column)
// Nothing is exchanged between threads; all per-thread state lives in registers.  Results go straight
// to their CCS positions through the per-cell slot tables (column `col` of jac_g / hess_l is a
// contiguous run in CCS, so each thread writes a contiguous run).
#pragma once
#include "common.cuh"

template <class PF>
__global__ void __launch_bounds__(RB_CELL_THREADS)
rk4_cells_kernel(const RbDev d, const RbBatch b) {
  constexpr int NZ = PF::NZ, NU = PF::NU, NX = PF::NX, NV = NX + 1, NL = NV + NU, NR = NZ + NU;
  constexpr int CPB = RB_CELL_THREADS / NV;
  const int lc = threadIdx.x / NV;
  const int col = threadIdx.x - lc * NV;
  if (lc >= CPB) return;
  const long long cell = (long long)blockIdx.x * CPB + lc;
  if (cell >= (long long)b.B * d.N) return;
  const int p = (int)(cell / d.N);
  const int n = (int)(cell - (long long)p * d.N);

  const double* __restrict__ w = b.x + (size_t)p * d.nw;
  const double h = w[n];
  const double* __restrict__ zu = w + d.N + (size_t)n * (NZ + 2 * NU);

  double x1[NX], du[NU];
#pragma unroll
  for (int i = 0; i < NX; ++i) x1[i] = zu[i];
#pragma unroll
  for (int j = 0; j < NU; ++j) du[j] = zu[NX + j];

  double fc[PF::NFC], vp[PF::NVP];
  if (PF::USES_FC) {
    const double* __restrict__ fcp =
        (b.fc_b ? b.fc_b + (size_t)p * d.N * PF::NFC : d.fc) + (size_t)n * PF::NFC;
#pragma unroll
    for (int i = 0; i < PF::NFC; ++i) fc[i] = fcp[i];
  } else {
#pragma unroll
    for (int i = 0; i < PF::NFC; ++i) fc[i] = 0.0;
  }
  {
    const double* __restrict__ vpp = b.vp + (size_t)p * b.vp_stride;
#pragma unroll
    for (int i = 0; i < PF::NVP; ++i) vp[i] = vpp[i];
  }

  // direction of this thread
  const double dh = (col == NX) ? 1.0 : 0.0;
  double e[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) e[i] = (i == col) ? 1.0 : 0.0;

  // ---------------------------------------------------------------- forward sweep (4 x jvp)
  double k1[NZ], k2[NZ], k3[NZ], k4[NZ], dk1[NZ], dk2[NZ], dk3[NZ], dk4[NZ];
  double xs[NX], dxs[NX];
  PF::jvp(x1, e, fc, vp, k1, dk1);
  const double hh = 0.5 * h, hdh = 0.5 * dh;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    xs[i] = x1[i];
    dxs[i] = e[i];
  }
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    xs[i] = x1[i] + hh * k1[i];
    dxs[i] = e[i] + hh * dk1[i] + hdh * k1[i];
  }
  PF::jvp(xs, dxs, fc, vp, k2, dk2);
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    xs[i] = x1[i] + hh * k2[i];
    dxs[i] = e[i] + hh * dk2[i] + hdh * k2[i];
  }
  PF::jvp(xs, dxs, fc, vp, k3, dk3);
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    xs[i] = x1[i] + h * k3[i];
    dxs[i] = e[i] + h * dk3[i] + dh * k3[i];
  }
  PF::jvp(xs, dxs, fc, vp, k4, dk4);

  const double h6 = h / 6.0, dh6 = dh / 6.0;
  double Ks[NZ], dKs[NZ], zn[NZ], dzn[NZ];
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    Ks[i] = k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i];
    dKs[i] = dk1[i] + 2.0 * dk2[i] + 2.0 * dk3[i] + dk4[i];
    zn[i] = x1[i] + h6 * Ks[i];
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
          uj = (i == j) ? x1[NZ + i] : uj;
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
      uj = (i == j) ? x1[NZ + i] : uj;
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
    for (int j = 0; j < NU; ++j) stage += d.R[j] * x1[NZ + j] * x1[NZ + j] + d.dR[j] * du[j] * du[j];
    if (b.grad_f) b.grad_f[(size_t)p * d.nw + n] = stage;
    if (b.fpart) b.fpart[(size_t)p * d.N + n] = stage * h;
  }

  if (!b.hess) return;

  // ---------------------------------------------------------------- reverse sweep (4 x hvp)
  double kb[NZ], dkb[NZ], xb[NX], dxb[NX];
  double gz[NZ], gu[NU];   // this thread's Hessian column: d/dv_col of (zbar, ubar); hbar below
  double dhb = 0.0;
#pragma unroll
  for (int i = 0; i < NZ; ++i) gz[i] = dmuz[i];
#pragma unroll
  for (int j = 0; j < NU; ++j) gu[j] = 0.0;
  // h-bar terms that involve only forward quantities: (1/6)(dmu.K + mu.dK)
#pragma unroll
  for (int i = 0; i < NZ; ++i) dhb += (dmuz[i] * Ks[i] + muz[i] * dKs[i]) * (1.0 / 6.0);

  // stage 4
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    kb[i] = h6 * muz[i];
    dkb[i] = dh6 * muz[i] + h6 * dmuz[i];
    xs[i] = x1[i] + h * k3[i];
    dxs[i] = e[i] + h * dk3[i] + dh * k3[i];
  }
  PF::hvp(xs, dxs, kb, dkb, fc, vp, xb, dxb);
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    gz[i] += dxb[i];
    dhb += dxb[i] * k3[i] + xb[i] * dk3[i];
  }
#pragma unroll
  for (int j = 0; j < NU; ++j) gu[j] += dxb[NZ + j];
  // stage 3
  const double h3 = h / 3.0, dh3 = dh / 3.0;
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    kb[i] = h3 * muz[i] + h * xb[i];
    dkb[i] = dh3 * muz[i] + h3 * dmuz[i] + dh * xb[i] + h * dxb[i];
    xs[i] = x1[i] + hh * k2[i];
    dxs[i] = e[i] + hh * dk2[i] + hdh * k2[i];
  }
  PF::hvp(xs, dxs, kb, dkb, fc, vp, xb, dxb);
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    gz[i] += dxb[i];
    dhb += 0.5 * (dxb[i] * k2[i] + xb[i] * dk2[i]);
  }
#pragma unroll
  for (int j = 0; j < NU; ++j) gu[j] += dxb[NZ + j];
  // stage 2
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    kb[i] = h3 * muz[i] + hh * xb[i];
    dkb[i] = dh3 * muz[i] + h3 * dmuz[i] + hdh * xb[i] + hh * dxb[i];
    xs[i] = x1[i] + hh * k1[i];
    dxs[i] = e[i] + hh * dk1[i] + hdh * k1[i];
  }
  PF::hvp(xs, dxs, kb, dkb, fc, vp, xb, dxb);
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    gz[i] += dxb[i];
    dhb += 0.5 * (dxb[i] * k1[i] + xb[i] * dk1[i]);
  }
#pragma unroll
  for (int j = 0; j < NU; ++j) gu[j] += dxb[NZ + j];
  // stage 1
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    kb[i] = h6 * muz[i] + hh * xb[i];
    dkb[i] = dh6 * muz[i] + h6 * dmuz[i] + hdh * xb[i] + hh * dxb[i];
  }
  PF::hvp(x1, e, kb, dkb, fc, vp, xb, dxb);
#pragma unroll
  for (int i = 0; i < NZ; ++i) gz[i] += dxb[i];
#pragma unroll
  for (int j = 0; j < NU; ++j) gu[j] += dxb[NZ + j];

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
    for (int j = 0; j < NU; ++j) gu[j] += sig * 2.0 * d.R[j] * x1[NZ + j];
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
