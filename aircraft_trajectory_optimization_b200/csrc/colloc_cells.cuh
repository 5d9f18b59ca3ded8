// Collocation interval-cell kernel: one warp per (problem, interval), K = 7 Legendre points; the generated point
// functions of the RB_COLLOC_WPB intervals of a CTA run side by side on consecutive threads.
//
// Reference definition of what is computed: drone3d/raceline/base_raceline.py:398-434 (ode rows:
// [sdot >= 0], defect f - sum_j C[j][k] Z_j / H, dU - sum_j C[j][k] U_j / H), :460-490 / :1132-1181
// (continuity through cont(sum_k D_k Z_k)), drone3d/raceline/drone_raceline.py:42-45 (cont = quaternion
// renormalisation), :47-104 (closure rows reuse the same end state), base_raceline.py:601-623 (cost).
//
// The path is HBM-bound (about 25 KB of jac_g / hess_l values per interval against ~10 kflop), so the
// kernel is organised around the stores: the interval's 169 variables and its multipliers are staged
// in shared memory, every contribution is accumulated into a shared-memory image of the cell's CCS
// entries (unique local entry ids ordered like the CCS), and the image is streamed out with
// consecutive lanes writing consecutive addresses.  Slot layout: structure_colloc.py::CollocLayout.
#pragma once
#include "common.cuh"

#define RB_COLLOC_WPB 4  // warps (= cells) per block
#define RB_KP 8

template <class PF>
struct CollocLayout {
  static constexpr int NZ = PF::NZ, NU = PF::NU, NX = PF::NX, S = NZ + 2 * NU, NJ = PF::NJ, NW = PF::NW;
  static constexpr int JS = 0;
  static constexpr int JF = JS + RB_KP * 9;
  static constexpr int JC = JF + 7 * NJ;
  static constexpr int JH = JC + 7 * NZ * RB_KP;
  static constexpr int JD = JH + 7 * NZ;
  static constexpr int JE = JD + RB_KP * NU * 10;
  static constexpr int JEP = JE + NZ * RB_KP * 4;
  static constexpr int JEU = JEP + NZ;
  static constexpr int JEUP = JEU + NU * RB_KP;
  static constexpr int NJS = JEUP + NU;
  static constexpr int HW = 0;
  static constexpr int HUU = HW + 7 * NW;
  static constexpr int HDD = HUU + RB_KP * NU;
  static constexpr int HHZ = HDD + RB_KP * NU;
  static constexpr int HHU = HHZ + RB_KP * NZ;
  static constexpr int HHD = HHU + RB_KP * NU;
  static constexpr int HHH = HHD + RB_KP * NU;
  static constexpr int HQ = HHH + 1;
  static constexpr int NHS = HQ + 32 * 32;
  static constexpr int RS = 0;
  static constexpr int RF = RB_KP;
  static constexpr int RD = RF + 7 * NZ;
  static constexpr int RE = RD + RB_KP * NU;
  static constexpr int REU = RE + NZ;
  static constexpr int NCR = REU + NU;
  static constexpr int NXS = 1 + RB_KP * S;  // staged variables: h, then the 8 points
};

#define RB_COLLOC_NCR(nz, nu) (RB_KP + 7 * (nz) + RB_KP * (nu) + (nz) + (nu))

// shared-memory doubles one cell needs
template <class PF>
__host__ __device__ constexpr int colloc_cell_doubles(int nju, int nhu) {
  using L = CollocLayout<PF>;
  return L::NXS + L::NCR + 7 * L::NZ + nju + nhu;
}

template <class PF>
__global__ void __launch_bounds__(RB_COLLOC_WPB * 32)
colloc_cells_kernel(const RbDev d, const RbBatch b) {
  using L = CollocLayout<PF>;
  constexpr int NZ = L::NZ, NU = L::NU, NX = L::NX, S = L::S, NJ = L::NJ, NW = L::NW, KP = RB_KP;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ncell = (long long)b.B * d.N;
  const long long cell_raw = (long long)blockIdx.x * RB_COLLOC_WPB + warp;
  const bool valid = cell_raw < ncell;                     // the last CTA may be short of cells: its spare warps
  const long long cell = valid ? cell_raw : ncell - 1;     // recompute the last cell and store nothing
  const int p = (int)(cell / d.N);
  const int n = (int)(cell - (long long)p * d.N);
  const int nju = d.cell_nj, nhu = d.cell_nh;

  double* xs = smem + (size_t)warp * colloc_cell_doubles<PF>(nju, nhu);
  double* lam_l = xs + L::NXS;       // multiplier of every local row (times the row coefficient)
  double* fv = lam_l + L::NCR;       // f at points 1..7
  double* Jst = fv + 7 * NZ;
  double* Hst = Jst + nju;

  const double* __restrict__ w = b.x + (size_t)p * d.nw;
  const double* __restrict__ lam = b.lam_g ? b.lam_g + (size_t)p * d.ng : nullptr;
  const int32_t* __restrict__ crow = d.cell_row + (size_t)n * L::NCR;
  const double* __restrict__ ccoef = d.cell_coef + (size_t)n * L::NCR;
  const double* __restrict__ C = d.colloc_C;   // C[j*8 + k]
  const double* __restrict__ D = d.colloc_D;
  const double* __restrict__ Bq = d.colloc_B;
  const int32_t* __restrict__ tj = d.tmpl_j;
  const int32_t* __restrict__ th = d.tmpl_h;
  const double sig = b.lam_f ? b.lam_f[p] : 1.0;
  const bool want_h = valid && b.hess != nullptr;
  const bool want_j = valid && b.jac != nullptr;

  // ---- stage variables, multipliers; clear the images -------------------------------------------
  const size_t wbase = (size_t)d.N + (size_t)n * KP * S;
  for (int i = lane; i < KP * S; i += 32) xs[1 + i] = w[wbase + i];
  if (lane == 0) xs[0] = w[n];
  for (int r = lane; r < L::NCR; r += 32) {
    const int row = crow[r];
    double v = (row >= 0 && lam) ? lam[row] : 0.0;
    if (r >= L::RE) v *= ccoef[r];
    lam_l[r] = v;
  }
  for (int i = lane; i < nju; i += 32) Jst[i] = 0.0;
  for (int i = lane; i < nhu; i += 32) Hst[i] = 0.0;
  __syncwarp();

  const double h = xs[0];
  const double hi = 1.0 / h, hi2 = hi * hi, hi3 = hi2 * hi;

  // ---- phase A: generated point functions.  The 7 interior points of the RB_COLLOC_WPB cells of this CTA are
  // spread over consecutive threads (28 of the 32 lanes of warp 0 busy) instead of 7 lanes in every warp.
  __syncthreads();                                   // every cell of the CTA has been staged
  if (threadIdx.x < RB_COLLOC_WPB * 7) {
    const int cl = threadIdx.x / 7, k = 1 + threadIdx.x - cl * 7;
    const long long cg_raw = (long long)blockIdx.x * RB_COLLOC_WPB + cl;
    const long long cg = cg_raw < ncell ? cg_raw : ncell - 1;
    const int pp = (int)(cg / d.N);
    const int nn = (int)(cg - (long long)pp * d.N);
    double* xs_c = smem + (size_t)cl * colloc_cell_doubles<PF>(nju, nhu);
    double* lam_c = xs_c + L::NXS;
    double* fv_c = lam_c + L::NCR;
    double* Jst_c = fv_c + 7 * NZ;
    double* Hst_c = Jst_c + nju;
    double x[NX], kb[NZ], f[NZ];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xs_c[1 + k * S + i];
#pragma unroll
    for (int i = 0; i < NZ; ++i) kb[i] = lam_c[L::RF + (k - 1) * NZ + i];
    const double* __restrict__ fcp =
        PF::USES_FC ? (b.fc_b ? b.fc_b + (size_t)pp * d.N * KP * PF::NFC : d.fc) + ((size_t)nn * KP + k) * PF::NFC : d.fc;
    const double* __restrict__ vpp = b.vp + (size_t)pp * b.vp_stride;
    PF::fJW_scatter(x, kb, fcp, vpp, f, Jst_c, tj + L::JF + (k - 1) * NJ, Hst_c, th + L::HW + (k - 1) * NW);
#pragma unroll
    for (int i = 0; i < NZ; ++i) fv_c[(k - 1) * NZ + i] = f[i];
  }
  __syncthreads();

  double* __restrict__ g = (valid && b.g) ? b.g + (size_t)p * d.ng : nullptr;

  // ---- phase B: transcription terms (lane-strided flat loops over the slot groups) ---------------
  // sdot rows (parametric frame):  P0_k / h
  for (int k = lane; k < KP; k += 32) {
    const int row = crow[L::RS + k];
    if (row < 0) continue;
    double P = 0.0;
    for (int j = 0; j < KP; ++j) P += C[j * KP + k] * xs[1 + j * S];
    if (g) g[row] = P * hi;
    for (int j = 0; j < KP; ++j) {
      const int u = tj[L::JS + k * 9 + j];
      if (u >= 0) Jst[u] += C[j * KP + k] * hi;
    }
    Jst[tj[L::JS + k * 9 + 8]] += -P * hi2;
  }
  // defect rows: f_i - P_i/h ; d/dZ_j[i] = -C/h ; d/dh = P_i/h^2
  for (int t = lane; t < 7 * NZ; t += 32) {
    const int k = 1 + t / NZ, i = t - (k - 1) * NZ;
    double P = 0.0;
    for (int j = 0; j < KP; ++j) P += C[j * KP + k] * xs[1 + j * S + i];
    const int row = crow[L::RF + t];
    if (g && row >= 0) g[row] = fv[t] - P * hi;
    for (int j = 0; j < KP; ++j) {
      const int u = tj[L::JC + t * KP + j];
      if (u >= 0) Jst[u] += -C[j * KP + k] * hi;
    }
    Jst[tj[L::JH + t]] += P * hi2;
  }
  // input-rate rows: dU_k[j] - Pu/h
  for (int t = lane; t < KP * NU; t += 32) {
    const int k = t / NU, j = t - k * NU;
    double P = 0.0;
    for (int m = 0; m < KP; ++m) P += C[m * KP + k] * xs[1 + m * S + NZ + j];
    const int row = crow[L::RD + t];
    if (g && row >= 0) g[row] = xs[1 + k * S + NX + j] - P * hi;
    const int b0 = L::JD + t * 10;
    for (int m = 0; m < KP; ++m) {
      const int u = tj[b0 + m];
      if (u >= 0) Jst[u] += -C[m * KP + k] * hi;
    }
    Jst[tj[b0 + 8]] += 1.0;
    Jst[tj[b0 + 9]] += P * hi2;
  }
  // end state cont(sum_k D_k Z_k), end input sum_k D_k U_k   (every lane computes the small sums it needs)
  double nq[4] = {0, 0, 0, 0}, ri = 1.0;
  if (PF::QUAT) {
    double r2 = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double s = 0.0;
      for (int k = 0; k < KP; ++k) s += D[k] * xs[1 + k * S + 3 + a];
      nq[a] = s;
      r2 += s * s;
    }
    ri = 1.0 / sqrt(r2);
#pragma unroll
    for (int a = 0; a < 4; ++a) nq[a] *= ri;
  }
  for (int c = lane; c < NX; c += 32) {
    const int lr = (c < NZ) ? L::RE + c : L::REU + (c - NZ);
    const int row = crow[lr];
    if (row < 0) continue;
    const bool isq = PF::QUAT && c >= 3 && c < 7;
    double out;
    if (isq) {
      out = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) out = (a == c - 3) ? nq[a] : out;
    } else {
      out = 0.0;
      for (int k = 0; k < KP; ++k) out += D[k] * xs[1 + k * S + c];
    }
    const int pv = d.cell_partner[(size_t)n * L::NCR + lr];
    const double pc = d.cell_pcoef[(size_t)n * L::NCR + lr];
    if (g) g[row] = ccoef[lr] * out + (pv >= 0 ? pc * w[pv] : 0.0) + d.cell_off[(size_t)n * L::NCR + lr];
    if (c < NZ) {
      for (int k = 0; k < KP; ++k) {
        if (isq) {
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            const int u = tj[L::JE + (c * KP + k) * 4 + bb];
            if (u >= 0) {
              double na = 0.0;
#pragma unroll
              for (int a = 0; a < 4; ++a) na = (a == c - 3) ? nq[a] : na;
              Jst[u] += ccoef[lr] * D[k] * (((bb == c - 3) ? 1.0 : 0.0) - na * nq[bb]) * ri;
            }
          }
        } else {
          const int u = tj[L::JE + (c * KP + k) * 4];
          if (u >= 0) Jst[u] += ccoef[lr] * D[k];
        }
      }
      const int u = tj[L::JEP + c];
      if (u >= 0 && pv >= 0) Jst[u] += pc;
    } else {
      const int j = c - NZ;
      for (int k = 0; k < KP; ++k) {
        const int u = tj[L::JEU + j * KP + k];
        if (u >= 0) Jst[u] += ccoef[lr] * D[k];
      }
      const int u = tj[L::JEUP + j];
      if (u >= 0 && pv >= 0) Jst[u] += pc;
    }
  }

  // ---- objective, grad_f -----------------------------------------------------------------------
  {
    double fsum = 0.0;   // sum_k stage_k B_k, reduced over lanes below
    for (int k = lane; k < KP; k += 32) {
      double stage = 1.0;
      for (int j = 0; j < NU; ++j) {
        const double u = xs[1 + k * S + NZ + j], du = xs[1 + k * S + NX + j];
        stage += d.R[j] * u * u + d.dR[j] * du * du;
      }
      fsum += stage * Bq[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
    if (lane == 0) {
      if (valid && b.fpart) b.fpart[(size_t)p * d.N + n] = fsum * h;
      if (valid && b.grad_f) b.grad_f[(size_t)p * d.nw + n] = fsum;
    }
    if (valid && b.grad_f) {
      double* __restrict__ gf = b.grad_f + (size_t)p * d.nw + wbase;
      for (int i = lane; i < KP * S; i += 32) {
        const int k = i / S, c = i - k * S;
        double v = 0.0;
        if (c >= NZ && c < NX) v = 2.0 * d.R[c - NZ] * xs[1 + i] * h * Bq[k];
        else if (c >= NX) v = 2.0 * d.dR[c - NX] * xs[1 + i] * h * Bq[k];
        gf[i] = v;
      }
    }
  }

  // ---- Hessian of the Lagrangian: transcription and objective terms -----------------------------
  if (want_h) {
    for (int t = lane; t < KP * NU; t += 32) {
      const int k = t / NU, j = t - k * NU;
      const double u = xs[1 + k * S + NZ + j], du = xs[1 + k * S + NX + j];
      int s = th[L::HUU + t];
      if (s >= 0) Hst[s] += sig * 2.0 * d.R[j] * h * Bq[k];
      s = th[L::HDD + t];
      if (s >= 0) Hst[s] += sig * 2.0 * d.dR[j] * h * Bq[k];
      s = th[L::HHD + t];
      if (s >= 0) Hst[s] += sig * 2.0 * d.dR[j] * du * Bq[k];
      s = th[L::HHU + t];       // (h, U_k[j]):  objective + sum_kk lam_d[kk][j] C[k][kk] / h^2
      if (s >= 0) {
        double acc = 0.0;
        for (int kk = 0; kk < KP; ++kk) acc += lam_l[L::RD + kk * NU + j] * C[k * KP + kk];
        Hst[s] += sig * 2.0 * d.R[j] * u * Bq[k] + acc * hi2;
      }
    }
    for (int t = lane; t < KP * NZ; t += 32) {
      const int j = t / NZ, i = t - j * NZ;
      const int s = th[L::HHZ + t];
      if (s < 0) continue;
      double acc = 0.0;
      for (int kk = 1; kk < KP; ++kk) acc += lam_l[L::RF + (kk - 1) * NZ + i] * C[j * KP + kk];
      if (i == 0)
        for (int kk = 0; kk < KP; ++kk) acc -= lam_l[L::RS + kk] * C[j * KP + kk];
      Hst[s] += acc * hi2;
    }
    {
      // (h, h): -2/h^3 [ sum lam_f P_f + sum lam_d P_u - sum lam_s P_0 ]
      double acc = 0.0;
      for (int t = lane; t < 7 * NZ; t += 32) {
        const int k = 1 + t / NZ, i = t - (k - 1) * NZ;
        double P = 0.0;
        for (int j = 0; j < KP; ++j) P += C[j * KP + k] * xs[1 + j * S + i];
        acc += lam_l[L::RF + t] * P;
      }
      for (int t = lane; t < KP * NU; t += 32) {
        const int k = t / NU, j = t - k * NU;
        double P = 0.0;
        for (int m = 0; m < KP; ++m) P += C[m * KP + k] * xs[1 + m * S + NZ + j];
        acc += lam_l[L::RD + t] * P;
      }
      for (int k = lane; k < KP; k += 32) {
        double P = 0.0;
        for (int j = 0; j < KP; ++j) P += C[j * KP + k] * xs[1 + j * S];
        acc -= lam_l[L::RS + k] * P;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) Hst[th[L::HHH]] += -2.0 * acc * hi3;
    }
    if (PF::QUAT) {
      // quaternion renormalisation block: Hq[a][b] D_k D_l over the 32 x 32 quaternion entries
      double mu[4], phi = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        mu[a] = lam_l[L::RE + 3 + a];
        phi += mu[a] * nq[a];
      }
      const double ri2 = ri * ri;
      for (int t = lane; t < 32 * 32; t += 32) {
        const int s = th[L::HQ + t];
        if (s < 0) continue;
        const int pq = t >> 5, qq = t & 31;
        const int k = pq >> 2, a = pq & 3, l = qq >> 2, bb = qq & 3;
        double ma = 0.0, mb = 0.0, na = 0.0, nb = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ma = (c == a) ? mu[c] : ma;
          na = (c == a) ? nq[c] : na;
          mb = (c == bb) ? mu[c] : mb;
          nb = (c == bb) ? nq[c] : nb;
        }
        const double hq = (-(ma * nb + na * mb) - ((a == bb) ? phi : 0.0) + 3.0 * phi * na * nb) * ri2;
        Hst[s] += hq * D[k] * D[l];
      }
    }
  }
  __syncwarp();

  // ---- phase C: stream the images out (consecutive uids are consecutive CCS positions) -----------
  if (want_j) {
    double* __restrict__ jac = b.jac + (size_t)p * d.nnzj;
    const int32_t* __restrict__ js = d.cell_jslot + (size_t)n * nju;
    // slot indices are fetched eight at a time so that their latency overlaps
    for (int i0 = 0; i0 < nju; i0 += 256) {
      int sl[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 32 + lane;
        sl[u] = i < nju ? js[i] : -1;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (sl[u] >= 0) jac[sl[u]] = Jst[i0 + u * 32 + lane];
    }
  }
  if (want_h) {
    double* __restrict__ H = b.hess + (size_t)p * d.nnzh;
    const int32_t* __restrict__ hs = d.cell_hslot + (size_t)n * nhu;
    for (int i0 = 0; i0 < nhu; i0 += 256) {
      int sl[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 32 + lane;
        sl[u] = i < nhu ? hs[i] : -1;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (sl[u] >= 0) H[sl[u]] = Hst[i0 + u * 32 + lane];
    }
  }
}
