// Collocation interval cells, output-driven form (round 2): two kernels per chunk of instances.
//
// What is computed is what colloc_cells.cuh computes (reference: drone3d/raceline/base_raceline.py:398-434 ode rows,
// :460-490 / :1132-1181 continuity through cont(sum_k D_k Z_k), drone3d/raceline/drone_raceline.py:42-45 quaternion
// renormalisation, :47-104 closure rows, base_raceline.py:601-623 cost), organised differently:
//
//   colloc_point_kernel<PF>  one thread per (instance, interval, interior collocation point): the generated point
//                            functions f, nnz(df/dx), nnz(sum_i lam_i d2f_i/dx2) written to a scratch that holds, per
//                            cell, a [entry][7 points] block -- consecutive threads write consecutive addresses, and a
//                            cell's block is one contiguous span for its consumer.  All lanes busy (the image kernel ran
//                            these functions on 28 of 128 threads while the rest of the CTA waited).
//   colloc_gather_kernel     one 128-thread CTA per (interval, instance).  Every CCS entry of the cell is produced by ONE
//                            thread from a short recipe (host-built at rb_problem_create from the slot -> unique-entry
//                            template of structure_colloc.py: CCS position per interval, argument, constant, grouped by
//                            kind): no shared-memory image, no accumulation; 16 KB of shared memory per cell (variables,
//                            multipliers, the stencil sums P / Q, the staged scratch block), ten CTAs per SM overlap one
//                            another's staging latency.  Variants measured and dropped (C1, evals/s): recipes copied to
//                            shared memory by a CTA that walks many instances of its interval (450 k: three CTAs per SM,
//                            staging latency exposed), four instances per CTA (533 k), one kind-sorted loop with a switch
//                            (569 k), entry-id -> slot indirection instead of per-interval slot lists (573 k).
#pragma once
#include "colloc_cells.cuh"

enum {
  CG_HI = 0,      // c / h
  CG_CONST,       // c
  CG_PHI2,        // c P[arg] / h^2
  CG_SCR,         // scratch[arg]
  CG_END,         // coef[arg] c
  CG_ENDQ,        // coef[lr] c (delta_ab - n_a n_b) / |q|
  CG_PARTNER,     // pcoef[arg]
  CG_SIGH,        // sigma c h
  CG_SIGX,        // sigma c x[arg]
  CG_HHU,         // sigma c x[arg & 4095] + Q[arg >> 12] / h^2
  CG_QHI2,        // Q[arg] / h^2
  CG_HHH,         // -2 acc / h^3
  CG_HQ           // c [-(mu_a n_b + n_a mu_b) - delta_ab phi + 3 phi n_a n_b] / |q|^2
};

#define CG_NKIND 13
// Entries with ONE contribution (almost all) are listed per kind -- tight loops, no divergence: kind k covers the records
// lptr[0][k] .. lptr[0][k+1] (Jacobian entries, then Hessian entries; index 1 of the two-element arrays is unused since the
// two lists were merged).  The few entries with several contributions (e.g. the own-point stencil term next to a df/dx
// entry) keep a CSR list of (kind << 24 | arg, constant) pairs: contributions mptr[0][i] .. mptr[0][i+1] of multi entry i.
struct CgTables {
  int nz, nu, NJ, NW, quat;
  int lptr[2][CG_NKIND + 1];
  int nmulti[2];
  const int32_t *mptr[2], *mka[2];
  const double* mc[2];
  // One 16-byte record per list element and interval: rec[0][n * lstride[0] + i] = (CCS position or -1 where the interval
  // lacks the row, kind << 24 | Hessian flag << 23 | argument, constant as two words); first the per-kind lists in order,
  // then the multi-contribution entries (position and flag only).  One LDG.128 per entry instead of three loads plus an
  // indirection (C1: 658 k -> 733 k evals/s).
  int lstride[2];
  const int4* rec[2];
};

#define RB_CG_THREADS 128
#define RB_CG_CHUNK 128          // instances per launch pair (bounds the scratch; it stays L2-resident)

__host__ __device__ inline int cg_nent(int nz, int NJ, int NW) { return nz + NJ + NW; }
__host__ __device__ inline size_t cg_scratch_doubles(long long cells, int nz, int NJ, int NW) {
  return (size_t)cells * cg_nent(nz, NJ, NW) * 7;
}
// shared-memory doubles of ONE cell context
__host__ __device__ inline size_t cg_cell_doubles(int nz, int nu, int NJ, int NW) {
  const int S = nz + 2 * nu;
  return (size_t)(1 + RB_KP * S) + RB_COLLOC_NCR(nz, nu) + (8 + 7 * nz + 8 * nu) + (8 * nu + 8 * nz) +
         (size_t)cg_nent(nz, NJ, NW) * 7 + 2 * (nz + nu) + 32;
}
// shared memory of a CTA: one cell context + the recipe lists of its interval (slot, argument: int32; constant: double)
// + the row tables of the interval
__host__ __device__ inline size_t cg_smem_bytes(int nz, int nu, int NJ, int NW, int lstride0, int lstride1, int nl0, int nl1) {
  const size_t ncr = RB_COLLOC_NCR(nz, nu);
  (void)lstride0; (void)lstride1; (void)nl0; (void)nl1;      // the recipe lists are read from global memory (L1 / L2)
  return (cg_cell_doubles(nz, nu, NJ, NW) + 2 * ncr + 2) * sizeof(double) + (2 * ncr + 8) * sizeof(int32_t);
}

template <class PF>
struct CgScrIdx {
  int tj[PF::NJ > 0 ? PF::NJ : 1], th[PF::NW > 0 ? PF::NW : 1];
  constexpr CgScrIdx() : tj(), th() {
    for (int e = 0; e < PF::NJ; ++e) tj[e] = (PF::NZ + e) * 7;
    for (int e = 0; e < PF::NW; ++e) th[e] = (PF::NZ + PF::NJ + e) * 7;
  }
};

template <class PF>
__global__ void __launch_bounds__(128)
colloc_point_kernel(const RbDev d, const RbBatch b, double* __restrict__ scr) {
  constexpr int NZ = PF::NZ, NU = PF::NU, NX = PF::NX, S = NZ + 2 * NU, KP = RB_KP;
  constexpr int NCR = RB_COLLOC_NCR(NZ, NU), RF = RB_KP;
  constexpr int NENT = NZ + PF::NJ + PF::NW;
  constexpr CgScrIdx<PF> idx{};
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ncell = (long long)b.B * d.N;
  if (t >= ncell * 7) return;
  const long long cell = t / 7;
  const int k = 1 + (int)(t - cell * 7);
  const int p = (int)(cell / d.N);
  const int n = (int)(cell - (long long)p * d.N);
  const double* __restrict__ w = b.x + (size_t)p * d.nw + d.N + ((size_t)n * KP + k) * S;
  const double* __restrict__ lam = b.lam_g ? b.lam_g + (size_t)p * d.ng : nullptr;
  const int32_t* __restrict__ crow = d.cell_row + (size_t)n * NCR + RF + (k - 1) * NZ;
  double x[NX], kb[NZ], f[NZ];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = w[i];
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    const int row = crow[i];
    kb[i] = (row >= 0 && lam) ? lam[row] : 0.0;
  }
  const double* __restrict__ fcp =
      PF::USES_FC ? (b.fc_b ? b.fc_b + (size_t)p * d.N * KP * PF::NFC : d.fc) + ((size_t)n * KP + k) * PF::NFC : d.fc;
  const double* __restrict__ vpp = b.vp + (size_t)p * b.vp_stride;
  double* __restrict__ out = scr + (size_t)cell * NENT * 7 + (k - 1);
  PF::fJW_scatter(x, kb, fcp, vpp, f, out, idx.tj, out, idx.th);
#pragma unroll
  for (int i = 0; i < NZ; ++i) out[i * 7] = f[i];
}

// offsets (doubles) of the parts of a cell context in shared memory
struct CgOff {
  int lam, P, Q, S, cco, pco, sc;
};
// scalars of a cell context
enum { SC_NQ = 0, SC_MU = 4, SC_RI = 8, SC_PHI, SC_ACC, SC_FSUM, SC_H, SC_HI, SC_HI2, SC_HI3, SC_SIG, SC_RI2 };

template <int KIND>
__device__ __forceinline__ double cg_term(int arg, double c, const double* __restrict__ cs, const CgOff& o) {
  const double* __restrict__ sc = cs + o.sc;
  if (KIND == CG_HI) return c * sc[SC_HI];
  if (KIND == CG_CONST) return c;
  if (KIND == CG_PHI2) return c * cs[o.P + arg] * sc[SC_HI2];
  if (KIND == CG_SCR) return cs[o.S + arg];
  if (KIND == CG_END) return cs[o.cco + arg] * c;
  if (KIND == CG_ENDQ) {
    const int lr = arg & 255, a = (arg >> 8) & 15, bq = (arg >> 12) & 15;
    return cs[o.cco + lr] * c * (((a == bq) ? 1.0 : 0.0) - sc[SC_NQ + a] * sc[SC_NQ + bq]) * sc[SC_RI];
  }
  if (KIND == CG_PARTNER) return cs[o.pco + arg];
  if (KIND == CG_SIGH) return sc[SC_SIG] * c * sc[SC_H];
  if (KIND == CG_SIGX) return sc[SC_SIG] * c * cs[arg];
  if (KIND == CG_HHU) return sc[SC_SIG] * c * cs[arg & 4095] + cs[o.Q + (arg >> 12)] * sc[SC_HI2];
  if (KIND == CG_QHI2) return cs[o.Q + arg] * sc[SC_HI2];
  if (KIND == CG_HHH) return -2.0 * sc[SC_ACC] * sc[SC_HI3];
  if (KIND == CG_HQ) {
    const int a = arg & 15, bq = arg >> 4;
    const double ma = sc[SC_MU + a], mb = sc[SC_MU + bq], na = sc[SC_NQ + a], nb = sc[SC_NQ + bq], phi = sc[SC_PHI];
    return (-(ma * nb + na * mb) - ((a == bq) ? phi : 0.0) + 3.0 * phi * na * nb) * sc[SC_RI2] * c;
  }
  return 0.0;
}

__device__ __forceinline__ double cg_eval(int ka, double c, const double* __restrict__ cs, const CgOff& o) {
  const int kind = ka >> 24, arg = ka & 0xffffff;
  switch (kind) {
    case CG_HI: return cg_term<CG_HI>(arg, c, cs, o);
    case CG_CONST: return cg_term<CG_CONST>(arg, c, cs, o);
    case CG_PHI2: return cg_term<CG_PHI2>(arg, c, cs, o);
    case CG_SCR: return cg_term<CG_SCR>(arg, c, cs, o);
    case CG_END: return cg_term<CG_END>(arg, c, cs, o);
    case CG_ENDQ: return cg_term<CG_ENDQ>(arg, c, cs, o);
    case CG_PARTNER: return cg_term<CG_PARTNER>(arg, c, cs, o);
    case CG_SIGH: return cg_term<CG_SIGH>(arg, c, cs, o);
    case CG_SIGX: return cg_term<CG_SIGX>(arg, c, cs, o);
    case CG_HHU: return cg_term<CG_HHU>(arg, c, cs, o);
    case CG_QHI2: return cg_term<CG_QHI2>(arg, c, cs, o);
    case CG_HHH: return cg_term<CG_HHH>(arg, c, cs, o);
    case CG_HQ: return cg_term<CG_HQ>(arg, c, cs, o);
  }
  return 0.0;
}

struct CgLists {      // the interval's recipe records
  const int4* rec[2];
};

// all single-contribution entries of one kind: out[CCS position of the entry] = term, Jacobian and Hessian entries in one
// list (bit 23 of the argument word selects the output).  One tight loop per kind (no divergence, no switch); measured
// faster than a single kind-sorted loop with a switch (636 k against 569 k evals/s on C1).
template <int KIND>
__device__ __forceinline__ void cg_kind_loop(const CgTables& t, const int4* __restrict__ rec, double* __restrict__ outJ,
                                             double* __restrict__ outH, const double* __restrict__ cs, const CgOff& o) {
#pragma unroll 2
  for (int i = t.lptr[0][KIND] + threadIdx.x; i < t.lptr[0][KIND + 1]; i += RB_CG_THREADS) {
    const int4 r = __ldg(rec + i);
    double* __restrict__ out = (r.y & 0x800000) ? outH : outJ;
    if (r.x >= 0 && out) out[r.x] = cg_term<KIND>(r.y & 0x7fffff, __hiloint2double(r.w, r.z), cs, o);
  }
}

__device__ __forceinline__ void cg_all_kinds(const CgTables& t, const int4* __restrict__ rec, double* __restrict__ outJ,
                                             double* __restrict__ outH, const double* __restrict__ cs, const CgOff& o) {
  cg_kind_loop<CG_HI>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_CONST>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_PHI2>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_SCR>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_END>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_ENDQ>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_PARTNER>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_SIGH>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_SIGX>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_HHU>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_QHI2>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_HHH>(t, rec, outJ, outH, cs, o);
  cg_kind_loop<CG_HQ>(t, rec, outJ, outH, cs, o);
  // entries with several contributions (few)
  const int nl = t.lptr[0][CG_NKIND];
  for (int i = threadIdx.x; i < t.nmulti[0]; i += RB_CG_THREADS) {
    const int4 r = __ldg(rec + nl + i);
    double* __restrict__ out = (r.y & 0x800000) ? outH : outJ;
    if (r.x < 0 || !out) continue;
    double v = 0.0;
    for (int e = t.mptr[0][i]; e < t.mptr[0][i + 1]; ++e) v += cg_eval(t.mka[0][e], t.mc[0][e], cs, o);
    out[r.x] = v;
  }
}

// grid (N, G): CTA (n, g) handles interval n of the instances g, g + G, g + 2 G, ...
#ifndef RB_CG_MINBLOCKS
#define RB_CG_MINBLOCKS 10
#endif
__global__ void __launch_bounds__(RB_CG_THREADS, RB_CG_MINBLOCKS)
colloc_gather_kernel(const RbDev d, const RbBatch b, const CgTables t, const double* __restrict__ scr) {
  extern __shared__ __align__(16) double cg_smem[];
  const int tid = threadIdx.x, nth = RB_CG_THREADS;
  const int nz = t.nz, nu = t.nu, nx = nz + nu, S = nz + 2 * nu, KP = RB_KP;
  const int NCR = RB_COLLOC_NCR(nz, nu);
  const int RS = 0, RF = KP, RD = RF + 7 * nz, RE = RD + KP * nu;
  const int nP = 8 + 7 * nz + 8 * nu, nQ = 8 * nu + 8 * nz, nS = cg_nent(nz, t.NJ, t.NW) * 7;
  const int n = blockIdx.x;
  CgOff o;
  o.lam = 1 + KP * S;          // xs at 0: h, then the 8 points
  o.P = o.lam + NCR;           // stencil sums: Ps[8], Pf[7 nz], Pu[8 nu]
  o.Q = o.P + nP;              // multiplier-weighted stencil sums: Qu[8 nu], Qz[8 nz]
  o.S = o.Q + nQ;              // the cell's scratch block [entry][7]
  o.cco = o.S + nS;            // coefficient / partner coefficient of the end rows
  o.pco = o.cco + nx;
  o.sc = o.pco + nx;
  // shared memory: cell context | row coefficients, offsets | row ids, partners
  double* cs = cg_smem;
  double* rcoef = cs + cg_cell_doubles(nz, nu, t.NJ, t.NW);     // [NCR] row coefficient
  double* roff = rcoef + NCR;                                   // [NCR] row offset
  int32_t* rrow = reinterpret_cast<int32_t*>(roff + NCR + 2);   // [NCR] global row of every local row
  int32_t* rpart = rrow + NCR;                                  // [NCR] partner variable
  for (int r = tid; r < NCR; r += nth) {
    rrow[r] = d.cell_row[(size_t)n * NCR + r];
    rcoef[r] = d.cell_coef[(size_t)n * NCR + r];
    rpart[r] = d.cell_partner[(size_t)n * NCR + r];
    roff[r] = d.cell_off[(size_t)n * NCR + r];
    if (r >= RE) {
      cs[o.cco + r - RE] = d.cell_coef[(size_t)n * NCR + r];
      cs[o.pco + r - RE] = d.cell_partner[(size_t)n * NCR + r] >= 0 ? d.cell_pcoef[(size_t)n * NCR + r] : 0.0;
    }
  }
  CgLists L;
  L.rec[0] = t.rec[0] + (size_t)n * t.lstride[0];
  L.rec[1] = nullptr;
  const double* __restrict__ C = d.colloc_C;
  const double* __restrict__ D = d.colloc_D;
  const double* __restrict__ Bq = d.colloc_B;
  const size_t wbase = (size_t)d.N + (size_t)n * KP * S;
  double* sc = cs + o.sc;
  double* lam_l = cs + o.lam;
  __syncthreads();

  for (int p = blockIdx.y; p < b.B; p += gridDim.y) {
    // ---- stage variables, multipliers, the scratch block -------------------------------------------------------------
    const double* __restrict__ w = b.x + (size_t)p * d.nw;
    const double* __restrict__ lam = b.lam_g ? b.lam_g + (size_t)p * d.ng : nullptr;
    for (int i = tid; i < KP * S; i += nth) cs[1 + i] = w[wbase + i];
    if (tid == 0) {
      const double h = w[n];
      cs[0] = h;
      sc[SC_H] = h;
      sc[SC_HI] = 1.0 / h;
      sc[SC_HI2] = sc[SC_HI] * sc[SC_HI];
      sc[SC_HI3] = sc[SC_HI2] * sc[SC_HI];
      sc[SC_SIG] = b.lam_f ? b.lam_f[p] : 1.0;
    }
    for (int r = tid; r < NCR; r += nth) {
      const int row = rrow[r];
      double v = (row >= 0 && lam) ? lam[row] : 0.0;
      if (r >= RE) v *= rcoef[r];
      lam_l[r] = v;
    }
    {
      const double* __restrict__ src = scr + ((size_t)p * d.N + n) * nS;
      for (int i = tid; i < nS; i += nth) cs[o.S + i] = src[i];
    }
    __syncthreads();
    // ---- stencil sums ----------------------------------------------------------------------------------------------
    for (int i = tid; i < nP; i += nth) {
      int k, off;
      if (i < 8) {
        k = i;
        off = 0;
      } else if (i < 8 + 7 * nz) {
        const int tt = i - 8;
        k = 1 + tt / nz;
        off = tt - (k - 1) * nz;
      } else {
        const int tt = i - 8 - 7 * nz;
        k = tt / nu;
        off = nz + tt - k * nu;
      }
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < RB_KP; ++j) acc += C[j * KP + k] * cs[1 + j * S + off];
      cs[o.P + i] = acc;
    }
    for (int i = tid; i < nQ; i += nth) {
      double acc = 0.0;
      if (i < 8 * nu) {
        const int k = i / nu, j = i - k * nu;
#pragma unroll
        for (int kk = 0; kk < RB_KP; ++kk) acc += lam_l[RD + kk * nu + j] * C[k * KP + kk];
      } else {
        const int tt = i - 8 * nu;
        const int j = tt / nz, c = tt - j * nz;
#pragma unroll
        for (int kk = 1; kk < RB_KP; ++kk) acc += lam_l[RF + (kk - 1) * nz + c] * C[j * KP + kk];
        if (c == 0)
          for (int kk = 0; kk < KP; ++kk) acc -= lam_l[RS + kk] * C[j * KP + kk];
      }
      cs[o.Q + i] = acc;
    }
    // end state cont(sum_k D_k Z_k): quaternion renormalisation (last warp), objective (second-last warp)
    if (tid == nth - 1) {
      double nq[4] = {0, 0, 0, 0}, ri = 1.0, phi = 0.0;
      if (t.quat) {
        double r2 = 0.0;
        for (int a = 0; a < 4; ++a) {
          double sacc = 0.0;
          for (int k = 0; k < KP; ++k) sacc += D[k] * cs[1 + k * S + 3 + a];
          nq[a] = sacc;
          r2 += sacc * sacc;
        }
        ri = 1.0 / sqrt(r2);
        for (int a = 0; a < 4; ++a) {
          nq[a] *= ri;
          sc[SC_MU + a] = lam_l[RE + 3 + a];
          phi += lam_l[RE + 3 + a] * nq[a];
        }
      } else {
        for (int a = 0; a < 4; ++a) sc[SC_MU + a] = 0.0;
      }
      for (int a = 0; a < 4; ++a) sc[SC_NQ + a] = nq[a];
      sc[SC_RI] = ri;
      sc[SC_RI2] = ri * ri;
      sc[SC_PHI] = phi;
    }
    if ((tid >> 5) == (nth >> 5) - 2) {
      const int lane = tid & 31;
      double fsum = 0.0;
      for (int k = lane; k < KP; k += 32) {
        double stage = 1.0;
        for (int j = 0; j < nu; ++j) {
          const double u = cs[1 + k * S + nz + j], du = cs[1 + k * S + nx + j];
          stage += d.R[j] * u * u + d.dR[j] * du * du;
        }
        fsum += stage * Bq[k];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, off);
      if (lane == 0) sc[SC_FSUM] = fsum;
    }
    __syncthreads();
    // (h, h) term: sum of multiplier x stencil sum (first warp; the others go on with g and the entries that do not need it)
    if (tid < 32) {
      double acc = 0.0;
      for (int i = tid; i < nP; i += 32) {
        const double l = i < 8 ? -lam_l[RS + i] : (i < 8 + 7 * nz ? lam_l[RF + i - 8] : lam_l[RD + i - 8 - 7 * nz]);
        acc += l * cs[o.P + i];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (tid == 0) sc[SC_ACC] = acc;
    }
    // ---- g -------------------------------------------------------------------------------------------------------------
    if (b.g) {
      double* __restrict__ g = b.g + (size_t)p * d.ng;
      for (int r = tid; r < NCR; r += nth) {
        const int row = rrow[r];
        if (row < 0) continue;
        double v;
        if (r < RF) {
          v = cs[o.P + r] * sc[SC_HI];
        } else if (r < RD) {
          const int tt = r - RF, k = 1 + tt / nz, i = tt - (k - 1) * nz;
          v = cs[o.S + i * 7 + (k - 1)] - cs[o.P + 8 + tt] * sc[SC_HI];
        } else if (r < RE) {
          const int tt = r - RD, k = tt / nu, j = tt - k * nu;
          v = cs[1 + k * S + nx + j] - cs[o.P + 8 + 7 * nz + tt] * sc[SC_HI];
        } else {
          const int c = r - RE;       // state component, or nz + input component
          double out = 0.0;
          if (t.quat && c >= 3 && c < 7) {
            out = sc[SC_NQ + c - 3];
          } else {
            for (int k = 0; k < KP; ++k) out += D[k] * cs[1 + k * S + c];
          }
          const int pv = rpart[r];
          v = rcoef[r] * out + (pv >= 0 ? cs[o.pco + c] * w[pv] : 0.0) + roff[r];
        }
        g[row] = v;
      }
    }
    // ---- objective pieces ----------------------------------------------------------------------------------------------
    if (tid == 32) {
      if (b.fpart) b.fpart[(size_t)p * d.N + n] = sc[SC_FSUM] * sc[SC_H];
      if (b.grad_f) b.grad_f[(size_t)p * d.nw + n] = sc[SC_FSUM];
    }
    if (b.grad_f) {
      double* __restrict__ gf = b.grad_f + (size_t)p * d.nw + wbase;
      const double h = cs[0];
      for (int i = tid; i < KP * S; i += nth) {
        const int k = i / S, c = i - k * S;
        double v = 0.0;
        if (c >= nz && c < nx) v = 2.0 * d.R[c - nz] * cs[1 + i] * h * Bq[k];
        else if (c >= nx) v = 2.0 * d.dR[c - nx] * cs[1 + i] * h * Bq[k];
        gf[i] = v;
      }
    }
    // ---- jac_g, hess_l: every unique entry is produced by one thread ---------------------------------------------------
    __syncthreads();      // SC_ACC (warp 0, needed by the (h, h) Hessian entry) is complete
    cg_all_kinds(t, L.rec[0], b.jac ? b.jac + (size_t)p * d.nnzj : nullptr, b.hess ? b.hess + (size_t)p * d.nnzh : nullptr,
                 cs, o);
    __syncthreads();      // the context is free for the next instance
  }
}
