// Block-tridiagonal + border KKT factor/solve for LARGE stage blocks (collocation intervals: a stage triple
// holds the 7 interior collocation points of an interval, ~310 unknowns -- it does not fit shared memory in
// fp64).  Same algorithm, tables and factor storage as kkt_factor_solve_kernel (kkt_blocks.cuh); the block
// matrix is inverted in place in its slot of the factor storage in global memory (L2-resident: 0.8 MB), the
// right-hand-side panels live in a per-instance global workspace, and only vectors sit in shared memory.
//
// One CTA of RB_KKTB_THREADS threads per instance.  Per Gauss-Jordan step: warp 0 picks the Bunch-Kaufman
// pivot from a shared-memory copy of the diagonal and from pivot ROWS (the partially inverted matrix satisfies
// M[j][p] = -M[p][j] for swept j and M[j][p] = M[p][j] for unswept j, so the pivot column never has to be read
// with a stride), then all threads apply the rank-1 / rank-2 update with coalesced row accesses.
#pragma once
#include "kkt_blocks.cuh"

// threads per CTA and Gauss-Jordan update vectors accumulated before one pass over the block applies them.  Measured on the
// condensed C1 factorisation (56 interiors of 276 unknowns, ms per factor + solve): 1024 threads / groups of 8: 4.05,
// 1024 / 16: 4.56, 512 / 8: 3.87, 256 / 16: 3.91, 384 / 24: 3.38, 512 / 16: 3.35 (groups of 32 do not fit shared memory)
#ifndef RB_KKTB_THREADS
#define RB_KKTB_THREADS 512
#endif
#ifndef RB_KKTB_GRP
#define RB_KKTB_GRP 16
#endif

struct KktBigSmem {
  double *Cg, *Rg;          // [RB_KKTB_GRP][nbb] update vectors of the current group
  double *rowA, *rowB;      // current rows k and r of the block (pivot choice, then the pivot rows)
  double *dg, *Lc, *carry, *rcarry, *xb;
  int* swept;
  int* ingrp;               // index pivoted inside the current group
  int* piv;                 // type, p, q, buffer of row p, k
};

__host__ __device__ inline size_t kktb_smem_bytes(int bmax, int nb, int mmax, int qmax) {
  const size_t nbb = nb > bmax ? nb : bmax, nrhs = 1 + nb;
  const size_t dbl = (2 * RB_KKTB_GRP + 4) * nbb + (size_t)mmax * qmax + (size_t)mmax * mmax + (size_t)mmax * nrhs;
  return dbl * sizeof(double) + (2 * nbb + 8) * sizeof(int);
}

__device__ inline KktBigSmem kktb_carve(double* s, int bmax, int nb, int mmax, int qmax) {
  const size_t nbb = nb > bmax ? nb : bmax, nrhs = 1 + nb;
  KktBigSmem k;
  k.Cg = s; s += RB_KKTB_GRP * nbb;
  k.Rg = s; s += RB_KKTB_GRP * nbb;
  k.rowA = s; s += nbb;
  k.rowB = s; s += nbb;
  k.dg = s; s += nbb;
  k.xb = s; s += nbb;
  k.Lc = s; s += (size_t)mmax * qmax;
  k.carry = s; s += (size_t)mmax * mmax;
  k.rcarry = s; s += (size_t)mmax * nrhs;
  k.swept = reinterpret_cast<int*>(s);
  k.ingrp = k.swept + nbb;
  k.piv = k.ingrp + nbb;
  return k;
}

// doubles of per-instance global workspace: two right-hand-side panels + YL panel of one block + border matrix
__host__ __device__ inline size_t kktb_work_doubles(int bmax, int nb, int mmax) {
  const size_t nbb = nb > bmax ? nb : bmax, nrhs = 1 + nb;
  return 2 * nbb * nrhs + (size_t)bmax * mmax + nbb * nbb;
}

// In-place inverse of the symmetric b x b matrix M (global memory, row-major, leading dimension LD) by
// Gauss-Jordan steps with Bunch-Kaufman pivoting, the updates applied in groups.
//
// A step on pivot p maps M to Z_p(M) - c r' where Z_p zeroes row and column p, c_i = a_ip (c_p = -1),
// r_j = a_pj / a_pp (r_p = 1 / a_pp); a 2x2 step on (p, q) subtracts two such terms.  After the steps s = 1..g of a
// group the block is  Z_P(M) - sum_s c~_s r~_s'  where P is the set of the group's pivots and c~_s, r~_s are
// c_s, r_s with the entries of the LATER pivots of the group zeroed.  The block itself (0.8 MB, L2-resident) is
// therefore read and written once per group of RB_KKTB_GRP vectors instead of once per pivot; in between, warp 0
// reconstructs the one or two current rows the pivot choice needs from the stored block and the group's
// vectors, and the current diagonal is carried in shared memory.  Pivot columns come from pivot rows: the
// partially inverted matrix satisfies M[j][p] = -M[p][j] for swept j and M[j][p] = M[p][j] otherwise.
__device__ int kktb_sym_invert(double* __restrict__ M, int LD, int b, const KktBigSmem& s, int* neg, int nbb) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
  const double alpha = 0.6403882032022076;
  int bad = 0, nneg = 0;
  for (int i = tid; i < b; i += nthreads) {
    s.swept[i] = 0;
    s.ingrp[i] = 0;
    s.dg[i] = M[(size_t)i * LD + i];
  }
  __syncthreads();
  int remaining = b, ng = 0;        // ng: vectors in the current group (uniform over the CTA)

  // current row i of the block -> buf (every thread one or more entries)
  auto current_row = [&](int i, double* __restrict__ buf) {
    const bool gi = s.ingrp[i] != 0;
    for (int j = tid; j < b; j += nthreads) {
      double v = (gi || s.ingrp[j]) ? 0.0 : M[(size_t)i * LD + j];
      for (int t = 0; t < ng; ++t) v -= s.Cg[t * nbb + i] * s.Rg[t * nbb + j];
      buf[j] = v;
    }
  };
  // M <- Z_P(M) - sum_t C_t R_t' : two rows per warp pass, the C entries of the rows in registers
  auto flush = [&]() {
    for (int i0 = 2 * warp; i0 < b; i0 += 2 * nwarps) {
      const int i1 = (i0 + 1 < b) ? i0 + 1 : i0;
      double c0[RB_KKTB_GRP], c1[RB_KKTB_GRP];
#pragma unroll
      for (int t = 0; t < RB_KKTB_GRP; ++t) {
        c0[t] = t < ng ? s.Cg[t * nbb + i0] : 0.0;
        c1[t] = t < ng ? s.Cg[t * nbb + i1] : 0.0;
      }
      const bool g0 = s.ingrp[i0] != 0, g1 = s.ingrp[i1] != 0;
      double* __restrict__ row0 = M + (size_t)i0 * LD;
      double* __restrict__ row1 = M + (size_t)i1 * LD;
      for (int j = lane; j < b; j += 32) {
        const bool gj = s.ingrp[j] != 0;
        double v0 = (g0 || gj) ? 0.0 : row0[j];
        double v1 = (g1 || gj) ? 0.0 : row1[j];
#pragma unroll
        for (int t = 0; t < RB_KKTB_GRP; ++t) {
          const double rj = t < ng ? s.Rg[t * nbb + j] : 0.0;
          v0 -= c0[t] * rj;
          v1 -= c1[t] * rj;
        }
        row0[j] = v0;
        if (i1 != i0) row1[j] = v1;
      }
    }
    __syncthreads();
    for (int i = tid; i < b; i += nthreads) s.ingrp[i] = 0;
    ng = 0;
    __syncthreads();
  };

  while (remaining > 0) {
    // ---- pivot choice: warp 0 scans, every thread helps to reconstruct the one or two current rows it needs
    if (warp == 0) {
      // k: largest remaining diagonal entry
      float v = -1.f;
      int k = 0x7fffffff;
      for (int i = lane; i < b; i += 32)
        if (!s.swept[i]) {
          const float a = kkt_mag(s.dg[i]);
          if (a > v) {
            v = a;
            k = i;
          }
        }
      int kk;
      kkt_argmax32(v, k, kk);
      if (lane == 0) s.piv[4] = kk;
    }
    __syncthreads();
    const int k = s.piv[4];
    current_row(k, s.rowA);
    __syncthreads();
    if (warp == 0) {
      int type = 1, p = k, q = k, r = k, need_r = 0;
      if (remaining > 1) {
        // lambda = largest remaining off-diagonal entry of row k (= column k by symmetry of the unswept part)
        float v = -1.f;
        r = 0x7fffffff;
        for (int j = lane; j < b; j += 32)
          if (!s.swept[j] && j != k) {
            const float a = kkt_mag(s.rowA[j]);
            if (a > v) {
              v = a;
              r = j;
            }
          }
        int rr;
        kkt_argmax32(v, r, rr);
        r = rr;
        const double akk = fabs(s.dg[k]), lam = fabs(s.rowA[r]);
        need_r = !(akk >= alpha * lam);
      }
      if (lane == 0) {
        s.piv[0] = need_r ? -1 : type;
        s.piv[1] = p;
        s.piv[2] = need_r ? r : q;
        s.piv[3] = 0;
      }
    }
    __syncthreads();
    if (s.piv[0] < 0) {
      const int r = s.piv[2];
      current_row(r, s.rowB);
      __syncthreads();
      if (warp == 0) {
        float v = -1.f;
        int t = 0x7fffffff;
        for (int j = lane; j < b; j += 32)
          if (!s.swept[j] && j != r) {
            const float a = kkt_mag(s.rowB[j]);
            if (a > v) {
              v = a;
              t = j;
            }
          }
        int tt;
        kkt_argmax32(v, t, tt);
        const double akk = fabs(s.dg[k]), lam = fabs(s.rowA[r]), sig = fabs(s.rowB[tt]);
        int type = 1, p, q;
        if (akk * sig >= alpha * lam * lam) {
          p = q = k;
        } else if (fabs(s.dg[r]) >= alpha * sig) {
          p = q = r;
        } else {
          type = 2;
          p = k < r ? k : r;
          q = k < r ? r : k;
        }
        if (lane == 0) {
          s.piv[0] = type;
          s.piv[1] = p;
          s.piv[2] = q;
          s.piv[3] = (p == k) ? 0 : 1;      // row p sits in rowA (0) or rowB (1); row q in the other one
        }
      }
      __syncthreads();
    }
    const int type = s.piv[0], p = s.piv[1], q = s.piv[2];
    const double* __restrict__ rowp = s.piv[3] ? s.rowB : s.rowA;
    const double* __restrict__ rowq = s.piv[3] ? s.rowA : s.rowB;
    if (type == 1) {
      double d = s.dg[p];
      if (!(fabs(d) > 1e-250)) {
        d = 1e-250;
        bad++;
      }
      if (d < 0) nneg++;
      const double di = 1.0 / d;
      double* __restrict__ C = s.Cg + (size_t)ng * nbb;
      double* __restrict__ R = s.Rg + (size_t)ng * nbb;
      for (int j = tid; j < b; j += nthreads) {
        const double a = rowp[j];
        const double c = (j == p) ? -1.0 : (s.swept[j] ? -a : a);     // column p from row p
        const double r = (j == p) ? di : a * di;
        C[j] = c;
        R[j] = r;
        s.dg[j] = ((j == p) ? 0.0 : s.dg[j]) - c * r;
        if (j == p)
          for (int t = 0; t < ng; ++t) {
            s.Cg[t * nbb + p] = 0.0;
            s.Rg[t * nbb + p] = 0.0;
          }
      }
      __syncthreads();
      if (tid == 0) {
        s.swept[p] = 1;
        s.ingrp[p] = 1;
      }
      ng += 1;
      remaining -= 1;
    } else {
      const double epp = s.dg[p], eqq = s.dg[q], epq = rowp[q];
      double det = epp * eqq - epq * epq;
      if (!(fabs(det) > 1e-250)) {
        det = -1e-250;
        bad++;
      }
      nneg += det < 0 ? 1 : (epp + eqq < 0 ? 2 : 0);
      const double i00 = eqq / det, i01 = -epq / det, i11 = epp / det;
      double* __restrict__ C1 = s.Cg + (size_t)ng * nbb;
      double* __restrict__ R1 = s.Rg + (size_t)ng * nbb;
      double* __restrict__ C2 = C1 + nbb;
      double* __restrict__ R2 = R1 + nbb;
      for (int j = tid; j < b; j += nthreads) {
        const double ap = rowp[j], aq = rowq[j];
        const double sg = s.swept[j] ? -1.0 : 1.0;
        double c1, c2, r1, r2;
        if (j == p) {
          c1 = -1.0; c2 = 0.0; r1 = i00; r2 = i01;
        } else if (j == q) {
          c1 = 0.0; c2 = -1.0; r1 = i01; r2 = i11;
        } else {
          c1 = sg * ap; c2 = sg * aq;
          r1 = i00 * ap + i01 * aq;
          r2 = i01 * ap + i11 * aq;
        }
        C1[j] = c1; C2[j] = c2; R1[j] = r1; R2[j] = r2;
        s.dg[j] = ((j == p || j == q) ? 0.0 : s.dg[j]) - c1 * r1 - c2 * r2;
        if (j == p || j == q)
          for (int t = 0; t < ng; ++t) {
            s.Cg[t * nbb + j] = 0.0;
            s.Rg[t * nbb + j] = 0.0;
          }
      }
      __syncthreads();
      if (tid == 0) {
        s.swept[p] = 1;
        s.swept[q] = 1;
        s.ingrp[p] = 1;
        s.ingrp[q] = 1;
      }
      ng += 2;
      remaining -= 2;
    }
    __syncthreads();
    if (ng > RB_KKTB_GRP - 2 || remaining == 0) flush();
  }
  *neg += nneg;
  return bad;
}

// out[i][r] (op)= sum_j A[i][j] * Bm[j][r]: A b x k (lda), Bm k x nr (ldb); 4 x 4 register tiles, columns across lanes
__device__ inline void kktb_gemm(const double* __restrict__ A, int lda, const double* __restrict__ Bm, int ldb,
                                 double* __restrict__ out, int ldo, int b, int k, int nr, double* __restrict__ out2,
                                 size_t ldo2) {
  const int nct = (nr + 3) >> 2, nrt = (b + 3) >> 2;
  for (int tile = threadIdx.x; tile < nrt * nct; tile += blockDim.x) {
    const int tr_ = tile / nct, tc_ = tile - tr_ * nct;
    int ri[4], rc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) ri[a] = (tr_ * 4 + a < b) ? tr_ * 4 + a : tr_ * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) rc[c] = (tc_ + c * nct < nr) ? tc_ + c * nct : tc_;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
    for (int j = 0; j < k; ++j) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = A[(size_t)ri[a] * lda + j];
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = Bm[(size_t)j * ldb + rc[c]];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] += av[a] * bv[c];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (tr_ * 4 + a < b && tc_ + c * nct < nr) {
          out[(size_t)ri[a] * ldo + rc[c]] = acc[a][c];
          if (out2) out2[(size_t)ri[a] * ldo2 + rc[c]] = acc[a][c];
        }
  }
}

__global__ void __launch_bounds__(RB_KKTB_THREADS, 1)
kkt_factor_solve_big_kernel(const RbKktDev d, const RbKktBatch bt, double* __restrict__ work, size_t work_stride) {
  extern __shared__ double kktb_smem[];
  const int p = blockIdx.x;
  if (p >= bt.B) return;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int nk = d.nw + d.ng, nrhs = 1 + d.nb, bmax = d.bmax, mmax = d.mmax, qmax = d.qmax, N = d.N, nb = d.nb;
  const int nbb = nb > bmax ? nb : bmax;
  KktBigSmem s = kktb_carve(kktb_smem, bmax, nb, mmax, qmax);
  KktVals v{bt.hess + (size_t)p * bt.nnzh, bt.jac + (size_t)p * bt.nnzj, bt.dx_diag + (size_t)p * d.nw,
            bt.neg_d + (size_t)p * d.ng};
  const double* __restrict__ rhs = bt.rhs + (size_t)p * nk;
  double* __restrict__ sol = bt.sol + (size_t)p * nk;
  double* __restrict__ Sinv_g = bt.Sinv + (size_t)p * N * bmax * d.ldS;
  double* __restrict__ YL_g = bt.YL + (size_t)p * N * bmax * d.ldY;
  double* __restrict__ X_g = bt.X + (size_t)p * N * bmax * nrhs;
  double* __restrict__ Yr = work + (size_t)p * work_stride;       // [nbb][nrhs]
  double* __restrict__ Zs = Yr + (size_t)nbb * nrhs;              // [nbb][nrhs]
  double* __restrict__ YLs = Zs + (size_t)nbb * nrhs;             // [bmax][mmax]
  double* __restrict__ Gm = YLs + (size_t)bmax * mmax;            // [nbb][nbb] border matrix
  int bad = 0, neg = 0;
  int m_prev = 0;
  const int32_t* cr_prev = nullptr;

  // ------------------------------------------------------------------------------ forward sweep
  for (int n = 0; n < N; ++n) {
    const int u0 = d.blk_ptr[n], b = d.blk_ptr[n + 1] - u0;
    const int32_t* __restrict__ unk = d.unk + u0;
    double* __restrict__ M = Sinv_g + (size_t)n * bmax * d.ldS;   // inverted in place, leading dimension ldS
    const int LD = d.ldS;
    for (int i = tid; i < b * LD; i += nthreads) M[i] = 0.0;
    for (int i = tid; i < b * nrhs; i += nthreads) Yr[i] = 0.0;
    __syncthreads();
    for (int e = d.dA_ptr[n] + tid; e < d.dA_ptr[n + 1]; e += nthreads) {
      const int pos = d.dA_pos[e];
      const int r = pos / bmax, c = pos - r * bmax;
      M[(size_t)r * LD + c] = kkt_val(v, d.dA_src[e]);
    }
    for (int e = d.bE_ptr[n] + tid; e < d.bE_ptr[n + 1]; e += nthreads)
      Yr[(size_t)d.bE_row[e] * nrhs + 1 + d.bE_col[e]] = kkt_val(v, d.bE_src[e]);
    __syncthreads();
    for (int i = tid; i < b; i += nthreads) {
      M[(size_t)i * LD + i] += kkt_diag(v, unk[i], d.nw);
      Yr[(size_t)i * nrhs] = rhs[unk[i]];
    }
    __syncthreads();
    if (n > 0) {
      for (int i = tid; i < m_prev * m_prev; i += nthreads) {
        const int a = i / m_prev, c = i - a * m_prev;
        M[(size_t)cr_prev[a] * LD + cr_prev[c]] -= s.carry[a * mmax + c];
      }
      for (int i = tid; i < m_prev * nrhs; i += nthreads) {
        const int a = i / nrhs, r = i - a * nrhs;
        Yr[(size_t)cr_prev[a] * nrhs + r] -= s.rcarry[a * nrhs + r];
      }
      __syncthreads();
    }
    bad += kktb_sym_invert(M, LD, b, s, &neg, nbb);
    // z = S^-1 y  -> Zs and the X buffer
    kktb_gemm(M, LD, Yr, nrhs, Zs, nrhs, b, b, nrhs, X_g + (size_t)n * bmax * nrhs, nrhs);
    if (n < N - 1) {
      const int m = d.cr_ptr[n + 1] - d.cr_ptr[n], q = d.cc_ptr[n + 1] - d.cc_ptr[n];
      const int32_t* __restrict__ cc = d.cc + d.cc_ptr[n];
      for (int i = tid; i < m * qmax; i += nthreads) s.Lc[i] = 0.0;
      __syncthreads();
      for (int e = d.cL_ptr[n] + tid; e < d.cL_ptr[n + 1]; e += nthreads) s.Lc[d.cL_pos[e]] = kkt_val(v, d.cL_src[e]);
      __syncthreads();
      // YL = S^-1[:, cc] Lc'   (b x m)
      for (int i = tid; i < b * m; i += nthreads) {
        const int r = i / m, a = i - r * m;
        double acc = 0.0;
        for (int t = 0; t < q; ++t) acc += M[(size_t)r * LD + cc[t]] * s.Lc[a * qmax + t];
        YLs[(size_t)r * mmax + a] = acc;
        YL_g[((size_t)n * bmax + r) * d.ldY + a] = acc;
      }
      __syncthreads();
      for (int i = tid; i < m * m; i += nthreads) {
        const int a = i / m, c = i - a * m;
        double acc = 0.0;
        for (int t = 0; t < q; ++t) acc += s.Lc[a * qmax + t] * YLs[(size_t)cc[t] * mmax + c];
        s.carry[a * mmax + c] = acc;
      }
      for (int i = tid; i < m * nrhs; i += nthreads) {
        const int a = i / nrhs, r = i - a * nrhs;
        double acc = 0.0;
        for (int t = 0; t < q; ++t) acc += s.Lc[a * qmax + t] * Zs[(size_t)cc[t] * nrhs + r];
        s.rcarry[a * nrhs + r] = acc;
      }
      m_prev = m;
      cr_prev = d.cr + d.cr_ptr[n];
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------------------ backward sweep
  // Zs holds x_{N-1}; walk down, the current block's solution alternates between Zs and Yr
  double* cur = Zs;
  double* nxt = Yr;
  for (int n = N - 2; n >= 0; --n) {
    const int b = d.blk_ptr[n + 1] - d.blk_ptr[n];
    const int m = d.cr_ptr[n + 1] - d.cr_ptr[n];
    const int32_t* __restrict__ cr = d.cr + d.cr_ptr[n];
    for (int i = tid; i < m * nrhs; i += nthreads) {
      const int a = i / nrhs, r = i - a * nrhs;
      s.rcarry[a * nrhs + r] = cur[(size_t)cr[a] * nrhs + r];
    }
    __syncthreads();
    for (int i = tid; i < b * nrhs; i += nthreads) {
      const int row = i / nrhs, r = i - row * nrhs;
      const double* __restrict__ yl = YL_g + ((size_t)n * bmax + row) * d.ldY;
      double acc = X_g[((size_t)n * bmax + row) * nrhs + r];
      for (int a = 0; a < m; ++a) acc -= yl[a] * s.rcarry[a * nrhs + r];
      nxt[(size_t)row * nrhs + r] = acc;
      X_g[((size_t)n * bmax + row) * nrhs + r] = acc;
    }
    __syncthreads();
    double* t = cur;
    cur = nxt;
    nxt = t;
  }

  // ------------------------------------------------------------------------------ border
  if (nb > 0) {
    const int32_t* __restrict__ unkb = d.unk + d.blk_ptr[N];
    const int LD = nbb;
    for (int i = tid; i < nb * LD; i += nthreads) Gm[i] = 0.0;
    __syncthreads();
    for (int e = tid; e < d.n_bG; e += nthreads) {
      const int pos = d.bG_pos[e];
      const int r = pos / nb, c = pos - r * nb;
      Gm[(size_t)r * LD + c] = kkt_val(v, d.bG_src[e]);
    }
    __syncthreads();
    for (int i = tid; i < nb; i += nthreads) {
      Gm[(size_t)i * LD + i] += kkt_diag(v, unkb[i], d.nw);
      s.rowB[i] = rhs[unkb[i]];     // rb (rowB is free outside the inversion)
    }
    __syncthreads();
    // G -= E' X_E, rb -= E' x_T : thread per (border row j, right-hand-side column r)
    for (int i = tid; i < nb * nrhs; i += nthreads) {
      const int j = i / nrhs, r = i - j * nrhs;
      double acc = 0.0;
      for (int t = d.bEc_ptr[j]; t < d.bEc_ptr[j + 1]; ++t) {
        const int e = d.bEc_idx[t];
        acc += kkt_val(v, d.bE_src[e]) * X_g[((size_t)d.bE_blk[e] * bmax + d.bE_row[e]) * nrhs + r];
      }
      if (r == 0) s.xb[j] = s.rowB[j] - acc;     // rb
      else Gm[(size_t)j * LD + (r - 1)] -= acc;
    }
    __syncthreads();
    bad += kktb_sym_invert(Gm, LD, nb, s, &neg, nbb);
    if (bt.SB)
      for (int i = tid; i < nb * nb; i += nthreads) {
        const int r = i / nb, c = i - r * nb;
        bt.SB[(size_t)p * nb * nb + i] = Gm[(size_t)r * LD + c];
      }
    for (int i = tid; i < nb; i += nthreads) {
      double acc = 0.0;
      for (int j = 0; j < nb; ++j) acc += Gm[(size_t)i * LD + j] * s.xb[j];
      s.rowA[i] = acc;
      sol[unkb[i]] = acc;
    }
    __syncthreads();
  }
  const int nchain = d.blk_ptr[N];
  for (int t = tid; t < nchain; t += nthreads) {
    int lo = 0, hi = N - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (d.blk_ptr[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int i = t - d.blk_ptr[lo];
    const double* __restrict__ x = X_g + ((size_t)lo * bmax + i) * nrhs;
    double acc = x[0];
    for (int j = 0; j < nb; ++j) acc -= x[1 + j] * s.rowA[j];
    sol[d.unk[t]] = acc;
  }
  if (tid == 0 && bt.status) {
    bt.status[2 * p] = bad;
    bt.status[2 * p + 1] = neg;
  }
}
