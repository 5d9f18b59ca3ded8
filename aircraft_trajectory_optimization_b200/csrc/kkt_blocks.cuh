// Batched block-tridiagonal + border factor/solve of the interior-point KKT system.
//
// What it replaces: the sparse symmetric-indefinite solve IPOPT performs on every iteration inside
// `self.solver(x0=..., ...)` (drone3d/raceline/base_raceline.py:160-165; linear solver chosen at
// :765-787).  The matrix is
//        [ W + diag(dx_diag)     J'        ]
//        [ J                     diag(neg_d) ]
// given by the CCS values of hess_l / jac_g and two diagonals; the grouping of unknowns into interval
// blocks + border is computed on the host (aircraft_trajectory_optimization_b200/kkt.py, which also
// documents the tables).
//
// One CTA per problem instance.  Forward sweep over the N interval blocks: assemble the diagonal block
// in shared memory, subtract the Schur carry of the previous block, invert it in place (Gauss-Jordan,
// partial pivoting, one pivot column per step), apply the inverse to the 1 + nb right-hand sides and to
// the coupling block, form the carry for the next block.  Backward sweep: x_n = z_n - YL_n x_{n+1}.
// Border: dense nb x nb Schur complement, inverted the same way.  Factors (block inverses, YL, border
// columns, border inverse) stay in HBM so that further right-hand sides (iterative refinement) cost a
// solve, not a factorisation.
#pragma once
#include "common.cuh"

#define RB_KKT_THREADS 256

struct RbKktDev {
  int nw, ng, N, nb, bmax, mmax, qmax, n_bG;
  int ldS, ldY;   // row strides of the stored block inverses / coupling solves: bmax, mmax rounded up to even, so that
                  // every block of the factor storage is a 16-byte multiple (TMA bulk copies in the re-solve)
  const int32_t *blk_ptr, *unk;
  const int32_t *dA_ptr, *dA_src, *dA_pos;
  const int32_t *cr_ptr, *cr, *cc_ptr, *cc, *cL_ptr, *cL_src, *cL_pos;
  const int32_t *bE_ptr, *bE_src, *bE_row, *bE_col;      // grouped by block
  const int32_t *bEc_ptr, *bEc_idx, *bE_blk;             // the same entries grouped by border column
  const int32_t *bG_src, *bG_pos;
  // K v products: CCS (column-wise) and CSR-ordered views of jac_g and of the upper triangle of hess_l
  const int32_t *j_colptr, *j_row, *j_rowptr, *j_col, *j_perm;
  const int32_t *h_colptr, *h_row, *h_rowptr, *h_col, *h_perm;
  // interface form of the border columns (kkt_chain.cuh; tables documented in kkt.py)
  int amax, smax;
  const int32_t *act, *sup_ptr, *sup, *crs, *bE_sup, *p_off, *q_off;
};

struct RbKktBatch {
  int B;
  int nnzh, nnzj;
  const double *hess, *jac, *dx_diag, *neg_d;   // [B][nnzh], [B][nnzj], [B][nw], [B][ng]
  const double* rhs;                            // [B][nw+ng]
  double* sol;                                  // [B][nw+ng]
  double* Sinv;                                 // [B][N][bmax*bmax]
  double* YL;                                   // [B][N][bmax*mmax]
  double* X;                                    // [B][N][bmax*(1+nb)]  z, then x; columns 1.. = T^-1 E
  double* Xr;                                   // [B][N][bmax]         single column for re-solves
  double* SB;                                   // [B][nb*nb] inverse of the border Schur complement
  int* status;                                  // [B][2] vanishing pivots met, negative eigenvalues
};

struct KktVals {
  const double *hess, *jac, *dx, *nd;
  const double* aux;      // source kind 4: values of a condensed system (kkt_condense.cuh), else unused
};

__device__ __forceinline__ double kkt_val(const KktVals& v, int32_t src) {
  const int kind = (src >> 28) & 7;
  const int idx = src & 0x0fffffff;
  const double* base = kind == 0 ? v.hess : (kind == 1 ? v.jac : (kind == 2 ? v.dx : (kind == 3 ? v.nd : v.aux)));
  return base[idx];
}

__device__ __forceinline__ double kkt_diag(const KktVals& v, int u, int nw) {
  return u < nw ? v.dx[u] : v.nd[u - nw];
}

// In-place inverse of the SYMMETRIC b x b matrix M (row-major, leading dimension LD, shared memory) by
// Gauss-Jordan sweeps with Bunch-Kaufman pivoting.
//
// Pivot choice (every warp computes it redundantly from shared memory, so no broadcast barrier is needed):
// k = remaining index with the largest diagonal entry, lambda = largest remaining off-diagonal entry of
// column k (row r), sigma = the same for column r; 1x1 pivot k when |a_kk| >= alpha lambda or
// |a_kk| sigma >= alpha lambda^2, 1x1 pivot r when |a_rr| >= alpha sigma, else the 2x2 pivot (k, r);
// alpha = (1 + sqrt 17) / 8.  The arg-max scans compare float magnitudes with one __reduce_max_sync each.
// Pivots are taken in place (symmetric pivoting only chooses the order): no permutation to undo.
//
// Update: a Gauss-Jordan step on pivot p is the rank-1 update M -= c r' (c_i = a_ip, r_j = a_pj / a_pp) away
// from the pivot row and column, which are assigned directly (folding them into the rank-1 form cancels
// catastrophically for large pivots); a 2x2 step on (p, q) is the rank-2 analogue.  Every thread keeps a
// fixed TR x TC tile of M in registers (columns strided across lanes) and mirrors it to shared memory; the
// pivot row / column cases are selects on per-tile flags, not branches.
//
// The pivot signs give the inertia: *neg accumulates the number of negative eigenvalues (a 2x2 pivot
// chosen this way has one of each sign); the return value counts vanishing pivots.
__device__ __forceinline__ void kkt_argmax32(float v, int i, int& arg) {
  // largest v over the warp (v >= 0 or -1), smallest index on ties; every lane gets the result
  const unsigned key = __float_as_uint(v < 0.f ? 0.f : v) + (v < 0.f ? 0u : 1u);
  const unsigned best = __reduce_max_sync(0xffffffffu, key);
  const unsigned cand = (key == best) ? (unsigned)i : 0x7fffffffu;
  arg = (int)__reduce_min_sync(0xffffffffu, cand);
}

__device__ __forceinline__ float kkt_mag(double a) {
  const float f = (float)fabs(a);
  return f != f ? 0.f : f;          // NaN -> 0
}

template <int TR, int TC>
__device__ int kkt_sym_invert(double* M, int LD, int b, double* c1, double* r1, double* c2, double* r2,
                              int* swept, int* neg) {
  const int tid = threadIdx.x, lane = tid & 31;
  const double alpha = 0.6403882032022076;
  int bad = 0, nneg = 0;
  const int ntc = (b + TC - 1) / TC, ntr = (b + TR - 1) / TR;
  const int ti = tid / ntc, tj = tid - ti * ntc;
  const bool owner = ti < ntr;
  const int i0 = ti * TR;
  double m[TR][TC];
#pragma unroll
  for (int a = 0; a < TR; ++a)
#pragma unroll
    for (int c = 0; c < TC; ++c) {
      const int i = i0 + a, j = tj + c * ntc;
      m[a][c] = (owner && i < b && j < b) ? M[i * LD + j] : 0.0;
    }
  unsigned swept_lo = 0u, swept_hi = 0u;   // bit i: index i / i + 32 has been pivoted
  (void)swept;
  __syncthreads();
  int remaining = b;
  while (remaining > 0) {
    // ---------------------------------------------------------------- pivot choice (per warp, redundant)
    // lane l looks at indices l and l + 32 (blocks have at most 64 unknowns); `alive` bits live in registers
    int k, r = 0, type = 1, p, q;
    const int ia = lane, ib = lane + 32;
    const bool la = ia < b && !((swept_lo >> lane) & 1u), lb = ib < b && !((swept_hi >> lane) & 1u);
    {
      const float va = la ? kkt_mag(M[ia * LD + ia]) : -1.f, vb = lb ? kkt_mag(M[ib * LD + ib]) : -1.f;
      kkt_argmax32(vb > va ? vb : va, vb > va ? ib : ia, k);
    }
    p = q = k;
    if (remaining > 1) {
      {
        const float va = (la && ia != k) ? kkt_mag(M[ia * LD + k]) : -1.f;
        const float vb = (lb && ib != k) ? kkt_mag(M[ib * LD + k]) : -1.f;
        kkt_argmax32(vb > va ? vb : va, vb > va ? ib : ia, r);
      }
      const double akk = fabs(M[k * LD + k]), lam = fabs(M[r * LD + k]);
      if (!(akk >= alpha * lam)) {
        const float va = (la && ia != r) ? kkt_mag(M[ia * LD + r]) : -1.f;
        const float vb = (lb && ib != r) ? kkt_mag(M[ib * LD + r]) : -1.f;
        int t;
        kkt_argmax32(vb > va ? vb : va, vb > va ? ib : ia, t);
        const double sig = fabs(M[t * LD + r]);
        if (akk * sig >= alpha * lam * lam) {
          p = k;
        } else if (fabs(M[r * LD + r]) >= alpha * sig) {
          p = r;
        } else {
          type = 2;
          p = k < r ? k : r;
          q = k < r ? r : k;
        }
      }
    }
    // every thread marks the pivots in its own copy of the bitmask
    if (p < 32) swept_lo |= 1u << p; else swept_hi |= 1u << (p - 32);
    if (type == 2) {
      if (q < 32) swept_lo |= 1u << q; else swept_hi |= 1u << (q - 32);
    }
    // ---------------------------------------------------------------- the update vectors
    if (type == 1) {
      double d = M[p * LD + p];
      if (!(fabs(d) > 1e-250)) {
        d = 1e-250;
        bad++;
      }
      if (d < 0) nneg++;
      const double di = 1.0 / d;
      if (tid < b) {
        const int j = tid;
        c1[j] = M[j * LD + p];
        r1[j] = (j == p) ? di : M[p * LD + j] * di;
      }
      __syncthreads();
      if (owner) {
        double cc[TR], rr[TC];
        bool rsel[TR], csel[TC];
#pragma unroll
        for (int a = 0; a < TR; ++a) {
          cc[a] = c1[i0 + a];
          rsel[a] = (i0 + a == p);
        }
#pragma unroll
        for (int c = 0; c < TC; ++c) {
          rr[c] = r1[tj + c * ntc];
          csel[c] = (tj + c * ntc == p);
        }
#pragma unroll
        for (int a = 0; a < TR; ++a)
#pragma unroll
          for (int c = 0; c < TC; ++c) {
            // pivot row: a_pj / d (1 / d on the pivot); pivot column: -a_ip / d; elsewhere the rank-1 update
            double v = csel[c] ? -cc[a] * di : m[a][c] - cc[a] * rr[c];
            v = rsel[a] ? rr[c] : v;
            m[a][c] = v;
            const int i = i0 + a, j = tj + c * ntc;
            if (i < b && j < b) M[i * LD + j] = v;
          }
      }
      remaining -= 1;
    } else {
      const double epp = M[p * LD + p], epq = M[p * LD + q], eqq = M[q * LD + q];
      double det = epp * eqq - epq * epq;
      if (!(fabs(det) > 1e-250)) {
        det = -1e-250;
        bad++;
      }
      nneg += det < 0 ? 1 : (epp + eqq < 0 ? 2 : 0);
      const double i00 = eqq / det, i01 = -epq / det, i11 = epp / det;
      if (tid < b) {
        const int j = tid;
        const double ap = M[p * LD + j], aq = M[q * LD + j];
        c1[j] = M[j * LD + p];
        c2[j] = M[j * LD + q];
        if (j == p) {
          r1[j] = i00;
          r2[j] = i01;
        } else if (j == q) {
          r1[j] = i01;
          r2[j] = i11;
        } else {
          r1[j] = i00 * ap + i01 * aq;
          r2[j] = i01 * ap + i11 * aq;
        }
      }
      __syncthreads();
      if (owner) {
        double ca[TR], cb[TR], ra[TC], rb[TC];
        int rsel[TR];
        bool csel[TC];
#pragma unroll
        for (int a = 0; a < TR; ++a) {
          ca[a] = c1[i0 + a];
          cb[a] = c2[i0 + a];
          rsel[a] = (i0 + a == p) ? 1 : ((i0 + a == q) ? 2 : 0);
        }
#pragma unroll
        for (int c = 0; c < TC; ++c) {
          const int j = tj + c * ntc;
          ra[c] = r1[j];
          rb[c] = r2[j];
          csel[c] = (j == p) || (j == q);
        }
#pragma unroll
        for (int a = 0; a < TR; ++a)
#pragma unroll
          for (int c = 0; c < TC; ++c) {
            // pivot rows: E^-1 A_Pj (E^-1 itself on the pivot block); pivot columns: -A_iP E^-1; else rank-2 update
            const double t = ca[a] * ra[c] + cb[a] * rb[c];
            double v = csel[c] ? -t : m[a][c] - t;
            v = rsel[a] == 1 ? ra[c] : (rsel[a] == 2 ? rb[c] : v);
            m[a][c] = v;
            const int i = i0 + a, j = tj + c * ntc;
            if (i < b && j < b) M[i * LD + j] = v;
          }
      }
      remaining -= 2;
    }
    __syncthreads();
  }
  *neg += nneg;
  return bad;
}

struct KktSmem {
  double *M, *Yr, *Zs, *Lc, *YLs, *carry, *rcarry, *colp, *rowp, *colq, *rowq;
  int* swept;
};
#define RB_KKT_VPAD 72   // length of the update vectors: largest block (64) + tile overhang

__host__ __device__ inline size_t kkt_smem_doubles(int bmax, int nb, int mmax, int qmax) {
  const int LD = bmax | 1, nrhs = 1 + nb, nbb = nb > bmax ? nb : bmax;
  const int LDm = nbb | 1;
  size_t n = (size_t)nbb * LDm;               // M (also hosts the border matrix)
  n += 2 * (size_t)nbb * nrhs;                // Yr, Zs
  n += (size_t)mmax * qmax;                   // Lc
  n += (size_t)bmax * mmax;                   // YLs
  n += (size_t)mmax * mmax;                   // carry
  n += (size_t)mmax * nrhs;                   // rcarry
  n += 4 * (size_t)RB_KKT_VPAD;               // the update vectors
  (void)LD;
  return n;
}

__host__ __device__ inline size_t kkt_smem_bytes(int bmax, int nb, int mmax, int qmax) {
  const int nbb = nb > bmax ? nb : bmax;
  return kkt_smem_doubles(bmax, nb, mmax, qmax) * sizeof(double) + ((size_t)nbb + 4) * sizeof(int);
}

__device__ inline KktSmem kkt_carve(double* s, int bmax, int nb, int mmax, int qmax) {
  const int nrhs = 1 + nb, nbb = nb > bmax ? nb : bmax, LDm = nbb | 1;
  KktSmem k;
  k.M = s; s += (size_t)nbb * LDm;
  k.Yr = s; s += (size_t)nbb * nrhs;
  k.Zs = s; s += (size_t)nbb * nrhs;
  k.Lc = s; s += (size_t)mmax * qmax;
  k.YLs = s; s += (size_t)bmax * mmax;
  k.carry = s; s += (size_t)mmax * mmax;
  k.rcarry = s; s += (size_t)mmax * nrhs;
  k.colp = s; s += RB_KKT_VPAD;
  k.rowp = s; s += RB_KKT_VPAD;
  k.colq = s; s += RB_KKT_VPAD;
  k.rowq = s; s += RB_KKT_VPAD;
  k.swept = reinterpret_cast<int*>(s);
  return k;
}

template <int TR, int TC>
__global__ void __launch_bounds__(RB_KKT_THREADS, TR * TC <= 8 ? 3 : 2)
kkt_factor_solve_kernel(const RbKktDev d, const RbKktBatch bt) {
  extern __shared__ double kkt_smem[];
  const int p = blockIdx.x;
  if (p >= bt.B) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int nk = d.nw + d.ng, nrhs = 1 + d.nb, bmax = d.bmax, mmax = d.mmax, qmax = d.qmax, N = d.N;
  const int nbb = d.nb > bmax ? d.nb : bmax;
  const int LD = nbb | 1;
  KktSmem s = kkt_carve(kkt_smem, bmax, d.nb, mmax, qmax);
  KktVals v{bt.hess + (size_t)p * bt.nnzh, bt.jac + (size_t)p * bt.nnzj, bt.dx_diag + (size_t)p * d.nw,
            bt.neg_d + (size_t)p * d.ng};
  const double* __restrict__ rhs = bt.rhs + (size_t)p * nk;
  double* __restrict__ sol = bt.sol + (size_t)p * nk;
  double* __restrict__ Sinv_g = bt.Sinv + (size_t)p * N * bmax * d.ldS;
  double* __restrict__ YL_g = bt.YL + (size_t)p * N * bmax * d.ldY;
  double* __restrict__ X_g = bt.X + (size_t)p * N * bmax * nrhs;
  int bad = 0, neg = 0;
  int m_prev = 0;
  const int32_t* cr_prev = nullptr;

  // ------------------------------------------------------------------------------ forward sweep
  for (int n = 0; n < N; ++n) {
    const int u0 = d.blk_ptr[n], b = d.blk_ptr[n + 1] - u0;
    const int32_t* __restrict__ unk = d.unk + u0;
    for (int i = tid; i < b * LD; i += blockDim.x) s.M[i] = 0.0;
    for (int i = tid; i < b * nrhs; i += blockDim.x) s.Yr[i] = 0.0;
    __syncthreads();
    for (int e = d.dA_ptr[n] + tid; e < d.dA_ptr[n + 1]; e += blockDim.x) {
      const int pos = d.dA_pos[e];
      const int r = pos / bmax, c = pos - r * bmax;
      s.M[r * LD + c] = kkt_val(v, d.dA_src[e]);
    }
    for (int e = d.bE_ptr[n] + tid; e < d.bE_ptr[n + 1]; e += blockDim.x)
      s.Yr[d.bE_row[e] * nrhs + 1 + d.bE_col[e]] = kkt_val(v, d.bE_src[e]);
    __syncthreads();
    for (int i = tid; i < b; i += blockDim.x) {
      s.M[i * LD + i] += kkt_diag(v, unk[i], d.nw);
      s.Yr[i * nrhs] = rhs[unk[i]];
    }
    __syncthreads();
    if (n > 0) {
      for (int i = tid; i < m_prev * m_prev; i += blockDim.x) {
        const int a = i / m_prev, c = i - a * m_prev;
        s.M[cr_prev[a] * LD + cr_prev[c]] -= s.carry[a * mmax + c];
      }
      for (int i = tid; i < m_prev * nrhs; i += blockDim.x) {
        const int a = i / nrhs, r = i - a * nrhs;
        s.Yr[cr_prev[a] * nrhs + r] -= s.rcarry[a * nrhs + r];
      }
      __syncthreads();
    }
    bad += kkt_sym_invert<TR, TC>(s.M, LD, b, s.colp, s.rowp, s.colq, s.rowq, s.swept, &neg);

    // z = S^-1 y: 2 x 4 output tiles per thread (rows i0, i0+1; right-hand sides strided across lanes)
    {
      const int nct = (nrhs + 3) >> 2, nrt = (b + 1) >> 1;
      for (int tile = tid; tile < nrt * nct; tile += blockDim.x) {
        const int tr_ = tile / nct, tc_ = tile - tr_ * nct;
        const int i0 = tr_ * 2, i1 = (i0 + 1 < b) ? i0 + 1 : i0;
        int rc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) rc[c] = (tc_ + c * nct < nrhs) ? tc_ + c * nct : tc_;
        double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
        for (int j = 0; j < b; ++j) {
          const double m0 = s.M[i0 * LD + j], m1 = s.M[i1 * LD + j];
          const double* __restrict__ yj = s.Yr + j * nrhs;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double y = yj[rc[c]];
            acc[0][c] += m0 * y;
            acc[1][c] += m1 * y;
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (tc_ + c * nct < nrhs) {
            s.Zs[i0 * nrhs + rc[c]] = acc[0][c];
            X_g[((size_t)n * bmax + i0) * nrhs + rc[c]] = acc[0][c];
            if (i0 + 1 < b) {
              s.Zs[i1 * nrhs + rc[c]] = acc[1][c];
              X_g[((size_t)n * bmax + i1) * nrhs + rc[c]] = acc[1][c];
            }
          }
        }
      }
    }
    // keep the inverse
    if (bt.Sinv) {
      for (int i = warp; i < b; i += nwarps)
        for (int j = lane; j < b; j += 32) Sinv_g[((size_t)n * bmax + i) * d.ldS + j] = s.M[i * LD + j];
    }
    if (n < N - 1) {
      const int m = d.cr_ptr[n + 1] - d.cr_ptr[n], q = d.cc_ptr[n + 1] - d.cc_ptr[n];
      const int32_t* __restrict__ cc = d.cc + d.cc_ptr[n];
      for (int i = tid; i < m * qmax; i += blockDim.x) s.Lc[i] = 0.0;
      __syncthreads();
      for (int e = d.cL_ptr[n] + tid; e < d.cL_ptr[n + 1]; e += blockDim.x) s.Lc[d.cL_pos[e]] = kkt_val(v, d.cL_src[e]);
      __syncthreads();
      // YL = S^-1[:, cc] Lc'   (b x m): rows across lanes, coupling rows across warps
      for (int a = warp; a < m; a += nwarps) {
        for (int i = lane; i < b; i += 32) {
          double acc = 0.0;
          for (int t = 0; t < q; ++t) acc += s.M[i * LD + cc[t]] * s.Lc[a * qmax + t];
          s.YLs[i * mmax + a] = acc;
          YL_g[((size_t)n * bmax + i) * d.ldY + a] = acc;
        }
      }
      __syncthreads();
      // carry = Lc YL[cc, :]  (m x m),  rcarry = Lc z[cc, :]  (m x nrhs)
      for (int i = tid; i < m * m; i += blockDim.x) {
        const int a = i / m, c = i - a * m;
        double acc = 0.0;
        for (int t = 0; t < q; ++t) acc += s.Lc[a * qmax + t] * s.YLs[cc[t] * mmax + c];
        s.carry[a * mmax + c] = acc;
      }
      for (int i = tid; i < m * nrhs; i += blockDim.x) {
        const int a = i / nrhs, r = i - a * nrhs;
        double acc = 0.0;
        for (int t = 0; t < q; ++t) acc += s.Lc[a * qmax + t] * s.Zs[cc[t] * nrhs + r];
        s.rcarry[a * nrhs + r] = acc;
      }
      m_prev = m;
      cr_prev = d.cr + d.cr_ptr[n];
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------------------ backward sweep
  // Zs holds x_{N-1} (= z_{N-1}); walk down, keeping the current block's solution in Zs
  for (int n = N - 2; n >= 0; --n) {
    const int b = d.blk_ptr[n + 1] - d.blk_ptr[n];
    const int m = d.cr_ptr[n + 1] - d.cr_ptr[n];
    const int32_t* __restrict__ cr = d.cr + d.cr_ptr[n];
    // gather the needed rows of x_{n+1}
    for (int i = tid; i < m * nrhs; i += blockDim.x) {
      const int a = i / nrhs, r = i - a * nrhs;
      s.rcarry[a * nrhs + r] = s.Zs[cr[a] * nrhs + r];
    }
    __syncthreads();
    // x_n = z_n - YL_n x_{n+1}[cr]: 2 x 4 output tiles per thread, YL_n rows read once per tile
    {
      const int nct = (nrhs + 3) >> 2, nrt = (b + 1) >> 1;
      for (int tile = tid; tile < nrt * nct; tile += blockDim.x) {
        const int tr_ = tile / nct, tc_ = tile - tr_ * nct;
        const int i0 = tr_ * 2, i1 = (i0 + 1 < b) ? i0 + 1 : i0;
        int rc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) rc[c] = (tc_ + c * nct < nrhs) ? tc_ + c * nct : tc_;
        const double* __restrict__ y0 = YL_g + ((size_t)n * bmax + i0) * d.ldY;
        const double* __restrict__ y1 = YL_g + ((size_t)n * bmax + i1) * d.ldY;
        double acc[2][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[0][c] = X_g[((size_t)n * bmax + i0) * nrhs + rc[c]];
          acc[1][c] = X_g[((size_t)n * bmax + i1) * nrhs + rc[c]];
        }
        for (int a = 0; a < m; ++a) {
          const double l0 = y0[a], l1 = y1[a];
          const double* __restrict__ xr = s.rcarry + a * nrhs;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double xv = xr[rc[c]];
            acc[0][c] -= l0 * xv;
            acc[1][c] -= l1 * xv;
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (tc_ + c * nct < nrhs) {
            s.Zs[i0 * nrhs + rc[c]] = acc[0][c];
            X_g[((size_t)n * bmax + i0) * nrhs + rc[c]] = acc[0][c];
            if (i0 + 1 < b) {
              s.Zs[i1 * nrhs + rc[c]] = acc[1][c];
              X_g[((size_t)n * bmax + i1) * nrhs + rc[c]] = acc[1][c];
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------------------ border
  const int nb = d.nb;
  double* xb = s.colp;   // reused after the inversion below
  if (nb > 0) {
    const int ub0 = d.blk_ptr[N];
    const int32_t* __restrict__ unkb = d.unk + ub0;
    for (int i = tid; i < nb * LD; i += blockDim.x) s.M[i] = 0.0;
    __syncthreads();
    for (int e = tid; e < d.n_bG; e += blockDim.x) {
      const int pos = d.bG_pos[e];
      const int r = pos / nb, c = pos - r * nb;
      s.M[r * LD + c] = kkt_val(v, d.bG_src[e]);
    }
    __syncthreads();
    for (int i = tid; i < nb; i += blockDim.x) {
      s.M[i * LD + i] += kkt_diag(v, unkb[i], d.nw);
      s.Yr[i] = rhs[unkb[i]];
    }
    __syncthreads();
    // G -= E' X_E, rb -= E' x_T : one border column j per warp, lanes over the nrhs columns of X
    for (int j = warp; j < nb; j += nwarps) {
      for (int r = lane; r < nrhs; r += 32) {
        double acc = 0.0;
        for (int t = d.bEc_ptr[j]; t < d.bEc_ptr[j + 1]; ++t) {
          const int e = d.bEc_idx[t];
          acc += kkt_val(v, d.bE_src[e]) * X_g[((size_t)d.bE_blk[e] * bmax + d.bE_row[e]) * nrhs + r];
        }
        if (r == 0) s.Yr[j] -= acc;
        else s.M[j * LD + (r - 1)] -= acc;
      }
    }
    __syncthreads();
    bad += kkt_sym_invert<TR, TC>(s.M, LD, nb, s.colp, s.rowp, s.colq, s.rowq, s.swept, &neg);
    if (bt.SB)
      for (int i = tid; i < nb * nb; i += blockDim.x) {
        const int r = i / nb, c = i - r * nb;
        bt.SB[(size_t)p * nb * nb + i] = s.M[r * LD + c];
      }
    __syncthreads();   // colp is free again
    for (int i = tid; i < nb; i += blockDim.x) {
      double acc = 0.0;
      for (int j = 0; j < nb; ++j) acc += s.M[i * LD + j] * s.Yr[j];
      xb[i] = acc;
      sol[unkb[i]] = acc;
    }
    __syncthreads();
  }
  // x = x_T - X_E x_b, scattered back to the (w, g) order
  const int nchain = d.blk_ptr[N];
  for (int t = tid; t < nchain; t += blockDim.x) {
    // block of unknown t: binary search over blk_ptr
    int lo = 0, hi = N - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (d.blk_ptr[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int i = t - d.blk_ptr[lo];
    const double* __restrict__ x = X_g + ((size_t)lo * bmax + i) * nrhs;
    double acc = x[0];
    for (int j = 0; j < nb; ++j) acc -= x[1 + j] * xb[j];
    sol[d.unk[t]] = acc;
  }
  if (tid == 0 && bt.status) {
    bt.status[2 * p] = bad;
    bt.status[2 * p + 1] = neg;
  }
}

// Another right-hand side with the stored factors (iterative refinement, corrections).
// E values are re-gathered from hess / jac, which must be unchanged since the factorisation.
__global__ void __launch_bounds__(RB_KKT_THREADS)
kkt_resolve_kernel(const RbKktDev d, const RbKktBatch bt) {
  extern __shared__ double kkt_smem[];
  const int p = blockIdx.x;
  if (p >= bt.B) return;
  const int tid = threadIdx.x;
  const int nk = d.nw + d.ng, nrhs = 1 + d.nb, bmax = d.bmax, mmax = d.mmax, N = d.N, nb = d.nb;
  const int nbb = nb > bmax ? nb : bmax;
  double* y = kkt_smem;             // [nbb] current block rhs / z
  double* xn = y + nbb;             // [nbb] solution of the block above (backward)
  double* rc = xn + nbb;            // [mmax] carry
  double* xb = rc + mmax;           // [nb]
  KktVals v{bt.hess + (size_t)p * bt.nnzh, bt.jac + (size_t)p * bt.nnzj, bt.dx_diag + (size_t)p * d.nw,
            bt.neg_d + (size_t)p * d.ng};
  const double* __restrict__ rhs = bt.rhs + (size_t)p * nk;
  double* __restrict__ sol = bt.sol + (size_t)p * nk;
  const double* __restrict__ Sinv_g = bt.Sinv + (size_t)p * N * bmax * d.ldS;
  const double* __restrict__ YL_g = bt.YL + (size_t)p * N * bmax * d.ldY;
  const double* __restrict__ X_g = bt.X + (size_t)p * N * bmax * nrhs;
  double* __restrict__ Xr = bt.Xr + (size_t)p * N * bmax;
  int m_prev = 0;
  const int32_t* cr_prev = nullptr;
  for (int n = 0; n < N; ++n) {
    const int u0 = d.blk_ptr[n], b = d.blk_ptr[n + 1] - u0;
    for (int i = tid; i < b; i += blockDim.x) y[i] = rhs[d.unk[u0 + i]];
    __syncthreads();
    if (n > 0) {
      for (int a = tid; a < m_prev; a += blockDim.x) y[cr_prev[a]] -= rc[a];
      __syncthreads();
    }
    for (int i = tid; i < b; i += blockDim.x) {
      const double* __restrict__ si = Sinv_g + ((size_t)n * bmax + i) * d.ldS;
      double acc = 0.0;
      for (int j = 0; j < b; ++j) acc += si[j] * y[j];
      Xr[(size_t)n * bmax + i] = acc;
    }
    if (n < N - 1) {
      const int m = d.cr_ptr[n + 1] - d.cr_ptr[n];
      // L_n S_n^-1 y_n = YL_n' y_n   (S_n is symmetric)
      for (int a = tid; a < m; a += blockDim.x) {
        double acc = 0.0;
        for (int i = 0; i < b; ++i) acc += YL_g[((size_t)n * bmax + i) * d.ldY + a] * y[i];
        rc[a] = acc;
      }
      m_prev = m;
      cr_prev = d.cr + d.cr_ptr[n];
    }
    __syncthreads();
  }
  {
    const int b = d.blk_ptr[N] - d.blk_ptr[N - 1];
    for (int i = tid; i < b; i += blockDim.x) xn[i] = Xr[(size_t)(N - 1) * bmax + i];
    __syncthreads();
  }
  for (int n = N - 2; n >= 0; --n) {
    const int b = d.blk_ptr[n + 1] - d.blk_ptr[n];
    const int m = d.cr_ptr[n + 1] - d.cr_ptr[n];
    const int32_t* __restrict__ cr = d.cr + d.cr_ptr[n];
    for (int a = tid; a < m; a += blockDim.x) rc[a] = xn[cr[a]];
    __syncthreads();
    for (int i = tid; i < b; i += blockDim.x) {
      const double* __restrict__ yl = YL_g + ((size_t)n * bmax + i) * d.ldY;
      double acc = Xr[(size_t)n * bmax + i];
      for (int a = 0; a < m; ++a) acc -= yl[a] * rc[a];
      xn[i] = acc;
      Xr[(size_t)n * bmax + i] = acc;
    }
    __syncthreads();
  }
  if (nb > 0) {
    const int32_t* __restrict__ unkb = d.unk + d.blk_ptr[N];
    for (int j = tid; j < nb; j += blockDim.x) {
      double acc = rhs[unkb[j]];
      for (int t = d.bEc_ptr[j]; t < d.bEc_ptr[j + 1]; ++t) {
        const int e = d.bEc_idx[t];
        acc -= kkt_val(v, d.bE_src[e]) * Xr[(size_t)d.bE_blk[e] * bmax + d.bE_row[e]];
      }
      y[j] = acc;
    }
    __syncthreads();
    for (int i = tid; i < nb; i += blockDim.x) {
      double acc = 0.0;
      for (int j = 0; j < nb; ++j) acc += bt.SB[(size_t)p * nb * nb + i * nb + j] * y[j];
      xb[i] = acc;
      sol[unkb[i]] = acc;
    }
    __syncthreads();
  }
  const int nchain = d.blk_ptr[N];
  for (int t = tid; t < nchain; t += blockDim.x) {
    int lo = 0, hi = N - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (d.blk_ptr[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int i = t - d.blk_ptr[lo];
    const double* __restrict__ x = X_g + ((size_t)lo * bmax + i) * nrhs;
    double acc = Xr[(size_t)lo * bmax + i];
    for (int j = 0; j < nb; ++j) acc -= x[1 + j] * xb[j];
    sol[d.unk[t]] = acc;
  }
}

// The same re-solve for stage blocks that fit shared memory: the block inverse S_n^-1 and the coupling solve YL_n
// of the next block are fetched by TMA bulk copies (cp.async.bulk + mbarrier, double-buffered) while the current
// block is applied, so the sequential walk over the chain is not exposed to global-memory latency; eight threads
// share every row of the matrix-vector products.
#define RB_KKT_RBUF 3   // factor blocks in flight per CTA in the re-solve (TMA pipeline depth)
__global__ void __launch_bounds__(RB_KKT_THREADS)
kkt_resolve_tma_kernel(const RbKktDev d, const RbKktBatch bt) {
  extern __shared__ __align__(128) double kkt_smem[];
  __shared__ __align__(8) unsigned long long bars[RB_KKT_RBUF];
  const int p = blockIdx.x;
  if (p >= bt.B) return;
  const int tid = threadIdx.x;
  const int nk = d.nw + d.ng, nrhs = 1 + d.nb, bmax = d.bmax, mmax = d.mmax, N = d.N, nb = d.nb;
  const int nbb = nb > bmax ? nb : bmax;
  const int SZ_S = bmax * d.ldS, SZ_Y = bmax * d.ldY;                // doubles per block of the two factor arrays (even)
  const int BUF = SZ_S + SZ_Y;
  double* buf0 = kkt_smem;
  double* y = kkt_smem + RB_KKT_RBUF * BUF;     // [nbb] current block rhs
  double* xn = y + nbb;               // [nbb] solution of the block above (backward)
  double* rc = xn + nbb;              // [mmax] carry
  double* xb = rc + mmax;             // [nb]
  KktVals v{bt.hess + (size_t)p * bt.nnzh, bt.jac + (size_t)p * bt.nnzj, bt.dx_diag + (size_t)p * d.nw,
            bt.neg_d + (size_t)p * d.ng};
  const double* __restrict__ rhs = bt.rhs + (size_t)p * nk;
  double* __restrict__ sol = bt.sol + (size_t)p * nk;
  const double* __restrict__ Sinv_g = bt.Sinv + (size_t)p * N * SZ_S;
  const double* __restrict__ YL_g = bt.YL + (size_t)p * N * SZ_Y;
  const double* __restrict__ X_g = bt.X + (size_t)p * N * bmax * nrhs;
  double* __restrict__ Xr = bt.Xr + (size_t)p * N * bmax;

  auto issue = [&](int step, int n, bool with_s) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[step % RB_KKT_RBUF]);
    double* dstS = buf0 + (size_t)(step % RB_KKT_RBUF) * BUF;
    const unsigned bytes = (unsigned)((with_s ? SZ_S : 0) + SZ_Y) * 8u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    if (with_s)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       (unsigned)__cvta_generic_to_shared(dstS)),
                   "l"(Sinv_g + (size_t)n * SZ_S), "r"((unsigned)SZ_S * 8u), "r"(bar)
                   : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dstS + SZ_S)),
                 "l"(YL_g + (size_t)n * SZ_Y), "r"((unsigned)SZ_Y * 8u), "r"(bar)
                 : "memory");
  };
  auto wait = [&](int step) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[step % RB_KKT_RBUF]);
    const unsigned parity = (step / RB_KKT_RBUF) & 1;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KKTR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KKTR_DONE_%=;\n"
        "bra KKTR_WAIT_%=;\n"
        "KKTR_DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
  };
  if (tid == 0) {
    for (int i = 0; i < RB_KKT_RBUF; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // steps 0..N-1: forward over blocks 0..N-1 (S^-1 and YL); steps N..2N-2: backward over blocks N-2..0 (YL only)
  const int nsteps = 2 * N - 1;
  auto block_of = [&](int step) { return step < N ? step : 2 * N - 2 - step; };
  if (tid == 0) {
    for (int s0 = 0; s0 < RB_KKT_RBUF && s0 < nsteps; ++s0) issue(s0, block_of(s0), s0 < N);
  }
  const int grp = tid >> 3, part = tid & 7;          // eight threads per row of a product
  int m_prev = 0;
  const int32_t* cr_prev = nullptr;
  // the right-hand-side entries of the next block are gathered one step ahead (blocks have at most 64 unknowns)
  double y_next = (tid < d.blk_ptr[1] - d.blk_ptr[0]) ? rhs[d.unk[d.blk_ptr[0] + tid]] : 0.0;
  for (int n = 0; n < N; ++n) {
    const int step = n;
    const int u0 = d.blk_ptr[n], b = d.blk_ptr[n + 1] - u0;
    if (tid < b) y[tid] = y_next;
    __syncthreads();
    if (n + 1 < N) {
      const int u1 = d.blk_ptr[n + 1], b1 = d.blk_ptr[n + 2] - u1;
      if (tid < b1) y_next = rhs[d.unk[u1 + tid]];
    }
    if (n > 0) {
      for (int a = tid; a < m_prev; a += blockDim.x) y[cr_prev[a]] -= rc[a];
      __syncthreads();
    }
    wait(step);
    const double* __restrict__ Sn = buf0 + (size_t)(step % RB_KKT_RBUF) * BUF;
    const double* __restrict__ Yn = Sn + SZ_S;
    // (loop bounds are uniform over the block: every lane takes part in the shuffles)
    for (int ib = 0; ib < b; ib += RB_KKT_THREADS / 8) {
      const int i = ib + grp;
      double acc = 0.0;
      if (i < b)
        for (int j = part; j < b; j += 8) acc += Sn[i * d.ldS + j] * y[j];
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (part == 0 && i < b) Xr[(size_t)n * bmax + i] = acc;
    }
    if (n < N - 1) {
      const int m = d.cr_ptr[n + 1] - d.cr_ptr[n];
      // L_n S_n^-1 y_n = YL_n' y_n   (S_n is symmetric)
      for (int ab = 0; ab < m; ab += RB_KKT_THREADS / 8) {
        const int a = ab + grp;
        double acc = 0.0;
        if (a < m)
          for (int i = part; i < b; i += 8) acc += Yn[i * d.ldY + a] * y[i];
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (part == 0 && a < m) rc[a] = acc;
      }
      m_prev = m;
      cr_prev = d.cr + d.cr_ptr[n];
    }
    __syncthreads();                                   // this buffer and y are free again
    if (tid == 0 && step + RB_KKT_RBUF < nsteps) issue(step + RB_KKT_RBUF, block_of(step + RB_KKT_RBUF), step + RB_KKT_RBUF < N);
  }
  {
    const int b = d.blk_ptr[N] - d.blk_ptr[N - 1];
    for (int i = tid; i < b; i += blockDim.x) xn[i] = Xr[(size_t)(N - 1) * bmax + i];
    __syncthreads();
  }
  // z_n of the next block down is read one step ahead
  double z_next = (N >= 2 && tid < d.blk_ptr[N - 1] - d.blk_ptr[N - 2]) ? Xr[(size_t)(N - 2) * bmax + tid] : 0.0;
  for (int n = N - 2; n >= 0; --n) {
    const int step = 2 * N - 2 - n;
    const int b = d.blk_ptr[n + 1] - d.blk_ptr[n];
    const double z_cur = z_next;
    if (n > 0 && tid < d.blk_ptr[n] - d.blk_ptr[n - 1]) z_next = Xr[(size_t)(n - 1) * bmax + tid];
    const int m = d.cr_ptr[n + 1] - d.cr_ptr[n];
    const int32_t* __restrict__ cr = d.cr + d.cr_ptr[n];
    for (int a = tid; a < m; a += blockDim.x) rc[a] = xn[cr[a]];
    wait(step);
    __syncthreads();
    const double* __restrict__ Yn = buf0 + (size_t)(step % RB_KKT_RBUF) * BUF + SZ_S;
    if (tid < b) {
      double acc = z_cur;
      for (int a = 0; a < m; ++a) acc -= Yn[tid * d.ldY + a] * rc[a];
      y[tid] = acc;
    }
    __syncthreads();
    for (int i = tid; i < b; i += blockDim.x) {
      xn[i] = y[i];
      Xr[(size_t)n * bmax + i] = y[i];
    }
    if (tid == 0 && step + RB_KKT_RBUF < nsteps) issue(step + RB_KKT_RBUF, block_of(step + RB_KKT_RBUF), false);
    __syncthreads();
  }
  if (nb > 0) {
    const int32_t* __restrict__ unkb = d.unk + d.blk_ptr[N];
    for (int j = tid; j < nb; j += blockDim.x) {
      double acc = rhs[unkb[j]];
      for (int t = d.bEc_ptr[j]; t < d.bEc_ptr[j + 1]; ++t) {
        const int e = d.bEc_idx[t];
        acc -= kkt_val(v, d.bE_src[e]) * Xr[(size_t)d.bE_blk[e] * bmax + d.bE_row[e]];
      }
      y[j] = acc;
    }
    __syncthreads();
    for (int i = tid; i < nb; i += blockDim.x) {
      double acc = 0.0;
      for (int j = 0; j < nb; ++j) acc += bt.SB[(size_t)p * nb * nb + i * nb + j] * y[j];
      xb[i] = acc;
      sol[unkb[i]] = acc;
    }
    __syncthreads();
  }
  const int nchain = d.blk_ptr[N];
  for (int t = tid; t < nchain; t += blockDim.x) {
    int lo = 0, hi = N - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (d.blk_ptr[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int i = t - d.blk_ptr[lo];
    const double* __restrict__ x = X_g + ((size_t)lo * bmax + i) * nrhs;
    double acc = Xr[(size_t)lo * bmax + i];
    for (int j = 0; j < nb; ++j) acc -= x[1 + j] * xb[j];
    sol[d.unk[t]] = acc;
  }
}

// out = K v  for the KKT matrix above (used for residuals in iterative refinement and for tests).
// One thread per output entry; fixed summation order (deterministic).
__global__ void kkt_matvec_kernel(const RbKktDev d, const RbKktBatch bt, const double* __restrict__ vec,
                                  double* __restrict__ out) {
  const int nk = d.nw + d.ng;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)bt.B * nk) return;
  const int p = (int)(t / nk);
  const int i = (int)(t - (long long)p * nk);
  const double* __restrict__ hess = bt.hess + (size_t)p * bt.nnzh;
  const double* __restrict__ jac = bt.jac + (size_t)p * bt.nnzj;
  const double* __restrict__ x = vec + (size_t)p * nk;
  double acc;
  if (i < d.nw) {
    acc = bt.dx_diag[(size_t)p * d.nw + i] * x[i];
    for (int k = d.h_colptr[i]; k < d.h_colptr[i + 1]; ++k) acc += hess[k] * x[d.h_row[k]];          // rows <= i
    for (int k = d.h_rowptr[i]; k < d.h_rowptr[i + 1]; ++k) {
      const int c = d.h_col[k];
      if (c != i) acc += hess[d.h_perm[k]] * x[c];                                                   // cols > i
    }
    for (int k = d.j_colptr[i]; k < d.j_colptr[i + 1]; ++k) acc += jac[k] * x[d.nw + d.j_row[k]];     // J' y
  } else {
    const int r = i - d.nw;
    acc = bt.neg_d[(size_t)p * d.ng + r] * x[i];
    for (int k = d.j_rowptr[r]; k < d.j_rowptr[r + 1]; ++k) acc += jac[d.j_perm[k]] * x[d.j_col[k]];
  }
  out[(size_t)p * nk + i] = acc;
}
