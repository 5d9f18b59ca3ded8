// Condensing of collocation intervals before the chain factorisation (host tables: kkt_condensed.py).
//
// What it replaces: the same sparse symmetric-indefinite solve as kkt_blocks.cuh / kkt_chain.cuh (IPOPT's linear solver
// inside `self.solver(x0=...)`, drone3d/raceline/base_raceline.py:160-165, :765-787) for problems transcribed with K = 7
// collocation points per interval (base_raceline.py:398-490).  The ~280 interior unknowns of an interval (collocation
// states / inputs / rates, step size, defect and rate rows) couple to the rest of the problem only through ~35 separator
// unknowns (node variables, continuity rows), so
//
//   kktc_interior_factor_kernel   one CTA per (interval, instance), all intervals at once: A_n^-1 in place (Gauss-Jordan
//                                 with Bunch-Kaufman pivoting, kktb_sym_invert), G_n = A_n^-1 B_n, T_n = B_n' G_n,
//                                 y_n = A_n^-1 r_n, t_n = B_n' y_n, inertia of A_n
//   kktc_gather_kernel            aux = C entries - sum T_n entries (fixed order: deterministic), rhs2 = rhs - sum t_n
//   kkt_factor_kernel / kkt_solve_kernel (kkt_chain.cuh) on the reduced multiple-shooting-shaped system, values from aux
//   kktc_interior_back_kernel     x_n = y_n - G_n x_sep
//
// and for further right-hand sides with stored factors kktc_interior_fwd_kernel (y_n = A_n^-1 r_n, t_n = G_n' r_n).
#pragma once
#include "kkt_big.cuh"

struct RbKktCondDev {
  int NI, amax, smax, ldA, ldG, n_aux, nk;
  const int32_t *iu_ptr, *iunk, *su_ptr, *sunk;
  const int32_t *iA_ptr, *iA_pos, *iA_src, *iB_ptr, *iB_pos, *iB_src;
  const int32_t *aux_orig, *aux_c_ptr, *aux_c_idx;
  const int32_t *rpos, *r_c_ptr, *r_c_idx;     // rpos[u]: index into r_c_ptr of KKT unknown u, -1 if nothing is subtracted
};

struct RbKktCondBatch {
  int B, nnzh, nnzj, nw, ng;
  const double *hess, *jac, *dx_diag, *neg_d;
  const double* rhs;
  double* sol;
  double* Ainv;   // [B][NI][amax * ldA]
  double* G;      // [B][NI][amax * ldG]
  double* Y;      // [B][NI][amax]
  double* Tbuf;   // [B][NI][smax * smax]
  double* tbuf;   // [B][NI][smax]
  double* aux;    // [B][n_aux]
  double* rhs2;   // [B][nk]
  int* istat;     // [B][2]
};

__host__ __device__ inline size_t kktc_cond_smem_bytes(int amax, int smax) {
  return kktb_smem_bytes(amax, 0, 1, 1) + ((size_t)amax * smax + 2 * (size_t)amax + 8) * sizeof(double);
}

// y[i] = sum_j M[i][j] r[j]: one warp per row
__device__ inline void kktc_matvec_rows(const double* __restrict__ M, int LD, int a, const double* __restrict__ r,
                                        double* __restrict__ y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = warp; i < a; i += nwarps) {
    const double* __restrict__ row = M + (size_t)i * LD;
    double acc = 0.0;
    for (int j = lane; j < a; j += 32) acc += row[j] * r[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[i] = acc;
  }
}

__global__ void __launch_bounds__(RB_KKTB_THREADS, 1)
kktc_interior_factor_kernel(const RbKktCondDev c, const RbKktCondBatch bt) {
  extern __shared__ double kktcond_smem[];
  const int n = blockIdx.x, p = blockIdx.y;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int amax = c.amax, smax = c.smax, LD = c.ldA, ldG = c.ldG;
  KktBigSmem s = kktb_carve(kktcond_smem, amax, 0, 1, 1);
  double* __restrict__ Bd = kktcond_smem + kktb_smem_bytes(amax, 0, 1, 1) / sizeof(double);   // [amax][smax]
  double* __restrict__ rI = Bd + (size_t)amax * smax;
  double* __restrict__ yv = rI + amax;
  KktVals v{bt.hess + (size_t)p * bt.nnzh, bt.jac + (size_t)p * bt.nnzj, bt.dx_diag + (size_t)p * bt.nw,
            bt.neg_d + (size_t)p * bt.ng, nullptr};
  const int u0 = c.iu_ptr[n], a = c.iu_ptr[n + 1] - u0;
  const int s0 = c.su_ptr[n], ns = c.su_ptr[n + 1] - s0;
  const int32_t* __restrict__ unk = c.iunk + u0;
  const size_t slot = (size_t)p * c.NI + n;
  double* __restrict__ M = bt.Ainv + slot * amax * LD;
  double* __restrict__ Gn = bt.G + slot * amax * ldG;
  const double* __restrict__ rhs = bt.rhs + (size_t)p * c.nk;
  for (int i = tid; i < a * LD; i += nthreads) M[i] = 0.0;
  for (int i = tid; i < a * smax; i += nthreads) Bd[i] = 0.0;
  __syncthreads();
  for (int e = c.iA_ptr[n] + tid; e < c.iA_ptr[n + 1]; e += nthreads) {
    const int pos = c.iA_pos[e];
    const int r = pos / amax, cc = pos - r * amax;
    M[(size_t)r * LD + cc] = kkt_val(v, c.iA_src[e]);
  }
  for (int e = c.iB_ptr[n] + tid; e < c.iB_ptr[n + 1]; e += nthreads) Bd[c.iB_pos[e]] = kkt_val(v, c.iB_src[e]);
  __syncthreads();
  for (int i = tid; i < a; i += nthreads) {
    M[(size_t)i * LD + i] += kkt_diag(v, unk[i], bt.nw);
    rI[i] = rhs[unk[i]];
  }
  __syncthreads();
  int neg = 0;
  const int bad = kktb_sym_invert(M, LD, a, s, &neg, amax);
  __syncthreads();
  // G = A^-1 B, y = A^-1 r
  kktb_gemm(M, LD, Bd, smax, Gn, ldG, a, a, ns, nullptr, 0);
  kktc_matvec_rows(M, LD, a, rI, yv);
  __syncthreads();
  double* __restrict__ Yg = bt.Y + slot * amax;
  for (int i = tid; i < a; i += nthreads) Yg[i] = yv[i];
  // T = B' G (upper triangle, mirrored), t = B' y
  double* __restrict__ T = bt.Tbuf + slot * smax * smax;
  for (int i = tid; i < ns * ns; i += nthreads) {
    const int ra = i / ns, cb = i - ra * ns;
    if (ra > cb) continue;
    double acc = 0.0;
    for (int j = 0; j < a; ++j) acc += Bd[j * smax + ra] * Gn[(size_t)j * ldG + cb];
    T[ra * smax + cb] = acc;
    T[cb * smax + ra] = acc;
  }
  double* __restrict__ tb = bt.tbuf + slot * smax;
  for (int ra = tid; ra < ns; ra += nthreads) {
    double acc = 0.0;
    for (int j = 0; j < a; ++j) acc += Bd[j * smax + ra] * yv[j];
    tb[ra] = acc;
  }
  if (tid == 0 && bt.istat) {
    atomicAdd(bt.istat + 2 * p, bad);
    atomicAdd(bt.istat + 2 * p + 1, neg);
  }
}

// aux and rhs2 of the reduced system; `with_aux` = 0 for re-solves (aux unchanged)
__global__ void kktc_gather_kernel(const RbKktCondDev c, const RbKktCondBatch bt, int with_aux) {
  const int p = blockIdx.y;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const double* __restrict__ Tb = bt.Tbuf + (size_t)p * c.NI * c.smax * c.smax;
  const double* __restrict__ tb = bt.tbuf + (size_t)p * c.NI * c.smax;
  if (with_aux && t < c.n_aux) {
    KktVals v{bt.hess + (size_t)p * bt.nnzh, bt.jac + (size_t)p * bt.nnzj, bt.dx_diag + (size_t)p * bt.nw,
              bt.neg_d + (size_t)p * bt.ng, nullptr};
    const int o = c.aux_orig[t];
    double acc = o >= 0 ? kkt_val(v, o) : 0.0;
    for (int k = c.aux_c_ptr[t]; k < c.aux_c_ptr[t + 1]; ++k) acc -= Tb[c.aux_c_idx[k]];
    bt.aux[(size_t)p * c.n_aux + t] = acc;
  }
  if (t < c.nk) {
    double acc = bt.rhs[(size_t)p * c.nk + t];
    const int rp = c.rpos[t];
    if (rp >= 0)
      for (int k = c.r_c_ptr[rp]; k < c.r_c_ptr[rp + 1]; ++k) acc -= tb[c.r_c_idx[k]];
    bt.rhs2[(size_t)p * c.nk + t] = acc;
  }
}

// re-solve, before the chain: y_n = A_n^-1 r_n, t_n = G_n' r_n
__global__ void __launch_bounds__(256)
kktc_interior_fwd_kernel(const RbKktCondDev c, const RbKktCondBatch bt) {
  extern __shared__ double kktcond_smem[];
  const int n = blockIdx.x, p = blockIdx.y;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int u0 = c.iu_ptr[n], a = c.iu_ptr[n + 1] - u0;
  const int ns = c.su_ptr[n + 1] - c.su_ptr[n];
  const int32_t* __restrict__ unk = c.iunk + u0;
  const size_t slot = (size_t)p * c.NI + n;
  const double* __restrict__ M = bt.Ainv + slot * c.amax * c.ldA;
  const double* __restrict__ Gn = bt.G + slot * c.amax * c.ldG;
  const double* __restrict__ rhs = bt.rhs + (size_t)p * c.nk;
  double* __restrict__ rI = kktcond_smem;
  for (int i = tid; i < a; i += nthreads) rI[i] = rhs[unk[i]];
  __syncthreads();
  kktc_matvec_rows(M, c.ldA, a, rI, bt.Y + slot * c.amax);
  double* __restrict__ tb = bt.tbuf + slot * c.smax;
  for (int ra = tid; ra < ns; ra += nthreads) {
    double acc = 0.0;
    for (int j = 0; j < a; ++j) acc += Gn[(size_t)j * c.ldG + ra] * rI[j];
    tb[ra] = acc;
  }
}

// after the chain: x_n = y_n - G_n x_sep
__global__ void __launch_bounds__(256)
kktc_interior_back_kernel(const RbKktCondDev c, const RbKktCondBatch bt) {
  extern __shared__ double kktcond_smem[];
  const int n = blockIdx.x, p = blockIdx.y;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int u0 = c.iu_ptr[n], a = c.iu_ptr[n + 1] - u0;
  const int s0 = c.su_ptr[n], ns = c.su_ptr[n + 1] - s0;
  const size_t slot = (size_t)p * c.NI + n;
  const double* __restrict__ Gn = bt.G + slot * c.amax * c.ldG;
  const double* __restrict__ Yg = bt.Y + slot * c.amax;
  double* __restrict__ sol = bt.sol + (size_t)p * c.nk;
  double* __restrict__ xs = kktcond_smem;
  for (int i = tid; i < ns; i += nthreads) xs[i] = sol[c.sunk[s0 + i]];
  __syncthreads();
  for (int i = tid; i < a; i += nthreads) {
    const double* __restrict__ g = Gn + (size_t)i * c.ldG;
    double acc = Yg[i];
    for (int k = 0; k < ns; ++k) acc -= g[k] * xs[k];
    sol[c.iunk[u0 + i]] = acc;
  }
}
