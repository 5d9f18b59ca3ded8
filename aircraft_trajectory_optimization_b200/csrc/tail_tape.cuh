// End rows of OPEN racelines: initial / terminal constraints and the gate on the end state
// (reference drone3d/raceline/base_raceline.py:516-543, :914-918, drone_raceline.py:110-148,
// point_raceline.py:15-45).  These rows are compositions phi(F(z, u, h)) of the last interval's end state; the host
// (aircraft_trajectory_optimization_b200/tail.py) differentiates their scalar expression graph at problem
// construction and uploads a levelised, register-allocated tape:
//
//   ins[i] = (op, a, b, dst)     work slots a, b -> dst;  loads: a = index into x / vp / lam_g / the constant table;
//                                stores: a = work slot, b = g row / CCS position in jac_g / hess_l
//   lvl_ptr[l] .. lvl_ptr[l+1]   the instructions of level l: independent of each other
//
// One CTA per problem instance, the work slots in shared memory, the threads of the CTA take the instructions of a
// level side by side, one barrier per level.  The levels come in three phases (g; + Jacobian; + Hessian) so an
// evaluation that wants g only stops after the first.  Runs AFTER the cell kernels and simple_hess_kernel: Hessian
// positions those already wrote are accumulated into (op T_ADD_H), the others assigned.
#pragma once
#include "common.cuh"

enum : int {
  T_CONST = 0, T_LOADX, T_LOADVP, T_LOADLAM, T_ADD, T_SUB, T_MUL, T_DIV, T_NEG, T_SQ, T_SQRT, T_SIN, T_COS, T_TAN,
  T_STORE_G, T_STORE_J, T_STORE_H, T_ADD_H
};

struct RbTail {
  const int4* ins;
  const int32_t* lvl_ptr;
  const double* cval;
  int n_slots;
  int n_levels[3];   // levels up to and including the g / Jacobian / Hessian phase
};

constexpr int RB_TAIL_THREADS = 512;

__device__ __forceinline__ void tail_exec(const int4 q, double* __restrict__ W, const RbTail& t, const double* __restrict__ x,
                                          const double* __restrict__ vp, const double* __restrict__ lam,
                                          double* __restrict__ g, double* __restrict__ jac, double* __restrict__ hess) {
  switch (q.x) {
    case T_CONST: W[q.w] = t.cval[q.y]; break;
    case T_LOADX: W[q.w] = x[q.y]; break;
    case T_LOADVP: W[q.w] = vp[q.y]; break;
    case T_LOADLAM: W[q.w] = lam ? lam[q.y] : 0.0; break;
    case T_ADD: W[q.w] = W[q.y] + W[q.z]; break;
    case T_SUB: W[q.w] = W[q.y] - W[q.z]; break;
    case T_MUL: W[q.w] = W[q.y] * W[q.z]; break;
    case T_DIV: W[q.w] = W[q.y] / W[q.z]; break;
    case T_NEG: W[q.w] = -W[q.y]; break;
    case T_SQ: { const double v = W[q.y]; W[q.w] = v * v; } break;
    case T_SQRT: W[q.w] = sqrt(W[q.y]); break;
    case T_SIN: W[q.w] = sin(W[q.y]); break;
    case T_COS: W[q.w] = cos(W[q.y]); break;
    case T_TAN: W[q.w] = tan(W[q.y]); break;
    case T_STORE_G: if (g) g[q.z] = W[q.y]; break;
    case T_STORE_J: if (jac) jac[q.z] = W[q.y]; break;
    case T_STORE_H: if (hess) hess[q.z] = W[q.y]; break;
    case T_ADD_H: if (hess) hess[q.z] += W[q.y]; break;
    default: break;
  }
}

constexpr int RB_TAIL_CHUNK = 1024;   // instructions staged per cp.async group (16 KB), double-buffered

__host__ __device__ inline size_t tail_smem_bytes(int n_slots, int n_levels) {
  return (size_t)n_slots * sizeof(double) + 2 * (size_t)RB_TAIL_CHUNK * sizeof(int4) + ((size_t)n_levels + 2) * sizeof(int);
}

__device__ __forceinline__ void tail_stage(int4* dst, const int4* __restrict__ src, int n, int tid) {
  for (int k = tid; k < n; k += RB_TAIL_THREADS)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + k)), "l"(src + k) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// dynamic shared memory: n_slots doubles (work slots), two instruction chunks, n_levels + 1 ints (level table).
// The instruction stream does not depend on the data, so it is staged through shared memory one chunk ahead of the
// interpreter (cp.async): fetching every instruction from L2 when it is needed costs ~700 cycles per instruction and
// thread (measured: 285 us for the 59 k-instruction tape of the RK4 quaternion drone; one-ahead register prefetch 252 us).
__global__ void __launch_bounds__(RB_TAIL_THREADS) tail_tape_kernel(const RbTail t, const RbDev d, const RbBatch b,
                                                                     const int n_levels) {
  extern __shared__ double tail_W[];
  const int p = blockIdx.x, tid = threadIdx.x;
  const double* __restrict__ x = b.x + (size_t)p * d.nw;
  const double* __restrict__ vp = b.vp + (size_t)p * b.vp_stride;
  const double* __restrict__ lam = b.lam_g ? b.lam_g + (size_t)p * d.ng : nullptr;
  double* __restrict__ g = b.g ? b.g + (size_t)p * d.ng : nullptr;
  double* __restrict__ jac = b.jac ? b.jac + (size_t)p * d.nnzj : nullptr;
  double* __restrict__ hess = b.hess ? b.hess + (size_t)p * d.nnzh : nullptr;
  double* W = tail_W;
  int4* ring = reinterpret_cast<int4*>(tail_W + t.n_slots + (t.n_slots & 1));
  int* lp = reinterpret_cast<int*>(ring + 2 * RB_TAIL_CHUNK);
  for (int l = tid; l <= n_levels; l += RB_TAIL_THREADS) lp[l] = t.lvl_ptr[l];
  __syncthreads();
  const int n_ins = lp[n_levels];
  const int n_chunks = (n_ins + RB_TAIL_CHUNK - 1) / RB_TAIL_CHUNK;
  auto chunk_len = [&](int c) { const int r = n_ins - c * RB_TAIL_CHUNK; return r < RB_TAIL_CHUNK ? r : RB_TAIL_CHUNK; };
  if (n_chunks > 0) tail_stage(ring, t.ins, chunk_len(0), tid);
  int l = 0;
  for (int c = 0; c < n_chunks; ++c) {
    const int a = c * RB_TAIL_CHUNK, e = a + chunk_len(c);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();            // chunk c has landed for everybody; everybody is done with chunk c - 1
    if (c + 1 < n_chunks) tail_stage(ring + ((c + 1) & 1) * RB_TAIL_CHUNK, t.ins + (size_t)(c + 1) * RB_TAIL_CHUNK, chunk_len(c + 1), tid);
    const int4* __restrict__ cur = ring + (c & 1) * RB_TAIL_CHUNK - a;
    while (l < n_levels && lp[l] < e) {
      const int lo = lp[l] > a ? lp[l] : a;
      const int hi = lp[l + 1] < e ? lp[l + 1] : e;
      for (int i = lo + tid; i < hi; i += RB_TAIL_THREADS) tail_exec(cur[i], W, t, x, vp, lam, g, jac, hess);
      if (lp[l + 1] > e) break;  // the level continues in the next chunk (its instructions are independent)
      __syncthreads();
      ++l;
    }
  }
}
