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

constexpr int RB_TAIL_THREADS = 256;

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

// dynamic shared memory: n_slots doubles (work slots) followed by n_levels + 1 ints (level table)
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
  int* lp = reinterpret_cast<int*>(tail_W + t.n_slots);
  for (int l = tid; l <= n_levels; l += RB_TAIL_THREADS) lp[l] = t.lvl_ptr[l];
  __syncthreads();
  // the instruction stream is independent of the data: every thread keeps its next instruction in flight (the next
  // one of this level, or its first one of the next level) while it executes the current one
  int hi = lp[1];
  int i = lp[0] + tid;
  bool has = i < hi;
  int4 cur = make_int4(-1, 0, 0, 0);
  if (has) cur = __ldg(t.ins + i);
  for (int l = 0; l < n_levels; ++l) {
    const bool more = l + 1 < n_levels;
    const int nhi = more ? lp[l + 2] : hi;
    const int pi = hi + tid;
    const bool pf_has = more && pi < nhi;
    int4 pf = make_int4(-1, 0, 0, 0);
    bool pf_done = false;
    while (has) {
      const int ni = i + RB_TAIL_THREADS;
      const bool nh = ni < hi;
      int4 nxt = make_int4(-1, 0, 0, 0);
      if (nh) {
        nxt = __ldg(t.ins + ni);
      } else {
        if (pf_has) pf = __ldg(t.ins + pi);
        pf_done = true;
      }
      tail_exec(cur, W, t, x, vp, lam, g, jac, hess);
      cur = nxt;
      i = ni;
      has = nh;
    }
    if (!pf_done && pf_has) pf = __ldg(t.ins + pi);
    __syncthreads();
    cur = pf;
    has = pf_has;
    i = pi;
    hi = nhi;
  }
}
