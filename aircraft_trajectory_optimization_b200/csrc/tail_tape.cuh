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

__global__ void __launch_bounds__(RB_TAIL_THREADS) tail_tape_kernel(const RbTail t, const RbDev d, const RbBatch b,
                                                                     const int n_levels) {
  extern __shared__ double tail_W[];
  const int p = blockIdx.x;
  const double* __restrict__ x = b.x + (size_t)p * d.nw;
  const double* __restrict__ vp = b.vp + (size_t)p * b.vp_stride;
  const double* __restrict__ lam = b.lam_g ? b.lam_g + (size_t)p * d.ng : nullptr;
  double* __restrict__ g = b.g ? b.g + (size_t)p * d.ng : nullptr;
  double* __restrict__ jac = b.jac ? b.jac + (size_t)p * d.nnzj : nullptr;
  double* __restrict__ hess = b.hess ? b.hess + (size_t)p * d.nnzh : nullptr;
  double* W = tail_W;
  int lo = t.lvl_ptr[0];
  for (int l = 0; l < n_levels; ++l) {
    const int hi = t.lvl_ptr[l + 1];
    for (int i = lo + threadIdx.x; i < hi; i += RB_TAIL_THREADS) {
      const int4 q = __ldg(t.ins + i);
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
    lo = hi;
    __syncthreads();
  }
}
