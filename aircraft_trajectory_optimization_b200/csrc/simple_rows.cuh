// Rows of g that are affine or sums of squared affine forms of w, and the objective reduction.
//
// Reference rows covered: equal-step rows H[n2]-H[n] (base_raceline.py:891-905),
// path-length pins Z[n,0][0]-s_n (:1059-1062, :1171-1181), regularity rows kn*y-ky*n (:1105-1129),
// gate rows (:545-595, CIRCLE = sum of two squares, SQUARE / axial = affine), tube discs
// (drone3d/obstacles/mesh_obstacle.py:219-237) and the point-mass thrust sphere u'u/T_max^2
// (drone3d/dynamics/point_model.py:122-129).  Their Jacobian entries are affine in w and their
// Hessians are constants times the row multiplier, so they are data, not generated code.
#pragma once
#include "common.cuh"

// one thread per (problem, simple row): g value and Jacobian entries
__global__ void simple_rows_kernel(const RbDev d, const RbBatch b) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)b.B * d.n_srow) return;
  const int p = (int)(t / d.n_srow);
  const int r = (int)(t - (long long)p * d.n_srow);
  const double* __restrict__ w = b.x + (size_t)p * d.nw;
  const int v0 = d.srow_var_ptr[r], nv = d.srow_var_ptr[r + 1] - v0;
  const int f0 = d.srow_form_ptr[r], nf = d.srow_form_ptr[r + 1] - f0;
  const int a0 = d.srow_coef_ptr[r];
  const int kind = d.srow_kind[r];
  double scale = 1.0;
  const int sk = d.srow_scale[r];
  if (sk >= 0) {
    const double s = b.vp[(size_t)p * b.vp_stride + sk];
    scale = 1.0 / (s * s);
  }
  double val = 0.0;
  double* __restrict__ jac = b.jac ? b.jac + (size_t)p * d.nnzj : nullptr;
  if (kind == 0) {
    double L = d.srow_c[f0];
    for (int i = 0; i < nv; ++i) L += d.srow_A[a0 + i] * w[d.srow_var[v0 + i]];
    val = L;
    if (jac)
      for (int i = 0; i < nv; ++i) jac[d.srow_jslot[v0 + i]] = d.srow_A[a0 + i];
  } else {
    // two passes over at most a handful of forms: values, then d/dw = 2 scale sum_m L_m A_m
    for (int m = 0; m < nf; ++m) {
      double L = d.srow_c[f0 + m];
      for (int i = 0; i < nv; ++i) L += d.srow_A[a0 + m * nv + i] * w[d.srow_var[v0 + i]];
      val += L * L;
    }
    val *= scale;
    if (jac) {
      for (int i = 0; i < nv; ++i) {
        double acc = 0.0;
        for (int m = 0; m < nf; ++m) {
          double L = d.srow_c[f0 + m];
          for (int k = 0; k < nv; ++k) L += d.srow_A[a0 + m * nv + k] * w[d.srow_var[v0 + k]];
          acc += L * d.srow_A[a0 + m * nv + i];
        }
        jac[d.srow_jslot[v0 + i]] = 2.0 * scale * acc;
      }
    }
  }
  if (b.g) b.g[(size_t)p * d.ng + d.srow_row[r]] = val;
}

// one thread per (problem, Hessian slot touched by simple rows); runs AFTER the cell kernel
__global__ void simple_hess_kernel(const RbDev d, const RbBatch b) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)b.B * d.n_shess) return;
  const int p = (int)(t / d.n_shess);
  const int k = (int)(t - (long long)p * d.n_shess);
  const double* __restrict__ lam = b.lam_g + (size_t)p * d.ng;
  double acc = 0.0;
  for (int e = d.shess_ptr[k]; e < d.shess_ptr[k + 1]; ++e) {
    double c = d.shess_coef[e] * lam[d.shess_row[e]];
    const int sk = d.shess_scale[e];
    if (sk >= 0) {
      const double s = b.vp[(size_t)p * b.vp_stride + sk];
      c /= (s * s);
    }
    acc += c;
  }
  double* __restrict__ H = b.hess + (size_t)p * d.nnzh + d.shess_slot[k];
  if (d.shess_add[k]) *H += acc; else *H = acc;
}

// f[p] = sum_n fpart[p][n]: one warp per problem, fixed summation order (deterministic)
__global__ void objective_sum_kernel(const RbDev d, const RbBatch b, int ncell) {
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= b.B) return;
  const double* __restrict__ fp = b.fpart + (size_t)warp * ncell;
  double acc = 0.0;
  for (int i = lane; i < ncell; i += 32) acc += fp[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) b.f[warp] = acc;
}
