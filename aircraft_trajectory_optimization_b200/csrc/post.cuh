// Data-parallel steps on either side of the solve.
//
//   centerline_frames_kernel  the per-point frame constants the evaluation kernels consume, for many path lengths and
//                             many tracks at once.  Reference: drone3d/centerlines/spline_centerline.py:232-322 (the 15
//                             `param_terms` xc, xc', xc'', ry, ry' of the two cubic splines, evaluated the way
//                             drone3d/utils/interp.py:55-84 evaluates them: right-continuous piece lookup, linear
//                             beyond the last knot) and :279-294 (es, ey, en, ks, ky, kn, |xc'|); optionally the global
//                             position xc + ey y + en n of parametric states (drone3d/dynamics/drone_models.py:306-328).
//   traj_interp_kernel        trajectory interpolants of B solutions at M query times each.  Reference:
//                             drone3d/utils/discretization_utils.py:53-137 (collocation: the degree-K Lagrange polynomial
//                             of the interval that holds t, its end value sum_k D_k X[N-1,k] after the last interval;
//                             shooting: pw_lin over the interval start states), used by _unpack_soln
//                             (drone3d/raceline/base_raceline.py:801-864).
#pragma once
#include "common.cuh"

#define RB_POST_THREADS 128

struct RbSplineTab {
  int nkx, nkr;            // knots of the xc / ry splines
  const double* kx;        // [nkx]
  const double* cx;        // [4][nkx-1][3] scipy CubicSpline.c (highest power first)
  const double* ex;        // [2][3] value and slope at the last knot
  const double* kr;        // [nkr]
  const double* cr;        // [4][nkr-1][3]
  const double* er;        // [2][3]
};

// value and derivatives (nd = 1: up to first, 2: up to second) of a piecewise cubic with linear extrapolation
__device__ inline void post_pwc(const double* __restrict__ k, const double* __restrict__ c, const double* __restrict__ e,
                                int nk, double s, int nd, double* v0, double* v1, double* v2) {
  // idx = searchsorted(k, s, 'right') - 1
  int lo = 0, hi = nk;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (k[mid] <= s) lo = mid + 1; else hi = mid;
  }
  const int idx = lo - 1;
  const bool above = idx >= nk - 1, below = idx < 0;
  const int in = idx < 0 ? 0 : (idx > nk - 2 ? nk - 2 : idx);
  const double rel = above ? s - k[nk - 1] : s - k[in];
  const int ns = nk - 1;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double c3 = c[(0 * ns + in) * 3 + d], c2 = c[(1 * ns + in) * 3 + d], c1 = c[(2 * ns + in) * 3 + d],
           c0 = c[(3 * ns + in) * 3 + d];
    if (above) {
      c0 = e[d];
      c1 = e[3 + d];
    }
    if (above || below) {
      c2 = 0.0;
      c3 = 0.0;
    }
    v0[d] = c0 + rel * (c1 + rel * (c2 + rel * c3));
    v1[d] = c1 + rel * (2 * c2 + 3 * rel * c3);
    if (nd > 1) v2[d] = 2 * c2 + 6 * rel * c3;
  }
}

// one thread per (track, query): fc [T][M][13], optional xc [T][M][3]; optional yn [T][M][2] -> xg [T][M][3]
__global__ void __launch_bounds__(RB_POST_THREADS)
centerline_frames_kernel(const RbSplineTab* __restrict__ tabs, int T, const double* __restrict__ s, int M, int s_stride,
                         double* __restrict__ fc, double* __restrict__ xc_out, const double* __restrict__ yn,
                         double* __restrict__ xg) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)T * M) return;
  const int tr = (int)(t / M), i = (int)(t - (long long)tr * M);
  const RbSplineTab tb = tabs[tr];
  const double sv = s[(size_t)tr * s_stride + i];
  double xc[3], xcs[3], xcss[3], ry[3], rys[3], dummy[3];
  post_pwc(tb.kx, tb.cx, tb.ex, tb.nkx, sv, 2, xc, xcs, xcss);
  post_pwc(tb.kr, tb.cr, tb.er, tb.nkr, sv, 1, ry, rys, dummy);
  const double mag = sqrt(xcs[0] * xcs[0] + xcs[1] * xcs[1] + xcs[2] * xcs[2]);
  double es[3], ey[3], en[3];
  double d = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    es[c] = xcs[c] / mag;
    d += es[c] * ry[c];
  }
  double n2 = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ey[c] = ry[c] - es[c] * d;
    n2 += ey[c] * ey[c];
  }
  const double ni = sqrt(n2);
#pragma unroll
  for (int c = 0; c < 3; ++c) ey[c] /= ni;
  en[0] = es[1] * ey[2] - es[2] * ey[1];
  en[1] = es[2] * ey[0] - es[0] * ey[2];
  en[2] = es[0] * ey[1] - es[1] * ey[0];
  double m00 = 0, m01 = 0, m10 = 0, m11 = 0, r0 = 0, r1 = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    m00 += xcs[c] * es[c];
    m01 += xcs[c] * ey[c];
    m10 += ry[c] * es[c];
    m11 += ry[c] * ey[c];
    r0 += xcss[c] * en[c];
    r1 += rys[c] * en[c];
  }
  const double det = m00 * m11 - m01 * m10;
  const double a = (m11 * r0 - m01 * r1) / det / mag;
  const double b = (-m10 * r0 + m00 * r1) / det / mag;
  // kn = -(xcss x xcs) . en / |xcs|^3
  const double cx0 = xcss[1] * xcs[2] - xcss[2] * xcs[1], cx1 = xcss[2] * xcs[0] - xcss[0] * xcs[2],
               cx2 = xcss[0] * xcs[1] - xcss[1] * xcs[0];
  const double kn = -(cx0 * en[0] + cx1 * en[1] + cx2 * en[2]) / (mag * mag * mag);
  double* __restrict__ o = fc + ((size_t)tr * M + i) * 13;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    o[r * 3] = es[r];
    o[r * 3 + 1] = ey[r];
    o[r * 3 + 2] = en[r];
  }
  o[9] = b;       // ks
  o[10] = -a;     // ky
  o[11] = kn;
  o[12] = mag;
  if (xc_out) {
#pragma unroll
    for (int c = 0; c < 3; ++c) xc_out[((size_t)tr * M + i) * 3 + c] = xc[c];
  }
  if (yn && xg) {
    const double y = yn[((size_t)tr * M + i) * 2], n = yn[((size_t)tr * M + i) * 2 + 1];
#pragma unroll
    for (int c = 0; c < 3; ++c) xg[((size_t)tr * M + i) * 3 + c] = xc[c] + ey[c] * y + en[c] * n;
  }
}

// cumulative interval start times tp [B][N+1] of B solutions (one thread per instance: N <= a few hundred)
__global__ void traj_times_kernel(const double* __restrict__ w, int B, int N, int nw, double* __restrict__ tp) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B) return;
  double acc = 0.0;
  tp[(size_t)p * (N + 1)] = 0.0;
  for (int n = 0; n < N; ++n) {
    acc += w[(size_t)p * nw + n];
    tp[(size_t)p * (N + 1) + n + 1] = acc;
  }
}

// one thread per (instance, query): out [B][M][S] = (z, u, du)(t).  P = K + 1 points per interval, tau / D the
// collocation nodes and end weights (ignored for P == 1: linear interpolation between interval start states)
__global__ void __launch_bounds__(RB_POST_THREADS)
traj_interp_kernel(const double* __restrict__ w, int B, int N, int P, int S, int nw, const double* __restrict__ tp,
                   const double* __restrict__ tau, const double* __restrict__ D, const double* __restrict__ tq, int M,
                   int tq_stride, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * M) return;
  const int p = (int)(t / M), i = (int)(t - (long long)p * M);
  const double tv = tq[(size_t)p * tq_stride + i];
  const double* __restrict__ T = tp + (size_t)p * (N + 1);
  const double* __restrict__ X = w + (size_t)p * nw + N;        // [N][P][S]
  double* __restrict__ o = out + ((size_t)p * M + i) * S;
  if (P == 1) {
    // pw_lin over the knots tp[0..N-1] (interval start times), linear extrapolation beyond the ends
    if (N == 1) {
      for (int c = 0; c < S; ++c) o[c] = X[c];
      return;
    }
    int lo = 0, hi = N;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (T[mid] <= tv) lo = mid + 1; else hi = mid;
    }
    int k = lo - 1;
    k = k < 0 ? 0 : (k > N - 2 ? N - 2 : k);
    const double a = (tv - T[k]) / (T[k + 1] - T[k]);
    for (int c = 0; c < S; ++c) {
      const double v0 = X[(size_t)k * S + c], v1 = X[(size_t)(k + 1) * S + c];
      o[c] = v0 + a * (v1 - v0);
    }
    return;
  }
  if (tv < T[0]) {
    for (int c = 0; c < S; ++c) o[c] = X[c];
    return;
  }
  if (tv >= T[N]) {
    for (int c = 0; c < S; ++c) {
      double acc = 0.0;
      for (int j = 0; j < P; ++j) acc += D[j] * X[((size_t)(N - 1) * P + j) * S + c];
      o[c] = acc;
    }
    return;
  }
  int lo = 0, hi = N + 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (T[mid] <= tv) lo = mid + 1; else hi = mid;
  }
  const int n = lo - 1;
  const double rel = (tv - T[n]) / (T[n + 1] - T[n]);
  double lw[16];
  for (int j = 0; j < P; ++j) {
    double v = 1.0;
    for (int r = 0; r < P; ++r)
      if (r != j) v *= (rel - tau[r]) / (tau[j] - tau[r]);
    lw[j] = v;
  }
  for (int c = 0; c < S; ++c) {
    double acc = 0.0;
    for (int j = 0; j < P; ++j) acc += lw[j] * X[((size_t)n * P + j) * S + c];
    o[c] = acc;
  }
}
