// Batched warm-start chain: point-mass raceline solutions -> drone initial guesses.
//
// Reference definition: drone3d/raceline/drone_raceline.py:158-277 (_guess_z / _guess_u of DroneRaceline: orientation
// from thrust and velocity of the point-mass solution, body velocity R' vg, body rate R' (T x dT) / |T|^2, rotor thrusts
// |T| / 4, quaternion sign / yaw wrap made continuous along the lap) with the point-mass helper functions
// drone3d/dynamics/point_model.py (f_T = R u, f_vg = R vb, R = I for global_r else the frame Rp) and scipy's
// Rotation.from_matrix(...).as_quat() / .as_euler('xyz') [third party] for the orientation parameters.
//
//   ws_point_kernel     one thread per (instance, collocation point): everything that is local to the point
//   ws_continuity_kernel one thread per instance: the sequential sign / wrap pass (point i depends on point i-1) and the
//                        closure information the builder needs (drone_raceline.py:82-95): |q_first - q_last| > 1, yaw wraps
#pragma once
#include "common.cuh"

struct RbWsArgs {
  int B, N, P;              // instances, intervals, points per interval (K + 1)
  int quat, closed, global_r, use_fc;
  int nw_pm, nw_dr;         // decision vector lengths (point mass: N + N P 12, drone: N + N P (nz + 8))
  const double* w_pm;       // [B][nw_pm]
  const double* fc;         // [N P][13] frame constants (Rp row-major first) or null
  double* w_dr;             // [B][nw_dr]
  int* info;                // [B][4]: flipped closure (quat), yaw wraps (euler), continuity failures, unused
};

#define RB_WS_THREADS 128

__device__ inline void ws_cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// scipy.spatial.transform.Rotation.from_matrix(R).as_quat() for an orthonormal R (row-major), scalar last
__device__ inline void ws_matrix_to_quat(const double* R, double* q) {
  const double tr = R[0] + R[4] + R[8];
  const double dec[4] = {R[0], R[4], R[8], tr};
  int c = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (dec[i] > dec[c]) c = i;
  if (c != 3) {
    const int i = c, j = (i + 1) % 3, k = (j + 1) % 3;
    q[i] = 1.0 - tr + 2.0 * R[i * 3 + i];
    q[j] = R[j * 3 + i] + R[i * 3 + j];
    q[k] = R[k * 3 + i] + R[i * 3 + k];
    q[3] = R[k * 3 + j] - R[j * 3 + k];
  } else {
    q[0] = R[7] - R[5];
    q[1] = R[2] - R[6];
    q[2] = R[3] - R[1];
    q[3] = 1.0 + tr;
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] /= n;
}

__global__ void __launch_bounds__(RB_WS_THREADS) ws_point_kernel(const RbWsArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int np = a.N * a.P;
  if (t >= (long long)a.B * np) return;
  const int p = (int)(t / np), i = (int)(t - (long long)p * np);
  const int nz = a.quat ? 13 : 12, S = nz + 8;
  const double* __restrict__ wp = a.w_pm + (size_t)p * a.nw_pm + a.N + (size_t)i * 12;
  double* __restrict__ wd = a.w_dr + (size_t)p * a.nw_dr + a.N + (size_t)i * S;
  if (i < a.N) a.w_dr[(size_t)p * a.nw_dr + i] = a.w_pm[(size_t)p * a.nw_pm + i];     // step sizes (_guess_h)
  double Rf[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};      // point-mass frame: I (global_r) or Rp
  if (!a.global_r && a.use_fc) {
#pragma unroll
    for (int e = 0; e < 9; ++e) Rf[e] = a.fc[(size_t)i * 13 + e];
  }
  double T[3], vg[3], dT[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    T[r] = Rf[r * 3] * wp[6] + Rf[r * 3 + 1] * wp[7] + Rf[r * 3 + 2] * wp[8];
    vg[r] = Rf[r * 3] * wp[3] + Rf[r * 3 + 1] * wp[4] + Rf[r * 3 + 2] * wp[5];
    dT[r] = Rf[r * 3] * wp[9] + Rf[r * 3 + 1] * wp[10] + Rf[r * 3 + 2] * wp[11];
  }
  const double Tn = sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]);
  double R[9];
  if (a.closed) {
    // orientation from thrust and velocity
    const double vn = sqrt(vg[0] * vg[0] + vg[1] * vg[1] + vg[2] * vg[2]);
    double e1[3], e2[3], e3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      e1[r] = vg[r] / vn;
      e3[r] = T[r] / Tn;
    }
    const double d = e1[0] * e3[0] + e1[1] * e3[1] + e1[2] * e3[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) e1[r] -= e3[r] * d;
    const double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
#pragma unroll
    for (int r = 0; r < 3; ++r) e1[r] /= n1;
    ws_cross(e3, e1, e2);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      R[r * 3] = e1[r];
      R[r * 3 + 1] = e2[r];
      R[r * 3 + 2] = e3[r];
    }
  } else {
    // orientation from thrust alone: Rodrigues rotation of e_z onto T / |T|
    const double tn[3] = {T[0] / Tn, T[1] / Tn, T[2] / Tn};
    const double b[3] = {0, 0, 1};
    double v[3];
    ws_cross(tn, b, v);
#pragma unroll
    for (int r = 0; r < 3; ++r) v[r] = -v[r];
    const double s2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], c = tn[2];
    const double H[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
    // exactly vertical thrust (the initial / terminal rows of an open point-mass raceline pin T_x = T_y = 0): the
    // reference's expression is 0 / 0 there; its limit (1 - cos) / sin^2 -> 1/2 gives R = I
    const double f = s2 > 0.0 ? (1.0 - c) / s2 : 0.5;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        double hh = 0.0;
#pragma unroll
        for (int m = 0; m < 3; ++m) hh += H[r * 3 + m] * H[m * 3 + cc];
        R[r * 3 + cc] = (r == cc ? 1.0 : 0.0) + H[r * 3 + cc] + hh * f;
      }
  }
  // body velocity and rate use the global R (drone_raceline.py:254-268 take them after the Rp' R step: R here is already
  // the relative one when global_r is false)
  if (!a.global_r && a.use_fc) {
    double Rl[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
        Rl[r * 3 + cc] = Rf[r] * R[cc] + Rf[3 + r] * R[3 + cc] + Rf[6 + r] * R[6 + cc];     // Rp' R
#pragma unroll
    for (int e = 0; e < 9; ++e) R[e] = Rl[e];
  }
  wd[0] = wp[0];
  wd[1] = wp[1];
  wd[2] = wp[2];
  int o = 3;
  if (a.quat) {
    double q[4];
    ws_matrix_to_quat(R, q);
#pragma unroll
    for (int r = 0; r < 4; ++r) wd[3 + r] = q[r];
    o = 7;
  } else {
    // np.flip(Rotation.from_matrix(R).as_euler('xyz')): yaw, pitch, roll of R = Ra(yaw) Rb(pitch) Rc(roll)
    wd[3] = atan2(R[3], R[0]);
    wd[4] = -asin(fmin(1.0, fmax(-1.0, R[6])));
    wd[5] = atan2(R[7], R[8]);
    o = 6;
  }
  double wg[3];
  ws_cross(T, dT, wg);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    wd[o + r] = R[r] * vg[0] + R[3 + r] * vg[1] + R[6 + r] * vg[2];                          // R' vg
    wd[o + 3 + r] = (R[r] * wg[0] + R[3 + r] * wg[1] + R[6 + r] * wg[2]) / (Tn * Tn);       // R' (T x dT) / |T|^2
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    wd[nz + j] = Tn / 4.0;
    wd[nz + 4 + j] = 0.0;
  }
}

__global__ void ws_continuity_kernel(const RbWsArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.B) return;
  const int nz = a.quat ? 13 : 12, S = nz + 8, np = a.N * a.P;
  double* __restrict__ w = a.w_dr + (size_t)p * a.nw_dr + a.N;
  int fails = 0;
  double last[4], first[4];
  const int nr = a.quat ? 4 : 3;
  for (int i = 0; i < np; ++i) {
    double* r = w + (size_t)i * S + 3;
    if (i > 0) {
      double d2 = 0.0;
      for (int c = 0; c < nr; ++c) d2 += (r[c] - last[c]) * (r[c] - last[c]);
      if (a.quat) {
        if (sqrt(d2) >= 1.0)
          for (int c = 0; c < 4; ++c) r[c] = -r[c];
      } else if (sqrt(d2) > 1.0) {
        const double kPi = 3.14159265358979323846;
        if (r[0] - last[0] > kPi) r[0] -= 2 * kPi;
        else if (r[0] - last[0] <= -kPi) r[0] += 2 * kPi;
        d2 = 0.0;
        for (int c = 0; c < 3; ++c) d2 += (r[c] - last[c]) * (r[c] - last[c]);
        if (sqrt(d2) > 1.0) fails++;
      }
    }
    for (int c = 0; c < nr; ++c) {
      last[c] = r[c];
      if (i == 0) first[c] = r[c];
    }
  }
  if (a.info) {
    double d2 = 0.0;
    for (int c = 0; c < nr; ++c) d2 += (first[c] - last[c]) * (first[c] - last[c]);
    a.info[4 * p] = (a.quat && sqrt(d2) > 1.0) ? 1 : 0;
    a.info[4 * p + 1] = a.quat ? 0 : (int)rint((last[0] - first[0]) / 2.0 / 3.14159265358979323846);
    a.info[4 * p + 2] = fails;
    a.info[4 * p + 3] = 0;
  }
}
