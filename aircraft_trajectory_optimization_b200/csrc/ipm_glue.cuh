// Fused element-wise / reduction kernels of the batched interior-point sweep (SURVEY.md s2.1 "K3").
//
// What they replace: the per-iteration vector work IPOPT does around its linear solves inside
// `self.solver(x0=..., ...)` (drone3d/raceline/base_raceline.py:160-165) -- optimality error, barrier gradients,
// condensed KKT right-hand side, fraction-to-the-boundary step lengths, trial points and their filter quantities,
// the primal-dual update with the kappa_sigma multiplier safeguard (Waechter & Biegler, Math. Prog. 106, 2006; the
// formulas are those of aircraft_trajectory_optimization_b200/ipm.py, which keeps the same arithmetic in torch for
// the CPU tests).  One CTA per problem instance; every kernel reads the state vectors of its instance once and reduces
// in shared memory -- HBM-bound, ~10 passes over the state per sweep instead of ~300.
//
// Layout: state vectors are [B][n] (variables) or [B][m] (constraint rows), row-major.  Bounds and flags may be shared
// by all instances (row stride 0) or per instance (row stride n / m).
//   xflag bit0: lower bound, bit1: upper bound.   sflag bit0: lower, bit1: upper (inequality rows only), bit2: equality row.
#pragma once
#include "common.cuh"

#define RB_IPM_THREADS 256

struct RbIpm {
  int B, n, m;
  long long sx, ss;              // row strides of the bound / flag arrays over variables / rows (0: shared)
  const double *xL, *xU, *sL, *sU, *ceq;
  const unsigned char *xflag, *sflag;
  double *x, *s, *y, *zL, *zU, *vL, *vU;          // state
  const double *grad_f, *g, *jty;                 // evaluation at x; J' y
  const double *mu, *delta_w, *delta_c;           // [B]
  const unsigned char* resto;                     // [B] in restoration (may be null)
  const double *x_R, *DR2;                        // [B][n] restoration reference / proximity weights (may be null)
  double kappa_d, rho;
};

__device__ __forceinline__ double blk_reduce(double v, int op, double* red) {
  // op 0: sum, 1: max, 2: min; result broadcast to all threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == 0 ? v + w : (op == 1 ? fmax(v, w) : fmin(v, w));
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double r = lane < (blockDim.x >> 5) ? red[lane] : (op == 0 ? 0.0 : (op == 1 ? -1e308 : 1e308));
    if (lane >= (blockDim.x >> 5)) r = op == 0 ? 0.0 : (op == 1 ? -INFINITY : INFINITY);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double w = __shfl_xor_sync(0xffffffffu, r, o);
      r = op == 0 ? r + w : (op == 1 ? fmax(r, w) : fmin(r, w));
    }
    if (lane == 0) red[32] = r;
  }
  __syncthreads();
  return red[32];
}

// ---- 1. optimality-error ingredients ---------------------------------------------------------------------------
// out[b][8] = { |grad_x L|_inf and |grad_s L|_inf (max), |c|_inf, sum z (bound multipliers with a bound), |y|_1,
//               min and max complementarity product over all bounds (+inf / -inf if none), |c|_1, unused }
__global__ void __launch_bounds__(RB_IPM_THREADS) ipm_error_kernel(const RbIpm a, double* __restrict__ out) {
  __shared__ double red[33];
  const int b = blockIdx.x;
  const size_t on = (size_t)b * a.n, om = (size_t)b * a.m;
  const size_t bx = (size_t)b * a.sx, bs = (size_t)b * a.ss;
  double dual = 0.0, zsum = 0.0, pmin = INFINITY, pmax = -INFINITY;
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    const unsigned f = a.xflag[bx + i];
    const double zl = a.zL[on + i], zu = a.zU[on + i], x = a.x[on + i];
    dual = fmax(dual, fabs(a.grad_f[on + i] + a.jty[on + i] - zl + zu));
    if (f & 1u) {
      const double p = (x - a.xL[bx + i]) * zl;
      zsum += zl;
      pmin = fmin(pmin, p);
      pmax = fmax(pmax, p);
    }
    if (f & 2u) {
      const double p = (a.xU[bx + i] - x) * zu;
      zsum += zu;
      pmin = fmin(pmin, p);
      pmax = fmax(pmax, p);
    }
  }
  double prim = 0.0, ysum = 0.0, th1 = 0.0;
  for (int j = threadIdx.x; j < a.m; j += blockDim.x) {
    const unsigned f = a.sflag[bs + j];
    const double y = a.y[om + j], s = a.s[om + j], vl = a.vL[om + j], vu = a.vU[om + j];
    const bool eq = f & 4u;
    const double c = eq ? a.g[om + j] - a.ceq[bs + j] : a.g[om + j] - s;
    prim = fmax(prim, fabs(c));
    th1 += fabs(c);
    ysum += fabs(y);
    if (!eq) dual = fmax(dual, fabs(-y - vl + vu));
    if (f & 1u) {
      const double p = (s - a.sL[bs + j]) * vl;
      zsum += vl;
      pmin = fmin(pmin, p);
      pmax = fmax(pmax, p);
    }
    if (f & 2u) {
      const double p = (a.sU[bs + j] - s) * vu;
      zsum += vu;
      pmin = fmin(pmin, p);
      pmax = fmax(pmax, p);
    }
  }
  dual = blk_reduce(dual, 1, red);
  prim = blk_reduce(prim, 1, red);
  zsum = blk_reduce(zsum, 0, red);
  ysum = blk_reduce(ysum, 0, red);
  pmin = blk_reduce(pmin, 2, red);
  pmax = blk_reduce(pmax, 1, red);
  th1 = blk_reduce(th1, 0, red);
  if (threadIdx.x == 0) {
    double* o = out + (size_t)b * 8;
    o[0] = dual; o[1] = prim; o[2] = zsum; o[3] = ysum; o[4] = pmin; o[5] = pmax; o[6] = th1; o[7] = 0.0;
  }
}

// ---- 2. condensed Newton system ------------------------------------------------------------------------------------
// dxd [B][n], negd [B][m], rhs [B][n+m]; saved for the step kernels: gphi_x [B][n], gphi_s, c, r_s (restoration: the
// barrier gradient of the slacks), Ssr (S_s + delta_w; restoration: S_s^R) [B][m];
// sc [B][4] = { theta = |c|_1, phi (barrier objective, needs f), phi_R (restoration merit), unused }
__global__ void __launch_bounds__(RB_IPM_THREADS)
ipm_newton_kernel(const RbIpm a, const double* __restrict__ f, double* __restrict__ dxd, double* __restrict__ negd,
                  double* __restrict__ rhs, double* __restrict__ gphi_x, double* __restrict__ gphi_s, double* __restrict__ cc,
                  double* __restrict__ r_s, double* __restrict__ Ssr, double* __restrict__ sc) {
  __shared__ double red[33];
  const int b = blockIdx.x;
  const size_t on = (size_t)b * a.n, om = (size_t)b * a.m, ok = (size_t)b * (a.n + a.m);
  const size_t bx = (size_t)b * a.sx, bs = (size_t)b * a.ss;
  const double mu = a.mu[b], dw = a.delta_w[b], dc = a.delta_c[b], kd = a.kappa_d;
  const bool R = a.resto && a.resto[b];
  const double zeta = sqrt(mu);
  double lg = 0.0, dmp = 0.0, prox = 0.0;
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    const unsigned fl = a.xflag[bx + i];
    const bool hL = fl & 1u, hU = fl & 2u;
    const double x = a.x[on + i];
    const double dL = hL ? x - a.xL[bx + i] : 1.0, dU = hU ? a.xU[bx + i] - x : 1.0;
    const double iL = hL ? 1.0 / dL : 0.0, iU = hU ? 1.0 / dU : 0.0;
    const double damp = (hL && !hU ? 1.0 : 0.0) - (hU && !hL ? 1.0 : 0.0);
    const double gb = -mu * iL + mu * iU + kd * mu * damp;
    lg += (hL ? log(dL) : 0.0) + (hU ? log(dU) : 0.0);
    dmp += (hL && !hU ? dL : 0.0) + (hU && !hL ? dU : 0.0);
    if (!R) {
      const double gp = a.grad_f[on + i] + gb;
      gphi_x[on + i] = gp;
      dxd[on + i] = a.zL[on + i] * iL + a.zU[on + i] * iU + dw;
      rhs[ok + i] = -(gp + a.jty[on + i]);
    } else {
      const double dr = a.DR2[on + i], xr = x - a.x_R[on + i];
      gphi_x[on + i] = gb;                                     // barrier part of the merit gradient
      dxd[on + i] = zeta * dr + mu * (iL * iL + iU * iU);
      rhs[ok + i] = -(zeta * dr * xr + gb);
      prox += dr * xr * xr;
    }
  }
  double th1 = 0.0, c2 = 0.0;
  for (int j = threadIdx.x; j < a.m; j += blockDim.x) {
    const unsigned fl = a.sflag[bs + j];
    const bool hL = fl & 1u, hU = fl & 2u, eq = fl & 4u;
    const double s = a.s[om + j];
    const double eL = hL ? s - a.sL[bs + j] : 1.0, eU = hU ? a.sU[bs + j] - s : 1.0;
    const double jL = hL ? 1.0 / eL : 0.0, jU = hU ? 1.0 / eU : 0.0;
    const double damp = (hL && !hU ? 1.0 : 0.0) - (hU && !hL ? 1.0 : 0.0);
    const double gbs = eq ? 0.0 : -mu * jL + mu * jU + kd * mu * damp;
    const double c = eq ? a.g[om + j] - a.ceq[bs + j] : a.g[om + j] - s;
    th1 += fabs(c);
    c2 += c * c;
    if (!eq) {
      lg += (hL ? log(eL) : 0.0) + (hU ? log(eU) : 0.0);
      dmp += (hL && !hU ? eL : 0.0) + (hU && !hL ? eU : 0.0);
    }
    cc[om + j] = c;
    gphi_s[om + j] = gbs;
    if (!R) {
      const double rs = eq ? 0.0 : gbs - a.y[om + j];
      const double S = eq ? 1.0 : a.vL[om + j] * jL + a.vU[om + j] * jU + dw;
      r_s[om + j] = rs;
      Ssr[om + j] = S;
      negd[om + j] = (eq ? 0.0 : -1.0 / S) - dc;
      rhs[ok + a.n + j] = eq ? -c : -c - rs / S;
    } else {
      const double S = eq ? 1.0 : mu * (jL * jL + jU * jU);
      r_s[om + j] = gbs;
      Ssr[om + j] = S;
      negd[om + j] = eq ? -1.0 / a.rho : -(1.0 / a.rho + 1.0 / S);
      rhs[ok + a.n + j] = eq ? -c : -c - gbs / S;
    }
  }
  lg = blk_reduce(lg, 0, red);
  dmp = blk_reduce(dmp, 0, red);
  th1 = blk_reduce(th1, 0, red);
  c2 = blk_reduce(c2, 0, red);
  prox = blk_reduce(prox, 0, red);
  if (threadIdx.x == 0) {
    const double bar = -mu * lg + kd * mu * dmp;
    sc[(size_t)b * 4 + 0] = th1;
    sc[(size_t)b * 4 + 1] = f[b] + bar;
    sc[(size_t)b * 4 + 2] = 0.5 * a.rho * c2 + 0.5 * zeta * prox + bar;
    sc[(size_t)b * 4 + 3] = 0.0;
  }
}

// ---- 3. search direction, fraction-to-the-boundary step lengths, slopes ---------------------------------------------
// sol [B][n+m] from the KKT solve; moved [B] (0: the instance does not step); tau [B].
// Writes dx [B][n], dy, ds [B][m], dzL, dzU [B][n], dvL, dvU [B][m];
// sc [B][4] = { alpha_pr_max, alpha_du_max, dphi (filter method), dphi_R (restoration merit) }
__global__ void __launch_bounds__(RB_IPM_THREADS)
ipm_direction_kernel(const RbIpm a, const double* __restrict__ sol, const unsigned char* __restrict__ moved,
                     const double* __restrict__ tau, const double* __restrict__ gphi_x, const double* __restrict__ gphi_s,
                     const double* __restrict__ cc, const double* __restrict__ r_s, const double* __restrict__ Ssr,
                     double* __restrict__ dx, double* __restrict__ dy, double* __restrict__ ds, double* __restrict__ dzL,
                     double* __restrict__ dzU, double* __restrict__ dvL, double* __restrict__ dvU, double* __restrict__ sc) {
  __shared__ double red[33];
  const int b = blockIdx.x;
  const size_t on = (size_t)b * a.n, om = (size_t)b * a.m, ok = (size_t)b * (a.n + a.m);
  const size_t bx = (size_t)b * a.sx, bs = (size_t)b * a.ss;
  const double mu = a.mu[b], t = tau[b];
  const bool mv = moved[b] != 0;
  const bool R = a.resto && a.resto[b];
  const double zeta = sqrt(mu);
  double apr = 1.0, adu = 1.0, dphi = 0.0, dphiR = 0.0;
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    const unsigned fl = a.xflag[bx + i];
    const bool hL = fl & 1u, hU = fl & 2u;
    const double x = a.x[on + i];
    const double dL = hL ? x - a.xL[bx + i] : 1.0, dU = hU ? a.xU[bx + i] - x : 1.0;
    const double iL = hL ? 1.0 / dL : 0.0, iU = hU ? 1.0 / dU : 0.0;
    const double d = mv ? sol[ok + i] : 0.0;
    const double zl = a.zL[on + i], zu = a.zU[on + i];
    double zdl = hL ? mu * iL - zl - zl * iL * d : 0.0, zdu = hU ? mu * iU - zu + zu * iU * d : 0.0;
    if (R) zdl = zdu = 0.0;
    dx[on + i] = d;
    dzL[on + i] = zdl;
    dzU[on + i] = zdu;
    if (hL && d < 0.0) apr = fmin(apr, -t * dL / d);
    if (hU && d > 0.0) apr = fmin(apr, t * dU / d);
    if (hL && zdl < 0.0) adu = fmin(adu, -t * zl / zdl);
    if (hU && zdu < 0.0) adu = fmin(adu, -t * zu / zdu);
    dphi += gphi_x[on + i] * d;
    if (R) dphiR += (zeta * a.DR2[on + i] * (x - a.x_R[on + i]) + gphi_x[on + i]) * d;
  }
  for (int j = threadIdx.x; j < a.m; j += blockDim.x) {
    const unsigned fl = a.sflag[bs + j];
    const bool hL = fl & 1u, hU = fl & 2u, eq = fl & 4u;
    const double s = a.s[om + j];
    const double eL = hL ? s - a.sL[bs + j] : 1.0, eU = hU ? a.sU[bs + j] - s : 1.0;
    const double jL = hL ? 1.0 / eL : 0.0, jU = hU ? 1.0 / eU : 0.0;
    const double w = mv ? sol[ok + a.n + j] : 0.0;
    const double S = Ssr[om + j], rs = r_s[om + j];
    const double d_s = (mv && !eq) ? (w - rs) / S : 0.0;        // filter method: rs = r_s; restoration: rs = barrier gradient
    const double vl = a.vL[om + j], vu = a.vU[om + j];
    double vdl = hL ? mu * jL - vl - vl * jL * d_s : 0.0, vdu = hU ? mu * jU - vu + vu * jU * d_s : 0.0;
    if (R) vdl = vdu = 0.0;
    dy[om + j] = R ? 0.0 : w;
    ds[om + j] = d_s;
    dvL[om + j] = vdl;
    dvU[om + j] = vdu;
    if (hL && d_s < 0.0) apr = fmin(apr, -t * eL / d_s);
    if (hU && d_s > 0.0) apr = fmin(apr, t * eU / d_s);
    if (hL && vdl < 0.0) adu = fmin(adu, -t * vl / vdl);
    if (hU && vdu < 0.0) adu = fmin(adu, -t * vu / vdu);
    dphi += gphi_s[om + j] * d_s;
    if (R) dphiR += cc[om + j] * (w - a.rho * cc[om + j]) + gphi_s[om + j] * d_s;
  }
  apr = blk_reduce(apr, 2, red);
  adu = blk_reduce(adu, 2, red);
  dphi = blk_reduce(dphi, 0, red);
  dphiR = blk_reduce(dphiR, 0, red);
  if (threadIdx.x == 0) {
    double* o = sc + (size_t)b * 4;
    o[0] = apr; o[1] = adu; o[2] = dphi; o[3] = dphiR;
  }
}

// ---- 4. trial points -----------------------------------------------------------------------------------------------
// rows [ns] instance of every search row, al [ns][Kw] step lengths; xt [ns][Kw][n] = x + al dx
__global__ void __launch_bounds__(RB_IPM_THREADS)
ipm_trial_kernel(int n, int Kw, const int* __restrict__ rows, const double* __restrict__ al, const double* __restrict__ x,
                 const double* __restrict__ dx, double* __restrict__ xt) {
  const int r = blockIdx.x / Kw, k = blockIdx.x - r * Kw;
  const size_t on = (size_t)rows[r] * n, ot = (size_t)blockIdx.x * n;
  const double alpha = al[(size_t)r * Kw + k];
  for (int i = threadIdx.x; i < n; i += blockDim.x) xt[ot + i] = x[on + i] + alpha * dx[on + i];
}

// ---- 5. filter quantities at the trial points ----------------------------------------------------------------------
// f_t [ns*Kw], g_t [ns*Kw][m]; out [ns*Kw][4] = { theta_t = |c_t|_1, phi_t (barrier objective), phi_R,t, finite flag }
__global__ void __launch_bounds__(RB_IPM_THREADS)
ipm_trial_merit_kernel(const RbIpm a, int Kw, const int* __restrict__ rows, const double* __restrict__ al,
                       const double* __restrict__ xt, const double* __restrict__ ds, const double* __restrict__ f_t,
                       const double* __restrict__ g_t, double* __restrict__ out) {
  __shared__ double red[33];
  const int r = blockIdx.x / Kw, k = blockIdx.x - r * Kw;
  const int b = rows[r];
  const size_t om = (size_t)b * a.m, on = (size_t)b * a.n;
  const size_t ot = (size_t)blockIdx.x * a.n, og = (size_t)blockIdx.x * a.m;
  const size_t bx = (size_t)b * a.sx, bs = (size_t)b * a.ss;
  const double alpha = al[(size_t)r * Kw + k], mu = a.mu[b], kd = a.kappa_d;
  const bool R = a.resto && a.resto[b];
  double lg = 0.0, dmp = 0.0, prox = 0.0;
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    const unsigned fl = a.xflag[bx + i];
    const bool hL = fl & 1u, hU = fl & 2u;
    const double x = xt[ot + i];
    const double dL = hL ? x - a.xL[bx + i] : 1.0, dU = hU ? a.xU[bx + i] - x : 1.0;
    lg += (hL ? log(dL) : 0.0) + (hU ? log(dU) : 0.0);
    dmp += (hL && !hU ? dL : 0.0) + (hU && !hL ? dU : 0.0);
    if (R) {
      const double xr = x - a.x_R[on + i];
      prox += a.DR2[on + i] * xr * xr;
    }
  }
  double th1 = 0.0, c2 = 0.0;
  for (int j = threadIdx.x; j < a.m; j += blockDim.x) {
    const unsigned fl = a.sflag[bs + j];
    const bool hL = fl & 1u, hU = fl & 2u, eq = fl & 4u;
    const double st = eq ? a.ceq[bs + j] : a.s[om + j] + alpha * ds[om + j];
    const double c = eq ? g_t[og + j] - a.ceq[bs + j] : g_t[og + j] - st;
    th1 += fabs(c);
    c2 += c * c;
    if (!eq) {
      const double eL = hL ? st - a.sL[bs + j] : 1.0, eU = hU ? a.sU[bs + j] - st : 1.0;
      lg += (hL ? log(eL) : 0.0) + (hU ? log(eU) : 0.0);
      dmp += (hL && !hU ? eL : 0.0) + (hU && !hL ? eU : 0.0);
    }
  }
  lg = blk_reduce(lg, 0, red);
  dmp = blk_reduce(dmp, 0, red);
  th1 = blk_reduce(th1, 0, red);
  c2 = blk_reduce(c2, 0, red);
  prox = blk_reduce(prox, 0, red);
  if (threadIdx.x == 0) {
    const double bar = -mu * lg + kd * mu * dmp;
    double* o = out + (size_t)blockIdx.x * 4;
    o[0] = th1;
    o[1] = f_t[blockIdx.x] + bar;
    o[2] = 0.5 * a.rho * c2 + 0.5 * sqrt(mu) * prox + bar;
    o[3] = (isfinite(th1) && isfinite(o[1])) ? 1.0 : 0.0;
  }
}

// ---- 6. primal-dual update with the kappa_sigma safeguard ------------------------------------------------------------
// alpha [B] accepted primal step (0: none), alpha_du [B] (already zero where the instance did not move)
__global__ void __launch_bounds__(RB_IPM_THREADS)
ipm_update_kernel(const RbIpm a, const double* __restrict__ alpha, const double* __restrict__ alpha_du,
                  const double* __restrict__ dx, const double* __restrict__ dy, const double* __restrict__ ds,
                  const double* __restrict__ dzL, const double* __restrict__ dzU, const double* __restrict__ dvL,
                  const double* __restrict__ dvU, double kappa_sigma) {
  const int b = blockIdx.x;
  const size_t on = (size_t)b * a.n, om = (size_t)b * a.m;
  const size_t bx = (size_t)b * a.sx, bs = (size_t)b * a.ss;
  const double al = alpha[b], ad = alpha_du[b], mu = a.mu[b];
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    const unsigned fl = a.xflag[bx + i];
    const double x = a.x[on + i] + al * dx[on + i];
    double zl = a.zL[on + i] + ad * dzL[on + i], zu = a.zU[on + i] + ad * dzU[on + i];
    if (fl & 1u) {
      const double d = x - a.xL[bx + i];
      zl = fmax(fmin(zl, kappa_sigma * mu / d), mu / (kappa_sigma * d));
    }
    if (fl & 2u) {
      const double d = a.xU[bx + i] - x;
      zu = fmax(fmin(zu, kappa_sigma * mu / d), mu / (kappa_sigma * d));
    }
    a.x[on + i] = x;
    a.zL[on + i] = zl;
    a.zU[on + i] = zu;
  }
  for (int j = threadIdx.x; j < a.m; j += blockDim.x) {
    const unsigned fl = a.sflag[bs + j];
    const bool eq = fl & 4u;
    const double s = eq ? a.ceq[bs + j] : a.s[om + j] + al * ds[om + j];
    double vl = a.vL[om + j] + ad * dvL[om + j], vu = a.vU[om + j] + ad * dvU[om + j];
    if (fl & 1u) {
      const double d = s - a.sL[bs + j];
      vl = fmax(fmin(vl, kappa_sigma * mu / d), mu / (kappa_sigma * d));
    }
    if (fl & 2u) {
      const double d = a.sU[bs + j] - s;
      vu = fmax(fmin(vu, kappa_sigma * mu / d), mu / (kappa_sigma * d));
    }
    a.s[om + j] = s;
    a.y[om + j] += al * dy[om + j];
    a.vL[om + j] = vl;
    a.vU[om + j] = vu;
  }
}
