// Batched block-tridiagonal + border KKT factorisation and solve, interface form (stage blocks that fit shared memory).
//
// What it replaces: the sparse symmetric-indefinite factor / solve IPOPT performs on every iteration inside
// `self.solver(x0=..., ...)` (drone3d/raceline/base_raceline.py:160-165; linear solver chosen at :765-787).
// Same matrix, same block tables as kkt_blocks.cuh (host side: aircraft_trajectory_optimization_b200/kkt.py); what
// differs is how the loop-closure border is carried:
//
//   [ T  E ] [x_T]   [r_T]        T block tridiagonal (diagonal blocks A_n, couplings L_n), E the border columns
//   [ E' G ] [x_b] = [r_b]
//
// Block LDL' with the border last: S_0 = A_0, S_{n+1} = A_{n+1} - L_n S_n^-1 L_n'.  The border rows of the factor,
// Y = L^-1 E, satisfy Y_{n+1} = E_{n+1} - (L_n S_n^-1) Y_n and are non-zero only on the support rows
// sup_n = cr_{n-1} U rows(E_n) of block n (17 of 43 for race.py), so the factorisation carries P_n = Y_n[sup_n, :act_n]
// (act_n = border columns that have appeared so far) instead of 1 + nb dense right-hand-side columns:
//
//   Q_n = S_n^-1[:, sup_n] P_n          border Schur complement  G - sum_n P_n' Q_n[sup_n]
//   solve:  forward  y_n = r_n - L_{n-1} z_{n-1},  z_n = S_n^-1 y_n,  r_b -= P_n' z_n[sup_n]
//           border   x_b = (G - ...)^-1 r_b
//           backward x_n = z_n - Q_n x_b - S_n^-1 L_n' x_{n+1}
//
// kkt_factor_kernel : one CTA per instance, no right-hand side.  The in-place Gauss-Jordan inverse of every diagonal
//                     block runs as a two-role pipeline (kf_sym_invert below); the values and index tables of the
//                     blocks to come arrive by cp.async one and two blocks ahead, so no global-memory latency is
//                     exposed on the sequential chain.
// kkt_solve_kernel  : one CTA per instance and right-hand side; the factor blocks stream in by TMA bulk copies
//                     (cp.async.bulk + mbarrier, three blocks in flight).
//
// Measured on B200 (profiles/microbench/lat.cu): dependent DFMA / DMUL / DADD 23 cycles, SHFL 26, REDUX 32, LDS ~30,
// bar.sync of 288 threads 40 -- and 190 when every thread has 64 bytes of shared-memory stores in flight (store
// bandwidth: 128 B / cycle).  The design below follows from those numbers: no conversions or divisions on the pivot
// chain, integer comparisons of magnitudes, half the matrix stored per step.
#pragma once
#include "kkt_blocks.cuh"

#ifndef RB_KF_EXP
#define RB_KF_EXP 0
#endif
#ifndef RB_KF_THREADS
#define RB_KF_THREADS 192
#endif
// RB_KF_THREADS: warp 0: pivot search; warps 1..5: 4 x 4 tiles of the upper triangle (up to 16 x 16 tiles)
#define RB_KF_VPAD 72       // update vectors: largest block (64) + tile overhang
#define RB_KS_THREADS 256
#ifndef RB_KS_STAGES
#define RB_KS_STAGES 3
#endif

struct __align__(16) KfVec {
  double c1[RB_KF_VPAD], r1[RB_KF_VPAD], c2[RB_KF_VPAD], r2[RB_KF_VPAD];
  int type, p, q, pad;
};

// per-block record (ints), built by rb_kkt_create from the tables of rb_kkt_desc: header, then the index lists
#define KF_REC_U0 0
#define KF_REC_B 1
#define KF_REC_SN 2
#define KF_REC_AN 3
#define KF_REC_M 4
#define KF_REC_Q 5
#define KF_REC_BE0 6
#define KF_REC_BE1 7
#define KF_REC_POFF 8
#define KF_REC_QOFF 9
#define KF_REC_HDR 12
struct RbKktRecDev {
  int R;                       // ints per block record (multiple of 4): header | sup[SP] | cc[QP] | cr[MP] | crs[MP]
  int SP, QP, MP;
  int T;                       // ints per block table (multiple of 4): dA_pos[SD] | dA_src[SD] | cL_pos[SL] | cL_src[SL] | unk[UB]
  int SD, SL, UB;
  const int32_t* rec;          // [N][R]
  const int32_t* tab;          // [N][T]   (pos / src / unk = -1: padding)
};

struct RbKktChainBatch {
  int B, nnzh, nnzj;
  const double *hess, *jac, *dx_diag, *neg_d;
  const double* rhs;
  double* sol;
  double *Sinv, *YL, *P, *Q, *Xr, *SB, *Gacc;   // factor storage, instance-major inside every region
  long long p_total, q_total;                   // doubles per instance of the P / Q regions
  int* status;                                  // [B][2] vanishing pivots, negative eigenvalues
  int pq_stage;                                 // doubles of the P part of a TMA stage buffer (solve kernel)
  RbKktRecDev rt;
  // condensed systems (kkt_condense.cuh): values of source kind 4, [B][naux]; pivot counts of the interiors to add
  const double* aux;
  int naux;
  const int* status_add;                        // [B][2] or null
  // solve kernel only: factor slot of batch row p (re-solves of a subset of the factorised instances); null = p
  const int* inst;
};

__device__ __forceinline__ double kf_upd1(double m, double c, double r) { return __fma_rn(-c, r, m); }
__device__ __forceinline__ double kf_upd2(double m, double ca, double ra, double cb, double rb) {
  return m - __fma_rn(cb, rb, __dmul_rn(ca, ra));
}

// 1 / d from the hardware approximation and two Newton steps (the full-precision division is longer, and it sits
// next to the pivot warp's critical path)
__device__ __forceinline__ double kf_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = __fma_rn(-d, r, 1.0);
  r = __fma_rn(r, e, r);
  e = __fma_rn(-d, r, 1.0);
  return __fma_rn(r, e, r);
}

// arg-max key: the high word of |a| (exponent + 13 mantissa bits survive) with the low 7 bits replaced by 64 - index,
// so that one __reduce_max_sync per scan yields the magnitude and the index (smallest index on ties); 0 = not a
// candidate.  Magnitudes are compared as integers (the high word of a non-negative double is monotone).
__device__ __forceinline__ unsigned kf_key(double a, int idx, bool ok) {
  unsigned hi = (unsigned)__double2hiint(a) & 0x7fffffffu;
  if (hi >= 0x7ff00000u) hi = (__double2loint(a) == 0 && hi == 0x7ff00000u) ? hi : 0u;      // NaN -> 0, inf stays
  return ok ? ((hi & 0xffffff80u) | (unsigned)(64 - idx)) : 0u;
}
__device__ __forceinline__ int kf_key_idx(unsigned key) { return 64 - (int)(key & 0x7fu); }
__device__ __forceinline__ unsigned kf_key_bits(unsigned key) { return key & 0xffffff80u; }
__device__ __forceinline__ double kf_bits_mag(unsigned bits) { return __hiloint2double((int)bits, 0); }
__device__ __forceinline__ unsigned kf_mag_bits(double a) { return (unsigned)__double2hiint(a) & 0x7fffff80u; }

__device__ __forceinline__ void kf_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void kf_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__host__ __device__ inline int kf_ld(int bmax, int nb) {
  const int nbb = nb > bmax ? nb : bmax;
  return ((nbb + 3) & ~3) + 2;          // >= 4 ceil(nbb / 4), = 2 mod 4: 16-byte rows, 4-way conflicts at worst on column reads
}

// In-place inverse of the symmetric b x b matrix (buf0 + diag(shift)) (row-major, leading dimension LD = kf_ld) by
// Gauss-Jordan sweeps with Bunch-Kaufman pivoting (the pivot rule of kkt_sym_invert: k = largest remaining diagonal,
// r = largest off-diagonal of its row; 1 x 1 pivot k if |a_kk| >= alpha lambda or |a_kk| sigma >= alpha lambda^2,
// 1 x 1 pivot r if |a_rr| >= alpha sigma, else the 2 x 2 pivot (k, r)).  State s of the matrix lives in buf[s & 1];
// returns the index of the buffer holding the inverse (both triangles).  Called by all RB_KF_THREADS threads.
// stat[0] += vanishing pivots, stat[1] += negative eigenvalues.
//
// Only the upper triangle is carried, in 4 x 4 register tiles (tile rows <= tile columns): a swept matrix stays
// symmetric up to signs, M[i][j] = -M[j][i] when exactly one of i, j has been pivoted, else +M[j][i].
//
// Two roles, decoupled by named barriers (producer / consumer, ids alternate with the step parity):
//   pivot warp, step u : waits for state u-1 in shared memory, reads the row(s) it needs, applies update u-1 to them on
//                        the fly (its vectors are still in registers), chooses pivot u, writes the vectors of update u,
//                        signals FULL_u.
//   tile warps, iter u : wait for FULL_u, apply update u to their register tiles, store state u+1 (16-byte stores),
//                        signal STATE_{u+1}.
// The tile warps trail the pivot warp by one step, so neither their store drain nor a block barrier sits on the
// pivot chain.
#define KF_BAR_FULL(u) (1 + ((u) & 1))
#define KF_BAR_STATE(u) (3 + ((u) & 1))
__device__ __noinline__ int kf_sym_invert(double* __restrict__ buf0, double* __restrict__ buf1,
                                          const double* __restrict__ shift, int LD, int b, KfVec* vec, int* stat) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double alpha = 0.6403882032022076, ralpha = 1.0 / 0.6403882032022076;
  // the pivot warp of co-resident CTAs should sit on different warp schedulers (warp index mod 4): CTAs that share
  // an SM usually differ by multiples of the SM count in blockIdx
  const int pw = (int)((blockIdx.x + blockIdx.x / 148u) & 3u);
  __syncthreads();          // buf0 and shift are complete
  int t = 0;
  if (warp == pw) {
    // ---------------------------------------------------------------- pivot warp
    // Single-warp critical path of the whole factorisation: kept to ~150 instructions per step in the common case
    // (1 x 1 pivot on the largest diagonal entry; previous update 1 x 1).  Everything else branches off.
    const int ja = lane, jb = lane + 32;
    const unsigned ibA = (unsigned)(64 - ja), ibB = (unsigned)(64 - jb);     // index bits of the arg-max keys
    double dgA = ja < b ? buf0[ja * LD + ja] + shift[ja] : 0.0;
    double dgB = jb < b ? buf0[jb * LD + jb] + shift[jb] : 0.0;
    double pc1A = 0, pc1B = 0, pr1A = 0, pr1B = 0, pc2A = 0, pc2B = 0, pr2A = 0, pr2B = 0;   // vectors of the last update
    int ptype = 0, pp = -1, pq = -1;
    bool swA = !(ja < b), swB = !(jb < b);      // pivoted so far (or out of range)
    bool pswA = false, pswB = false;            // pivoted before the last update (the state in shared memory)
    int remaining = b, bad = 0, nneg = 0;
    auto key = [](double a, unsigned ib, bool ok) -> unsigned {
      const unsigned kk = ((unsigned)__double2hiint(a) & 0x7fffff80u) | ib;     // NaN ranks above everything: it
      return ok ? kk : 0u;                                                     // surfaces in the result / status
    };
    auto flip = [](double a, bool f) -> double {
      return __hiloint2double(__double2hiint(a) ^ (f ? (int)0x80000000 : 0), __double2loint(a));
    };
    for (;; ++t) {
      KfVec& V = vec[t & 1];
      if (t >= 2) kf_bar_sync(KF_BAR_STATE(t - 1), RB_KF_THREADS);     // state t-1 is in shared memory
      if (remaining == 0) {
        if (lane == 0) V.type = 0;
        kf_bar_arrive(KF_BAR_FULL(t), RB_KF_THREADS);
        if (t >= 1) kf_bar_sync(KF_BAR_STATE(t), RB_KF_THREADS);       // the last state (balances the arrivals)
        break;
      }
      const double* __restrict__ Mo = (t == 0 || ((t - 1) & 1) == 0) ? buf0 : buf1;
      // logical row k (k not pivoted) of the current state: stored state t-1 + update t-1.  Stored entries: M[k][j]
      // for j in or right of k's tile column, else M[j][k] = -/+ M[k][j].  A pivot column j of update t-1 holds
      // -(c_k r_j) afterwards: the stored value is dropped (s = 0) and the same multiply-add yields it.
      auto rowread = [&](int k, double& oA, double& oB) {
        const int kb0 = k & ~3;
        double sA = Mo[ja >= kb0 ? k * LD + ja : ja * LD + k];      // (lanes beyond b read inside the buffer: LD rows)
        double sB = jb < b ? Mo[jb >= kb0 ? k * LD + jb : jb * LD + k] : 0.0;
        const double ck1 = __shfl_sync(0xffffffffu, k < 32 ? pc1A : pc1B, k & 31);
        sA = flip(sA, ja < kb0 && pswA);
        sB = flip(sB, jb < kb0 && pswB);
        if (ja == pp || ja == pq) sA = 0.0;
        if (jb == pp || jb == pq) sB = 0.0;
        if (ptype == 2) {
          const double ck2 = __shfl_sync(0xffffffffu, k < 32 ? pc2A : pc2B, k & 31);
          oA = sA - __fma_rn(ck2, pr2A, __dmul_rn(ck1, pr1A));
          oB = sB - __fma_rn(ck2, pr2B, __dmul_rn(ck1, pr1B));
        } else {
          oA = kf_upd1(sA, ck1, pr1A);
          oB = kf_upd1(sB, ck1, pr1B);
        }
      };
      const bool la = !swA, lb = !swB;
      int k, r, type = 1, p, q;
      {
        const unsigned ka = key(dgA, ibA, la), kb = key(dgB, ibB, lb);
        const unsigned best = __reduce_max_sync(0xffffffffu, ka > kb ? ka : kb);
        k = kf_key_idx(best);
      }
      p = q = k;
      double rkA, rkB, rrA = 0.0, rrB = 0.0;
      rowread(k, rkA, rkB);
      // next to the row scan: the reciprocal of the most likely pivot, |a_kk| / alpha as comparison bits
      double dk = __shfl_sync(0xffffffffu, k < 32 ? dgA : dgB, k & 31);
      if (!(fabs(dk) > 1e-250)) dk = 1e-250;
      double di = kf_rcp(dk);
      const unsigned thr_bits = kf_mag_bits(fabs(dk) * ralpha);
      unsigned lam_bits;
      {
        const unsigned ka = key(rkA, ibA, la && ja != k), kb = key(rkB, ibB, lb && jb != k);
        const unsigned best = __reduce_max_sync(0xffffffffu, ka > kb ? ka : kb);
        r = kf_key_idx(best);
        lam_bits = kf_key_bits(best);
      }
#if RB_KF_EXP == 2
      if (false) {
#else
      if (remaining > 1 && !(thr_bits >= lam_bits)) {      // |a_kk| < alpha lambda (13-bit mantissas)
#endif
        rowread(r, rrA, rrB);
        const unsigned ka = key(rrA, ibA, la && ja != r), kb = key(rrB, ibB, lb && jb != r);
        const unsigned best = __reduce_max_sync(0xffffffffu, ka > kb ? ka : kb);
        const double sig = kf_bits_mag(kf_key_bits(best)), lam = kf_bits_mag(lam_bits), akk = fabs(dk);
        const double arr = fabs(__shfl_sync(0xffffffffu, r < 32 ? dgA : dgB, r & 31));
        if (akk * sig >= alpha * lam * lam) {
          p = k;
        } else if (arr >= alpha * sig) {
          p = r;
        } else {
          type = 2;
          p = k < r ? k : r;
          q = k < r ? r : k;
        }
      }
      double c1A, c1B, r1A, r1B, c2A = 0.0, c2B = 0.0, r2A = 0.0, r2B = 0.0;
      if (type == 1) {
        const bool is_k = (p == k);
        double d = dk;
        if (!is_k) {
          d = __shfl_sync(0xffffffffu, p < 32 ? dgA : dgB, p & 31);
          if (!(fabs(d) > 1e-250)) d = 1e-250;
          di = kf_rcp(d);
        }
        if (d == 1e-250) bad++;
        if (d < 0) nneg++;
        const double rpA = is_k ? rkA : rrA, rpB = is_k ? rkB : rrB;     // M[p][j]
        c1A = flip(rpA, swA);                                            // M[j][p] = -/+ M[p][j]
        c1B = flip(rpB, swB);
        r1A = (ja == p) ? di : rpA * di;
        r1B = (jb == p) ? di : rpB * di;
        V.c1[ja] = c1A;       // (entries beyond b are written too: the tile warps read whole 4-vectors)
        V.r1[ja] = r1A;
        V.c1[jb] = c1B;
        V.r1[jb] = r1B;
      } else {
        const bool pk = (p == k);
        const double epp = __shfl_sync(0xffffffffu, p < 32 ? dgA : dgB, p & 31);
        const double eqq = __shfl_sync(0xffffffffu, q < 32 ? dgA : dgB, q & 31);
        const double apA = pk ? rkA : rrA, apB = pk ? rkB : rrB;         // M[p][j]
        const double aqA = pk ? rrA : rkA, aqB = pk ? rrB : rkB;         // M[q][j]
        const double epq = __shfl_sync(0xffffffffu, q < 32 ? apA : apB, q & 31);
        double det = epp * eqq - epq * epq;
        if (!(fabs(det) > 1e-250)) {
          det = -1e-250;
          bad++;
        }
        nneg += det < 0 ? 1 : (epp + eqq < 0 ? 2 : 0);
        const double rdet = kf_rcp(det);
        const double i00 = eqq * rdet, i01 = -epq * rdet, i11 = epp * rdet;
        c1A = flip(apA, swA);
        c1B = flip(apB, swB);
        c2A = flip(aqA, swA);
        c2B = flip(aqB, swB);
        r1A = (ja == p) ? i00 : ((ja == q) ? i01 : i00 * apA + i01 * aqA);
        r2A = (ja == p) ? i01 : ((ja == q) ? i11 : i01 * apA + i11 * aqA);
        r1B = (jb == p) ? i00 : ((jb == q) ? i01 : i00 * apB + i01 * aqB);
        r2B = (jb == p) ? i01 : ((jb == q) ? i11 : i01 * apB + i11 * aqB);
        V.c1[ja] = c1A;
        V.r1[ja] = r1A;
        V.c2[ja] = c2A;
        V.r2[ja] = r2A;
        V.c1[jb] = c1B;
        V.r1[jb] = r1B;
        V.c2[jb] = c2B;
        V.r2[jb] = r2B;
      }
      if (lane == 0) *reinterpret_cast<int4*>(&V.type) = make_int4(type, p, type == 2 ? q : -1, 0);
      kf_bar_arrive(KF_BAR_FULL(t), RB_KF_THREADS);
      pswA = swA;
      pswB = swB;
      swA = swA || ja == p || (type == 2 && ja == q);
      swB = swB || jb == p || (type == 2 && jb == q);
      // the diagonal of the next state (entries that are still alive), and this step's vectors for the lazy rows
      if (type == 1) {
        dgA = kf_upd1(dgA, c1A, r1A);
        dgB = kf_upd1(dgB, c1B, r1B);
      } else {
        dgA = kf_upd2(dgA, c1A, r1A, c2A, r2A);
        dgB = kf_upd2(dgB, c1B, r1B, c2B, r2B);
      }
      pc1A = c1A; pc1B = c1B; pr1A = r1A; pr1B = r1B;
      pc2A = c2A; pc2B = c2B; pr2A = r2A; pr2B = r2B;
      ptype = type;
      pp = p;
      pq = (type == 2) ? q : -1;
      remaining -= type;
    }
    if (lane == 0) {
      stat[0] += bad;
      stat[1] += nneg;
    }
  } else {
    // ---------------------------------------------------------------- tile warps
    // tile (I, J), I <= J: rows 4 I .. 4 I + 3, columns 4 J .. 4 J + 3; tiles numbered row by row over the upper triangle
    const int w = (warp < pw ? warp : warp - 1) * 32 + lane;
    const int nt = (b + 3) >> 2;
    int I = 0, off = 0;
    while (I < nt && w >= off + (nt - I)) {
      off += nt - I;
      ++I;
    }
    const bool owner = I < nt;
    const int J = owner ? I + (w - off) : 0;
    const int i0 = 4 * I, j0 = 4 * J;
    double m[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = i0 + a, j = j0 + c;
        m[a][c] = (owner && i < b && j < b) ? buf0[i * LD + j] + (i == j ? shift[i] : 0.0) : 0.0;
      }
    for (;; ++t) {
      kf_bar_sync(KF_BAR_FULL(t), RB_KF_THREADS);
      const KfVec& V = vec[t & 1];
      const int type = V.type;
      if (type == 0) break;
#if RB_KF_EXP == 1
      if (false) {
#else
      if (owner) {
#endif
        double* __restrict__ Mw = ((t + 1) & 1) ? buf1 : buf0;
        const int p = V.p, q = V.q;
        const double2 ra0 = *reinterpret_cast<const double2*>(&V.r1[j0]), ra1 = *reinterpret_cast<const double2*>(&V.r1[j0 + 2]);
        const double2 ca0 = *reinterpret_cast<const double2*>(&V.c1[i0]), ca1 = *reinterpret_cast<const double2*>(&V.c1[i0 + 2]);
        const double rr[4] = {ra0.x, ra0.y, ra1.x, ra1.y}, cc[4] = {ca0.x, ca0.y, ca1.x, ca1.y};
        if (type == 1) {
          const double di = V.r1[p];
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const double ccdi = -cc[a] * di;     // pivot column: -a_ip / d
            const bool rsel = (i0 + a == p);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              // pivot row: a_pj / d (1 / d on the pivot); elsewhere the rank-1 update
              double v = (j0 + c == p) ? ccdi : kf_upd1(m[a][c], cc[a], rr[c]);
              m[a][c] = rsel ? rr[c] : v;
            }
          }
        } else {
          const double2 rb0 = *reinterpret_cast<const double2*>(&V.r2[j0]), rb1 = *reinterpret_cast<const double2*>(&V.r2[j0 + 2]);
          const double2 cb0 = *reinterpret_cast<const double2*>(&V.c2[i0]), cb1 = *reinterpret_cast<const double2*>(&V.c2[i0 + 2]);
          const double rb[4] = {rb0.x, rb0.y, rb1.x, rb1.y}, cb[4] = {cb0.x, cb0.y, cb1.x, cb1.y};
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int rsel = (i0 + a == p) ? 1 : ((i0 + a == q) ? 2 : 0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              // pivot rows: E^-1 A_Pj (E^-1 itself on the pivot block); pivot columns: -A_iP E^-1; else rank-2 update
              const double tt = __fma_rn(cb[a], rb[c], __dmul_rn(cc[a], rr[c]));
              double v = (j0 + c == p || j0 + c == q) ? -tt : m[a][c] - tt;
              m[a][c] = rsel == 1 ? rr[c] : (rsel == 2 ? rb[c] : v);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = i0 + a;
          if (i < b) {
            *reinterpret_cast<double2*>(&Mw[i * LD + j0]) = make_double2(m[a][0], m[a][1]);
            *reinterpret_cast<double2*>(&Mw[i * LD + j0 + 2]) = make_double2(m[a][2], m[a][3]);
          }
        }
      }
      kf_bar_arrive(KF_BAR_STATE(t + 1), RB_KF_THREADS);
    }
  }
  __syncthreads();
  // mirror the tiles below the diagonal (everything is pivoted now: the inverse is symmetric)
  {
    double* __restrict__ Mr = (t & 1) ? buf1 : buf0;
    for (int i = 4 + warp; i < b; i += RB_KF_THREADS / 32)
      for (int j = lane; j < (i & ~3); j += 32) Mr[i * LD + j] = Mr[j * LD + i];
  }
  __syncthreads();
  return t & 1;
}

struct KfSmem {
  double *M0, *M1, *carry, *PA, *PB, *Lc, *shift, *vals;
  int32_t *rec0, *rec1, *tab0, *tab1;
  KfVec* vec;
  int* stat;
};

__host__ __device__ inline size_t kf_mbuf_doubles(int bmax, int nb) {
  const int nbb = nb > bmax ? nb : bmax;
  return (size_t)(((nbb + 3) & ~3)) * kf_ld(bmax, nb);
}

__host__ __device__ inline size_t kf_even(size_t x) { return (x + 1) & ~(size_t)1; }

__host__ __device__ inline size_t kf_smem_bytes(int bmax, int nb, int mmax, int qmax, int amax, int smax, int R, int T,
                                                int SD, int SL, int UB) {
  const int ldq = (amax + 1) / 2 * 2;
  size_t n = 2 * kf_mbuf_doubles(bmax, nb);          // two states of the matrix
  n += kf_even((size_t)mmax * mmax);                 // carry (every region an even number of doubles: 16-byte aligned)
  n += (size_t)(smax > 0 ? smax : 1) * (ldq > 0 ? ldq : 2);   // P_n
  n += (size_t)mmax * (ldq > 0 ? ldq : 2);           // the rows of P_{n+1} that come from block n
  n += kf_even((size_t)mmax * qmax);                 // L_n
  n += RB_KF_VPAD;                                   // diagonal shift of the block
  n += kf_even((size_t)SD + SL + UB);                // staged values of the next block
  return n * sizeof(double) + 2 * (size_t)(R + T) * sizeof(int32_t) + 2 * sizeof(KfVec) + 4 * sizeof(int);
}

// the scratch (the matrix buffer that does not hold the inverse) hosts YL_n and Q_n side by side
__host__ inline bool kf_scratch_fits(int bmax, int nb, int mmax, int N, const int32_t* act, const int32_t* cr_ptr) {
  const size_t cap = kf_mbuf_doubles(bmax, nb);
  for (int n = 0; n < N; ++n) {
    const size_t ldq = (size_t)(act[n] + 1) / 2 * 2;
    const bool coupled = n < N - 1 && cr_ptr[n + 1] > cr_ptr[n];
    if ((coupled ? (size_t)bmax * mmax : 0) + (size_t)bmax * ldq > cap) return false;
  }
  return true;
}

__device__ __forceinline__ void kf_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void kf_cp8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void kf_cp_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ const double* kkt_val_ptr(const KktVals& v, int32_t src) {
  const int kind = (src >> 28) & 7;
  const int idx = src & 0x0fffffff;
  const double* base = kind == 0 ? v.hess : (kind == 1 ? v.jac : (kind == 2 ? v.dx : (kind == 3 ? v.nd : v.aux)));
  return base + idx;
}

__global__ void __launch_bounds__(RB_KF_THREADS, 3)
kkt_factor_kernel(const RbKktDev d, const RbKktChainBatch bt) {
  extern __shared__ __align__(16) double kf_smem[];
  const int p = blockIdx.x;
  if (p >= bt.B) return;
  const int tid = threadIdx.x, nth = RB_KF_THREADS;
  const int bmax = d.bmax, mmax = d.mmax, qmax = d.qmax, N = d.N, nb = d.nb;
  const int LD = kf_ld(bmax, nb);
  const int ldqmax = (d.amax + 1) / 2 * 2;
  const RbKktRecDev rt = bt.rt;
  KfSmem s;
  {
    double* q = kf_smem;
    s.M0 = q; q += kf_mbuf_doubles(bmax, nb);
    s.M1 = q; q += kf_mbuf_doubles(bmax, nb);
    s.carry = q; q += kf_even((size_t)mmax * mmax);
    s.PA = q; q += (size_t)(d.smax > 0 ? d.smax : 1) * (ldqmax > 0 ? ldqmax : 2);
    s.PB = q; q += (size_t)mmax * (ldqmax > 0 ? ldqmax : 2);
    s.Lc = q; q += kf_even((size_t)mmax * qmax);
    s.shift = q; q += RB_KF_VPAD;
    s.vals = q; q += kf_even((size_t)rt.SD + rt.SL + rt.UB);
    int32_t* qi = reinterpret_cast<int32_t*>(q);
    s.rec0 = qi; qi += rt.R;
    s.rec1 = qi; qi += rt.R;
    s.tab0 = qi; qi += rt.T;
    s.tab1 = qi; qi += rt.T;
    s.vec = reinterpret_cast<KfVec*>(qi);
    s.stat = reinterpret_cast<int*>(s.vec + 2);
  }
  KktVals v{bt.hess + (size_t)p * bt.nnzh, bt.jac + (size_t)p * bt.nnzj, bt.dx_diag + (size_t)p * d.nw,
            bt.neg_d + (size_t)p * d.ng, bt.aux ? bt.aux + (size_t)p * bt.naux : nullptr};
  double* __restrict__ Sinv_g = bt.Sinv + (size_t)p * N * bmax * d.ldS;
  double* __restrict__ YL_g = bt.YL + (size_t)p * N * bmax * d.ldY;
  double* __restrict__ P_g = bt.P + (size_t)p * bt.p_total;
  double* __restrict__ Q_g = bt.Q + (size_t)p * bt.q_total;
  double* __restrict__ Gacc = bt.Gacc + (size_t)p * nb * nb;
  if (tid < 4) s.stat[tid] = 0;
  for (int i = tid; i < nb * nb; i += nth) Gacc[i] = 0.0;
#ifdef RB_KF_PROFILE
  long long tk[5] = {0, 0, 0, 0, 0};
  long long t0 = clock64(), t1;
#define KF_TICK(i) do { __syncthreads(); t1 = clock64(); tk[i] += t1 - t0; t0 = t1; } while (0)
#else
#define KF_TICK(i)
#endif
  auto rec_of = [&](int n) { return (n & 1) ? s.rec1 : s.rec0; };
  auto tab_of = [&](int n) { return (n & 1) ? s.tab1 : s.tab0; };
  // asynchronous copies (each thread later consumes exactly what it copied, after kf_cp_wait)
  auto fetch_rec = [&](int n) {      // the record of block n
    for (int i = tid * 4; i < rt.R; i += nth * 4) kf_cp16(rec_of(n) + i, rt.rec + (size_t)n * rt.R + i);
  };
  auto fetch_tab = [&](int n) {      // the entry tables of block n
    for (int i = tid * 4; i < rt.T; i += nth * 4) kf_cp16(tab_of(n) + i, rt.tab + (size_t)n * rt.T + i);
  };
  auto fetch_vals = [&](int n) {     // the values of block n (its tables must have landed and been synchronised)
    const int32_t* __restrict__ tb = tab_of(n);
    for (int e = tid; e < rt.SD; e += nth) {
      const int src = tb[rt.SD + e];
      if (src >= 0) kf_cp8(s.vals + e, kkt_val_ptr(v, src));
    }
    for (int e = tid; e < rt.SL; e += nth) {
      const int src = tb[2 * rt.SD + rt.SL + e];
      if (src >= 0) kf_cp8(s.vals + rt.SD + e, kkt_val_ptr(v, src));
    }
    for (int e = tid; e < rt.UB; e += nth) {
      const int u = tb[2 * rt.SD + 2 * rt.SL + e];
      if (u >= 0) kf_cp8(s.vals + rt.SD + rt.SL + e, u < d.nw ? v.dx + u : v.nd + (u - d.nw));
    }
  };
  // zero the buffers block n is assembled in (M0, L_n, P_n)
  auto zero_for = [&](int n) {
    const int32_t* __restrict__ rc = rec_of(n);
    const int b = rc[KF_REC_B], sn = rc[KF_REC_SN], ldq = (rc[KF_REC_AN] + 1) / 2 * 2, m = rc[KF_REC_M];
    for (int i = tid; i < b * LD; i += nth) s.M0[i] = 0.0;
    for (int i = tid; i < sn * ldq; i += nth) s.PA[i] = 0.0;
    for (int i = tid; i < m * qmax; i += nth) s.Lc[i] = 0.0;
  };
  // block n into M0 (without its diagonal shift), shift, L_n, rows of P_n from block n - 1 (values staged by this thread)
  auto scatter = [&](int n, int m_prev, int a_prev, int ldq_prev, const int32_t* crs_prev) {
    const int32_t* __restrict__ tb = tab_of(n);
    const int32_t* __restrict__ rc = rec_of(n);
    for (int e = tid; e < rt.SD; e += nth) {
      const int pos = tb[e];
      if (pos >= 0) {
        const int r = pos / bmax, c = pos - r * bmax;
        s.M0[r * LD + c] = s.vals[e];
      }
    }
    for (int e = tid; e < rt.SL; e += nth) {
      const int pos = tb[2 * rt.SD + e];
      if (pos >= 0) s.Lc[pos] = s.vals[rt.SD + e];
    }
    for (int e = tid; e < RB_KF_VPAD; e += nth) s.shift[e] = (e < rt.UB && tb[2 * rt.SD + 2 * rt.SL + e] >= 0) ? s.vals[rt.SD + rt.SL + e] : 0.0;
    const int ldq = (rc[KF_REC_AN] + 1) / 2 * 2;
    for (int i = tid; i < m_prev * a_prev; i += nth) {
      const int a = i / a_prev, j = i - a * a_prev;
      s.PA[crs_prev[a] * ldq + j] = s.PB[a * ldq_prev + j];
    }
  };

  // ---- prologue: records / tables of blocks 0 and 1, values of block 0
  fetch_rec(0);
  fetch_tab(0);
  if (N > 1) fetch_tab(1);
  kf_cp_wait();
  __syncthreads();
  fetch_vals(0);
  kf_cp_wait();
  zero_for(0);
  __syncthreads();
  scatter(0, 0, 0, 2, nullptr);
  __syncthreads();

  for (int n = 0; n < N; ++n) {
    const int32_t* __restrict__ rc = rec_of(n);
    const int b = rc[KF_REC_B], sn = rc[KF_REC_SN], an = rc[KF_REC_AN], m = rc[KF_REC_M], q = rc[KF_REC_Q];
    const int ldq = (an + 1) / 2 * 2;
    const int32_t* __restrict__ sup = rc + KF_REC_HDR;
    const int32_t* __restrict__ cc = sup + rt.SP;
    const int32_t* __restrict__ cr = cc + rt.QP;
    const int32_t* __restrict__ crs = cr + rt.MP;
    // border entries of this block (few blocks have any)
    for (int e = rc[KF_REC_BE0] + tid; e < rc[KF_REC_BE1]; e += nth)
      s.PA[d.bE_sup[e] * ldq + d.bE_col[e]] += kkt_val(v, d.bE_src[e]);     // (row, column) pairs are unique
    // in flight during the inversion: values of block n + 1, tables of block n + 2, record of block n + 1
    if (n + 1 < N) {
      fetch_vals(n + 1);
      fetch_rec(n + 1);
    }
    if (n + 2 < N) fetch_tab(n + 2);
    asm volatile("cp.async.commit_group;" ::: "memory");
    KF_TICK(0);
    const int res = kf_sym_invert(s.M0, s.M1, s.shift, LD, b, s.vec, s.stat);
    KF_TICK(1);
    const double* __restrict__ Si = res ? s.M1 : s.M0;     // S_n^-1 (symmetric)
    double* __restrict__ YLs = res ? s.M0 : s.M1;          // scratch: YL_n (b x mmax), then Q_n (b x ldq)
    double* __restrict__ Qs = YLs + (m > 0 ? (size_t)bmax * mmax : 0);
    const double* __restrict__ Lc = s.Lc;

    // ---- phase A: Q_n = S^-1[:, sup] P_n (b x an) and YL_n = S^-1[:, cc] L_n' (b x m).  One row and KA columns per
    // item: KA independent multiply-add chains per thread (a dependent DFMA takes 23 cycles), S^-1 read through its
    // symmetry (lanes along a row).  Factor blocks go to global memory as they appear.
    {
      constexpr int KA = 9;
      const int njg = (an + KA - 1) / KA, nag = (m + KA - 1) / KA;
      for (int item = tid; item < b * (njg + nag); item += nth) {
        const int g = item / b, i = item - g * b;
        double acc[KA];
#pragma unroll
        for (int c = 0; c < KA; ++c) acc[c] = 0.0;
        if (g < njg) {
          const int j0 = g * KA, nc = an - j0 < KA ? an - j0 : KA;
          const double* __restrict__ pr = s.PA + j0;            // columns beyond an stay inside the padded row
          for (int t = 0; t < sn; ++t) {
            const double sv = Si[sup[t] * LD + i];
#pragma unroll
            for (int c = 0; c < KA; ++c) acc[c] = __fma_rn(sv, pr[t * ldq + (c < nc ? c : 0)], acc[c]);
          }
          double* __restrict__ qg = Q_g + rc[KF_REC_QOFF] + (size_t)i * ldq + j0;
#pragma unroll
          for (int c = 0; c < KA; ++c)
            if (c < nc) {
              Qs[i * ldq + j0 + c] = acc[c];
              qg[c] = acc[c];
            }
        } else {
          const int a0 = (g - njg) * KA, nc = m - a0 < KA ? m - a0 : KA;
          for (int t = 0; t < q; ++t) {
            const double sv = Si[cc[t] * LD + i];
#pragma unroll
            for (int c = 0; c < KA; ++c) acc[c] = __fma_rn(sv, Lc[(a0 + (c < nc ? c : 0)) * qmax + t], acc[c]);
          }
          double* __restrict__ yg = YL_g + ((size_t)n * bmax + i) * d.ldY + a0;
#pragma unroll
          for (int c = 0; c < KA; ++c)
            if (c < nc) {
              YLs[i * mmax + a0 + c] = acc[c];
              yg[c] = acc[c];
            }
        }
      }
      for (int i = tid; i < sn * ldq; i += nth) P_g[rc[KF_REC_POFF] + i] = s.PA[i];
      for (int r = tid >> 5; r < b; r += nth / 32)
        for (int c = tid & 31; c < d.ldS; c += 32) Sinv_g[((size_t)n * bmax + r) * d.ldS + c] = c < b ? Si[r * LD + c] : 0.0;
    }
    __syncthreads();
    KF_TICK(2);
    // ---- phase B: border Schur complement Gacc[:an, :an] += P_n' Q_n[sup];  carry = L_n YL_n[cc] (m x m);
    // rows of P_{n+1} that come from this block: -(YL_n[sup])' P_n (m x an).  KB adjacent outputs per item.
    {
      constexpr int KB = 6;
      const int gG = (an + KB - 1) / KB, gC = (m + KB - 1) / KB;
      const int nG = an * gG, nC = m * gC, nP = m * gG;
      for (int o = tid; o < nG + nC + nP; o += nth) {
        double acc[KB];
#pragma unroll
        for (int c = 0; c < KB; ++c) acc[c] = 0.0;
        if (o < nG) {
          const int j1 = o / gG, j2 = (o - j1 * gG) * KB, nc = an - j2 < KB ? an - j2 : KB;
          double g0[KB];
#pragma unroll
          for (int c = 0; c < KB; ++c) g0[c] = c < nc ? Gacc[j1 * nb + j2 + c] : 0.0;
          for (int t = 0; t < sn; ++t) {
            const double pv = s.PA[t * ldq + j1];
            const double* __restrict__ qr = Qs + sup[t] * ldq + j2;
#pragma unroll
            for (int c = 0; c < KB; ++c) acc[c] = __fma_rn(pv, qr[c < nc ? c : 0], acc[c]);
          }
#pragma unroll
          for (int c = 0; c < KB; ++c)
            if (c < nc) Gacc[j1 * nb + j2 + c] = g0[c] + acc[c];
        } else if (o < nG + nC) {
          const int i = o - nG;
          const int a = i / gC, c0 = (i - a * gC) * KB, nc = m - c0 < KB ? m - c0 : KB;
          for (int t = 0; t < q; ++t) {
            const double lv = Lc[a * qmax + t];
            const double* __restrict__ yr = YLs + cc[t] * mmax + c0;
#pragma unroll
            for (int c = 0; c < KB; ++c) acc[c] = __fma_rn(lv, yr[c < nc ? c : 0], acc[c]);
          }
#pragma unroll
          for (int c = 0; c < KB; ++c)
            if (c < nc) s.carry[a * mmax + c0 + c] = acc[c];
        } else {
          const int i = o - nG - nC;
          const int a = i / gG, j0 = (i - a * gG) * KB, nc = an - j0 < KB ? an - j0 : KB;
          for (int t = 0; t < sn; ++t) {
            const double yv = YLs[sup[t] * mmax + a];
            const double* __restrict__ pr = s.PA + t * ldq + j0;
#pragma unroll
            for (int c = 0; c < KB; ++c) acc[c] = __fma_rn(yv, pr[c < nc ? c : 0], acc[c]);
          }
#pragma unroll
          for (int c = 0; c < KB; ++c)
            if (c < nc) s.PB[a * ldq + j0 + c] = -acc[c];
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");    // this thread's copies for block n + 1 have landed
    __syncthreads();
    // ---- the next block: zero, scatter the staged values, subtract the Schur carry
    if (n + 1 < N) {
      zero_for(n + 1);
      __syncthreads();
      scatter(n + 1, m, an, ldq, crs);
      __syncthreads();
      for (int i = tid; i < m * m; i += nth) {
        const int a = i / m, c = i - a * m;
        s.M0[cr[a] * LD + cr[c]] -= s.carry[a * mmax + c];
      }
    }
    __syncthreads();
    KF_TICK(3);
  }
#ifdef RB_KF_PROFILE
  if (tid == 0 && p == 0)
    printf("kf cycles: assemble %lld invert %lld phaseA %lld phaseB+next %lld\n", tk[0], tk[1], tk[2], tk[3]);
#endif

  // ------------------------------------------------------------------------------ border
  if (nb > 0) {
    const int32_t* __restrict__ unkb = d.unk + d.blk_ptr[N];
    for (int i = tid; i < nb * LD; i += nth) s.M0[i] = 0.0;
    __syncthreads();
    for (int e = tid; e < d.n_bG; e += nth) {
      const int pos = d.bG_pos[e];
      const int r = pos / nb, c = pos - r * nb;
      s.M0[r * LD + c] = kkt_val(v, d.bG_src[e]);
    }
    for (int e = tid; e < RB_KF_VPAD; e += nth) s.shift[e] = e < nb ? kkt_diag(v, unkb[e], d.nw) : 0.0;
    __syncthreads();
    for (int i = tid; i < nb * nb; i += nth) {
      const int r = i / nb, c = i - r * nb;
      s.M0[r * LD + c] -= Gacc[i];
    }
    const int res = kf_sym_invert(s.M0, s.M1, s.shift, LD, nb, s.vec, s.stat);
    const double* __restrict__ Gi = res ? s.M1 : s.M0;
    for (int i = tid; i < nb * nb; i += nth) {
      const int r = i / nb, c = i - r * nb;
      bt.SB[(size_t)p * nb * nb + i] = Gi[r * LD + c];
    }
  }
  __syncthreads();
  if (tid == 0 && bt.status) {
    bt.status[2 * p] = s.stat[0] + (bt.status_add ? bt.status_add[2 * p] : 0);
    bt.status[2 * p + 1] = s.stat[1] + (bt.status_add ? bt.status_add[2 * p + 1] : 0);
  }
}

// One right-hand side with the stored factors.  Steps 0..N-1: forward over blocks 0..N-1 (S^-1, YL, P staged by TMA);
// steps N..2N-1: backward over blocks N-1..0 (YL, and Q in the place of S^-1).  P / Q blocks larger than their part of the
// stage buffer (the first / last block of the chain, where the loop closure enters) are read from global memory directly.
__global__ void __launch_bounds__(RB_KS_THREADS)
kkt_solve_kernel(const RbKktDev d, const RbKktChainBatch bt) {
  extern __shared__ __align__(128) double ks_smem[];
  __shared__ __align__(8) unsigned long long bars[RB_KS_STAGES];
  const int p = blockIdx.x;
  if (p >= bt.B) return;
  const int tid = threadIdx.x;
  const int nk = d.nw + d.ng, bmax = d.bmax, mmax = d.mmax, N = d.N, nb = d.nb;
  const int nbb = nb > bmax ? nb : bmax;
  const int SZ_S = bmax * d.ldS, SZ_Y = bmax * d.ldY, SZ_PQ = bt.pq_stage;
  const int BUF = SZ_S + SZ_Y + SZ_PQ;
  double* buf0 = ks_smem;
  double* y = ks_smem + (size_t)RB_KS_STAGES * BUF;   // [nbb] current block rhs
  double* xn = y + nbb;                               // [nbb] z of this block (forward) / x of the block above (backward)
  double* rc = xn + nbb;                              // [mmax] carry
  double* xb = rc + mmax;                             // [nb] border solution
  double* racc = xb + nb;                             // [nb] P' z accumulated over the chain
  const double* __restrict__ rhs = bt.rhs + (size_t)p * nk;
  double* __restrict__ sol = bt.sol + (size_t)p * nk;
  const int pf = bt.inst ? bt.inst[p] : p;            // factor slot
  const double* __restrict__ Sinv_g = bt.Sinv + (size_t)pf * N * SZ_S;
  const double* __restrict__ YL_g = bt.YL + (size_t)pf * N * SZ_Y;
  const double* __restrict__ P_g = bt.P + (size_t)pf * bt.p_total;
  const double* __restrict__ Q_g = bt.Q + (size_t)pf * bt.q_total;
  double* __restrict__ Xr = bt.Xr + (size_t)pf * N * bmax;

  auto bulk = [&](double* dst, const double* src, unsigned doubles, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(doubles * 8u), "r"(bar)
                 : "memory");
  };
  auto issue = [&](int step) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[step % RB_KS_STAGES]);
    double* dst = buf0 + (size_t)(step % RB_KS_STAGES) * BUF;
    if (step < N) {
      const int n = step;
      const int pq = d.p_off[n + 1] - d.p_off[n];
      const bool st_p = pq > 0 && pq <= SZ_PQ;
      const bool with_y = n < N - 1;
      const unsigned bytes = (unsigned)(SZ_S + (with_y ? SZ_Y : 0) + (st_p ? pq : 0)) * 8u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      bulk(dst, Sinv_g + (size_t)n * SZ_S, SZ_S, bar);
      if (with_y) bulk(dst + SZ_S, YL_g + (size_t)n * SZ_Y, SZ_Y, bar);
      if (st_p) bulk(dst + SZ_S + SZ_Y, P_g + d.p_off[n], pq, bar);
    } else {
      const int n = 2 * N - 1 - step;
      const int qq = d.q_off[n + 1] - d.q_off[n];
      const bool st_q = qq > 0 && qq <= SZ_S;            // Q_n takes the place of S_n^-1 in the stage buffer
      const bool with_y = n < N - 1;
      const unsigned bytes = (unsigned)((with_y ? SZ_Y : 0) + (st_q ? qq : 0)) * 8u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      if (with_y) bulk(dst + SZ_S, YL_g + (size_t)n * SZ_Y, SZ_Y, bar);
      if (st_q) bulk(dst, Q_g + d.q_off[n], qq, bar);
    }
  };
  auto wait = [&](int step) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[step % RB_KS_STAGES]);
    const unsigned parity = (step / RB_KS_STAGES) & 1;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KKTS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KKTS_DONE_%=;\n"
        "bra KKTS_WAIT_%=;\n"
        "KKTS_DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
  };
  if (tid == 0) {
    for (int i = 0; i < RB_KS_STAGES; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < nb; i += blockDim.x) racc[i] = 0.0;
  __syncthreads();
  const int nsteps = 2 * N;
  if (tid == 0)
    for (int s0 = 0; s0 < RB_KS_STAGES && s0 < nsteps; ++s0) issue(s0);
  const int grp = tid >> 3, part = tid & 7;          // eight threads per row of a product
  // The table entries of a step (block extent, coupling / support extents, border columns, panel offsets) are loaded
  // one step ahead: a dozen dependent L2 reads per step on the critical path of a 43 x 43 product otherwise.
  struct Step { int u0, b, cr0, m, sup0, sn, an, po, pq; };
  auto load_step = [&](int n, bool fwd) {
    Step t;
    t.u0 = d.blk_ptr[n];
    t.b = d.blk_ptr[n + 1] - t.u0;
    t.cr0 = d.cr_ptr[n];
    t.m = n < N - 1 ? d.cr_ptr[n + 1] - t.cr0 : 0;
    t.sup0 = d.sup_ptr[n];
    t.sn = d.sup_ptr[n + 1] - t.sup0;
    t.an = d.act[n];
    t.po = fwd ? d.p_off[n] : d.q_off[n];
    t.pq = (fwd ? d.p_off[n + 1] : d.q_off[n + 1]) - t.po;
    return t;
  };
  int m_prev = 0;
  int cr_prev_t = 0;        // this thread's coupling row of the previous block (its carry entry lands there)
  Step nx = load_step(0, true);
  // the right-hand-side entries of the next block are gathered one step ahead (blocks have at most 64 unknowns)
  double y_next = (tid < nx.b) ? rhs[d.unk[nx.u0 + tid]] : 0.0;
  for (int n = 0; n < N; ++n) {
    const int step = n;
    const Step c = nx;
    const int b = c.b;
    if (n + 1 < N) nx = load_step(n + 1, true);
    if (tid < b) y[tid] = y_next;
    // support rows of this block (for the border product below) and its coupling rows (for the carry of the next
    // step), one per thread: index loads off the critical path
    const int sup_t = tid < c.sn ? d.sup[c.sup0 + tid] : 0;
    const int cr_t = tid < c.m ? d.cr[c.cr0 + tid] : 0;
    __syncthreads();
    if (n + 1 < N && tid < nx.b) y_next = rhs[d.unk[nx.u0 + tid]];
    if (n > 0) {
      if (tid < m_prev) y[cr_prev_t] -= rc[tid];
      __syncthreads();
    }
    wait(step);
    const double* __restrict__ Sn = buf0 + (size_t)(step % RB_KS_STAGES) * BUF;
    const double* __restrict__ Yn = Sn + SZ_S;
    // z = S^-1 y   (loop bounds are uniform over the block: every lane takes part in the shuffles)
    for (int ib = 0; ib < b; ib += RB_KS_THREADS / 8) {
      const int i = ib + grp;
      double acc = 0.0;
      if (i < b)
        for (int j = part; j < b; j += 8) acc += Sn[i * d.ldS + j] * y[j];
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (part == 0 && i < b) {
        Xr[(size_t)n * bmax + i] = acc;
        xn[i] = acc;
      }
    }
    if (n < N - 1) {
      const int m = c.m;
      // L_n S_n^-1 y_n = YL_n' y_n   (S_n is symmetric)
      for (int ab = 0; ab < m; ab += RB_KS_THREADS / 8) {
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
      cr_prev_t = cr_t;
    }
    __syncthreads();                                   // z of this block is complete
    // z on the support rows, in the place of y (y is rebuilt at the top of the next step)
    if (tid < c.sn) y[tid] = xn[sup_t];
    __syncthreads();
    {
      // r_b -= P_n' z_n[sup]: eight threads per border column
      const int sn = c.sn, an = c.an;
      const int ldq = (an + 1) / 2 * 2;
      const double* __restrict__ Pn = (c.pq <= SZ_PQ) ? Yn + SZ_Y : P_g + c.po;
      for (int jb = 0; jb < an; jb += RB_KS_THREADS / 8) {
        const int j = jb + grp;
        double acc = 0.0;
        if (j < an)
          for (int t = part; t < sn; t += 8) acc += Pn[t * ldq + j] * y[t];
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (part == 0 && j < an) racc[j] += acc;
      }
    }
    __syncthreads();                                   // this buffer, y and xn are free again
    if (tid == 0 && step + RB_KS_STAGES < nsteps) issue(step + RB_KS_STAGES);
  }
  // ---- border
  if (nb > 0) {
    const int32_t* __restrict__ unkb = d.unk + d.blk_ptr[N];
    for (int j = tid; j < nb; j += blockDim.x) y[j] = rhs[unkb[j]] - racc[j];
    __syncthreads();
    for (int i = tid; i < nb; i += blockDim.x) {
      double acc = 0.0;
      for (int j = 0; j < nb; ++j) acc += bt.SB[(size_t)pf * nb * nb + i * nb + j] * y[j];
      xb[i] = acc;
      sol[unkb[i]] = acc;
    }
    __syncthreads();
  }
  // ---- backward: x_n = z_n - Q_n x_b - YL_n x_{n+1}[cr]
  nx = load_step(N - 1, false);
  double z_next = (tid < nx.b) ? Xr[(size_t)(N - 1) * bmax + tid] : 0.0;
  int cr_nx = tid < nx.m ? d.cr[nx.cr0 + tid] : 0;         // coupling row / unknown index of the next step, per thread
  int unk_nx = tid < nx.b ? d.unk[nx.u0 + tid] : 0;
  for (int n = N - 1; n >= 0; --n) {
    const int step = 2 * N - 1 - n;
    const Step c = nx;
    const int b = c.b;
    const int cr_t = cr_nx, unk_t = unk_nx;
    if (n > 0) {
      nx = load_step(n - 1, false);
      cr_nx = tid < nx.m ? d.cr[nx.cr0 + tid] : 0;
      unk_nx = tid < nx.b ? d.unk[nx.u0 + tid] : 0;
    }
    const double z_cur = z_next;
    const int m = c.m;
    if (tid < m) rc[tid] = xn[cr_t];
    if (tid < b) y[tid] = z_cur;
    wait(step);
    __syncthreads();
    if (n > 0 && tid < nx.b) z_next = Xr[(size_t)(n - 1) * bmax + tid];
    const double* __restrict__ Yn = buf0 + (size_t)(step % RB_KS_STAGES) * BUF + SZ_S;
    const int an = c.an;
    const int ldq = (an + 1) / 2 * 2;
    const double* __restrict__ Qn = (c.pq <= SZ_S) ? Yn - SZ_S : Q_g + c.po;
    for (int ib = 0; ib < b; ib += RB_KS_THREADS / 8) {
      const int i = ib + grp;
      double acc = 0.0;
      if (i < b) {
        for (int j = part; j < an; j += 8) acc += Qn[i * ldq + j] * xb[j];
        for (int a = part; a < m; a += 8) acc += Yn[i * d.ldY + a] * rc[a];
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (part == 0 && i < b) y[i] -= acc;
    }
    __syncthreads();
    if (tid < b) {
      const double xv = y[tid];
      xn[tid] = xv;
      sol[unk_t] = xv;
    }
    if (tid == 0 && step + RB_KS_STAGES < nsteps) issue(step + RB_KS_STAGES);
    __syncthreads();
  }
}
