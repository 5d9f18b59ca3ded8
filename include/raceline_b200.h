/*
 * raceline_b200.h -- C ABI of libraceline_b200.so (sm_100a).
 *
 * Drop-in boundary for the hot path of thomasfork/aircraft_trajectory_optimization: the five NLP
 * functions CasADi creates when the reference calls
 *     self.solver = ca.nlpsol('solver', 'ipopt', prob, opts)      drone3d/raceline/base_raceline.py:799
 * and that IPOPT calls on every iteration of
 *     sol = self.solver(x0=..., lbx=..., ubx=..., lbg=..., ubg=...)   drone3d/raceline/base_raceline.py:160-165
 * i.e. nlp_f, nlp_g, nlp_grad_f, nlp_jac_g, nlp_hess_l, plus batched device-pointer variants and
 * the block-structured KKT solve used by the batched interior-point driver.
 *
 * A *problem* (rb_problem) is the structured description of one raceline NLP that the Python
 * builders produce in place of the reference's SX graph (same decision-vector layout and g row
 * order: base_raceline.py:232-239, :670-717).  All index tables are int32,
 * all values fp64.  Plain pointers and sizes only; no exceptions cross this boundary.
 *
 * Return value: 0 on success, non-zero on failure (CasADi convention); rb_last_error() describes
 * the last failure on the calling thread.  NaN/Inf in outputs are not errors.
 */
#ifndef RACELINE_B200_H
#define RACELINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rb_problem rb_problem; /* opaque, owns device copies of the tables below */

/* transcription kinds */
#define RB_RK4 0     /* direct multiple shooting, classic RK4 (base_raceline.py:363-391, :1052-1112) */
#define RB_COLLOC 1  /* direct orthogonal collocation, Legendre K points (base_raceline.py:398-490) */

/* Host-side description of a problem.  Arrays are copied; the caller may free them after
 * rb_problem_create returns.  Sections that do not apply are passed as NULL / 0. */
typedef struct rb_problem_desc {
  int transcription;      /* RB_RK4 | RB_COLLOC */
  const char* variant;    /* generated model variant, e.g. "drone_quat_param_gr" */
  int N, K;               /* intervals, collocation order (K = 0 for RK4) */
  int nw, ng, nnz_jac, nnz_hess;
  const double* R;        /* [nu] diagonal of the input cost            (base_raceline.py:616-623) */
  const double* dR;       /* [nu] diagonal of the input-rate cost */
  const double* fc;       /* [N*(K+1)][13] frame constants Rp|ks|ky|kn||xcs| (NULL: global frame) */

  /* ---- interval cells: rows produced by the dynamics of interval n ---------------------- */
  const int32_t* cell_row;       /* rk4: [N][nz+nu] g row of out_c / input row j (-1: none)
                                    colloc: [N][ncr] see aircraft_trajectory_optimization_b200/structure.py */
  const double* cell_coef;       /* same shape: sign / coefficient applied to the cell output */
  const int32_t* cell_partner;   /* same shape: index in w of the linear partner variable (-1: none) */
  const double* cell_pcoef;      /* same shape: coefficient of the partner variable */
  const double* cell_off;        /* same shape: constant offset of the row */
  const double* cell_par;        /* [N][ncp] per-cell scalars (rk4: du_coef; colloc: unused) */
  const int32_t* cell_jslot;     /* flattened local-Jacobian-slot -> CCS position (-1: not stored) */
  const int32_t* cell_hslot;     /* flattened local-Hessian-slot  -> CCS position (-1: not stored) */
  int cell_nj, cell_nh;          /* local slots per cell in the two tables */
  int cell_ncp;
  /* collocation only (NULL / 0 for RK4): local contribution slot -> unique local entry id, and the
   * Legendre coefficient tables of drone3d/utils/discretization_utils.py:6-34 */
  const int32_t* tmpl_j;         /* [n_tmpl_j] */
  const int32_t* tmpl_h;         /* [n_tmpl_h] */
  int n_tmpl_j, n_tmpl_h;
  const double* colloc_C;        /* [(K+1)*(K+1)] C[j][k] = l_j'(tau_k), row-major */
  const double* colloc_D;        /* [K+1] */
  const double* colloc_B;        /* [K+1] */

  /* ---- simple rows: g = scale * sum_m (sum_i A[m][i] w[idx_i] + c[m])^(1 or 2) ------------- */
  int n_srow;
  const int32_t* srow_row;       /* [n_srow] g row */
  const int32_t* srow_kind;      /* [n_srow] 0 affine (one form), 1 sum of squares */
  const int32_t* srow_scale;     /* [n_srow] -1: scale 1; k >= 0: scale = 1 / vp[k]^2 */
  const int32_t* srow_var_ptr;   /* [n_srow+1] into srow_var / srow_jslot */
  const int32_t* srow_var;       /* w indices */
  const int32_t* srow_jslot;     /* CCS position of (row, var) */
  const int32_t* srow_form_ptr;  /* [n_srow+1] into srow_c; coefficient block of form m of row r starts at
                                    srow_coef_ptr[r] + (m - srow_form_ptr[r]) * nvar_r */
  const int32_t* srow_coef_ptr;  /* [n_srow+1] into srow_A */
  const double* srow_A;
  const double* srow_c;

  /* ---- Hessian contributions of the simple rows (constant * lam_g[row] * scale) ----------- */
  int n_shess;
  const int32_t* shess_slot;     /* [n_shess] CCS position in hess_l */
  const int32_t* shess_add;      /* [n_shess] 1: add to what the cell kernel wrote, 0: assign */
  const int32_t* shess_ptr;      /* [n_shess+1] into the three arrays below */
  const int32_t* shess_row;
  const double* shess_coef;
  const int32_t* shess_scale;

  /* ---- CCS patterns (returned by the sparsity queries; CasADi layout) ---------------------- */
  const int64_t* jac_colind;     /* [nw+1] */
  const int64_t* jac_row;        /* [nnz_jac] */
  const int64_t* hess_colind;    /* [nw+1] */
  const int64_t* hess_row;       /* [nnz_hess] upper triangle */
} rb_problem_desc;

const char* rb_last_error(void);
int rb_device_count(int* count);
int rb_set_device(int device);

int rb_problem_create(const rb_problem_desc* desc, rb_problem** out);
void rb_problem_destroy(rb_problem* p);
int rb_problem_nvp(const rb_problem* p); /* vehicle parameters per problem instance */

/* ---- end rows of OPEN racelines -------------------------------------------------
 * `_enforce_initial_constraints` / `_enforce_terminal_constraints` and the gate on the end state
 * (drone3d/raceline/base_raceline.py:516-543, :914-918, drone_raceline.py:110-148, point_raceline.py:15-45):
 * a dozen rows that are compositions of the last interval's end state.  CasADi evaluates them as part of the
 * one SX graph of the NLP; here the host differentiates their expression graph and hands the library a
 * levelised tape (aircraft_trajectory_optimization_b200/tail.py, csrc/tail_tape.cuh) that every
 * evaluation runs after the interval kernels.  Optional: closed tracks have no such rows. */
typedef struct rb_tail_desc {
  int n_ins;               /* instructions */
  const int32_t* ins;      /* [n_ins][4] = (op, a, b, dst), opcodes in csrc/tail_tape.cuh */
  int n_levels[3];         /* levels up to and including the g / Jacobian / Hessian phase */
  const int32_t* lvl_ptr;  /* [n_levels[2] + 1] first instruction of every level */
  int n_const;
  const double* cval;      /* [n_const] constants referenced by T_CONST */
  int n_slots;             /* work slots (doubles of shared memory per instance) */
} rb_tail_desc;
int rb_problem_set_tail(rb_problem* p, const rb_tail_desc* tail);

/* CasADi compressed sparsity of jac_g / hess_l: [nrow, ncol, colind[ncol+1], row[nnz]] as long long.
 * Replaces Function::sparsity_out of nlp_jac_g / nlp_hess_l [CasADi, third party]. */
int rb_sparsity_size(const rb_problem* p, int which /*0 jac_g, 1 hess_l*/, size_t* n_entries);
int rb_sparsity_get(const rb_problem* p, int which, long long* out);

/* ---- batched evaluation on DEVICE pointers (asynchronous on `stream`) ------------------------
 * One *eval* = f, grad_f, g, jac_g values, hess_l values of one problem instance at (x, lam_g, lam_f).
 *   x      [B][nw]        lam_g [B][ng]      lam_f [B]       vp [B or 1][nvp] (vp_stride 0: shared)
 *   fc_b   optional per-instance frame constants [B][N*(K+1)][13] (NULL: the problem's own)
 * outputs (any may be NULL = not wanted):
 *   f [B]   grad_f [B][nw]   g [B][ng]   jac [B][nnz_jac]   hess [B][nnz_hess]
 * scratch: device buffer of rb_eval_scratch_bytes(p, B) bytes.
 * `stream` is a cudaStream_t passed as void*. */
size_t rb_eval_scratch_bytes(const rb_problem* p, int B);
int rb_eval_batch(const rb_problem* p, int B, const double* x, const double* lam_g, const double* lam_f,
                  const double* vp, int vp_stride, const double* fc_b, double* f, double* grad_f,
                  double* g, double* jac, double* hess, void* scratch, void* stream);

/* ---- CasADi-shaped HOST entry points (synchronous; copies in and out) ------------------------
 * Same argument meaning as CasADi's nlp_* oracle functions [third party]; p (parameters) is the
 * vehicle-parameter vector.  Any output pointer may be NULL.  B instances at once (B = 1 is the
 * reference's call pattern). */
int rb_nlp_f(const rb_problem* p, int B, const double* x, const double* vp, double* f);
int rb_nlp_g(const rb_problem* p, int B, const double* x, const double* vp, double* g);
int rb_nlp_grad_f(const rb_problem* p, int B, const double* x, const double* vp, double* f, double* grad_f);
int rb_nlp_jac_g(const rb_problem* p, int B, const double* x, const double* vp, double* g, double* jac);
int rb_nlp_hess_l(const rb_problem* p, int B, const double* x, const double* vp, const double* lam_f,
                  const double* lam_g, double* hess);
/* everything in one pass: the bench's "jac_g + hess_lag eval" through host buffers */
int rb_nlp_eval_all(const rb_problem* p, int B, const double* x, const double* vp, const double* lam_f,
                    const double* lam_g, double* f, double* grad_f, double* g, double* jac, double* hess);

/* ---- CasADi generated-code / `external` shaped entry points ----------------------------------------
 * Replace the five Function objects CasADi's nlpsol derives from prob = {x, g, f, p}
 * (drone3d/raceline/base_raceline.py:752-799, called inside self.solver(...) at :160-165) with the C
 * signature CasADi's code generator emits and `ca.external(name, lib)` binds [CasADi, third party]:
 *     int F(const double** arg, double** res, long long* iw, double* w, int mem)
 * plus F_n_in/F_n_out, F_name_in/out, F_sparsity_in/out (compressed CCS [nrow, ncol, colind, row]),
 * F_default_in, F_work and the reference-counting / memory-slot functions.
 *   nlp_f(x, p) -> f;  nlp_g(x, p) -> g;  nlp_grad_f(x, p) -> (f, grad_f_x);
 *   nlp_jac_g(x, p) -> (g, jac_g_x);  nlp_hess_l(x, p, lam_f, lam_g) -> triu_hess_gamma_x_x.
 * arg[i] == NULL reads as zeros (p: the vehicle parameters given at bind time); res[i] == NULL = not
 * wanted.  CasADi externals carry no context argument: rb_casadi_bind() selects the problem they
 * evaluate (NULL unbinds); sparsity pointers stay valid until the next bind. */
int rb_casadi_bind(rb_problem* p, const double* default_vp /* [nvp] or NULL */);
#define RB_CASADI_DECLARE(NAME)                                                          \
  int NAME(const double** arg, double** res, long long* iw, double* w, int mem);         \
  long long NAME##_n_in(void);                                                           \
  long long NAME##_n_out(void);                                                          \
  const char* NAME##_name_in(long long i);                                               \
  const char* NAME##_name_out(long long i);                                              \
  double NAME##_default_in(long long i);                                                 \
  const long long* NAME##_sparsity_in(long long i);                                      \
  const long long* NAME##_sparsity_out(long long i);                                     \
  int NAME##_work(long long* sz_arg, long long* sz_res, long long* sz_iw, long long* sz_w); \
  int NAME##_alloc_mem(void);                                                            \
  int NAME##_init_mem(int mem);                                                          \
  void NAME##_free_mem(int mem);                                                         \
  int NAME##_checkout(void);                                                             \
  void NAME##_release(int mem);                                                          \
  void NAME##_incref(void);                                                              \
  void NAME##_decref(void);
RB_CASADI_DECLARE(nlp_f)
RB_CASADI_DECLARE(nlp_g)
RB_CASADI_DECLARE(nlp_grad_f)
RB_CASADI_DECLARE(nlp_jac_g)
RB_CASADI_DECLARE(nlp_hess_l)

/* ---- KKT factor / solve of the interior-point step ---------------------------------------------
 * Replaces the sparse symmetric-indefinite solve IPOPT performs on every iteration of
 *     sol = self.solver(x0=..., lbx=..., ubx=..., lbg=..., ubg=...)   drone3d/raceline/base_raceline.py:160-165
 * (linear solver MA97 / MUMPS selected at base_raceline.py:765-787) for the matrix
 *     [ W + diag(dx_diag)   J'           ]      W = hess_l values (upper-triangular CCS), J = jac_g values (CCS)
 *     [ J                   diag(neg_d)  ]
 * The unknowns are grouped into interval blocks + a dense border by the host
 * (aircraft_trajectory_optimization_b200/kkt.py documents every table).  All device pointers;
 * asynchronous on `stream`. */
typedef struct rb_kkt rb_kkt;

typedef struct rb_kkt_desc {
  int nw, ng, N, nb, bmax, mmax, qmax, nnz_hess, nnz_jac, n_bG;
  const int32_t *blk_ptr, *unk;                           /* [N+2], [nw+ng] */
  const int32_t *dA_ptr, *dA_src, *dA_pos;                /* diagonal blocks */
  const int32_t *cr_ptr, *cr, *cc_ptr, *cc;               /* coupling rows / columns */
  const int32_t *cL_ptr, *cL_src, *cL_pos;                /* coupling entries */
  const int32_t *bE_ptr, *bE_src, *bE_row, *bE_col;       /* border columns */
  const int32_t *bG_src, *bG_pos;                         /* border x border */
  /* interface form of the border columns (stage blocks that fit shared memory): active columns per block,
   * support rows of block n, position of the coupling rows / border entries inside the support, offsets of the
   * P_n / Q_n factor blocks */
  int amax, smax;
  const int32_t *act, *sup_ptr, *sup, *crs, *bE_sup, *p_off, *q_off;
  const int64_t *jac_colind, *jac_row, *hess_colind, *hess_row;   /* CCS patterns (for K v products) */
} rb_kkt_desc;

int rb_kkt_create(const rb_kkt_desc* desc, rb_kkt** out);
void rb_kkt_destroy(rb_kkt* k);

/* Collocation transcriptions (drone3d/raceline/base_raceline.py:398-490, K = 7 points per interval): the interior
 * unknowns of every interval are eliminated first, all intervals at once, and `desc` of rb_kkt_create then describes the
 * reduced (multiple-shooting-shaped) system whose values come from the aux array (source kind 4).  Tables:
 * aircraft_trajectory_optimization_b200/kkt_condensed.py.  Call once, right after rb_kkt_create. */
typedef struct rb_kkt_interior_desc {
  int NI, amax, smax, n_aux, n_rsep;
  const int32_t *iu_ptr, *iunk, *su_ptr, *sunk;            /* interior / touched separator unknowns per interval */
  const int32_t *iA_ptr, *iA_pos, *iA_src;                 /* A_n entries */
  const int32_t *iB_ptr, *iB_pos, *iB_src;                 /* B_n entries */
  const int32_t *aux_orig, *aux_c_ptr, *aux_c_idx;         /* reduced-system values: original source + T_n entries */
  const int32_t *rsep, *r_c_ptr, *r_c_idx;                 /* reduced right-hand side */
} rb_kkt_interior_desc;
int rb_kkt_set_interiors(rb_kkt* k, const rb_kkt_interior_desc* desc);
/* bytes of factor storage (block inverses, coupling solves, border columns) for a batch of B */
size_t rb_kkt_factor_bytes(const rb_kkt* k, int B);
/* factor K and solve K sol = rhs for B instances.  hess [B][nnz_hess], jac [B][nnz_jac], dx_diag [B][nw],
 * neg_d [B][ng], rhs / sol [B][nw+ng] in (w, g) order; status [B][2] (optional): vanishing pivots met, and
 * the number of negative eigenvalues of K (the interior-point driver expects exactly ng). */
int rb_kkt_factor_solve(const rb_kkt* k, int B, const double* hess, const double* jac, const double* dx_diag,
                        const double* neg_d, const double* rhs, double* sol, void* factors, int* status,
                        void* stream);
/* solve with the stored factors (the value arrays must be unchanged since rb_kkt_factor_solve) */
int rb_kkt_resolve(const rb_kkt* k, int B, const double* hess, const double* jac, const double* dx_diag,
                   const double* neg_d, const double* rhs, double* sol, void* factors, void* stream);
/* the same for a subset of the factorised instances: row p of rhs / sol belongs to factor slot inst[p] (device int array)
 * of the last rb_kkt_factor_solve call, made with B_factored instances.  Shooting-sized stage blocks only. */
int rb_kkt_resolve_rows(const rb_kkt* k, int B, int B_factored, const int* inst, const double* rhs, double* sol,
                        void* factors, void* stream);
/* out = K vec */
int rb_kkt_matvec(const rb_kkt* k, int B, const double* hess, const double* jac, const double* dx_diag,
                  const double* neg_d, const double* vec, double* out, void* stream);

/* ---- centerline frame tables and trajectory interpolants ------------------------------
 * rb_centerline_frames replaces the per-point host evaluation of the 15 `param_terms` and of the Darboux frame
 * (drone3d/centerlines/spline_centerline.py:232-322, :279-294; spline evaluation of drone3d/utils/interp.py:55-84) for M
 * path lengths on each of T tracks: fc [T][M][13] = Rp (row-major, columns es|ey|en), ks, ky, kn, |xc'| -- the
 * per-point constants rb_eval_batch takes as fc_b, so track sweeps build their tables on the device.  Optional: xc
 * [T][M][3]; yn [T][M][2] -> xg [T][M][3] = xc + ey y + en n (global position of a parametric state,
 * drone3d/dynamics/drone_models.py:306-328).  s is [T][s_stride >= M], or one row shared by all tracks (s_stride 0).
 * `tabs` is a DEVICE array of T records whose pointers are device pointers (scipy CubicSpline tables: knots, c[4][n-1][3],
 * value and slope at the last knot). */
typedef struct rb_spline_tab {
  int nkx, nkr;
  const double *kx, *cx, *ex, *kr, *cr, *er;
} rb_spline_tab;
int rb_centerline_frames(const rb_spline_tab* tabs, int T, const double* s, int M, int s_stride, double* fc, double* xc,
                         const double* yn, double* xg, void* stream);
/* rb_traj_interp replaces the interpolants _unpack_soln builds (drone3d/raceline/base_raceline.py:801-864,
 * drone3d/utils/discretization_utils.py:53-137): w [B][N + N*P*S] solutions, P = K + 1 points per interval, S values per
 * point (z, u, du); tq [B][tq_stride >= M] query times (tq_stride 0: shared); tp [B][N+1] workspace; out [B][M][S]. */
int rb_traj_interp(const double* w, int B, int N, int P, int S, const double* tau, const double* D, const double* tq, int M,
                   int tq_stride, double* tp, double* out, void* stream);

/* ---- batched warm-start chain --------------------------------------------------------------------
 * Replaces the per-point Python loop DroneRaceline._guess_z / _guess_u (drone3d/raceline/drone_raceline.py:158-277): B
 * point-mass raceline solutions [B][N + N*P*12] (decision-vector layout of base_raceline.py:681-713, state (p, v), input
 * thrust vector, input rate) become B drone initial guesses [B][N + N*P*(nz+8)] (orientation from thrust and velocity,
 * body velocity and rate, rotor thrusts |T|/4), with the quaternion sign / yaw wrap made continuous along the lap.
 * info [B][4] (optional): closure flipped (|q_first - q_last| > 1, drone_raceline.py:82-95), yaw wraps, continuity
 * failures (the reference raises NotImplementedError), reserved.  All pointers are device pointers. */
typedef struct rb_ws_args {
  int B, N, P;
  int quat, closed, global_r, reserved;
  int nw_pm, nw_dr;
  const double* w_pm;
  const double* fc;       /* [N*P][13] frame constants, needed when global_r == 0 */
  double* w_dr;
  int* info;
} rb_ws_args;
int rb_ws_drone_guess(const rb_ws_args* args, void* stream);

/* ---- fused vector kernels of the interior-point sweep ("K3" of SURVEY.md s2.1) ---------------------------------------
 * They replace the per-iteration vector work IPOPT does around its linear solves inside `self.solver(x0=...)`
 * (drone3d/raceline/base_raceline.py:160-165): optimality error, condensed KKT right-hand side, fraction-to-the-boundary
 * rule, trial points and their filter quantities, primal-dual update.  B instances, n variables, m constraint rows;
 * all pointers are DEVICE pointers; state vectors are [B][n] / [B][m].  Bounds and flags have row strides sx / ss (0: one
 * row shared by all instances).  xflag bit0 lower bound, bit1 upper bound; sflag bit0 / bit1 the same for the slack of an
 * inequality row, bit2 equality row.  resto / x_R / DR2 may be NULL (no instance in feasibility restoration). */
typedef struct rb_ipm_args {
  int B, n, m;
  long long sx, ss;
  const double *xL, *xU, *sL, *sU, *ceq;
  const unsigned char *xflag, *sflag;
  double *x, *s, *y, *zL, *zU, *vL, *vU;
  const double *grad_f, *g, *jty;
  const double *mu, *delta_w, *delta_c;
  const unsigned char* resto;
  const double *x_R, *DR2;
  double kappa_d, rho;
} rb_ipm_args;
/* out [B][8]: |grad L|_inf, |c|_inf, sum of bound multipliers, |y|_1, min / max complementarity product, |c|_1, 0 */
int rb_ipm_error(const rb_ipm_args* a, double* out, void* stream);
/* dxd [B][n], negd [B][m], rhs [B][n+m] of the condensed Newton system; gphi_x, gphi_s, c, r_s, Ssr are kept for
 * rb_ipm_direction; sc [B][4]: theta, barrier objective (needs f [B]), restoration merit, 0 */
int rb_ipm_newton(const rb_ipm_args* a, const double* f, double* dxd, double* negd, double* rhs, double* gphi_x,
                  double* gphi_s, double* c, double* r_s, double* Ssr, double* sc, void* stream);
/* search direction from the KKT solution sol [B][n+m]; sc [B][4]: alpha_pr_max, alpha_du_max, slope of the barrier
 * objective, slope of the restoration merit */
int rb_ipm_direction(const rb_ipm_args* a, const double* sol, const unsigned char* moved, const double* tau,
                     const double* gphi_x, const double* gphi_s, const double* c, const double* r_s, const double* Ssr,
                     double* dx, double* dy, double* ds, double* dzL, double* dzU, double* dvL, double* dvU, double* sc,
                     void* stream);
/* xt [ns][Kw][n] = x[rows[r]] + al[r][k] dx[rows[r]] */
int rb_ipm_trial(int n, int Kw, int ns, const int* rows, const double* al, const double* x, const double* dx, double* xt,
                 void* stream);
/* out [ns*Kw][4]: theta, barrier objective, restoration merit, finite flag at the trial points (f_t, g_t evaluated there) */
int rb_ipm_trial_merit(const rb_ipm_args* a, int Kw, int ns, const int* rows, const double* al, const double* xt,
                       const double* ds, const double* f_t, const double* g_t, double* out, void* stream);
/* x, s, y, z, v updated in place with the accepted step lengths; multipliers kept within kappa_sigma of mu / slack */
int rb_ipm_update(const rb_ipm_args* a, const double* alpha, const double* alpha_du, const double* dx, const double* dy,
                  const double* ds, const double* dzL, const double* dzU, const double* dvL, const double* dvU,
                  double kappa_sigma, void* stream);

/* Signed distance of np points to a triangle mesh, positive OUTSIDE (the convention of the reference's
 * MeshObstacle.signed_distance, drone3d/obstacles/mesh_obstacle.py:38-42, which wraps trimesh.proximity), and the closest
 * point on the mesh (optional, may be NULL) -- what the obstacle-free tube search needs (:110-145).
 * tri [nt][9] vertex triples, pts [np][3], dist [np], closest [np][3]: DEVICE pointers. */
int rb_mesh_sdf(const double* tri, int nt, const double* pts, int np, double* dist, double* closest, void* stream);

/* instrumentation: when enabled, every rb_eval_batch brackets the interval-cell kernel (the dominant
 * kernel) with CUDA events on the launching stream; rb_profile_cell_ms waits for them, returns the
 * summed duration and the number of bracketed launches, and resets the counters. */
int rb_profile_enable(int on);
int rb_profile_cell_ms(double* total_ms, int* launches);
/* measured FP64 FMA throughput of the current device in TFLOP/s (fma = 2 flop) */
int rb_fp64_peak(double* tflops);

/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
long long rb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RACELINE_B200_H */
