// Shared device-side structures of libraceline_b200 (see include/raceline_b200.h for the ABI).
#pragma once
#include <cstdint>
#include <cstddef>

#define RB_CELL_THREADS 128

// device copies of the problem tables (rb_problem_desc), passed to kernels by value
struct RbDev {
  int transcription, N, K, nw, ng, nnzj, nnzh;
  const double* R;
  const double* dR;
  const double* fc;
  const int32_t* cell_row;
  const double* cell_coef;
  const int32_t* cell_partner;
  const double* cell_pcoef;
  const double* cell_off;
  const double* cell_par;
  const int32_t* cell_jslot;
  const int32_t* cell_hslot;
  int cell_nj, cell_nh, cell_ncp;
  // collocation only: slot -> unique-entry template and the Legendre coefficient tables
  const int32_t* tmpl_j;
  const int32_t* tmpl_h;
  const double* colloc_C;  // [8][8]  C[j][k] = l_j'(tau_k)
  const double* colloc_D;  // [8]
  const double* colloc_B;  // [8]
  int n_srow;
  const int32_t* srow_row;
  const int32_t* srow_kind;
  const int32_t* srow_scale;
  const int32_t* srow_var_ptr;
  const int32_t* srow_var;
  const int32_t* srow_jslot;
  const int32_t* srow_form_ptr;
  const int32_t* srow_coef_ptr;
  const double* srow_A;
  const double* srow_c;
  int n_shess;
  const int32_t* shess_slot;
  const int32_t* shess_add;
  const int32_t* shess_ptr;
  const int32_t* shess_row;
  const double* shess_coef;
  const int32_t* shess_scale;
};

// one batched evaluation request (all device pointers)
struct RbBatch {
  int B;
  const double* x;
  const double* lam_g;
  const double* lam_f;
  const double* vp;
  int vp_stride;
  const double* fc_b;
  double* f;
  double* grad_f;
  double* g;
  double* jac;
  double* hess;
  double* fpart;   // scratch [B][N] per-interval objective terms
  double* cell_scr;  // scratch of the shooting-cell kernels (one chunk of instances)
};
