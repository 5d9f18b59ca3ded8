/*
 * Flat-tape interpreter for scalar expression graphs.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).
 *
 * This is the execution model of CasADi's SX functions [third party, not in the reference tree]:
 * the reference's nlp_f / nlp_g / nlp_grad_f / nlp_jac_g / nlp_hess_l (created by ca.nlpsol at
 * drone3d/raceline/base_raceline.py:799) run as one interpreted scalar instruction per graph node
 * over a work vector whose slots are reused once a value is dead.  Same here.
 *
 * Instruction i:  w[dst[i]] = op[i](w[a[i]], w[b[i]])   (CONST: consts[a[i]];  INPUT: in[a[i]])
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { OP_CONST = 0, OP_INPUT, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SQ, OP_SQRT, OP_SIN, OP_COS, OP_TAN };

/* Linear-scan slot allocation.  a/b hold instruction indices on entry and are rewritten to slot
 * numbers; dst receives the slot of each instruction.  keep[i] != 0 pins a value (function
 * output) until the end.  Returns the work-vector length. */
int sxvm_allocate(int n, const int* op, int* a, int* b, const unsigned char* keep, int* dst) {
  int* last = (int*)malloc(sizeof(int) * (size_t)n);
  int* freelist = (int*)malloc(sizeof(int) * (size_t)n);
  int nfree = 0, nslots = 0;
  for (int i = 0; i < n; ++i) last[i] = keep[i] ? n : i;
  for (int i = 0; i < n; ++i) {
    if (op[i] >= OP_ADD) {
      if (last[a[i]] < i) last[a[i]] = i;
      if (op[i] <= OP_DIV && last[b[i]] < i) last[b[i]] = i;
    }
  }
  for (int i = 0; i < n; ++i) {
    int sa = -1, sb = -1, ia = -1, ib = -1;
    if (op[i] >= OP_ADD) {
      ia = a[i]; sa = dst[ia];
      if (op[i] <= OP_DIV) { ib = b[i]; sb = dst[ib]; }
    }
    /* operands that die here free their slot before the result is placed (in-place update is safe) */
    if (ia >= 0 && last[ia] == i) freelist[nfree++] = sa;
    if (ib >= 0 && ib != ia && last[ib] == i) freelist[nfree++] = sb;
    dst[i] = nfree ? freelist[--nfree] : nslots++;
    if (op[i] >= OP_ADD) { a[i] = sa; if (op[i] <= OP_DIV) b[i] = sb; }
    if (last[i] == i) freelist[nfree++] = dst[i]; /* never used */
  }
  free(last); free(freelist);
  return nslots;
}

void sxvm_eval(int n, const int* op, const int* a, const int* b, const int* dst, const double* consts,
               const double* in, double* w, int nout, const int* out_slot, double* out) {
  for (int i = 0; i < n; ++i) {
    double r;
    switch (op[i]) {
      case OP_CONST: r = consts[a[i]]; break;
      case OP_INPUT: r = in[a[i]]; break;
      case OP_ADD: r = w[a[i]] + w[b[i]]; break;
      case OP_SUB: r = w[a[i]] - w[b[i]]; break;
      case OP_MUL: r = w[a[i]] * w[b[i]]; break;
      case OP_DIV: r = w[a[i]] / w[b[i]]; break;
      case OP_NEG: r = -w[a[i]]; break;
      case OP_SQ: r = w[a[i]] * w[a[i]]; break;
      case OP_SQRT: r = sqrt(w[a[i]]); break;
      case OP_SIN: r = sin(w[a[i]]); break;
      case OP_COS: r = cos(w[a[i]]); break;
      case OP_TAN: r = tan(w[a[i]]); break;
      default: r = NAN;
    }
    w[dst[i]] = r;
  }
  for (int k = 0; k < nout; ++k) out[k] = out_slot[k] < 0 ? 0.0 : w[out_slot[k]];
}

/* batch of independent evaluations, one problem per thread (pthreads; nthreads <= 0: all cores) */
#include <pthread.h>
#include <unistd.h>

typedef struct {
  int n; const int *op, *a, *b, *dst; const double* consts; int nin; const double* in; int nslots;
  int nout; const int* out_slot; double* out; int batch; int* next; pthread_mutex_t* mu;
} sxvm_job;

static void* sxvm_worker(void* arg) {
  sxvm_job* j = (sxvm_job*)arg;
  double* w = (double*)malloc(sizeof(double) * (size_t)(j->nslots > 0 ? j->nslots : 1));
  for (;;) {
    pthread_mutex_lock(j->mu);
    int p = (*j->next)++;
    pthread_mutex_unlock(j->mu);
    if (p >= j->batch) break;
    sxvm_eval(j->n, j->op, j->a, j->b, j->dst, j->consts, j->in + (size_t)p * j->nin, w, j->nout,
              j->out_slot, j->out + (size_t)p * j->nout);
  }
  free(w);
  return NULL;
}

int sxvm_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

void sxvm_eval_batch(int n, const int* op, const int* a, const int* b, const int* dst, const double* consts,
                     int nin, const double* in, int nslots, int nout, const int* out_slot, double* out,
                     int batch, int nthreads) {
  if (nthreads <= 0) nthreads = sxvm_max_threads();
  if (nthreads > batch) nthreads = batch;
  if (nthreads > 256) nthreads = 256;
  int next = 0;
  pthread_mutex_t mu;
  pthread_mutex_init(&mu, NULL);
  sxvm_job job = {n, op, a, b, dst, consts, nin, in, nslots, nout, out_slot, out, batch, &next, &mu};
  pthread_t th[256];
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, sxvm_worker, &job);
  sxvm_worker(&job);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  pthread_mutex_destroy(&mu);
}
