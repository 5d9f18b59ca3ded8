// libraceline_b200: C ABI + kernel dispatch (see include/raceline_b200.h).
//
// Replaces, for the raceline NLP, what CasADi's nlpsol derives from the SX graph the reference
// builds (drone3d/raceline/base_raceline.py:752-799) and evaluates on every IPOPT iteration
// (:160-165): nlp_f, nlp_g, nlp_grad_f, nlp_jac_g, nlp_hess_l.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/raceline_b200.h"
#include "common.cuh"
#include "generated/pf_drone_quat_global.cuh"
#include "generated/pf_drone_quat_param_gr.cuh"
#include "generated/pf_drone_ypr_global.cuh"
#include "generated/pf_drone_ypr_param_gr.cuh"
#include "generated/pf_point_pm_global.cuh"
#include "generated/pf_point_pm_param_gr.cuh"
#include "generated/pf_drone_quat_param_lr.cuh"
#include "generated/pf_drone_ypr_param_lr.cuh"
#include "generated/pf_point_pm_param_lr.cuh"
#include "generated/pf_drone_quat_global_drag.cuh"
#include "generated/pf_drone_quat_param_gr_drag.cuh"
#include "rk4_cells2.cuh"
#include "colloc_cells.cuh"
#include "colloc_gather.cuh"
#include "simple_rows.cuh"
#include "tail_tape.cuh"
#include "kkt_blocks.cuh"
#include "kkt_big.cuh"
#include "kkt_chain.cuh"
#include "kkt_condense.cuh"
#include "mesh_sdf.cuh"
#include "ipm_glue.cuh"
#include "warm_start.cuh"
#include "post.cuh"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

// optional instrumentation: CUDA-event pairs around the interval-cell kernel (the dominant one)
struct CellTimer {
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  size_t used = 0;
} g_timer;
std::mutex g_timer_mu;

int fail(const std::string& m) {
  g_err = m;
  return 1;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return 1;
}
#define CK(call)                                        \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

struct VariantInfo {
  const char* name;
  int nz, nu, nvp;
  int rk4_ns;   // scratch doubles per shooting cell (Rk4Scratch<PF>::NS)
  int rk4_cpb;  // cells per CTA of the direction kernel (scratch group size)
  int nj, nwh, quat;   // non-zeros of df/dx and of sum_i lam_i d2f_i/dx2 per point; quaternion orientation
};
#ifndef RB_RK4_CHUNK
#define RB_RK4_CHUNK 128
#endif
//   // instances per launch pair of the shooting kernels (bounds the scratch: NS * 8 B per cell)
const VariantInfo kVariants[] = {
    {"drone_quat_global", PF_drone_quat_global::NZ, PF_drone_quat_global::NU, PF_drone_quat_global::NVP, Rk4Scratch<PF_drone_quat_global>::NS, PF_drone_quat_global::CPB, PF_drone_quat_global::NJ, PF_drone_quat_global::NW, PF_drone_quat_global::QUAT ? 1 : 0},
    {"drone_quat_param_gr", PF_drone_quat_param_gr::NZ, PF_drone_quat_param_gr::NU, PF_drone_quat_param_gr::NVP, Rk4Scratch<PF_drone_quat_param_gr>::NS, PF_drone_quat_param_gr::CPB, PF_drone_quat_param_gr::NJ, PF_drone_quat_param_gr::NW, PF_drone_quat_param_gr::QUAT ? 1 : 0},
    {"point_pm_global", PF_point_pm_global::NZ, PF_point_pm_global::NU, PF_point_pm_global::NVP, Rk4Scratch<PF_point_pm_global>::NS, PF_point_pm_global::CPB, PF_point_pm_global::NJ, PF_point_pm_global::NW, PF_point_pm_global::QUAT ? 1 : 0},
    {"point_pm_param_gr", PF_point_pm_param_gr::NZ, PF_point_pm_param_gr::NU, PF_point_pm_param_gr::NVP, Rk4Scratch<PF_point_pm_param_gr>::NS, PF_point_pm_param_gr::CPB, PF_point_pm_param_gr::NJ, PF_point_pm_param_gr::NW, PF_point_pm_param_gr::QUAT ? 1 : 0},
    {"drone_ypr_param_gr", PF_drone_ypr_param_gr::NZ, PF_drone_ypr_param_gr::NU, PF_drone_ypr_param_gr::NVP, Rk4Scratch<PF_drone_ypr_param_gr>::NS, PF_drone_ypr_param_gr::CPB, PF_drone_ypr_param_gr::NJ, PF_drone_ypr_param_gr::NW, PF_drone_ypr_param_gr::QUAT ? 1 : 0},
    {"drone_ypr_global", PF_drone_ypr_global::NZ, PF_drone_ypr_global::NU, PF_drone_ypr_global::NVP, Rk4Scratch<PF_drone_ypr_global>::NS, PF_drone_ypr_global::CPB, PF_drone_ypr_global::NJ, PF_drone_ypr_global::NW, PF_drone_ypr_global::QUAT ? 1 : 0},
#define RB_VARIANT(PF, name) {name, PF::NZ, PF::NU, PF::NVP, Rk4Scratch<PF>::NS, PF::CPB, PF::NJ, PF::NW, PF::QUAT ? 1 : 0}
    RB_VARIANT(PF_drone_quat_param_lr, "drone_quat_param_lr"),
    RB_VARIANT(PF_drone_ypr_param_lr, "drone_ypr_param_lr"),
    RB_VARIANT(PF_point_pm_param_lr, "point_pm_param_lr"),
    RB_VARIANT(PF_drone_quat_global_drag, "drone_quat_global_drag"),
    RB_VARIANT(PF_drone_quat_param_gr_drag, "drone_quat_param_gr_drag"),
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

// Per-device launch state: the opt-in for large dynamic shared memory is a per-device function attribute, and the side
// streams / events of the chunked shooting launches belong to the device they were created on.  One entry per CUDA
// device, created on first use under a mutex (launches on one device from several host threads share the entry;
// stream-ordered work is serialised by the events, the entry itself is immutable after creation).
constexpr int kMaxDevices = 64;
struct DeviceLaunchState {
  bool ready = false;
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_in = nullptr, ev_out[2] = {nullptr, nullptr};
};
DeviceLaunchState g_dev_state[kMaxDevices];
std::mutex g_dev_mu;

cudaError_t device_state(DeviceLaunchState** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(g_dev_mu);
  DeviceLaunchState& s = g_dev_state[dev];
  if (!s.ready) {
    for (int i = 0; i < 2; ++i) {
      e = cudaStreamCreateWithFlags(&s.side[i], cudaStreamNonBlocking);
      if (e != cudaSuccess) return e;
      e = cudaEventCreateWithFlags(&s.ev_out[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    e = cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    s.ready = true;
  }
  *out = &s;
  return cudaSuccess;
}

// largest dynamic shared memory size configured so far for kernel instantiation `Tag`, per device
template <class Tag>
cudaError_t ensure_dyn_smem(const void* func, size_t bytes) {
  static size_t configured[kMaxDevices] = {};
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(mu);
  if (bytes > configured[dev]) {
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    configured[dev] = bytes;
  }
  return cudaSuccess;
}
template <class PF> struct DirTag {};
template <class PF> struct CollocTag {};
struct GatherTag {};

// RB_COLLOC_IMAGE=1 in the environment selects the round-1 shared-memory-image kernel (colloc_cells.cuh) for A/B runs
bool colloc_image_path() {
  static const bool on = [] {
    const char* e = std::getenv("RB_COLLOC_IMAGE");
    return e && e[0] == '1';
  }();
  return on;
}

template <class PF>
cudaError_t launch_cells_t(const RbDev& d, const RbBatch& b, const CgTables* cg, cudaStream_t st) {
  const long long cells = (long long)b.B * d.N;
  if (d.transcription == RB_RK4) {
    constexpr int CPB = RB_CELL_THREADS / (PF::NX + 1);
    if (!b.cell_scr) return cudaErrorInvalidValue;
    const size_t dir_smem = Rk4Scratch<PF>::dir_smem_bytes(RB_CELL_THREADS);
    cudaError_t e0 = ensure_dyn_smem<DirTag<PF>>((const void*)rk4_dir_kernel<PF>, dir_smem);
    if (e0 != cudaSuccess) return e0;
    // Chunks alternate between two side streams and the two halves of the scratch, so that the point kernel of
    // chunk i+1 (bound by its scratch writes) overlaps the direction kernel of chunk i (bound by arithmetic).
    DeviceLaunchState* ds = nullptr;
    e0 = device_state(&ds);
    if (e0 != cudaSuccess) return e0;
    cudaStream_t* side = ds->side;
    cudaEvent_t ev_in = ds->ev_in;
    cudaEvent_t* ev_out = ds->ev_out;
    const int nchunk = (b.B + RB_RK4_CHUNK - 1) / RB_RK4_CHUNK;
    const bool fork = nchunk > 1;
    const size_t half = Rk4Scratch<PF>::doubles((long long)RB_RK4_CHUNK * d.N);
    if (fork) {
      cudaEventRecord(ev_in, st);
      cudaStreamWaitEvent(side[0], ev_in, 0);
      cudaStreamWaitEvent(side[1], ev_in, 0);
    }
    int ci = 0;
    for (int p0 = 0; p0 < b.B; p0 += RB_RK4_CHUNK, ++ci) {
      RbBatch c = b;
      c.B = b.B - p0 < RB_RK4_CHUNK ? b.B - p0 : RB_RK4_CHUNK;
      c.x += (size_t)p0 * d.nw;
      if (c.lam_g) c.lam_g += (size_t)p0 * d.ng;
      if (c.lam_f) c.lam_f += p0;
      c.vp += (size_t)p0 * b.vp_stride;
      if (c.fc_b) c.fc_b += (size_t)p0 * d.N * PF::NFC;
      if (c.grad_f) c.grad_f += (size_t)p0 * d.nw;
      if (c.g) c.g += (size_t)p0 * d.ng;
      if (c.jac) c.jac += (size_t)p0 * d.nnzj;
      if (c.hess) c.hess += (size_t)p0 * d.nnzh;
      if (c.fpart) c.fpart += (size_t)p0 * d.N;
      const long long cc = (long long)c.B * d.N;
      cudaStream_t cs = fork ? side[ci & 1] : st;
      double* scr = b.cell_scr + (fork ? (size_t)(ci & 1) * half : 0);
      rk4_point_kernel<PF><<<(unsigned)((cc + 127) / 128), 128, 0, cs>>>(d, c, scr);
      rk4_dir_kernel<PF><<<(unsigned)((cc + CPB - 1) / CPB), RB_CELL_THREADS, dir_smem, cs>>>(d, c, scr);
      g_launches += 2;
    }
    if (fork) {
      for (int i = 0; i < 2; ++i) {
        cudaEventRecord(ev_out[i], side[i]);
        cudaStreamWaitEvent(st, ev_out[i], 0);
      }
    }
    return cudaGetLastError();
  } else if (cg && b.cell_scr && !colloc_image_path()) {
    // output-driven form (colloc_gather.cuh): point functions of a chunk of instances, then CTA (n, g) per interval
    const size_t smem = cg_smem_bytes(PF::NZ, PF::NU, PF::NJ, PF::NW, cg->lstride[0], cg->lstride[1], cg->lptr[0][CG_NKIND],
                                      cg->lptr[1][CG_NKIND]);
    cudaError_t e0 = ensure_dyn_smem<GatherTag>((const void*)colloc_gather_kernel, smem);
    if (e0 != cudaSuccess) return e0;
    for (int p0 = 0; p0 < b.B; p0 += RB_CG_CHUNK) {
      RbBatch c = b;
      c.B = b.B - p0 < RB_CG_CHUNK ? b.B - p0 : RB_CG_CHUNK;
      c.x += (size_t)p0 * d.nw;
      if (c.lam_g) c.lam_g += (size_t)p0 * d.ng;
      if (c.lam_f) c.lam_f += p0;
      c.vp += (size_t)p0 * b.vp_stride;
      if (c.fc_b) c.fc_b += (size_t)p0 * d.N * (d.K + 1) * PF::NFC;
      if (c.grad_f) c.grad_f += (size_t)p0 * d.nw;
      if (c.g) c.g += (size_t)p0 * d.ng;
      if (c.jac) c.jac += (size_t)p0 * d.nnzj;
      if (c.hess) c.hess += (size_t)p0 * d.nnzh;
      if (c.fpart) c.fpart += (size_t)p0 * d.N;
      const long long cc = (long long)c.B * d.N;
      colloc_point_kernel<PF><<<(unsigned)((cc * 7 + 127) / 128), 128, 0, st>>>(d, c, b.cell_scr);
      // one CTA per cell: ~10 CTAs per SM overlap each other's staging latency
      const int G = c.B;
      colloc_gather_kernel<<<dim3((unsigned)d.N, (unsigned)G), RB_CG_THREADS, smem, st>>>(d, c, *cg, b.cell_scr);
      g_launches += 2;
    }
    return cudaGetLastError();
  } else {
    const long long blocks = (cells + RB_COLLOC_WPB - 1) / RB_COLLOC_WPB;
    const size_t smem = sizeof(double) * RB_COLLOC_WPB * (size_t)colloc_cell_doubles<PF>(d.cell_nj, d.cell_nh);
    cudaError_t e0 = ensure_dyn_smem<CollocTag<PF>>((const void*)colloc_cells_kernel<PF>, smem);
    if (e0 != cudaSuccess) return e0;
    colloc_cells_kernel<PF><<<(unsigned)blocks, RB_COLLOC_WPB * 32, smem, st>>>(d, b);
  }
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_cells(int variant, const RbDev& d, const RbBatch& b, const CgTables* cg, cudaStream_t st) {
  switch (variant) {
    case 0: return launch_cells_t<PF_drone_quat_global>(d, b, cg, st);
    case 1: return launch_cells_t<PF_drone_quat_param_gr>(d, b, cg, st);
    case 2: return launch_cells_t<PF_point_pm_global>(d, b, cg, st);
    case 3: return launch_cells_t<PF_point_pm_param_gr>(d, b, cg, st);
    case 4: return launch_cells_t<PF_drone_ypr_param_gr>(d, b, cg, st);
    case 5: return launch_cells_t<PF_drone_ypr_global>(d, b, cg, st);
    case 6: return launch_cells_t<PF_drone_quat_param_lr>(d, b, cg, st);
    case 7: return launch_cells_t<PF_drone_ypr_param_lr>(d, b, cg, st);
    case 8: return launch_cells_t<PF_point_pm_param_lr>(d, b, cg, st);
    case 9: return launch_cells_t<PF_drone_quat_global_drag>(d, b, cg, st);
    case 10: return launch_cells_t<PF_drone_quat_param_gr_drag>(d, b, cg, st);
  }
  return cudaErrorInvalidValue;
}

struct Workspace {
  int B = 0;
  double *x = nullptr, *lam_g = nullptr, *lam_f = nullptr, *vp = nullptr;
  double *f = nullptr, *grad_f = nullptr, *g = nullptr, *jac = nullptr, *hess = nullptr;
  void* scratch = nullptr;
  void* scratch2 = nullptr;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev_done2 = nullptr;
};

}  // namespace

struct rb_problem {
  RbDev d{};
  int variant = -1;
  int nz = 0, nu = 0, nvp = 0;
  std::vector<void*> owned;  // device allocations
  std::vector<long long> jac_sp, hess_sp;
  Workspace ws;
  std::mutex mu;
  CgTables cg{};            // collocation: per-entry recipes of the output-driven kernel (colloc_gather.cuh)
  bool has_cg = false;
  RbTail tail{};            // open tracks: tape of the end rows (tail_tape.cuh)
  bool has_tail = false;
};

namespace {

template <class T>
int upload(rb_problem* p, const T* src, size_t n, const T** dst) {
  *dst = nullptr;
  if (!src || n == 0) return 0;
  void* dev = nullptr;
  CK(cudaMalloc(&dev, n * sizeof(T)));
  p->owned.push_back(dev);
  CK(cudaMemcpy(dev, src, n * sizeof(T), cudaMemcpyHostToDevice));
  *dst = static_cast<const T*>(dev);
  return 0;
}

void free_ws(Workspace& w) {
  for (double* q : {w.x, w.lam_g, w.lam_f, w.vp, w.f, w.grad_f, w.g, w.jac, w.hess})
    if (q) cudaFree(q);
  if (w.scratch) cudaFree(w.scratch);
  if (w.scratch2) cudaFree(w.scratch2);
  w = Workspace{.stream = w.stream, .stream2 = w.stream2, .ev_done2 = w.ev_done2};
}

int ensure_ws(rb_problem* p, int B) {
  Workspace& w = p->ws;
  if (!w.stream) {
    CK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&w.stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&w.ev_done2, cudaEventDisableTiming));
  }
  if (w.B >= B) return 0;
  free_ws(w);
  const RbDev& d = p->d;
  auto alloc = [&](double** q, size_t n) -> cudaError_t { return cudaMalloc((void**)q, n * sizeof(double)); };
  CK(alloc(&w.x, (size_t)B * d.nw));
  CK(alloc(&w.lam_g, (size_t)B * (d.ng > 0 ? d.ng : 1)));
  CK(alloc(&w.lam_f, (size_t)B));
  CK(alloc(&w.vp, (size_t)B * p->nvp));
  CK(alloc(&w.f, (size_t)B));
  CK(alloc(&w.grad_f, (size_t)B * d.nw));
  CK(alloc(&w.g, (size_t)B * (d.ng > 0 ? d.ng : 1)));
  CK(alloc(&w.jac, (size_t)B * (d.nnzj > 0 ? d.nnzj : 1)));
  CK(alloc(&w.hess, (size_t)B * (d.nnzh > 0 ? d.nnzh : 1)));
  CK(cudaMalloc(&w.scratch, rb_eval_scratch_bytes(p, B)));
  CK(cudaMalloc(&w.scratch2, rb_eval_scratch_bytes(p, B)));
  w.B = B;
  return 0;
}

}  // namespace


// per-entry recipes of colloc_gather_kernel from the slot -> unique-entry template (structure_colloc.py: CollocLayout,
// _template): every local contribution slot has a fixed formula; the contributions of a unique entry are listed together
static int build_gather_tables(rb_problem* p, const rb_problem_desc* s) {
  const VariantInfo& vi = kVariants[p->variant];
  const int nz = p->nz, nu = p->nu, nx = nz + nu, S = nz + 2 * nu, NJ = vi.nj, NW = vi.nwh, KP = 8;
  const int JS = 0, JF = JS + KP * 9, JC = JF + 7 * NJ, JH = JC + 7 * nz * KP, JD = JH + 7 * nz, JE = JD + KP * nu * 10,
            JEP = JE + nz * KP * 4, JEU = JEP + nz, JEUP = JEU + nu * KP, NJS = JEUP + nu;
  const int HW = 0, HUU = HW + 7 * NW, HDD = HUU + KP * nu, HHZ = HDD + KP * nu, HHU = HHZ + KP * nz, HHD = HHU + KP * nu,
            HHH = HHD + KP * nu, HQ = HHH + 1, NHS = HQ + 32 * 32;
  if (s->n_tmpl_j != NJS || s->n_tmpl_h != NHS) return fail("collocation template sizes do not match the model variant");
  const double *C = s->colloc_C, *D = s->colloc_D, *B = s->colloc_B;
  struct Rec { int ka; double c; };
  std::vector<Rec> jr(NJS, Rec{-1, 0.0}), hr(NHS, Rec{-1, 0.0});
  auto ka = [](int kind, int arg) { return (kind << 24) | arg; };
  for (int k = 0; k < KP; ++k) {
    for (int j = 0; j < KP; ++j) jr[JS + k * 9 + j] = Rec{ka(CG_HI, 0), C[j * KP + k]};
    jr[JS + k * 9 + 8] = Rec{ka(CG_PHI2, k), -1.0};
  }
  for (int k = 1; k < KP; ++k) {
    for (int e = 0; e < NJ; ++e) jr[JF + (k - 1) * NJ + e] = Rec{ka(CG_SCR, (nz + e) * 7 + (k - 1)), 0.0};
    for (int i = 0; i < nz; ++i) {
      for (int j = 0; j < KP; ++j) jr[JC + ((k - 1) * nz + i) * KP + j] = Rec{ka(CG_HI, 0), -C[j * KP + k]};
      jr[JH + (k - 1) * nz + i] = Rec{ka(CG_PHI2, 8 + (k - 1) * nz + i), 1.0};
    }
  }
  for (int k = 0; k < KP; ++k)
    for (int j = 0; j < nu; ++j) {
      const int t = k * nu + j, b0 = JD + t * 10;
      for (int m = 0; m < KP; ++m) jr[b0 + m] = Rec{ka(CG_HI, 0), -C[m * KP + k]};
      jr[b0 + 8] = Rec{ka(CG_CONST, 0), 1.0};
      jr[b0 + 9] = Rec{ka(CG_PHI2, 8 + 7 * nz + t), 1.0};
    }
  for (int c = 0; c < nz; ++c) {
    const bool isq = vi.quat && c >= 3 && c < 7;
    for (int k = 0; k < KP; ++k) {
      if (isq) {
        for (int bb = 0; bb < 4; ++bb) jr[JE + (c * KP + k) * 4 + bb] = Rec{ka(CG_ENDQ, c | ((c - 3) << 8) | (bb << 12)), D[k]};
      } else {
        jr[JE + (c * KP + k) * 4] = Rec{ka(CG_END, c), D[k]};
      }
    }
    jr[JEP + c] = Rec{ka(CG_PARTNER, c), 0.0};
  }
  for (int j = 0; j < nu; ++j) {
    for (int k = 0; k < KP; ++k) jr[JEU + j * KP + k] = Rec{ka(CG_END, nz + j), D[k]};
    jr[JEUP + j] = Rec{ka(CG_PARTNER, nz + j), 0.0};
  }
  for (int k = 1; k < KP; ++k)
    for (int e = 0; e < NW; ++e) hr[HW + (k - 1) * NW + e] = Rec{ka(CG_SCR, (nz + NJ + e) * 7 + (k - 1)), 0.0};
  for (int k = 0; k < KP; ++k)
    for (int j = 0; j < nu; ++j) {
      const int t = k * nu + j;
      hr[HUU + t] = Rec{ka(CG_SIGH, 0), 2.0 * s->R[j] * B[k]};
      hr[HDD + t] = Rec{ka(CG_SIGH, 0), 2.0 * s->dR[j] * B[k]};
      hr[HHD + t] = Rec{ka(CG_SIGX, 1 + k * S + nx + j), 2.0 * s->dR[j] * B[k]};
      hr[HHU + t] = Rec{ka(CG_HHU, (1 + k * S + nz + j) | (t << 12)), 2.0 * s->R[j] * B[k]};
    }
  for (int j = 0; j < KP; ++j)
    for (int i = 0; i < nz; ++i) hr[HHZ + j * nz + i] = Rec{ka(CG_QHI2, 8 * nu + j * nz + i), 0.0};
  hr[HHH] = Rec{ka(CG_HHH, 0), 0.0};
  for (int pq = 0; pq < 32; ++pq)
    for (int qq = 0; qq < 32; ++qq)
      hr[HQ + pq * 32 + qq] = Rec{ka(CG_HQ, (pq & 3) | ((qq & 3) << 4)), D[pq >> 2] * D[qq >> 2]};
  CgTables& t = p->cg;
  t.nz = nz;
  t.nu = nu;
  t.NJ = NJ;
  t.NW = NW;
  t.quat = vi.quat;
  // One list for both outputs: per kind the Jacobian entries, then the Hessian entries (flag bit 23 of the argument word)
  struct Ent { int u, ka, h; double c; };
  std::vector<std::vector<int>> byJ(s->cell_nj), byH(s->cell_nh);
  for (int sidx = 0; sidx < NJS; ++sidx)
    if (s->tmpl_j[sidx] >= 0) {
      if (s->tmpl_j[sidx] >= s->cell_nj || jr[sidx].ka < 0) return fail("collocation template: entry without a recipe");
      byJ[s->tmpl_j[sidx]].push_back(sidx);
    }
  for (int sidx = 0; sidx < NHS; ++sidx)
    if (s->tmpl_h[sidx] >= 0) {
      if (s->tmpl_h[sidx] >= s->cell_nh || hr[sidx].ka < 0) return fail("collocation template: entry without a recipe");
      byH[s->tmpl_h[sidx]].push_back(sidx);
    }
  std::vector<Ent> single, multi;
  std::vector<int32_t> mptr(1, 0), mka;
  std::vector<double> mc;
  for (int kind = 0; kind < CG_NKIND; ++kind) {
    t.lptr[0][kind] = (int)single.size();
    for (int u = 0; u < s->cell_nj; ++u)
      if (byJ[u].size() == 1 && (jr[byJ[u][0]].ka >> 24) == kind) single.push_back(Ent{u, jr[byJ[u][0]].ka, 0, jr[byJ[u][0]].c});
    for (int u = 0; u < s->cell_nh; ++u)
      if (byH[u].size() == 1 && (hr[byH[u][0]].ka >> 24) == kind) single.push_back(Ent{u, hr[byH[u][0]].ka, 1, hr[byH[u][0]].c});
  }
  t.lptr[0][CG_NKIND] = (int)single.size();
  for (int h = 0; h < 2; ++h) {
    const auto& by = h ? byH : byJ;
    const auto& rec = h ? hr : jr;
    for (int u = 0; u < (int)by.size(); ++u)
      if (by[u].size() > 1) {
        multi.push_back(Ent{u, 0, h, 0.0});
        for (int sidx : by[u]) {
          mka.push_back(rec[sidx].ka);
          mc.push_back(rec[sidx].c);
        }
        mptr.push_back((int32_t)mka.size());
      }
  }
  t.nmulti[0] = (int)multi.size();
  t.lstride[0] = (int)(single.size() + multi.size());
  for (int k = 0; k <= CG_NKIND; ++k) t.lptr[1][k] = 0;
  t.nmulti[1] = 0;
  t.lstride[1] = 0;
  {
    std::vector<int4> recs((size_t)s->N * t.lstride[0]);
    auto slot_of = [&](int n, const Ent& e) {
      return e.h ? s->cell_hslot[(size_t)n * s->cell_nh + e.u] : s->cell_jslot[(size_t)n * s->cell_nj + e.u];
    };
    for (int n = 0; n < s->N; ++n) {
      int4* row = recs.data() + (size_t)n * t.lstride[0];
      for (size_t i = 0; i < single.size(); ++i) {
        long long bits;
        std::memcpy(&bits, &single[i].c, sizeof(bits));
        row[i] = make_int4(slot_of(n, single[i]), single[i].ka | (single[i].h << 23), (int)(bits & 0xffffffffLL), (int)(bits >> 32));
      }
      for (size_t i = 0; i < multi.size(); ++i) row[single.size() + i] = make_int4(slot_of(n, multi[i]), multi[i].h << 23, 0, 0);
    }
    int rc = 0;
    rc |= upload(p, recs.data(), recs.size(), &t.rec[0]);
    rc |= upload(p, mptr.data(), mptr.size(), &t.mptr[0]);
    rc |= upload(p, mka.data(), mka.size(), &t.mka[0]);
    rc |= upload(p, mc.data(), mc.size(), &t.mc[0]);
    if (rc) return 1;
  }
  p->has_cg = true;
  return 0;
}

static void cas_forget(rb_problem* p);

extern "C" {

const char* rb_last_error(void) { return g_err.c_str(); }

int rb_device_count(int* count) {
  CK(cudaGetDeviceCount(count));
  return 0;
}

int rb_set_device(int device) {
  CK(cudaSetDevice(device));
  return 0;
}

long long rb_launch_count(void) { return g_launches.load(); }

int rb_problem_create(const rb_problem_desc* s, rb_problem** out) {
  if (!s || !out) return fail("rb_problem_create: null argument");
  *out = nullptr;
  int variant = -1;
  for (int i = 0; i < kNumVariants; ++i)
    if (s->variant && std::strcmp(s->variant, kVariants[i].name) == 0) variant = i;
  if (variant < 0) return fail(std::string("model variant not compiled into libraceline_b200: ") + (s->variant ? s->variant : "(null)"));
  if (s->transcription != RB_RK4 && s->transcription != RB_COLLOC) return fail("unknown transcription");
  if (s->transcription == RB_COLLOC && s->K != 7) return fail("collocation kernels are built for K = 7");
  auto* p = new rb_problem();
  p->variant = variant;
  p->nz = kVariants[variant].nz;
  p->nu = kVariants[variant].nu;
  p->nvp = kVariants[variant].nvp;
  RbDev& d = p->d;
  d.transcription = s->transcription;
  d.N = s->N;
  d.K = s->K;
  d.nw = s->nw;
  d.ng = s->ng;
  d.nnzj = s->nnz_jac;
  d.nnzh = s->nnz_hess;
  d.cell_nj = s->cell_nj;
  d.cell_nh = s->cell_nh;
  d.cell_ncp = s->cell_ncp;
  d.n_srow = s->n_srow;
  d.n_shess = s->n_shess;
  const int P = s->K + 1;
  const int S = p->nz + 2 * p->nu;
  if (s->nw != s->N + s->N * P * S) {
    delete p;
    return fail("nw does not match N + N*(K+1)*(nz+2nu)");
  }
  int rc = 0;
  rc |= upload(p, s->R, (size_t)p->nu, &d.R);
  rc |= upload(p, s->dR, (size_t)p->nu, &d.dR);
  rc |= upload(p, s->fc, (size_t)s->N * P * 13, &d.fc);
  const size_t nrow_tab = s->transcription == RB_RK4 ? (size_t)s->N * (p->nz + p->nu) : (size_t)s->N * RB_COLLOC_NCR(p->nz, p->nu);
  rc |= upload(p, s->cell_row, nrow_tab, &d.cell_row);
  rc |= upload(p, s->cell_coef, nrow_tab, &d.cell_coef);
  rc |= upload(p, s->cell_partner, nrow_tab, &d.cell_partner);
  rc |= upload(p, s->cell_pcoef, nrow_tab, &d.cell_pcoef);
  rc |= upload(p, s->cell_off, nrow_tab, &d.cell_off);
  rc |= upload(p, s->cell_par, (size_t)s->N * s->cell_ncp, &d.cell_par);
  rc |= upload(p, s->cell_jslot, (size_t)s->N * s->cell_nj, &d.cell_jslot);
  rc |= upload(p, s->cell_hslot, (size_t)s->N * s->cell_nh, &d.cell_hslot);
  if (s->transcription == RB_COLLOC) {
    if (!s->tmpl_j || !s->tmpl_h || !s->colloc_C || !s->colloc_D || !s->colloc_B) {
      rb_problem_destroy(p);
      return fail("collocation problem without template / coefficient tables");
    }
    rc |= upload(p, s->tmpl_j, (size_t)s->n_tmpl_j, &d.tmpl_j);
    rc |= upload(p, s->tmpl_h, (size_t)s->n_tmpl_h, &d.tmpl_h);
    rc |= upload(p, s->colloc_C, (size_t)P * P, &d.colloc_C);
    rc |= upload(p, s->colloc_D, (size_t)P, &d.colloc_D);
    rc |= upload(p, s->colloc_B, (size_t)P, &d.colloc_B);
    if (!rc && build_gather_tables(p, s)) {
      rb_problem_destroy(p);
      return 1;
    }
  }
  if (s->n_srow > 0) {
    rc |= upload(p, s->srow_row, (size_t)s->n_srow, &d.srow_row);
    rc |= upload(p, s->srow_kind, (size_t)s->n_srow, &d.srow_kind);
    rc |= upload(p, s->srow_scale, (size_t)s->n_srow, &d.srow_scale);
    rc |= upload(p, s->srow_var_ptr, (size_t)s->n_srow + 1, &d.srow_var_ptr);
    rc |= upload(p, s->srow_form_ptr, (size_t)s->n_srow + 1, &d.srow_form_ptr);
    rc |= upload(p, s->srow_coef_ptr, (size_t)s->n_srow + 1, &d.srow_coef_ptr);
    const size_t nvar = (size_t)s->srow_var_ptr[s->n_srow];
    rc |= upload(p, s->srow_var, nvar, &d.srow_var);
    rc |= upload(p, s->srow_jslot, nvar, &d.srow_jslot);
    rc |= upload(p, s->srow_A, (size_t)s->srow_coef_ptr[s->n_srow], &d.srow_A);
    rc |= upload(p, s->srow_c, (size_t)s->srow_form_ptr[s->n_srow], &d.srow_c);
  }
  if (s->n_shess > 0) {
    rc |= upload(p, s->shess_slot, (size_t)s->n_shess, &d.shess_slot);
    rc |= upload(p, s->shess_add, (size_t)s->n_shess, &d.shess_add);
    rc |= upload(p, s->shess_ptr, (size_t)s->n_shess + 1, &d.shess_ptr);
    const size_t ne = (size_t)s->shess_ptr[s->n_shess];
    rc |= upload(p, s->shess_row, ne, &d.shess_row);
    rc |= upload(p, s->shess_coef, ne, &d.shess_coef);
    rc |= upload(p, s->shess_scale, ne, &d.shess_scale);
  }
  if (rc) {
    rb_problem_destroy(p);
    return 1;
  }
  auto pack = [&](std::vector<long long>& v, int nrow, int ncol, const int64_t* colind, const int64_t* row, int nnz) {
    v.clear();
    v.push_back(nrow);
    v.push_back(ncol);
    for (int i = 0; i <= ncol; ++i) v.push_back(colind ? colind[i] : 0);
    for (int i = 0; i < nnz; ++i) v.push_back(row ? row[i] : 0);
  };
  pack(p->jac_sp, s->ng, s->nw, s->jac_colind, s->jac_row, s->nnz_jac);
  pack(p->hess_sp, s->nw, s->nw, s->hess_colind, s->hess_row, s->nnz_hess);
  *out = p;
  return 0;
}

void rb_problem_destroy(rb_problem* p) {
  cas_forget(p);   // the CasADi-style symbols must not keep a dangling problem
  if (!p) return;
  for (void* q : p->owned) cudaFree(q);
  cudaStream_t st = p->ws.stream, st2 = p->ws.stream2;
  cudaEvent_t ev = p->ws.ev_done2;
  free_ws(p->ws);
  if (st) cudaStreamDestroy(st);
  if (st2) cudaStreamDestroy(st2);
  if (ev) cudaEventDestroy(ev);
  delete p;
}

int rb_problem_nvp(const rb_problem* p) { return p ? p->nvp : 0; }

int rb_sparsity_size(const rb_problem* p, int which, size_t* n) {
  if (!p || !n) return fail("rb_sparsity_size: null argument");
  *n = which == 0 ? p->jac_sp.size() : p->hess_sp.size();
  return 0;
}

int rb_sparsity_get(const rb_problem* p, int which, long long* out) {
  if (!p || !out) return fail("rb_sparsity_get: null argument");
  const auto& v = which == 0 ? p->jac_sp : p->hess_sp;
  std::memcpy(out, v.data(), v.size() * sizeof(long long));
  return 0;
}

int rb_problem_set_tail(rb_problem* p, const rb_tail_desc* s) {
  if (!p || !s) return fail("rb_problem_set_tail: null argument");
  if (s->n_ins <= 0 || !s->ins || !s->lvl_ptr || !s->cval || s->n_slots <= 0 || s->n_levels[2] <= 0)
    return fail("rb_problem_set_tail: empty tape");
  if (s->n_levels[0] > s->n_levels[1] || s->n_levels[1] > s->n_levels[2] || s->lvl_ptr[s->n_levels[2]] != s->n_ins)
    return fail("rb_problem_set_tail: inconsistent level table");
  if (tail_smem_bytes(s->n_slots, s->n_levels[2]) + 16 > 200 * 1024) return fail("rb_problem_set_tail: tape needs more work slots than one CTA's shared memory holds");
  for (int i = 0; i < s->n_ins; ++i) {
    const int32_t* q = s->ins + 4 * (size_t)i;
    const bool store = q[0] >= T_STORE_G;
    const bool load = q[0] <= T_LOADLAM;
    bool ok = q[0] >= 0 && q[0] <= T_ADD_H;
    if (ok && store) ok = q[1] >= 0 && q[1] < s->n_slots && q[2] >= 0 &&
                          q[2] < (q[0] == T_STORE_G ? p->d.ng : q[0] == T_STORE_J ? p->d.nnzj : p->d.nnzh);
    if (ok && !store) ok = q[3] >= 0 && q[3] < s->n_slots;
    if (ok && load) ok = q[1] >= 0 && q[1] < (q[0] == T_CONST ? s->n_const : q[0] == T_LOADX ? p->d.nw : q[0] == T_LOADVP ? p->nvp : p->d.ng);
    if (ok && !store && !load) ok = q[1] >= 0 && q[1] < s->n_slots && q[2] >= 0 && q[2] < s->n_slots;
    if (!ok) return fail("rb_problem_set_tail: instruction " + std::to_string(i) + " out of range");
  }
  const int32_t* ins = nullptr;
  int rc = upload(p, s->ins, (size_t)s->n_ins * 4, &ins);
  rc |= upload(p, s->lvl_ptr, (size_t)s->n_levels[2] + 1, &p->tail.lvl_ptr);
  rc |= upload(p, s->cval, (size_t)(s->n_const > 0 ? s->n_const : 1), &p->tail.cval);
  if (rc) return 1;
  p->tail.ins = reinterpret_cast<const int4*>(ins);
  p->tail.n_slots = s->n_slots;
  for (int k = 0; k < 3; ++k) p->tail.n_levels[k] = s->n_levels[k];
  CK(cudaFuncSetAttribute(tail_tape_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  p->has_tail = true;
  return 0;
}

static size_t fpart_bytes(const rb_problem* p, int B) {
  return (((size_t)B * p->d.N * sizeof(double)) + 255) / 256 * 256;
}

size_t rb_eval_scratch_bytes(const rb_problem* p, int B) {
  if (!p || B <= 0) return 0;
  size_t n = fpart_bytes(p, B) + 256;
  if (p->d.transcription == RB_RK4) {
    const long long cells = (long long)(B < RB_RK4_CHUNK ? B : RB_RK4_CHUNK) * p->d.N;
    const int cpb = kVariants[p->variant].rk4_cpb;
    // two halves: consecutive chunks run on two streams
    n += 2 * (size_t)((cells + cpb - 1) / cpb) * kVariants[p->variant].rk4_ns * cpb * sizeof(double);
  } else if (p->has_cg) {
    const long long cells = (long long)(B < RB_CG_CHUNK ? B : RB_CG_CHUNK) * p->d.N;
    n += cg_scratch_doubles(cells, p->nz, kVariants[p->variant].nj, kVariants[p->variant].nwh) * sizeof(double);
  }
  return n;
}

int rb_eval_batch(const rb_problem* p, int B, const double* x, const double* lam_g, const double* lam_f,
                  const double* vp, int vp_stride, const double* fc_b, double* f, double* grad_f,
                  double* g, double* jac, double* hess, void* scratch, void* stream) {
  if (!p) return fail("rb_eval_batch: null problem");
  if (B <= 0) return 0;     // an empty batch is a no-op
  if (!x || !vp) return fail("rb_eval_batch: null x / vp");
  if (hess && !lam_g && p->d.ng > 0) return fail("rb_eval_batch: hess requested without lam_g");
  if ((f || p->d.transcription == RB_RK4) && !scratch) return fail("rb_eval_batch: scratch buffer required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RbBatch b{};
  b.B = B;
  b.x = x;
  b.lam_g = lam_g;
  b.lam_f = lam_f;
  b.vp = vp;
  b.vp_stride = vp_stride;
  b.fc_b = fc_b;
  b.f = f;
  b.grad_f = grad_f;
  b.g = g;
  b.jac = jac;
  b.hess = hess;
  b.fpart = f ? static_cast<double*>(scratch) : nullptr;
  b.cell_scr = scratch ? reinterpret_cast<double*>(static_cast<char*>(scratch) + fpart_bytes(p, B)) : nullptr;
  const RbDev& d = p->d;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (g_timer.on) {
    std::lock_guard<std::mutex> lock(g_timer_mu);
    if (g_timer.used == g_timer.ev.size()) {
      cudaEvent_t a, c;
      CK(cudaEventCreate(&a));
      CK(cudaEventCreate(&c));
      g_timer.ev.emplace_back(a, c);
    }
    ev0 = g_timer.ev[g_timer.used].first;
    ev1 = g_timer.ev[g_timer.used].second;
    g_timer.used++;
    CK(cudaEventRecord(ev0, st));
  }
  CK(launch_cells(p->variant, d, b, p->has_cg ? &p->cg : nullptr, st));
  if (ev1) CK(cudaEventRecord(ev1, st));
  if (d.n_srow > 0 && (g || jac)) {
    const long long t = (long long)B * d.n_srow;
    simple_rows_kernel<<<(unsigned)((t + 127) / 128), 128, 0, st>>>(d, b);
    g_launches++;
    CK(cudaGetLastError());
  }
  if (d.n_shess > 0 && hess) {
    const long long t = (long long)B * d.n_shess;
    simple_hess_kernel<<<(unsigned)((t + 127) / 128), 128, 0, st>>>(d, b);
    g_launches++;
    CK(cudaGetLastError());
  }
  if (p->has_tail && (g || jac || hess)) {
    const int nl = p->tail.n_levels[hess ? 2 : jac ? 1 : 0];
    tail_tape_kernel<<<(unsigned)B, RB_TAIL_THREADS, tail_smem_bytes(p->tail.n_slots, nl) + 16, st>>>(p->tail, d, b, nl);
    g_launches++;
    CK(cudaGetLastError());
  }
  if (f) {
    objective_sum_kernel<<<(unsigned)(((long long)B * 32 + 127) / 128), 128, 0, st>>>(d, b, d.N);
    g_launches++;
    CK(cudaGetLastError());
  }
  return 0;
}

int rb_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_timer_mu);
  g_timer.on = on != 0;
  g_timer.used = 0;
  return 0;
}

int rb_profile_cell_ms(double* total_ms, int* launches) {
  std::lock_guard<std::mutex> lock(g_timer_mu);
  double tot = 0.0;
  for (size_t i = 0; i < g_timer.used; ++i) {
    CK(cudaEventSynchronize(g_timer.ev[i].second));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, g_timer.ev[i].first, g_timer.ev[i].second));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int)g_timer.used;
  g_timer.used = 0;
  return 0;
}

// FP64 FMA throughput of this device (the denominator for the compute-bound shooting kernel)
__global__ void fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int rb_fp64_peak(double* tflops) {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 1 << 16;
  double* buf = nullptr;
  CK(cudaMalloc(&buf, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  fp64_peak_kernel<<<blocks, threads>>>(buf, 1024);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, threads>>>(buf, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (fl > best) best = fl;
  }
  g_launches += 4;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  if (tflops) *tflops = best;
  return 0;
}

// ---- host-buffer entry points -------------------------------------------------------------------
static int host_eval(const rb_problem* cp, int B, const double* x, const double* vp, const double* lam_f,
                     const double* lam_g, double* f, double* grad_f, double* g, double* jac, double* hess) {
  if (!cp) return fail("null problem");
  if (B <= 0) return 0;     // an empty batch is a no-op
  if (!x || !vp) return fail("null x / vp");
  rb_problem* p = const_cast<rb_problem*>(cp);
  std::lock_guard<std::mutex> lock(p->mu);
  if (ensure_ws(p, B)) return 1;
  Workspace& w = p->ws;
  const RbDev& d = p->d;
  if (hess && !lam_g && d.ng > 0) return fail("nlp_hess_l needs lam_g");
  // Chunks of instances alternate between two streams: the device->host copies of one chunk (the bulk of the
  // bytes: jac_g and hess_l values) overlap the host->device copies and kernels of the next one.
  const int CH = 64;
  const size_t nw = d.nw, ng = d.ng, nj = d.nnzj, nh = d.nnzh, nvp = p->nvp;
  int ci = 0;
  for (int p0 = 0; p0 < B; p0 += CH, ++ci) {
    const int Bc = B - p0 < CH ? B - p0 : CH;
    cudaStream_t st = (ci & 1) ? w.stream2 : w.stream;
    void* scr = (ci & 1) ? w.scratch2 : w.scratch;
    CK(cudaMemcpyAsync(w.x + p0 * nw, x + p0 * nw, Bc * nw * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w.vp + p0 * nvp, vp + p0 * nvp, Bc * nvp * sizeof(double), cudaMemcpyHostToDevice, st));
    const double* dl = nullptr;
    const double* dlf = nullptr;
    if (hess) {
      if (ng > 0) CK(cudaMemcpyAsync(w.lam_g + p0 * ng, lam_g + p0 * ng, Bc * ng * sizeof(double), cudaMemcpyHostToDevice, st));
      dl = w.lam_g + p0 * ng;
      if (lam_f) {
        CK(cudaMemcpyAsync(w.lam_f + p0, lam_f + p0, Bc * sizeof(double), cudaMemcpyHostToDevice, st));
        dlf = w.lam_f + p0;
      }
    }
    if (rb_eval_batch(p, Bc, w.x + p0 * nw, dl, dlf, w.vp + p0 * nvp, p->nvp, nullptr, f ? w.f + p0 : nullptr,
                      grad_f ? w.grad_f + p0 * nw : nullptr, g ? w.g + p0 * ng : nullptr, jac ? w.jac + p0 * nj : nullptr,
                      hess ? w.hess + p0 * nh : nullptr, scr, st))
      return 1;
    if (f) CK(cudaMemcpyAsync(f + p0, w.f + p0, Bc * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (grad_f) CK(cudaMemcpyAsync(grad_f + p0 * nw, w.grad_f + p0 * nw, Bc * nw * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (g) CK(cudaMemcpyAsync(g + p0 * ng, w.g + p0 * ng, Bc * ng * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (jac) CK(cudaMemcpyAsync(jac + p0 * nj, w.jac + p0 * nj, Bc * nj * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (hess) CK(cudaMemcpyAsync(hess + p0 * nh, w.hess + p0 * nh, Bc * nh * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(w.stream));
  CK(cudaStreamSynchronize(w.stream2));
  return 0;
}

int rb_nlp_f(const rb_problem* p, int B, const double* x, const double* vp, double* f) {
  return host_eval(p, B, x, vp, nullptr, nullptr, f, nullptr, nullptr, nullptr, nullptr);
}
int rb_nlp_g(const rb_problem* p, int B, const double* x, const double* vp, double* g) {
  return host_eval(p, B, x, vp, nullptr, nullptr, nullptr, nullptr, g, nullptr, nullptr);
}
int rb_nlp_grad_f(const rb_problem* p, int B, const double* x, const double* vp, double* f, double* grad_f) {
  return host_eval(p, B, x, vp, nullptr, nullptr, f, grad_f, nullptr, nullptr, nullptr);
}
int rb_nlp_jac_g(const rb_problem* p, int B, const double* x, const double* vp, double* g, double* jac) {
  return host_eval(p, B, x, vp, nullptr, nullptr, nullptr, nullptr, g, jac, nullptr);
}
int rb_nlp_hess_l(const rb_problem* p, int B, const double* x, const double* vp, const double* lam_f,
                  const double* lam_g, double* hess) {
  return host_eval(p, B, x, vp, lam_f, lam_g, nullptr, nullptr, nullptr, nullptr, hess);
}
int rb_nlp_eval_all(const rb_problem* p, int B, const double* x, const double* vp, const double* lam_f,
                    const double* lam_g, double* f, double* grad_f, double* g, double* jac, double* hess) {
  return host_eval(p, B, x, vp, lam_f, lam_g, f, grad_f, g, jac, hess);
}

// signed distance (positive outside) and closest point of np points to a triangle mesh; device pointers
int rb_mesh_sdf(const double* tri, int nt, const double* pts, int np, double* dist, double* closest, void* stream) {
  if (!tri || !pts || !dist) return fail("rb_mesh_sdf: null argument");
  if (np <= 0) return 0;
  if (nt <= 0) return fail("rb_mesh_sdf: empty mesh");
  mesh_sdf_kernel<<<(unsigned)((np + RB_SDF_THREADS - 1) / RB_SDF_THREADS), RB_SDF_THREADS, 0,
                    static_cast<cudaStream_t>(stream)>>>(tri, nt, pts, np, dist, closest);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

// frame constants of many path lengths on T tracks (csrc/post.cuh); `tabs` is a DEVICE array of T rb_spline_tab records
static_assert(sizeof(rb_spline_tab) == sizeof(RbSplineTab), "rb_spline_tab mirrors RbSplineTab");
int rb_centerline_frames(const rb_spline_tab* tabs, int T, const double* s, int M, int s_stride, double* fc, double* xc,
                         const double* yn, double* xg, void* stream) {
  if (!tabs || !s || !fc) return fail("rb_centerline_frames: null argument");
  if ((yn == nullptr) != (xg == nullptr)) return fail("rb_centerline_frames: yn and xg go together");
  if (T <= 0 || M <= 0) return 0;
  const long long t = (long long)T * M;
  centerline_frames_kernel<<<(unsigned)((t + RB_POST_THREADS - 1) / RB_POST_THREADS), RB_POST_THREADS, 0,
                             static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const RbSplineTab*>(tabs), T, s, M,
                                                                  s_stride, fc, xc, yn, xg);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

// trajectory interpolants of B solutions at M times each (csrc/post.cuh); tp is a [B][N+1] workspace
int rb_traj_interp(const double* w, int B, int N, int P, int S, const double* tau, const double* D, const double* tq, int M,
                   int tq_stride, double* tp, double* out, void* stream) {
  if (!w || !tq || !tp || !out) return fail("rb_traj_interp: null argument");
  if (P > 1 && (!tau || !D)) return fail("rb_traj_interp: collocation nodes / end weights missing");
  if (P < 1 || P > 16) return fail("rb_traj_interp: 1 <= K + 1 <= 16");
  if (B <= 0 || M <= 0) return 0;
  const int nw = N + N * P * S;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  traj_times_kernel<<<(B + 63) / 64, 64, 0, st>>>(w, B, N, nw, tp);
  CK(cudaGetLastError());
  const long long t = (long long)B * M;
  traj_interp_kernel<<<(unsigned)((t + RB_POST_THREADS - 1) / RB_POST_THREADS), RB_POST_THREADS, 0, st>>>(
      w, B, N, P, S, nw, tp, tau, D, tq, M, tq_stride, out);
  g_launches += 2;
  CK(cudaGetLastError());
  return 0;
}

// batched warm-start chain: point-mass solutions -> drone initial guesses (csrc/warm_start.cuh); device pointers
static_assert(sizeof(rb_ws_args) == sizeof(RbWsArgs), "rb_ws_args mirrors RbWsArgs");
int rb_ws_drone_guess(const rb_ws_args* a, void* stream) {
  if (!a || !a->w_pm || !a->w_dr) return fail("rb_ws_drone_guess: null argument");
  if (a->B <= 0) return 0;
  if (!a->global_r && !a->fc) return fail("rb_ws_drone_guess: frame constants are needed when global_r is false");
  const int nz = a->quat ? 13 : 12;
  if (a->nw_pm != a->N + a->N * a->P * 12 || a->nw_dr != a->N + a->N * a->P * (nz + 8))
    return fail("rb_ws_drone_guess: decision vector lengths do not match N, K and the orientation parameters");
  RbWsArgs r;
  std::memcpy(&r, a, sizeof(RbWsArgs));
  r.use_fc = a->fc != nullptr;
  const long long t = (long long)a->B * a->N * a->P;
  ws_point_kernel<<<(unsigned)((t + RB_WS_THREADS - 1) / RB_WS_THREADS), RB_WS_THREADS, 0,
                    static_cast<cudaStream_t>(stream)>>>(r);
  CK(cudaGetLastError());
  ws_continuity_kernel<<<(a->B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(r);
  g_launches += 2;
  CK(cudaGetLastError());
  return 0;
}

// ---- fused kernels of the interior-point sweep (csrc/ipm_glue.cuh); all pointers are device pointers -------------
static_assert(sizeof(rb_ipm_args) == sizeof(RbIpm), "rb_ipm_args mirrors RbIpm");
static RbIpm ipm_args(const rb_ipm_args* a) {
  RbIpm r;
  std::memcpy(&r, a, sizeof(RbIpm));
  return r;
}
#define IPM_STREAM static_cast<cudaStream_t>(stream)
int rb_ipm_error(const rb_ipm_args* a, double* out, void* stream) {
  if (!a || !out) return fail("rb_ipm_error: null argument");
  if (a->B <= 0) return 0;
  ipm_error_kernel<<<a->B, RB_IPM_THREADS, 0, IPM_STREAM>>>(ipm_args(a), out);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}
int rb_ipm_newton(const rb_ipm_args* a, const double* f, double* dxd, double* negd, double* rhs, double* gphi_x,
                  double* gphi_s, double* c, double* r_s, double* Ssr, double* sc, void* stream) {
  if (!a || !f || !dxd || !negd || !rhs || !gphi_x || !gphi_s || !c || !r_s || !Ssr || !sc) return fail("rb_ipm_newton: null argument");
  if (a->B <= 0) return 0;
  ipm_newton_kernel<<<a->B, RB_IPM_THREADS, 0, IPM_STREAM>>>(ipm_args(a), f, dxd, negd, rhs, gphi_x, gphi_s, c, r_s, Ssr, sc);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}
int rb_ipm_direction(const rb_ipm_args* a, const double* sol, const unsigned char* moved, const double* tau,
                     const double* gphi_x, const double* gphi_s, const double* c, const double* r_s, const double* Ssr,
                     double* dx, double* dy, double* ds, double* dzL, double* dzU, double* dvL, double* dvU, double* sc,
                     void* stream) {
  if (!a || !sol || !moved || !tau || !dx || !dy || !ds || !dzL || !dzU || !dvL || !dvU || !sc) return fail("rb_ipm_direction: null argument");
  if (a->B <= 0) return 0;
  ipm_direction_kernel<<<a->B, RB_IPM_THREADS, 0, IPM_STREAM>>>(ipm_args(a), sol, moved, tau, gphi_x, gphi_s, c, r_s, Ssr, dx,
                                                                 dy, ds, dzL, dzU, dvL, dvU, sc);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}
int rb_ipm_trial(int n, int Kw, int ns, const int* rows, const double* al, const double* x, const double* dx, double* xt,
                 void* stream) {
  if (ns <= 0) return 0;
  if (!rows || !al || !x || !dx || !xt) return fail("rb_ipm_trial: null argument");
  ipm_trial_kernel<<<ns * Kw, RB_IPM_THREADS, 0, IPM_STREAM>>>(n, Kw, rows, al, x, dx, xt);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}
int rb_ipm_trial_merit(const rb_ipm_args* a, int Kw, int ns, const int* rows, const double* al, const double* xt,
                       const double* ds, const double* f_t, const double* g_t, double* out, void* stream) {
  if (ns <= 0) return 0;
  if (!a || !rows || !al || !xt || !ds || !f_t || !g_t || !out) return fail("rb_ipm_trial_merit: null argument");
  ipm_trial_merit_kernel<<<ns * Kw, RB_IPM_THREADS, 0, IPM_STREAM>>>(ipm_args(a), Kw, rows, al, xt, ds, f_t, g_t, out);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}
int rb_ipm_update(const rb_ipm_args* a, const double* alpha, const double* alpha_du, const double* dx, const double* dy,
                  const double* ds, const double* dzL, const double* dzU, const double* dvL, const double* dvU,
                  double kappa_sigma, void* stream) {
  if (!a || !alpha || !alpha_du || !dx || !dy || !ds || !dzL || !dzU || !dvL || !dvU) return fail("rb_ipm_update: null argument");
  if (a->B <= 0) return 0;
  ipm_update_kernel<<<a->B, RB_IPM_THREADS, 0, IPM_STREAM>>>(ipm_args(a), alpha, alpha_du, dx, dy, ds, dzL, dzU, dvL, dvU,
                                                              kappa_sigma);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"

#include "casadi_abi.inc"
#include "kkt_host.inc"
