// spk_solver.cu -- handle, operator orchestration, Chambolle-Pock / SuperMann loops and the C ABI
// (include/spock_b200.h).  Reference citations are file:line under /root/reference/src.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/spock_b200.h"
#include "spk_comm.h"
#include "spk_fused.cuh"
#include "spk_kernels.cuh"
#include "spk_offline.h"
#include "spk_offline_dev.cuh"
#include "spk_peer.cuh"
#include "spk_sp.cuh"

#ifdef SPK_HOST_EMU
thread_local spk_dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local std::vector<unsigned char> spk_emu_smem;
#endif

// ------------------------------------------------------------------------------------------------ errors

static thread_local std::string g_last_error;
const char *spk_last_error(void) { return g_last_error.c_str(); }
int spk_version(void) { return 100; }

static int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

#define SPK_TRY(expr)            \
  do {                           \
    const int rc_ = (expr);      \
    if (rc_ != SPK_OK) return rc_; \
  } while (0)

// ------------------------------------------------------------------------------------------------ runtime shims

#ifndef SPK_HOST_EMU
#define RT_CHECK(call)                                                                          \
  do {                                                                                          \
    const cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                      \
      return fail(SPK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
  } while (0)
static int rt_malloc(void **p, size_t bytes) {
  const cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the (non-sticky) error: the next handle's launch checks must not see this failure
    return fail(SPK_ERR_ALLOC, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  return SPK_OK;
}
static void rt_free(void *p) { cudaFree(p); }
static int rt_h2d(void *dst, const void *src, size_t bytes, spk_stream_t s) {
  RT_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
  RT_CHECK(cudaStreamSynchronize(s));
  return SPK_OK;
}
static int rt_d2h(void *dst, const void *src, size_t bytes, spk_stream_t s) {
  RT_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
  RT_CHECK(cudaStreamSynchronize(s));
  return SPK_OK;
}
static int rt_d2d(void *dst, const void *src, size_t bytes, spk_stream_t s) {
  RT_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
  return SPK_OK;
}
static int rt_memset(void *dst, int val, size_t bytes, spk_stream_t s) {
  RT_CHECK(cudaMemsetAsync(dst, val, bytes, s));
  return SPK_OK;
}
static int rt_sync(spk_stream_t s) {
  RT_CHECK(cudaStreamSynchronize(s));
  return SPK_OK;
}
static int rt_launch_check() {
  RT_CHECK(cudaGetLastError());
  return SPK_OK;
}
#else
static int rt_malloc(void **p, size_t bytes) {
  *p = calloc(bytes ? bytes : 16, 1);
  return *p ? SPK_OK : fail(SPK_ERR_ALLOC, "calloc failed");
}
static void rt_free(void *p) { free(p); }
static int rt_h2d(void *dst, const void *src, size_t bytes, spk_stream_t) { memcpy(dst, src, bytes); return SPK_OK; }
static int rt_d2h(void *dst, const void *src, size_t bytes, spk_stream_t) { memcpy(dst, src, bytes); return SPK_OK; }
static int rt_d2d(void *dst, const void *src, size_t bytes, spk_stream_t) { memmove(dst, src, bytes); return SPK_OK; }
static int rt_memset(void *dst, int val, size_t bytes, spk_stream_t) { memset(dst, val, bytes); return SPK_OK; }
static int rt_sync(spk_stream_t) { return SPK_OK; }
static int rt_launch_check() { return SPK_OK; }
#endif

static double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------ handle

enum { PC_S1_BWD = 0, PC_S1_FWD, PC_LLT, PC_PROX_H, PC_S2, PC_VEC, PC_FUSED_BWD, PC_FUSED_FWD, PC_COUNT };
#define AA_M 3  // ANDERSON_BUFFER_SIZE (anderson.jl:6, spock.jl:108)

struct ProfSection {
  int cls;
#ifndef SPK_HOST_EMU
  cudaEvent_t e0, e1;
#else
  double t0, t1;
#endif
};

// Single-process multi-GPU (SURVEY.md section 8(b) "Threading": the Julia host is ONE process): a group handle owns one
// sub-handle per device -- rank g of the subtree partition, with its own stream and NCCL rank -- and one persistent host
// thread per device that issues every call of that sub-handle (the CUDA current device is per thread, and the ranks'
// collectives must be issued concurrently).  Every C ABI call on the group handle fans out to the workers and joins.
struct SpkWorker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, stop = false, done = false;
  int rc = 0;
  std::string err;
};
struct spk_handle;
struct SpkGroup {
  std::vector<spk_handle *> subs;
  std::vector<SpkWorker *> workers;
};

struct spk_handle {
  SpkGroup *group = nullptr;  // non-NULL: this handle is the front of a device group (only D and `group` are used)
  SpkDims D;
  SpkPtrs P;
  int alg = SPK_CP;
  int sp_options = 0;  // SPK_SP_LINESEARCH | SPK_SP_K0 (spk_set_sp_options); 0 = the reference as shipped
  double L_norm = 10.0;
  int device = 0;
  spk_stream_t stream = 0;
  std::vector<void *> allocs;
  int64_t bytes = 0;
  int64_t launches = 0;
  SpkFactorTable table;      // per-stage factors (host copy, also serves spk_get_factors)
  std::vector<double> ls_host;
  // iterate vectors: every "zv" buffer holds [z (nz) | v (nv)] contiguously
  double *zv = nullptr, *zvbar = nullptr, *zv_old = nullptr, *dzv = nullptr, *xi12 = nullptr, *zw = nullptr,
         *xi = nullptr, *ric_q = nullptr, *ric_d = nullptr, *x0 = nullptr;
  // SuperMann / Anderson (model.jl:137-226)
  double *rzv = nullptr, *rzv_old = nullptr, *drzv = nullptr, *dir = nullptr, *wu = nullptr, *wubar = nullptr,
         *rwu = nullptr, *ws_rho = nullptr, *ws_L = nullptr;
  double *MP[AA_M] = {nullptr, nullptr, nullptr}, *MR[AA_M] = {nullptr, nullptr, nullptr},
         *AQ[AA_M] = {nullptr, nullptr, nullptr};
  double aaR[AA_M * AA_M];  // column-major upper triangular (spock.jl:200)
  double res_0[2];
  double last_xi1 = NAN, last_xi2 = NAN;
  // subtree partition: levels < t_cut are replicated, rank g owns subtrees g*m_sub .. (g+1)*m_sub-1
  SpkComm comm;
  int rank = 0, world = 1, t_cut = 0;
  long m_sub = 1;
  bool custom_modes = false;   // spk_problem.mode_of_node: per-node Riccati recursion, no per-stage table
  bool scratch_ready = false;  // z_old, delta, xi_1/xi_2, xi and the modular operators' workspaces exist (ensure_scratch)
  bool in_group = false;  // sub-handle of a device group
  // per-node factors under the partition: only this rank's families are stored, packed level by level (replicated top
  // first); fac_start[t] = first packed family slot of level t
  long fac_start[SPK_MAX_LEVELS] = {0};
  SpkDims D_tab;          // the same dimensions with per_node = 0 and
  SpkPtrs P_tab;          // the per-stage factor table: what the modular S1 kernels (spk_apply) read
  bool gathered = false;  // the device iterate is complete on this rank (spk_gather_iterate since the last pass)
  // fused path (spk_fused.cuh)
  bool fused_ok = false, shape_specialised = false;
  void (*k_bwd)(SpkDims, SpkPtrs, SpkFusedArgs) = nullptr;
  void (*k_fwd)(SpkDims, SpkPtrs, SpkFusedArgs) = nullptr;
  void (*k_dual)(SpkDims, SpkPtrs, SpkFusedArgs) = nullptr;
  // top-of-tree kernels: every level of a sweep with at most top_cap families (one resident wave) in one launch
  void (*k_bwd_top)(SpkDims, SpkPtrs, SpkFusedArgs, SpkTopPlan) = nullptr;
  void (*k_fwd_top)(SpkDims, SpkPtrs, SpkFusedArgs, SpkTopPlan) = nullptr;
  bool use_topk = false;
  long top_cap_bwd = 0, top_cap_fwd = 0;
  unsigned *d_topbar = nullptr;
  void (*k_dual2)(SpkDims, SpkPtrs, SpkFusedArgs, int, long, long) = nullptr;
  void (*k_dual3[2][2])(SpkDims, SpkPtrs, SpkFusedArgs, long, long) = {{nullptr, nullptr}, {nullptr, nullptr}};  // [children are leaves][with <r,Pr>]
  void (*k_dual3_flat[2])(SpkDims, SpkPtrs, SpkFusedArgs, long, long) = {nullptr, nullptr};  // flat owned-family indices, [with <r,Pr>]
  int occ_dual3 = 4;
  // per-stage factors on the fp64 tensor pipe (spk_mm.cuh)
  bool use_mm = false, mm_vec2 = false;
  void (*k_bwd_mm[2])(SpkDims, SpkPtrs, SpkFusedArgs, const double *) = {nullptr, nullptr};
  void (*k_fwd_mm)(SpkDims, SpkPtrs, SpkFusedArgs, const double *) = nullptr;
  void (*k_top_mm)(SpkDims, SpkPtrs, SpkFusedArgs, const double *, const double *, int) = nullptr;
  long mm_min_count = 512;  // levels with fewer families run the CTA-per-family kernels (lower fixed cost per launch)
  int t_top = -1;  // levels 0 .. t_top of both sweeps run in the single-CTA top-of-tree kernel (-1: none)
  double *d_frag_bwd = nullptr, *d_frag_fwd = nullptr;
  long mm_bwd_frags = 0, mm_fwd_frags = 0;
  int occ_bwd_mm = 1, occ_fwd_mm = 1;
  std::vector<double> B_host;
  int dual2_F = 1;          // families per batch of the batched dual kernel
  size_t dual2_smem = 0;
  bool use_dual2 = true;
  int fused_threads_bwd = 0, fused_threads_fwd = 0, dual_warps = 4, occ_bwd = 1, occ_fwd = 1;
  int num_sms = 148;   // replaced by the device's multiProcessorCount in spk_create
  long max_ctas = 0;   // SPK_MAX_CTAS (test switch): cap of every persistent grid, so that small trees reach the steady-state loop
  size_t fused_smem_bwd = 0, fused_smem_fwd = 0, fused_smem_dual = 0;
  double *d_pnorm = nullptr;
  // CUDA graphs of the fused operator evaluation, keyed by (input, output, gamma, sigma, pnorm slot)
  bool use_graphs = false;
  // cut-level exchanges over peer memory (spk_peer.cuh): this rank's mailbox and every rank's mailbox as mapped here
  double *mbox = nullptr;
  double *peer_box[SPK_PEER_MAX] = {nullptr};
  std::vector<void *> ipc_opened;
  bool peer_ok = false;
  long lenA = 0, lenB = 0;
  bool graph_nccl = false;  // SPK_GRAPH_NCCL=1: capture the cut-level exchanges inside the evaluation's CUDA graph (one graph per evaluation)
  bool use_s2_staged = true;  // SPK_NO_S2_STAGED=1: the direct thread-per-node S2 kernel everywhere (cross-check)
  bool use_pdl = true;  // SPK_NO_PDL=1: plain stream order between the level launches of the S1 sweeps
#ifndef SPK_HOST_EMU
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t kernels = 0; };
  std::map<std::tuple<const void *, void *, double, double, int, int>, GraphEntry> graphs;
#endif
  // device-resident loop control (spk_sp.cuh): termination test, SuperMann / Anderson scalars, per-pass log
  SpkCtl *d_ctl = nullptr;
  SpkCtl h_ctl;
  bool dev_loop = false;         // CP and SP loops run without a host synchronisation per pass
  int sp_rot = 0;                // rotations of the Anderson column slots so far (logical column j = slot (j + 2 rot) % 3)
  double *d_log = nullptr, *d_sp_partial = nullptr;
  unsigned *d_ticket = nullptr;
  int log_cap = 0;
#ifndef SPK_HOST_EMU
  cudaGraphExec_t sp_graph = nullptr;
  double sp_graph_gamma = 0, sp_graph_sigma = 0;
  int64_t sp_graph_kernels = 0;
#endif
  // SuperMann over the subtree partition: entries this rank counts in the global sums / maxima (1 byte per entry of (z|v)),
  // and which scalar slots hold a rank-local value that the next fetch must all-reduce (1 = sum, 2 = max)
  unsigned char *d_mask = nullptr;
  int slot_dirty[32] = {0};
  // reductions
  double *d_partial = nullptr, *d_scal = nullptr;
  double h_scal[32];
  int red_blocks = 0;
  // profiling
  bool prof_on = false;
  double prof_ms[PC_COUNT];
  int64_t prof_launches[PC_COUNT];
  std::vector<ProfSection> prof_open;
#ifndef SPK_HOST_EMU
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
#endif
  double *z() { return zv; }
  double *v() { return zv + D.nz; }
};

static int ensure_scratch(spk_handle *h);

static int dev_alloc(spk_handle *h, double **p, long count, bool zero = true) {
  void *q = nullptr;
  const size_t bytes = (size_t)(count > 0 ? count : 1) * sizeof(double);
  SPK_TRY(rt_malloc(&q, bytes));
  h->allocs.push_back(q);
  h->bytes += (int64_t)bytes;
  *p = (double *)q;
  if (zero) SPK_TRY(rt_memset(q, 0, bytes, h->stream));
  return SPK_OK;
}

static int upload(spk_handle *h, double **p, const double *src, long count) {
  SPK_TRY(dev_alloc(h, p, count, false));
  return rt_h2d(*p, src, (size_t)count * sizeof(double), h->stream);
}

// ------------------------------------------------------------------------------------------------ profiling sections

struct ProfScope {
  spk_handle *h;
  bool on;
  ProfSection s;
  ProfScope(spk_handle *h_, int cls) : h(h_), on(h_->prof_on) {
    if (!on) return;
    s.cls = cls;
#ifndef SPK_HOST_EMU
    auto get = [&]() {
      cudaEvent_t e;
      if (!h->ev_pool.empty()) { e = h->ev_pool.back(); h->ev_pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    s.e0 = get();
    s.e1 = get();
    cudaEventRecord(s.e0, h->stream);
#else
    s.t0 = wall_ms();
#endif
  }
  ~ProfScope() {
    if (!on) return;
#ifndef SPK_HOST_EMU
    cudaEventRecord(s.e1, h->stream);
#else
    s.t1 = wall_ms();
#endif
    h->prof_open.push_back(s);
  }
};

static void prof_harvest(spk_handle *h) {
#ifndef SPK_HOST_EMU
  if (!h->prof_open.empty()) cudaStreamSynchronize(h->stream);
#endif
  for (auto &s : h->prof_open) {
#ifndef SPK_HOST_EMU
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s.e0, s.e1);
    h->prof_ms[s.cls] += ms;
    h->ev_pool.push_back(s.e0);
    h->ev_pool.push_back(s.e1);
#else
    h->prof_ms[s.cls] += s.t1 - s.t0;
#endif
  }
  h->prof_open.clear();
}

// ------------------------------------------------------------------------------------------------ launch helpers

static inline int pow2_lanes(int rows) {
#ifdef SPK_HOST_EMU
  (void)rows;
  return 1;
#else
  int l = 1;
  while (l < rows && l < 32) l <<= 1;
  return l;
#endif
}
static inline int elem_block() { return SPK_EMU ? 1 : 256; }
static inline unsigned grid_for(long n, int block) { return (unsigned)((n + block - 1) / block); }

#define SPK_LANES_SWITCH(L, ...)                                   \
  switch (L) {                                                     \
    case 1: { constexpr int LANES = 1; __VA_ARGS__; } break;        \
    case 2: { constexpr int LANES = 2; __VA_ARGS__; } break;        \
    case 4: { constexpr int LANES = 4; __VA_ARGS__; } break;        \
    case 8: { constexpr int LANES = 8; __VA_ARGS__; } break;        \
    case 16: { constexpr int LANES = 16; __VA_ARGS__; } break;      \
    default: { constexpr int LANES = 32; __VA_ARGS__; } break;      \
  }

#define COUNT_LAUNCH(h, cls) \
  do { (h)->launches++; if ((h)->prof_on) (h)->prof_launches[cls]++; } while (0)

// out = alpha * L' v + beta * in   (spock_mul! with trans = true, implicit_l.jl:451-481)
static int op_Lt(spk_handle *h, const double *v, double alpha, double beta, const double *in, double *out) {
  ProfScope ps(h, PC_LLT);
  const int b = elem_block();
  SPK_LAUNCH(spk_k_Lt, grid_for(h->D.nz, b), b, 0, h->stream, h->D, h->P, v, alpha, beta, in, out);
  COUNT_LAUNCH(h, PC_LLT);
  return rt_launch_check();
}
// out = alpha * L z + beta * in
static int op_L(spk_handle *h, const double *z, double alpha, double beta, const double *in, double *out) {
  ProfScope ps(h, PC_LLT);
  const int b = elem_block();
  SPK_LAUNCH(spk_k_L, grid_for(h->D.nv, b), b, 0, h->stream, h->D, h->P, z, alpha, beta, in, out);
  COUNT_LAUNCH(h, PC_LLT);
  return rt_launch_check();
}

static int op_axpby(spk_handle *h, long n, double a, const double *x, double b, const double *y, double *out) {
  ProfScope ps(h, PC_VEC);
  const int blk = elem_block();
  long g = (n + blk - 1) / blk;
  if (g > (long)h->num_sms * 32) g = (long)h->num_sms * 32;
  if (g < 1) g = 1;
  SPK_LAUNCH(spk_k_axpby, (unsigned)g, blk, 0, h->stream, n, a, x, b, y, out);
  COUNT_LAUNCH(h, PC_VEC);
  return rt_launch_check();
}

// mode 0: d_scal[slot] (+)= sum x.*y ; mode 1: d_scal[slot] = max|x|
// moff: offset of x[0] inside the (z|v) index space (selects the part of the ownership mask; multi-rank SuperMann only)
static int op_reduce(spk_handle *h, long n, int mode, const double *x, const double *y, int slot, int accumulate, long moff = 0) {
  ProfScope ps(h, PC_VEC);
  const int blk = SPK_EMU ? 1 : 256;
  long g = (n + blk - 1) / blk;
  if (g > h->red_blocks) g = h->red_blocks;
  if (g < 1) g = 1;
  const unsigned char *mask = (h->d_mask && mode != 2) ? h->d_mask + moff : nullptr;
  if (h->world > 1) h->slot_dirty[slot] = (mode == 1) ? 2 : 1;
  SPK_LAUNCH(spk_k_reduce_partial, (unsigned)g, blk, blk * sizeof(double), h->stream, n, mode, x, y, h->d_partial, mask);
  SPK_LAUNCH(spk_k_reduce_final, 1, blk, blk * sizeof(double), h->stream, (int)g, mode, h->d_partial, h->d_scal, slot,
             accumulate);
  COUNT_LAUNCH(h, PC_VEC);
  COUNT_LAUNCH(h, PC_VEC);
  return rt_launch_check();
}
static int fetch_scalars(spk_handle *h, int count) {
  if (h->world > 1) {
    std::string cerr;
    if (h->alg == SPK_CP) {  // xi1, xi2 are maxima over the ranks' owned families (left in d_scal[0..1] by the dual kernels)
      if (!h->comm.allreduce_max(h->d_scal, 2, h->stream, cerr)) return fail(SPK_ERR_COMM, cerr);
    } else {  // SuperMann: every slot reduced since the last fetch holds a rank-local sum / maximum
      for (int k = 0; k < count && k < 32;) {  // one collective per run of consecutive slots of the same kind
        const int kind = h->slot_dirty[k];
        int e = k + 1;
        while (e < count && e < 32 && h->slot_dirty[e] == kind) ++e;
        if (kind == 1 && !h->comm.allreduce_sum(h->d_scal + k, e - k, h->stream, cerr)) return fail(SPK_ERR_COMM, cerr);
        if (kind == 2 && !h->comm.allreduce_max(h->d_scal + k, e - k, h->stream, cerr)) return fail(SPK_ERR_COMM, cerr);
        for (int j = k; j < e; ++j) h->slot_dirty[j] = 0;
        k = e;
      }
    }
  }
  return rt_d2h(h->h_scal, h->d_scal, sizeof(double) * count, h->stream);
}

// projection_S1! (implicit_l.jl:559-686) in place on arg = (x | u)
static int op_proj_S1(spk_handle *h, double *arg) {
  // (the modular kernels visit EVERY node: they read the per-stage table, which holds the same numbers as the per-node slots)
  const SpkDims &D = h->D_tab;
  const SpkPtrs &Ptab = h->P_tab;
  const int lanes = pow2_lanes(std::max(D.nx, D.nu));
  int block = SPK_EMU ? 1 : 128;
  const size_t per_group_b = (size_t)(D.d * D.nx + 3 * D.nu) * sizeof(double);
  const size_t per_group_f = (size_t)(D.nx + D.nu) * sizeof(double);
  while (block > lanes && (block / lanes) * per_group_b > 40000) block >>= 1;
  const int gpb = block / lanes;
  for (int t = D.N - 2; t >= 0; --t) {
    ProfScope ps(h, PC_S1_BWD);
    const long count = spk_level_count(D.d, t);
    const unsigned grid = (unsigned)((count + gpb - 1) / gpb);
    SPK_LANES_SWITCH(lanes, SPK_LAUNCH(spk_k_s1_bwd<LANES>, grid, block, gpb * per_group_b, h->stream, D, Ptab, t,
                                       (int)(t == D.N - 2), arg, h->ric_q, h->ric_d));
    COUNT_LAUNCH(h, PC_S1_BWD);
  }
  for (int t = 0; t <= D.N - 2; ++t) {
    ProfScope ps(h, PC_S1_FWD);
    const long count = spk_level_count(D.d, t);
    const unsigned grid = (unsigned)((count + gpb - 1) / gpb);
    SPK_LANES_SWITCH(lanes, SPK_LAUNCH(spk_k_s1_fwd<LANES>, grid, block, gpb * per_group_f, h->stream, D, Ptab, t, arg,
                                       h->ric_d));
    COUNT_LAUNCH(h, PC_S1_FWD);
  }
  return rt_launch_check();
}

// projection_S2! (implicit_l.jl:688-727) in place on the full z-sized vector
static int op_proj_S2(spk_handle *h, double *arg) {
  ProfScope ps(h, PC_S2);
  const int b = SPK_EMU ? 1 : 128;
  SPK_LAUNCH(spk_k_s2, grid_for(h->D.nnl, b), b, 0, h->stream, h->D, h->P, arg);
  COUNT_LAUNCH(h, PC_S2);
  return rt_launch_check();
}

// prox_f! (implicit_l.jl:729-748) in place
static int op_prox_f(spk_handle *h, double *arg, double gamma) {
  {
    ProfScope ps(h, PC_VEC);
    SPK_LAUNCH(spk_k_add_scalar, 1, 1, 0, h->stream, arg, h->D.zo[2], -gamma);  // :741
    COUNT_LAUNCH(h, PC_VEC);
  }
  SPK_TRY(op_proj_S1(h, arg));
  return op_proj_S2(h, arg);
}

// prox_h_conj! (implicit_l.jl:911-951); in == out allowed
static int op_prox_hconj(spk_handle *h, const double *in, double sigma, double *out) {
  ProfScope ps(h, PC_PROX_H);
  const SpkDims &D = h->D;
  const int b = elem_block();
  SPK_LAUNCH(spk_k_prox_own, grid_for(D.nnl, b), b, 0, h->stream, D, sigma, in, out);
  const long nbox = D.nnl * (long)(D.nx + D.nu) + D.nl * (long)D.nx;
  SPK_LAUNCH(spk_k_prox_box, grid_for(nbox, b), b, 0, h->stream, D, sigma, in, out);
  const int lanes = pow2_lanes(D.nx + D.nu);
  const int block = SPK_EMU ? 1 : 128;
  const int gpb = block / lanes;
  SPK_LANES_SWITCH(lanes, SPK_LAUNCH(spk_k_prox_soc_child<LANES>, grid_for(D.n - 1, gpb), block, 0, h->stream, D, sigma,
                                     in, out));
  const int lanes_l = pow2_lanes(D.nx);
  const int gpb_l = block / lanes_l;
  SPK_LANES_SWITCH(lanes_l, SPK_LAUNCH(spk_k_prox_soc_leaf<LANES>, grid_for(D.nl, gpb_l), block, 0, h->stream, D, sigma,
                                       in, out));
  for (int k = 0; k < 4; ++k) COUNT_LAUNCH(h, PC_PROX_H);
  return rt_launch_check();
}

// T: (z,v) -> (zbar,vbar)   update_zbar!/update_vbar! (cp.jl:5-32 == sp.jl:5-32 == sp.jl:133-153)
static int op_T(spk_handle *h, const double *zv_in, double *zv_out, double gamma, double sigma) {
  const long nz = h->D.nz;
  const double *z = zv_in, *v = zv_in + nz;
  double *zbar = zv_out, *vbar = zv_out + nz;
  SPK_TRY(op_Lt(h, v, -gamma, 1.0, z, zbar));                 // zbar = -gamma L'v + z
  SPK_TRY(op_prox_f(h, zbar, gamma));
  SPK_TRY(op_axpby(h, nz, 2.0, zbar, -1.0, z, h->zw));        // z_workspace = 2 zbar - z
  SPK_TRY(op_L(h, h->zw, sigma, 1.0, v, vbar));               // vbar = sigma L(zw) + v
  return op_prox_hconj(h, vbar, sigma, vbar);
}

// ---- fused operator evaluation (spk_fused.cuh): T(z,v) in 2(N-1) family-kernel launches.
// Leaves xi1, xi2 of (zbar,vbar)-(z,v) in d_scal[0..1] and, if pnorm_slot >= 0, <r,Pr> in d_scal[pnorm_slot].
// positions [lo, lo+cnt) of level t processed by this rank
static void level_range(const spk_handle *h, int t, long &lo, long &cnt) {
  const long full = spk_level_count(h->D.d, t);
  if (h->world <= 1 || t < h->t_cut) { lo = 0; cnt = full; return; }
  const long sub = spk_level_count(h->D.d, t - h->t_cut);  // nodes of level t below one cut-level node
  lo = (long)h->rank * h->m_sub * sub;
  cnt = h->m_sub * sub;
}
// owned non-leaf families as contiguous node ranges (kernels 3 and 4)
static void fill_ranges(const spk_handle *h, SpkFusedArgs &A) {
  const SpkDims &D = h->D;
  if (h->world <= 1) {
    A.n_ranges = 1; A.r_start[0] = 0; A.r_prefix[0] = 0; A.r_prefix[1] = D.nnl;
    return;
  }
  int k = 0;
  long pre = 0;
  if (h->t_cut > 0) {
    A.r_start[k] = 0; A.r_prefix[k] = 0;
    pre = spk_level_first0(D.d, h->t_cut);
    ++k;
  }
  for (int t = h->t_cut; t <= D.N - 2; ++t) {
    long lo, cnt;
    level_range(h, t, lo, cnt);
    A.r_start[k] = spk_level_first0(D.d, t) + lo;
    A.r_prefix[k] = pre;
    pre += cnt;
    ++k;
  }
  A.n_ranges = k;
  A.r_prefix[k] = pre;
}
static int comm_fail(const std::string &err) { return fail(SPK_ERR_COMM, err); }

// Shape-specialised instantiations of the two factor-streaming kernels (nx, nu, d, threads); anything else runs
// the runtime-size instantiation <0,0,0,0>.
#ifndef SPK_HOST_EMU
#include "spk_shapes.h"
#define SPK_SHAPE_X(NX, NU, DD, TB, TF) extern "C" const SpkShapeEntry *spk_shape_entry_##NX##_##NU##_##DD();
#include "spk_shapes.def"
#undef SPK_SHAPE_X
typedef const SpkShapeEntry *(*spk_shape_getter_t)();
#define SPK_SHAPE_X(NX, NU, DD, TB, TF) spk_shape_entry_##NX##_##NU##_##DD,
static const spk_shape_getter_t g_shapes[] = {
#include "spk_shapes.def"
};
#undef SPK_SHAPE_X
#else
typedef void (*spk_fused_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs);
typedef void (*spk_dual2_kernel_t)(SpkDims, SpkPtrs, SpkFusedArgs, int, long, long);
#endif

#ifndef SPK_HOST_EMU
// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the FUNCTION, shared by every handle of the process (the
// runtime-size instantiations serve all unspecialised shapes): only ever raise it
static cudaError_t set_max_smem(const void *fn, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> cur;  // (function, device): the attribute is per device
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  size_t &c = cur[std::make_pair(fn, dev)];
  if (bytes <= c) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) c = bytes;
  return e;
}
#endif

static int fused_setup(spk_handle *h) {
  const SpkDims &D = h->D;
  h->fused_ok = false;
  const char *env = getenv("SPK_DISABLE_FUSED");
  if (env && env[0] == '1') return SPK_OK;
  if (std::max(D.nx, D.nu) > 256) return SPK_OK;
  // threads per family: enough slices of the inner dimension to keep every mat-vec short
  int Tb = 32, Tf = 32;
  while (Tb < 4 * std::max(D.nx, D.nu) && Tb < 256) Tb <<= 1;
  while (Tf < 2 * (D.nu + D.d * D.nx) && Tf < 256) Tf <<= 1;
  h->k_bwd = spk_k_bwd<0, 0, 0, 0>;
  h->k_fwd = spk_k_fwd_s1<0, 0, 0, 0>;
  h->k_dual = spk_k_dual<0, 0, 0>;
  h->k_dual2 = spk_k_dual2<0, 0, 0>;
  h->k_bwd_top = spk_k_bwd_top<0, 0, 0, 0>;
  h->k_fwd_top = spk_k_fwd_top<0, 0, 0, 0>;
  if (h->custom_modes) {
    h->k_bwd = spk_k_bwd<0, 0, 0, 0, true>;
    h->k_fwd = spk_k_fwd_s1<0, 0, 0, 0, true>;
    h->k_bwd_top = spk_k_bwd_top<0, 0, 0, 0, true>;
    h->k_fwd_top = spk_k_fwd_top<0, 0, 0, 0, true>;
  }
  const char *gen = getenv("SPK_GENERIC_KERNELS");
#ifndef SPK_HOST_EMU
  if (!(gen && gen[0] == '1'))
    for (const spk_shape_getter_t getter : g_shapes) {
      const SpkShapeEntry &e = *getter();
      if (e.nx == D.nx && e.nu == D.nu && e.d == D.d) {
        h->k_dual = e.dual; h->k_dual2 = e.dual2;
        if (!h->custom_modes) {  // (a mode map runs the runtime-size sweeps: see spk_mode_c)
          Tb = e.Tb; Tf = e.Tf; h->k_bwd = e.bwd; h->k_fwd = e.fwd;
          h->k_bwd_top = e.bwd_top; h->k_fwd_top = e.fwd_top;
        }
        h->shape_specialised = true;
        // register-resident dual kernel: Diagonal sqrt-costs, risk cone and branching within one warp
        const char *e3 = getenv("SPK_NO_DUAL3");
        const bool uc = D.n_sqrtQ == 1 && D.n_sqrtQN == 1;
        if (!D.cost_dense && D.ny <= spk_dual3_group(D.nx, D.nu, D.d) && D.d <= 32 && !(e3 && e3[0] == '1') && (uc || (D.n_sqrtQ > 1 && D.n_sqrtQN > 1))) {
          for (int lv = 0; lv < 2; ++lv)
            for (int pn = 0; pn < 2; ++pn) h->k_dual3[lv][pn] = e.dual3[lv][uc ? 1 : 0][pn];
          for (int pn = 0; pn < 2; ++pn) h->k_dual3_flat[pn] = e.dual3_flat[uc ? 1 : 0][pn];
        }
        const char *emm = getenv("SPK_NO_MM");
        if (!D.per_node && !D.cost_dense && !(emm && emm[0] == '1')) {
          h->k_bwd_mm[0] = e.bwd_mm[0]; h->k_bwd_mm[1] = e.bwd_mm[1]; h->k_fwd_mm = e.fwd_mm; h->k_top_mm = e.top_mm;
          h->use_mm = true;
        }
      }
    }
#else
  (void)gen;
#endif
  if (SPK_EMU) Tb = Tf = 1;
  h->fused_threads_bwd = Tb;
  h->fused_threads_fwd = Tf;
  h->fused_smem_bwd = (size_t)spk_bwd_smem_doubles(D, Tb) * sizeof(double);
  h->fused_smem_fwd = (size_t)spk_fwd_smem_doubles(D, Tf) * sizeof(double);

  {  // batched dual kernel: as many families per CTA as fit in ~44 KB of shared memory
    int F = 16;
    while (F > 1 && spk_dual_map(D.nx, D.nu, D.d, D.ny, F, true).total * 8 > 44 * 1024) F >>= 1;
    h->dual2_F = F;
    h->dual2_smem = (size_t)spk_dual_map(D.nx, D.nu, D.d, D.ny, F, true).total * sizeof(double);
    const char *e1 = getenv("SPK_DUAL_V1");
    h->use_dual2 = !(e1 && e1[0] == '1') && h->dual2_smem <= 200 * 1024;
  }
  h->dual_warps = SPK_EMU ? 1 : 4;
  h->fused_smem_dual = (size_t)spk_dual_warp_doubles(D) * h->dual_warps * sizeof(double);
#ifndef SPK_HOST_EMU
  if (std::max(h->fused_smem_bwd, h->fused_smem_fwd) > 227 * 1024 || h->fused_smem_dual > 200 * 1024) return SPK_OK;
  RT_CHECK(set_max_smem((const void *)h->k_bwd, h->fused_smem_bwd));
  RT_CHECK(set_max_smem((const void *)h->k_fwd, h->fused_smem_fwd));
  RT_CHECK(set_max_smem((const void *)h->k_dual, h->fused_smem_dual));
  if (h->use_dual2)
    RT_CHECK(set_max_smem((const void *)h->k_dual2, h->dual2_smem));
  RT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_bwd, (const void *)h->k_bwd, Tb, h->fused_smem_bwd));
  RT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_fwd, (const void *)h->k_fwd, Tf, h->fused_smem_fwd));
  h->occ_bwd = std::max(1, h->occ_bwd);
  {  // top-of-tree kernels: their grid must be co-resident (software grid barrier)
    // Measured (profiles/r02c_topk_ab.txt): with the level launches chained by programmatic dependent launch a level costs
    // ~3.1 us, a level step behind the grid barrier ~4.6 us (N=10, nx=20: 65.6 -> 88.1 us per pass; N=13, nx=50: 185 -> 217 us),
    // because the floor is the dependent chain INSIDE a level, not the launch.  The kernels stay as an opt-in (SPK_TOPK=1).
    const char *nt = getenv("SPK_TOPK");
    int ob = 0, of = 0;
    RT_CHECK(set_max_smem((const void *)h->k_bwd_top, h->fused_smem_bwd));
    RT_CHECK(set_max_smem((const void *)h->k_fwd_top, h->fused_smem_fwd));
    RT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ob, (const void *)h->k_bwd_top, Tb, h->fused_smem_bwd));
    RT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&of, (const void *)h->k_fwd_top, Tf, h->fused_smem_fwd));
    h->top_cap_bwd = (long)h->num_sms * ob;
    h->top_cap_fwd = (long)h->num_sms * of;
    h->use_topk = (nt && nt[0] == '1') && ob >= 1 && of >= 1;
    double *tb = nullptr;
    SPK_TRY(dev_alloc(h, &tb, 2));
    h->d_topbar = reinterpret_cast<unsigned *>(tb);
  }
  if (h->k_dual3[1][0]) {
    RT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_dual3, (const void *)h->k_dual3[1][0], 128, 0));
    h->occ_dual3 = std::max(1, h->occ_dual3);
  }
  if (h->use_mm) {  // fragment-ordered per-level weight tables (spk_mm_layout.h)
    const SpkMmLayout L = spk_mm_layout(D.nx, D.nu, D.d);
    h->mm_bwd_frags = L.bwd_frags; h->mm_fwd_frags = L.fwd_frags;
    {  // 16-byte loads of consecutive entry pairs need even sizes and even segment offsets (reference layout, 8-byte entries)
      const char *nv2 = getenv("SPK_NO_VEC2");
      h->mm_vec2 = !(nv2 && nv2[0] == '1') && D.nx % 2 == 0 && D.nu % 2 == 0 && D.nz % 2 == 0 && D.zo[1] % 2 == 0 &&
                   D.vo[2] % 2 == 0 && D.vo[3] % 2 == 0 && D.vo[6] % 2 == 0 && D.vo[7] % 2 == 0 && D.vo[10] % 2 == 0;
    }
    const size_t sb = spk_mm_smem_bytes(L.bwd_frags), sf = spk_mm_smem_bytes(L.fwd_frags);
    if (std::max(sb, sf) > 227 * 1024) {
      h->use_mm = false;
    } else {
      std::vector<double> fb, ff;
      spk_mm_fragment_tables(h->table, h->B_host.data(), fb, ff);
      SPK_TRY(upload(h, &h->d_frag_bwd, fb.data(), (long)fb.size()));
      SPK_TRY(upload(h, &h->d_frag_fwd, ff.data(), (long)ff.size()));
      for (int k = 0; k < 2; ++k)
        RT_CHECK(set_max_smem((const void *)h->k_bwd_mm[k], sb));
      RT_CHECK(set_max_smem((const void *)h->k_fwd_mm, sf));
      RT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_bwd_mm, (const void *)h->k_bwd_mm[1], 256, sb));
      RT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_fwd_mm, (const void *)h->k_fwd_mm, 256, sf));
      h->occ_bwd_mm = std::max(1, h->occ_bwd_mm);
      h->occ_fwd_mm = std::max(1, h->occ_fwd_mm);
      RT_CHECK(set_max_smem((const void *)h->k_top_mm, std::max(sb, sf)));
      const char *nt = getenv("SPK_NO_TOP");
      h->t_top = -1;
      // (measured: with large fragment tables -- nx = 50: 93-168 KB per level -- a level in the single-CTA kernel costs
      // ~9 us against ~5 us for a CTA-per-family launch, so the top-of-tree kernel is used for the small shapes only)
      if (!(nt && nt[0] == '1') && std::max(sb, sf) <= 64 * 1024)
        for (int t = 0; t <= D.N - 3 && spk_level_count(D.d, t) <= 64; ++t) h->t_top = t;  // never the leaf-parent level
      const char *mc = getenv("SPK_MM_MIN_COUNT");
      if (mc) h->mm_min_count = atol(mc);
    }
  }
  h->occ_fwd = std::max(1, h->occ_fwd);
#endif
  h->fused_ok = true;
  {
    const char *ns2 = getenv("SPK_NO_S2_STAGED");
    h->use_s2_staged = !(ns2 && ns2[0] == '1');
    const char *np = getenv("SPK_NO_PDL");
    h->use_pdl = !(np && np[0] == '1');
  }
  {
    const char *mc = getenv("SPK_MAX_CTAS");
    h->max_ctas = mc ? atol(mc) : 0;
  }
  {
    const char *ng = getenv("SPK_NO_GRAPH");
    h->use_graphs = !SPK_EMU && !(ng && ng[0] == '1');
  }
  return SPK_OK;
}

#ifndef SPK_HOST_EMU
// launch of a level kernel of the S1 sweeps, chained to the previous launch on the stream by programmatic dependent launch
static cudaError_t launch_level(spk_handle *h, spk_fused_kernel_t k, unsigned grid, unsigned block, size_t smem, const SpkDims &D,
                                const SpkFusedArgs &A) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = h->use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, D, h->P, A);
}
#endif

static unsigned fused_grid(const spk_handle *h, long count, int occ) {
  // persistent CTAs: one resident wave strides over the families of the level, so the per-thread index arithmetic is
  // paid once per kernel; with per-stage factors the level's matrices are also fetched only once per CTA
  long g = std::min(count, (long)h->num_sms * occ);
  if (SPK_EMU) g = std::min(count, 3L);
  if (h->max_ctas > 0) g = std::min(g, h->max_ctas);
  return (unsigned)std::max(1L, g);
}

// One cut-level exchange (kind 0: the subtree roots' q; kind 1: their vbar2 and the residual maxima): a peer-memory kernel
// (spk_peer.cuh) when the ranks' mailboxes are mapped, else an in-place NCCL all-gather issued by the host
static int launch_exchange(spk_handle *h, int kind, double *vec, long vlen) {
  ProfScope ps(h, PC_VEC);
#ifndef SPK_HOST_EMU
  if (h->peer_ok) {
    SpkPeerArgs X;
    X.rank = h->rank; X.world = h->world; X.kind = kind;
    X.len = kind == 0 ? h->lenA : h->lenB;
    X.off_payload = kind == 0 ? 0 : 2L * h->world * h->lenA;
    X.off_flags = 2L * h->world * (h->lenA + h->lenB);
    X.mine = h->mbox;
    for (int r = 0; r < SPK_PEER_MAX; ++r) X.peer[r] = r < h->world ? h->peer_box[r] : nullptr;
    X.gate = h->d_ctl ? &h->d_ctl->done : nullptr;
    X.vec = vec; X.vlen = vlen; X.xi = h->d_scal;
    spk_k_peer_exchange<<<1, 128, 0, h->stream>>>(X);
    COUNT_LAUNCH(h, PC_VEC);
    return rt_launch_check();
  }
#endif
  std::string cerr;
  if (!h->comm.allgather_inplace(vec, vlen, h->stream, cerr)) return fail(SPK_ERR_COMM, cerr);
  return SPK_OK;
}

// `parts`: 1 = kernels up to the cut-level exchange, 2 = that exchange (all-gather of the subtree roots' q), 4 = the remaining
// kernels, 8 = the exchange of the subtree roots' vbar2.  All (15) for a plain launch; the multi-GPU graph replay captures the
// kernel segments 1 and 4 separately and issues the two NCCL exchanges between them on the stream.
#ifndef SPK_HOST_EMU
// S2 staged through shared memory (spk_k_s2t_st): the projector of the one risk measure travels as a kernel parameter
template <int NY_, int DD_>
static void launch_s2t_staged(spk_handle *h, const SpkDims &D, const SpkFusedArgs &A, long owned) {
  SpkLsMat<NY_, DD_> M;
  std::copy(h->ls_host.begin(), h->ls_host.begin() + (size_t)(NY_ + 2 * DD_) * (NY_ + 2 * DD_), M.m);
  spk_k_s2t_st<NY_, DD_><<<grid_for(owned, 128), 128, 0, h->stream>>>(D, h->P, A, M);
}
#endif

static int op_T_fused_launch(spk_handle *h, const double *zv_in, double *zv_out, double gamma, double sigma, int pnorm_slot,
                             int parts = 15) {
  const SpkDims &D = h->D;
  SpkFusedArgs A;
  A.gamma = gamma; A.sigma = sigma; A.inv_gamma = 1.0 / gamma; A.inv_sigma = 1.0 / sigma;
  A.z = zv_in; A.v = zv_in + D.nz; A.zbar = zv_out; A.vbar = zv_out + D.nz;
  A.q = h->ric_q; A.dd = h->ric_d; A.xi = h->d_scal; A.pnorm = pnorm_slot >= 0 ? h->d_pnorm : nullptr;
  A.t = 0; A.children_are_leaves = 0; A.reset_xi = 0; A.p_lo = 0; A.p_cnt = 0; A.kshift = 0;
  A.gate = h->d_ctl ? &h->d_ctl->done : nullptr;
  A.vec2 = h->mm_vec2 ? 1 : 0;
  A.topbar = h->d_topbar;
  fill_ranges(h, A);
  const long owned = A.r_prefix[A.n_ranges];
  std::string cerr;
  const bool seg_a = (parts & 1) != 0, seg_b = (parts & 4) != 0;
  if (seg_a) {  // 3. S2 (independent of the sweeps)
    ProfScope ps(h, PC_S2);
    const int b = SPK_EMU ? 1 : 128;
    void (*ks2)(SpkDims, SpkPtrs, SpkFusedArgs) = spk_k_s2t<0, 0>;
    if (!SPK_EMU) {  // the common risk measures: AV@R (ny = 2d+1), l2-ball / TV style (ny = 2d+2)
      if (D.ny == 5 && D.d == 2) ks2 = spk_k_s2t<5, 2>;
      else if (D.ny == 7 && D.d == 3) ks2 = spk_k_s2t<7, 3>;
      else if (D.ny == 9 && D.d == 4) ks2 = spk_k_s2t<9, 4>;
      else if (D.ny == 6 && D.d == 2) ks2 = spk_k_s2t<6, 2>;
      else if (D.ny == 8 && D.d == 3) ks2 = spk_k_s2t<8, 3>;
      else if (D.ny == 10 && D.d == 4) ks2 = spk_k_s2t<10, 4>;
    }
    bool staged = false;
#ifndef SPK_HOST_EMU
    // one contiguous node range and one risk measure for all nodes: the shared-memory-staged version (spk_k_s2t_st)
    // (not on small trees: its 2.6 KB of extra kernel parameters cost a launch ~1 us, more than the kernel saves there;
    // SPK_MAX_CTAS, the test switch that makes small trees exercise the steady-state paths, keeps it on)
    if (h->use_s2_staged && A.n_ranges == 1 && !D.rms_per_node && (long)h->ls_host.size() == (long)D.nls * D.nls &&
        (owned >= 4096 || h->max_ctas > 0)) {
      staged = true;
      if (D.ny == 5 && D.d == 2) launch_s2t_staged<5, 2>(h, D, A, owned);
      else if (D.ny == 7 && D.d == 3) launch_s2t_staged<7, 3>(h, D, A, owned);
      else if (D.ny == 9 && D.d == 4) launch_s2t_staged<9, 4>(h, D, A, owned);
      else if (D.ny == 6 && D.d == 2) launch_s2t_staged<6, 2>(h, D, A, owned);
      else if (D.ny == 8 && D.d == 3) launch_s2t_staged<8, 3>(h, D, A, owned);
      else if (D.ny == 10 && D.d == 4) launch_s2t_staged<10, 4>(h, D, A, owned);
      else staged = false;
    }
#endif
    if (!staged) SPK_LAUNCH(ks2, grid_for(owned, b), b, 0, h->stream, D, h->P, A);
    COUNT_LAUNCH(h, PC_S2);
  }
  // levels 0 .. t_top of both sweeps run in one single-CTA launch (single GPU, tensor-pipe kernels)
  const int t_top = (h->use_mm && h->world == 1) ? h->t_top : -1;
  // levels 0 .. ts_b / ts_f of a sweep are narrower than one resident wave: one top-of-tree launch each (a second one for
  // the backward sweep under the partition, where the cut-level exchange sits between two of its levels)
  int ts_b = -1, ts_f = -1;
#ifndef SPK_HOST_EMU
  if (h->use_topk && t_top < 0) {
    auto narrow = [&](int t, long cap) {
      long lo, cnt;
      level_range(h, t, lo, cnt);
      return cnt <= cap && !(h->use_mm && cnt >= h->mm_min_count);
    };
    // (SPK_MAX_CTAS test switch: the top-of-tree grid is capped too, so its CTAs stride over up to 64 families per level,
    // and wider levels keep exercising the steady-state loop of the per-level kernels)
    const long cb = h->max_ctas > 0 ? std::min(h->top_cap_bwd, h->max_ctas * 64) : h->top_cap_bwd;
    const long cf = h->max_ctas > 0 ? std::min(h->top_cap_fwd, h->max_ctas * 64) : h->top_cap_fwd;
    while (ts_b + 1 <= D.N - 2 && narrow(ts_b + 1, cb)) ++ts_b;
    while (ts_f + 1 <= D.N - 2 && narrow(ts_f + 1, cf)) ++ts_f;
    if (ts_b < 1) ts_b = -1;  // a single level gains nothing
    if (ts_f < 1) ts_f = -1;
  }
  auto launch_top = [&](bool bwd, int t_first, int t_last, int bar_idx) -> int {  // levels t_first .. t_last in execution order
    SpkTopPlan plan;
    plan.n = 0;
    plan.bar = h->d_topbar + bar_idx;
    long widest = 1;
    for (int t = t_first; bwd ? t >= t_last : t <= t_last; t += bwd ? -1 : 1) {
      const int k = plan.n++;
      plan.t[k] = t; plan.leaves[k] = (t == D.N - 2);
      level_range(h, t, plan.p_lo[k], plan.p_cnt[k]);
      plan.kshift[k] = spk_level_first0(D.d, t) + plan.p_lo[k] - h->fac_start[t];
      widest = std::max(widest, plan.p_cnt[k]);
    }
    SpkFusedArgs At = A;
    At.reset_xi = (bwd && t_first == D.N - 2) ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    long grid = std::min(widest, bwd ? h->top_cap_bwd : h->top_cap_fwd);
    if (h->max_ctas > 0) grid = std::min(grid, h->max_ctas);
    cfg.gridDim = dim3((unsigned)std::max(1L, grid), 1, 1);
    cfg.blockDim = dim3((unsigned)(bwd ? h->fused_threads_bwd : h->fused_threads_fwd), 1, 1);
    cfg.dynamicSmemBytes = bwd ? h->fused_smem_bwd : h->fused_smem_fwd;
    cfg.stream = h->stream;
    // cooperative launch: the grid is gang-scheduled, so the software grid barrier cannot deadlock against kernels of
    // other handles / streams that share the GPU (batched independent problems)
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    RT_CHECK(cudaLaunchKernelEx(&cfg, bwd ? h->k_bwd_top : h->k_fwd_top, D, h->P, At, plan));
    COUNT_LAUNCH(h, bwd ? PC_FUSED_BWD : PC_FUSED_FWD);
    return SPK_OK;
  };
#endif
  for (int t = D.N - 2; t > t_top; --t) {  // 1. backward sweep
    const bool in_a = h->world <= 1 || t >= h->t_cut;  // levels down to the cut belong to the first kernel segment
#ifndef SPK_HOST_EMU
    if (t <= ts_b) {  // the narrow levels t .. stop in one launch
      const int stop = (h->world > 1 && t >= h->t_cut) ? h->t_cut : 0;
      if (in_a ? seg_a : seg_b) {
        ProfScope ps(h, PC_FUSED_BWD);
        SPK_TRY(launch_top(true, t, stop, in_a ? 0 : 2));
      }
      t = stop;
      if (h->world > 1 && t == h->t_cut && (parts & 2))
        SPK_TRY(launch_exchange(h, 0, h->ric_q + spk_level_first0(D.d, t) * D.nx, h->m_sub * D.nx));
      continue;
    }
#endif
    if (in_a ? seg_a : seg_b) {
      ProfScope ps(h, PC_FUSED_BWD);
      A.t = t; A.children_are_leaves = (t == D.N - 2); A.reset_xi = (t == D.N - 2);
      level_range(h, t, A.p_lo, A.p_cnt);
      A.kshift = spk_level_first0(D.d, t) + A.p_lo - h->fac_start[t];
#ifndef SPK_HOST_EMU
      if (h->use_mm && A.p_cnt >= h->mm_min_count) {  // per-stage factors: the level is a chain of mat-mats on the fp64 tensor pipe
        // small levels: fewer warps (tiles) per CTA so that the tiles spread over the SMs (a DMMA chain is bound by
        // the tensor pipe of its SM sub-partition)
        long cap = (long)h->num_sms * h->occ_bwd_mm;
        if (h->max_ctas > 0) cap = std::min(cap, h->max_ctas);
        const long tiles = (A.p_cnt + 7) / 8;
        const long wpb = std::min(8L, std::max(1L, (tiles + h->num_sms - 1) / h->num_sms));  // one tile per SM sub-partition first
        const long ctas = std::min((tiles + wpb - 1) / wpb, cap);
        spk_mm_kernel_t kmm = h->k_bwd_mm[A.children_are_leaves ? 1 : 0];
        kmm<<<(unsigned)std::max(1L, ctas), (unsigned)(32 * wpb), spk_mm_smem_bytes((int)h->mm_bwd_frags), h->stream>>>(
            D, h->P, A, h->d_frag_bwd + (size_t)t * h->mm_bwd_frags * 32);
      } else
#endif
      {
        spk_fused_kernel_t kfn = h->k_bwd;
#ifndef SPK_HOST_EMU
        RT_CHECK(launch_level(h, kfn, fused_grid(h, A.p_cnt, h->occ_bwd), h->fused_threads_bwd, h->fused_smem_bwd, D, A));
#else
        SPK_LAUNCH(kfn, fused_grid(h, A.p_cnt, h->occ_bwd), h->fused_threads_bwd, h->fused_smem_bwd, h->stream, D, h->P, A);
#endif
      }
      COUNT_LAUNCH(h, PC_FUSED_BWD);
    }
    if (h->world > 1 && t == h->t_cut && (parts & 2))  // the replicated levels above the cut need every subtree root's q
      SPK_TRY(launch_exchange(h, 0, h->ric_q + spk_level_first0(D.d, t) * D.nx, h->m_sub * D.nx));
  }
  A.reset_xi = 0;
  if (!seg_b) {
    if (h->world > 1 && (parts & 8))  // the second exchange on its own (between graph segments)
      SPK_TRY(launch_exchange(h, 1, A.vbar + D.vo[1] + spk_level_first0(D.d, h->t_cut), h->m_sub));
    return rt_launch_check();
  }
#ifndef SPK_HOST_EMU
  if (t_top >= 0) {
    ProfScope ps(h, PC_FUSED_BWD);
    A.t = 0; A.children_are_leaves = 0; A.p_lo = 0; A.p_cnt = 1;
    h->k_top_mm<<<1, 256, spk_mm_smem_bytes((int)std::max(h->mm_bwd_frags, h->mm_fwd_frags)), h->stream>>>(
        D, h->P, A, h->d_frag_bwd, h->d_frag_fwd, t_top);
    COUNT_LAUNCH(h, PC_FUSED_BWD);
  }
#endif
#ifndef SPK_HOST_EMU
  if (ts_f >= 1) {
    ProfScope ps(h, PC_FUSED_FWD);
    SPK_TRY(launch_top(false, 0, ts_f, 1));
  }
#endif
  for (int t = std::max(t_top, ts_f) + 1; t <= D.N - 2; ++t) {  // 2. forward sweep
    ProfScope ps(h, PC_FUSED_FWD);
    A.t = t; A.children_are_leaves = (t == D.N - 2);
    level_range(h, t, A.p_lo, A.p_cnt);
    A.kshift = spk_level_first0(D.d, t) + A.p_lo - h->fac_start[t];
#ifndef SPK_HOST_EMU
    if (h->use_mm && A.p_cnt >= h->mm_min_count) {
      long cap = (long)h->num_sms * h->occ_fwd_mm;
      if (h->max_ctas > 0) cap = std::min(cap, h->max_ctas);
      const long tiles = (A.p_cnt + 7) / 8;
      const long wpb = std::min(8L, std::max(1L, (tiles + h->num_sms - 1) / h->num_sms));
      const long ctas = std::min((tiles + wpb - 1) / wpb, cap);
      h->k_fwd_mm<<<(unsigned)std::max(1L, ctas), (unsigned)(32 * wpb), spk_mm_smem_bytes((int)h->mm_fwd_frags), h->stream>>>(
          D, h->P, A, h->d_frag_fwd + (size_t)t * h->mm_fwd_frags * 32);
    } else
#endif
    {
      spk_fused_kernel_t kfn = h->k_fwd;
#ifndef SPK_HOST_EMU
      RT_CHECK(launch_level(h, kfn, fused_grid(h, A.p_cnt, h->occ_fwd), h->fused_threads_fwd, h->fused_smem_fwd, D, A));
#else
      SPK_LAUNCH(kfn, fused_grid(h, A.p_cnt, h->occ_fwd), h->fused_threads_fwd, h->fused_smem_fwd, h->stream, D, h->P, A);
#endif
    }
    COUNT_LAUNCH(h, PC_FUSED_FWD);
  }
  long pn_parts = owned;  // number of <r,Pr> partials in d_pnorm
#ifndef SPK_HOST_EMU
  if (h->use_dual2 && h->world > 1 && h->k_dual3_flat[0] && h->k_dual3[1][0]) {
    // 4c under the subtree partition: the register-resident kernel over FLAT owned-family indices -- one launch for all
    // families with non-leaf children (the slices of every level at once) and one for the leaf parents (the last range)
    ProfScope ps(h, PC_PROX_H);
    const long fpc = 4L * (32 / spk_dual3_group(D.nx, D.nu, D.d));  // families per CTA (4 warps)
    const int leaf_rng = A.n_ranges - 1;
    const long F_all = A.r_prefix[A.n_ranges], F_nl = A.r_prefix[leaf_rng];
    const long top = h->t_cut > 0 ? A.r_prefix[1] : 0;
    const bool skip_top_pn = A.pnorm && h->rank != 0 && top > 0;  // <r, P r> of the replicated top is counted by rank 0 only
    long pn_off = 0;
    auto launch3 = [&](int part, long lo_arg, long cnt, bool with_pn) {  // part 0: lo_arg = first FLAT index; part 1: first node
      if (cnt <= 0) return;
      SpkFusedArgs A2 = A;
      A2.children_are_leaves = part;
      A2.pnorm = with_pn ? A.pnorm + pn_off : nullptr;
      const long g = (cnt + fpc - 1) / fpc;
      spk_dual3_kernel_t k3 = part == 0 ? h->k_dual3_flat[with_pn ? 1 : 0] : h->k_dual3[part][with_pn ? 1 : 0];
      k3<<<(unsigned)g, 128, 0, h->stream>>>(D, h->P, A2, lo_arg, cnt);
      COUNT_LAUNCH(h, PC_PROX_H);
      if (with_pn) pn_off += g;
    };
    if (skip_top_pn) {
      launch3(0, 0, std::min(top, F_nl), false);
      launch3(0, top, F_nl - top, true);
    } else {
      launch3(0, 0, F_nl, A.pnorm != nullptr);
    }
    launch3(1, A.r_start[leaf_rng], F_all - F_nl, A.pnorm != nullptr);
    pn_parts = pn_off;
  } else
#endif
  if (h->use_dual2) {  // 4b. batched dual kernel: one launch per owned range, split where the children become leaves
    ProfScope ps(h, PC_PROX_H);
    const long leaf_parents0 = spk_level_first0(D.d, D.N - 2);  // first family whose children are leaves
    const int T = SPK_EMU ? 1 : 256;
    long pn_off = 0;
    for (int r = 0; r < A.n_ranges; ++r) {
      const long lo = A.r_start[r], hi = lo + (A.r_prefix[r + 1] - A.r_prefix[r]);
      for (int part = 0; part < 2; ++part) {
        const long a = part == 0 ? lo : std::max(lo, leaf_parents0), b = part == 0 ? std::min(hi, leaf_parents0) : hi;
        if (b <= a) continue;
        SpkFusedArgs A2 = A;
        A2.children_are_leaves = part;
        // <r, P r> of the replicated families above the cut is counted by rank 0 only
        const bool count_pn = A.pnorm && !(h->world > 1 && h->rank != 0 && h->t_cut > 0 && r == 0);
        A2.pnorm = count_pn ? A.pnorm + pn_off : nullptr;
        long g;
#ifndef SPK_HOST_EMU
        if (h->k_dual3[part][0]) {  // 4c. one group of lanes per family, everything in registers
          const long fpc = 4L * (32 / spk_dual3_group(D.nx, D.nu, D.d));  // families per CTA (4 warps)
          g = (b - a + fpc - 1) / fpc;
          spk_dual3_kernel_t k3 = h->k_dual3[part][A2.pnorm ? 1 : 0];
          k3<<<(unsigned)g, 128, 0, h->stream>>>(D, h->P, A2, a, b - a);
        } else
#endif
        {
          const long nbatch = (b - a + h->dual2_F - 1) / h->dual2_F;
          g = std::min(nbatch, (long)h->num_sms * 8);
          if (SPK_EMU) g = std::min(g, 3L);
          if (h->max_ctas > 0) g = std::min(g, h->max_ctas);
          spk_dual2_kernel_t kfn = h->k_dual2;
          SPK_LAUNCH(kfn, (unsigned)g, T, h->dual2_smem, h->stream, D, h->P, A2, h->dual2_F, a, b - a);
        }
        COUNT_LAUNCH(h, PC_PROX_H);
        if (count_pn) pn_off += g;
      }
    }
    pn_parts = pn_off;
  } else {  // 4. dual blocks and residuals, all owned families in one launch (one warp per family)
    ProfScope ps(h, PC_PROX_H);
    const int T = SPK_EMU ? 1 : 32 * h->dual_warps;
    long g = (owned + h->dual_warps - 1) / h->dual_warps;
    g = std::min(g, (long)h->num_sms * 32);
    if (SPK_EMU) g = std::min(g, 4L);
    spk_fused_kernel_t kfn = h->k_dual;
    SPK_LAUNCH(kfn, (unsigned)std::max(1L, g), T, h->fused_smem_dual, h->stream, D, h->P, A);
    COUNT_LAUNCH(h, PC_PROX_H);
  }
  if (h->world > 1 && (parts & 8))  // v2 of the subtree roots is read by the replicated family above the cut in the next evaluation
    SPK_TRY(launch_exchange(h, 1, A.vbar + D.vo[1] + spk_level_first0(D.d, h->t_cut), h->m_sub));
  SPK_TRY(rt_launch_check());
  if (pnorm_slot >= 0) SPK_TRY(op_reduce(h, pn_parts, 2, h->d_pnorm, nullptr, pnorm_slot, 0));
  return SPK_OK;
}

#ifndef SPK_HOST_EMU
// replay (capturing it on first use) the kernel segment `parts` of one operator evaluation
static int graph_segment(spk_handle *h, const double *zv_in, double *zv_out, double gamma, double sigma, int pnorm_slot, int parts) {
  const auto key = std::make_tuple((const void *)zv_in, (void *)zv_out, gamma, sigma, pnorm_slot, parts);
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    if (h->graphs.size() >= 96) {  // bounded cache
      for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
      h->graphs.clear();
    }
    const int64_t l0 = h->launches;
    RT_CHECK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = op_T_fused_launch(h, zv_in, zv_out, gamma, sigma, pnorm_slot, parts);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    if (rc != SPK_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(SPK_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    spk_handle::GraphEntry ge;
    ge.kernels = h->launches - l0;
    h->launches = l0;
    const cudaError_t ie = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return fail(SPK_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
    it = h->graphs.emplace(key, ge).first;
  }
  RT_CHECK(cudaGraphLaunch(it->second.exec, h->stream));
  h->launches += it->second.kernels;
  return SPK_OK;
}
#endif

// One operator evaluation.  The launch sequence (2(N-1)+3 kernels, more with the P-norm) is captured once per (input buffer,
// output buffer, step sizes) into CUDA graphs and replayed: for small and medium trees a pass is launch-bound and a graph
// node costs a fraction of a stream launch.  With several GPUs the two kernel segments around the cut-level exchange are
// separate graphs and the NCCL all-gathers are issued on the stream between them (a collective inside a capture hung).
static int op_T_fused(spk_handle *h, const double *zv_in, double *zv_out, double gamma, double sigma, int pnorm_slot) {
  if (h->world > 1 && pnorm_slot >= 0) h->slot_dirty[pnorm_slot] = 1;  // (a graph replay does not pass through op_reduce)
#ifndef SPK_HOST_EMU
  if (h->use_graphs && !h->prof_on) {
    if (h->world == 1 || h->graph_nccl || h->peer_ok) return graph_segment(h, zv_in, zv_out, gamma, sigma, pnorm_slot, 15);
    SPK_TRY(graph_segment(h, zv_in, zv_out, gamma, sigma, pnorm_slot, 1));
    SPK_TRY(op_T_fused_launch(h, zv_in, zv_out, gamma, sigma, pnorm_slot, 2));
    SPK_TRY(graph_segment(h, zv_in, zv_out, gamma, sigma, pnorm_slot, 4));
    return op_T_fused_launch(h, zv_in, zv_out, gamma, sigma, pnorm_slot, 8);
  }
#endif
  return op_T_fused_launch(h, zv_in, zv_out, gamma, sigma, pnorm_slot);
}

// spock_dot (implicit_l.jl:485-516): d_scal[slot] = <a, P b>, P(b) = (b_z - gamma L'b_v, b_v - sigma L b_z)
static int op_pdot(spk_handle *h, const double *a_zv, const double *b_zv, double gamma, double sigma, int slot) {
  const long nz = h->D.nz, nv = h->D.nv;
  SPK_TRY(op_Lt(h, b_zv + nz, -gamma, 1.0, b_zv, h->ws_L));
  SPK_TRY(op_L(h, b_zv, -sigma, 1.0, b_zv + nz, h->ws_L + nz));
  return op_reduce(h, nz + nv, 0, a_zv, h->ws_L, slot, 0);
}

// ------------------------------------------------------------------------------------------------ loops

struct LoopState {
  double gamma, sigma, tol;
  int verbose_log;
  double r_safe = INFINITY;
  double eta = INFINITY;  // sp.jl:381
  int iter = 0, broyden_k = 0, k1 = 0, k2 = 0, fb = 0, k0 = 0;
};

// should_terminate! (cp.jl:54-123 / sp.jl:270-344).  Leaves xi1, xi2 (and xi in LOG mode) in d_scal[0..2].
static int launch_termination(spk_handle *h, const LoopState &S) {
  const long nz = h->D.nz, nv = h->D.nv;
  SPK_TRY(op_axpby(h, nz + nv, 1.0, h->zv, -1.0, h->zv_old, h->dzv));                 // delta = (z,v) - (z,v)_old
  SPK_TRY(op_Lt(h, h->dzv + nz, 1.0, -1.0 / S.gamma, h->dzv, h->xi12));               // xi_1 = L'dv - dz/gamma
  SPK_TRY(op_L(h, h->dzv, 1.0, -1.0 / S.sigma, h->dzv + nz, h->xi12 + nz));           // xi_2 = L dz - dv/sigma
  SPK_TRY(op_reduce(h, nz, 1, h->xi12, nullptr, 0, 0));
  SPK_TRY(op_reduce(h, nv, 1, h->xi12 + nz, nullptr, 1, 0, nz));
  if (S.verbose_log && h->alg == SPK_CP) {  // cp.jl:82-86
    SPK_TRY(op_Lt(h, h->xi12 + nz, 1.0, 1.0, h->xi12, h->xi));
    SPK_TRY(op_reduce(h, nz, 1, h->xi, nullptr, 2, 0));
  }
  return SPK_OK;
}
// host side of should_terminate!; expects h_scal[0..2] fetched
static bool decide_termination(spk_handle *h, const LoopState &S, double *xi, double *xi1, double *xi2) {
  *xi1 = h->h_scal[0];
  *xi2 = h->h_scal[1];
  *xi = NAN;
  bool res = false;
  if (S.verbose_log) {
    *xi = (h->alg == SPK_CP) ? h->h_scal[2] : spk_max_nn(*xi1, *xi2);  // sp.jl:304 (Julia's max propagates NaN)
    res = *xi <= S.tol;
  } else {
    const double t1 = std::max(S.tol * h->res_0[0], S.tol), t2 = std::max(S.tol * h->res_0[1], S.tol);
    res = (*xi1 <= t1) && (*xi2 <= t2);
  }
  if (h->res_0[0] == -INFINITY) h->res_0[0] = *xi1;
  if (h->res_0[1] == -INFINITY) h->res_0[1] = *xi2;
  h->last_xi1 = *xi1;
  h->last_xi2 = *xi2;
  return res;
}

// one pass of run_cp! (cp.jl:205-228) up to and including the launches of should_terminate!
static int cp_pass_launch(spk_handle *h, const LoopState &S) {
  const long nzv = h->D.nz + h->D.nv;
  if (!(h->fused_ok && !S.verbose_log)) SPK_TRY(ensure_scratch(h));
  if (h->fused_ok && !S.verbose_log) {
    // z_old is the input buffer itself and should_terminate!'s xi1, xi2 come out of the forward sweep (d_scal[0..1])
    SPK_TRY(op_T_fused(h, h->zv, h->zvbar, S.gamma, S.sigma, -1));
    std::swap(h->zv, h->zvbar);
    return SPK_OK;
  }
  SPK_TRY(rt_d2d(h->zv_old, h->zv, sizeof(double) * nzv, h->stream));  // :207-208
  if (h->fused_ok) SPK_TRY(op_T_fused(h, h->zv, h->zvbar, S.gamma, S.sigma, -1));
  else SPK_TRY(op_T(h, h->zv, h->zvbar, S.gamma, S.sigma));            // :210-211
  std::swap(h->zv, h->zvbar);                                          // :213-214 with lambda = 1
  return launch_termination(h, S);
}

// anderson! (anderson.jl:1-123) with rotating column buffers instead of physical shifts
static int anderson(spk_handle *h, int iter) {
  const long nr = h->D.nz + h->D.nv;
  double *R = h->aaR;
  auto rot = [](double **a) { double *last = a[AA_M - 1]; for (int j = AA_M - 1; j >= 1; --j) a[j] = a[j - 1]; a[0] = last; };
  rot(h->MP); rot(h->MR); rot(h->AQ);                                             // :13-24, :52-56
  SPK_TRY(rt_d2d(h->MR[0], h->drzv, sizeof(double) * nr, h->stream));             // :26-33
  SPK_TRY(op_axpby(h, nr, 1.0, h->dzv, -1.0, h->MR[0], h->MP[0]));
  for (int j = AA_M - 1; j >= 1; --j)                                             // :46-50
    for (int i = AA_M - 1; i >= 1; --i)
      if (i <= j) R[i + j * AA_M] = R[(i - 1) + (j - 1) * AA_M];
  SPK_TRY(rt_d2d(h->AQ[0], h->MR[0], sizeof(double) * nr, h->stream));            // :63-65
  if (iter > 1) {
    const int jmax = std::min(iter, AA_M);
    for (int j = 2; j <= jmax; ++j) {                                             // :78-91
      SPK_TRY(op_reduce(h, nr, 0, h->AQ[j - 1], h->AQ[0], 8, 0));
      SPK_TRY(op_reduce(h, nr, 0, h->AQ[j - 1], h->AQ[j - 1], 9, 0));
      SPK_TRY(fetch_scalars(h, 10));
      const double r = h->h_scal[8] / std::sqrt(h->h_scal[9]);
      R[0 + (j - 1) * AA_M] = r;
      SPK_TRY(op_axpby(h, nr, 1.0, h->AQ[0], -r, h->AQ[j - 1], h->AQ[0]));
    }
  }
  SPK_TRY(op_reduce(h, nr, 0, h->AQ[0], h->AQ[0], 8, 0));                         // :68-76 / :93-98
  SPK_TRY(fetch_scalars(h, 9));
  const double r = std::sqrt(h->h_scal[8]);
  SPK_TRY(op_axpby(h, nr, 1.0 / r, h->AQ[0], 0.0, nullptr, h->AQ[0]));
  R[0] = r;
  for (int j = 0; j < AA_M; ++j) SPK_TRY(op_reduce(h, nr, 0, h->AQ[j], h->rzv, 8 + j, 0));  // :111 gamma = Q' r
  SPK_TRY(fetch_scalars(h, 8 + AA_M));
  double g[AA_M];
  for (int j = 0; j < AA_M; ++j) g[j] = h->h_scal[8 + j];
  for (int j = AA_M - 1; j >= 0; --j) {                                           // :113 ldiv!(UpperTriangular)
    g[j] = g[j] / R[j + j * AA_M];
    for (int i = 0; i < j; ++i) g[i] -= R[i + j * AA_M] * g[j];
  }
  // d = -r - MP*gamma (:116-122)
  SPK_TRY(op_axpby(h, nr, -1.0, h->rzv, -g[0], h->MP[0], h->dir));
  for (int j = 1; j < AA_M; ++j) SPK_TRY(op_axpby(h, nr, 1.0, h->dir, -g[j], h->MP[j], h->dir));
  return SPK_OK;
}

// one pass of run_sp! (sp.jl:388-463); *bc = backtrack_count.  Ends with the launches of should_terminate!.
static int sp_pass_launch(spk_handle *h, LoopState &S, int *bc) {
  const long nz = h->D.nz, nv = h->D.nv, nr = nz + nv;
  const double c1 = 0.99, q = 0.99, sigma_k2 = 0.1, lambda_sp = 1.0;  // sp.jl:366-372
  if (h->fused_ok) {  // T and <r,Pr> in one pair of sweeps (r = -(zbar - z, vbar - v))
    SPK_TRY(op_T_fused(h, h->zv, h->zvbar, S.gamma, S.sigma, 4));
    SPK_TRY(op_axpby(h, nr, 1.0, h->zv, -1.0, h->zvbar, h->rzv));
  } else {
    SPK_TRY(op_T(h, h->zv, h->zvbar, S.gamma, S.sigma));                               // :389-390
    SPK_TRY(op_axpby(h, nr, 1.0, h->zv, -1.0, h->zvbar, h->rzv));                      // update_residual! :34-51
    SPK_TRY(op_pdot(h, h->rzv, h->rzv, S.gamma, S.sigma, 4));
  }
  SPK_TRY(op_axpby(h, nr, 1.0, h->rzv, -1.0, h->rzv_old, h->drzv));                    // update_delta_r! :257-268
  SPK_TRY(rt_d2d(h->zv_old, h->zv, sizeof(double) * nr, h->stream));                   // update_sy! :232-255
  SPK_TRY(rt_d2d(h->rzv_old, h->rzv, sizeof(double) * nr, h->stream));
  S.broyden_k += 1;                                                                    // spock.jl:212-224
  SPK_TRY(anderson(h, S.broyden_k));  // its last fetch also brought slot 4 (launched earlier on the stream)
  const double rnorm = std::sqrt(h->h_scal[4]);
  // Reference as shipped (sp_options == 0): tau is always 1 because perform_linesearch!'s result is discarded (sp.jl:439), so
  // the up-to-9 "backtracks" recompute the same candidate; it is evaluated once and a failed K1/K2 test reports
  // backtrack_count = 9 like the reference.  SPK_SP_LINESEARCH assigns tau <- beta tau (up to 9 different candidates);
  // SPK_SP_K0 restores the blind update whose test ends in "&& false" (sp.jl:80-106).
  const double c0 = 0.99, beta = 0.5;  // sp.jl:364,369
  const int MAX_BACKTRACK_COUNT = 8;
  double tau = 1.0;
  int backtrack_count = 0;
  bool loop = true;
  if ((h->sp_options & SPK_SP_K0) && rnorm <= c0 * S.eta) {                              // K0 :72-106
    S.eta = rnorm;
    SPK_TRY(op_axpby(h, nr, 1.0, h->zv, 1.0, h->dir, h->zv));
    loop = false;
    S.k0++;
  }
  while (loop && backtrack_count <= MAX_BACKTRACK_COUNT) {                               // :411-441
    SPK_TRY(op_axpby(h, nr, 1.0, h->zv, tau, h->dir, h->wu));                            // generate_candidate_update! :118-131
    if (h->fused_ok) {
      SPK_TRY(op_T_fused(h, h->wu, h->wubar, S.gamma, S.sigma, 5));
      SPK_TRY(op_axpby(h, nr, 1.0, h->wu, -1.0, h->wubar, h->rwu));
    } else {
      SPK_TRY(op_T(h, h->wu, h->wubar, S.gamma, S.sigma));                               // update_candidate_residual! :133-164
      SPK_TRY(op_axpby(h, nr, 1.0, h->wu, -1.0, h->wubar, h->rwu));
      SPK_TRY(op_pdot(h, h->rwu, h->rwu, S.gamma, S.sigma, 5));
    }
    SPK_TRY(fetch_scalars(h, 6));
    const double rtilde_norm = std::sqrt(h->h_scal[5]);
    if (rnorm <= S.r_safe && rtilde_norm <= c1 * rnorm) {                                // K1 :166-191
      std::swap(h->zv, h->wu);
      S.r_safe = rtilde_norm + std::pow(q, (double)S.iter);
      S.k1++;
      loop = false;
      continue;
    }
    SPK_TRY(op_axpby(h, nr, 1.0, h->rwu, -tau, h->dir, h->ws_rho));                      // :419-424
    SPK_TRY(op_pdot(h, h->rwu, h->ws_rho, S.gamma, S.sigma, 6));
    SPK_TRY(fetch_scalars(h, 7));
    double rho = h->h_scal[6];
    if (rho >= sigma_k2 * rnorm * rtilde_norm) {                                         // K2 :193-222
      rho /= rtilde_norm * rtilde_norm;
      SPK_TRY(op_axpby(h, nr, 1.0, h->zv, -(lambda_sp * rho), h->rwu, h->zv));
      S.k2++;
      loop = false;
      continue;
    }
    if (h->sp_options & SPK_SP_LINESEARCH) {
      tau *= beta;                                                                       // :224-230 with the result assigned
      backtrack_count += 1;
    } else {
      backtrack_count = MAX_BACKTRACK_COUNT + 1;  // the remaining 8 candidates are identical
    }
  }
  if (loop) {                                                                            // fallback :443-446
    std::swap(h->zv, h->zvbar);
    S.fb++;
  }
  *bc = backtrack_count;
  return launch_termination(h, S);
}

static void default_steps(const spk_handle *h, double &gamma, double &sigma) {
  if (!(sigma > 0) || !(gamma > 0)) {  // cp.jl:198-201
    sigma = 0.99 / std::sqrt(h->L_norm);
    gamma = sigma;
  }
}


// ------------------------------------------------------------------------------------------------ device-resident loops
// (spk_sp.cuh) The loop scalars live in SpkCtl on the device; the host launches passes in batches and polls `done`.

static int ctl_begin(spk_handle *h, double tol, int verbose_log, int ignore_done) {
  SpkCtl &c = h->h_ctl;
  memset(&c, 0, sizeof(c));
  c.tol = tol;
  c.res0[0] = h->res_0[0]; c.res0[1] = h->res_0[1];
  c.xi1 = c.xi2 = NAN;
  c.ignore_done = ignore_done; c.verbose_log = verbose_log; c.log_cap = h->log_cap;
  for (int k = 0; k < AA_M * AA_M; ++k) c.R[k] = h->aaR[k];
  c.r_safe = INFINITY;  // sp.jl:373
  c.rot = h->sp_rot;
  return rt_h2d(h->d_ctl, &c, sizeof(SpkCtl), h->stream);
}
static int ctl_poll(spk_handle *h) { return rt_d2h(&h->h_ctl, h->d_ctl, sizeof(SpkCtl), h->stream); }
static void ctl_end(spk_handle *h) {
  const SpkCtl &c = h->h_ctl;
  for (int k = 0; k < AA_M * AA_M; ++k) h->aaR[k] = c.R[k];
  h->sp_rot = c.rot % 3;
  h->res_0[0] = c.res0[0]; h->res_0[1] = c.res0[1];
  h->last_xi1 = c.xi1; h->last_xi2 = c.xi2;
}

static unsigned vec_grid(const spk_handle *h, long n) {
  const long blk = SPK_EMU ? 1 : SPK_SP_THREADS;
  long g = (n + blk - 1) / blk;
  return (unsigned)std::max(1L, std::min(g, (long)h->red_blocks));
}
// should_terminate! on the device for one finished pass (CP: xi1, xi2 come out of the dual kernel in d_scal[0..1])
static int launch_terminate(spk_handle *h, int xi_slot, int is_sp) {
  if (h->world > 1 && !h->peer_ok) {  // (the peer-memory exchange after the dual kernel already left the global maxima in d_scal[0..1])
    std::string cerr;
    if (!h->comm.allreduce_max(h->d_scal, 2, h->stream, cerr)) return fail(SPK_ERR_COMM, cerr);
  }
  SPK_LAUNCH_PDL(h->use_pdl, spk_k_terminate, 1, 1, 0, h->stream, h->d_ctl, h->d_scal, xi_slot, is_sp, h->d_log);
  COUNT_LAUNCH(h, PC_VEC);
  return rt_launch_check();
}

// one pass of run_sp! (sp.jl:388-463) as a fixed launch sequence without any host decision
static int sp_pass_device_launch(spk_handle *h, double gamma, double sigma) {
  const SpkDims &D = h->D;
  const long nz = D.nz, nv = D.nv, nr = nz + nv;
  SpkSpVecs V;
  V.zv = h->zv; V.zvbar = h->zvbar; V.zv_old = h->zv_old; V.dzv = h->dzv; V.rzv = h->rzv; V.rzv_old = h->rzv_old;
  V.dir = h->dir; V.wu = h->wu; V.wubar = h->wubar; V.rwu = h->rwu; V.ws_rho = h->ws_rho;
  for (int j = 0; j < AA_M; ++j) { V.MP[j] = h->MP[j]; V.MR[j] = h->MR[j]; V.AQ[j] = h->AQ[j]; }
  SpkCtl *c = h->d_ctl;
  double *scal = h->d_scal, *part = h->d_sp_partial;
  unsigned *tick = h->d_ticket;
  const unsigned g = vec_grid(h, nr), gz = vec_grid(h, nz), gv = vec_grid(h, nv);
  const int T = SPK_EMU ? 1 : SPK_SP_THREADS;
  auto kLt = spk_k_Lop_ctl<true>;
  auto kL = spk_k_Lop_ctl<false>;
  const double *nodp = nullptr;
  double *nop = nullptr;
  const int *always = nullptr;
  SPK_TRY(op_T_fused_launch(h, h->zv, h->zvbar, gamma, sigma, 4));                       // :389-390 and <r, P r>
  {
    ProfScope ps(h, PC_VEC);
    SPK_LAUNCH_PDL(h->use_pdl, spk_k_sp_a, g, T, 0, h->stream, nr, V, c, scal, part, tick);
    SPK_LAUNCH_PDL(h->use_pdl, spk_k_sp_b, g, T, 0, h->stream, nr, V, c, part, tick);
    SPK_LAUNCH_PDL(h->use_pdl, spk_k_sp_c, g, T, 0, h->stream, nr, V, c, part, tick);
    SPK_LAUNCH_PDL(h->use_pdl, spk_k_sp_d, g, T, 0, h->stream, nr, V, c, part, tick);
    SPK_LAUNCH_PDL(h->use_pdl, spk_k_sp_e, g, T, 0, h->stream, nr, V, c);
    for (int k = 0; k < 5; ++k) COUNT_LAUNCH(h, PC_VEC);
  }
  SPK_TRY(op_T_fused_launch(h, h->wu, h->wubar, gamma, sigma, 5));                       // :133-164
  {
    ProfScope ps(h, PC_VEC);
    SPK_LAUNCH_PDL(h->use_pdl, spk_k_sp_f_decide, 1, 1, 0, h->stream, c, scal);
    SPK_LAUNCH_PDL(h->use_pdl, spk_k_sp_f, g, T, 0, h->stream, nr, V, c);
    // K2: rho = <r~, P (r~ - d)> (spock_dot :485-516), only when K1 failed
    SPK_LAUNCH_PDL(h->use_pdl, kLt, gz, T, 0, h->stream, D, h->P, h->ws_rho + nz, -gamma, 1.0, h->ws_rho, nop, h->rwu, scal + 6, 0, nop, c, &c->need_k2,
               part, tick);
    SPK_LAUNCH_PDL(h->use_pdl, kL, gv, T, 0, h->stream, D, h->P, h->ws_rho, -sigma, 1.0, h->ws_rho + nz, nop, h->rwu + nz, scal + 6, 1, nop, c,
               &c->need_k2, part, tick);
    SPK_LAUNCH_PDL(h->use_pdl, spk_k_sp_g_decide, 1, 1, 0, h->stream, c, scal);
    SPK_LAUNCH_PDL(h->use_pdl, spk_k_sp_g, g, T, 0, h->stream, nr, V, c);
    // should_terminate! (sp.jl:270-344): xi1 = ||L'dv - dz/gamma||_inf, xi2 = ||L dz - dv/sigma||_inf
    SPK_LAUNCH_PDL(h->use_pdl, kLt, gz, T, 0, h->stream, D, h->P, h->dzv + nz, 1.0, -1.0 / gamma, h->dzv, nop, nodp, nop, 0, scal + 10, c, always, part,
               tick);
    SPK_LAUNCH_PDL(h->use_pdl, kL, gv, T, 0, h->stream, D, h->P, h->dzv, 1.0, -1.0 / sigma, h->dzv + nz, nop, nodp, nop, 0, scal + 11, c, always, part,
               tick);
    for (int k = 0; k < 8; ++k) COUNT_LAUNCH(h, PC_VEC);
  }
  return launch_terminate(h, 10, 1);
}
static int sp_pass_device(spk_handle *h, double gamma, double sigma) {
#ifndef SPK_HOST_EMU
  if (h->use_graphs && !h->prof_on) {
    if (h->sp_graph && (h->sp_graph_gamma != gamma || h->sp_graph_sigma != sigma)) {
      cudaGraphExecDestroy(h->sp_graph);
      h->sp_graph = nullptr;
    }
    if (!h->sp_graph) {
      const int64_t l0 = h->launches;
      RT_CHECK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      const int rc = sp_pass_device_launch(h, gamma, sigma);
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
      if (rc != SPK_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess) return fail(SPK_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
      h->sp_graph_kernels = h->launches - l0;
      h->launches = l0;
      const cudaError_t ie = cudaGraphInstantiate(&h->sp_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) return fail(SPK_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
      h->sp_graph_gamma = gamma; h->sp_graph_sigma = sigma;
    }
    RT_CHECK(cudaGraphLaunch(h->sp_graph, h->stream));
    h->launches += h->sp_graph_kernels;
    return SPK_OK;
  }
#endif
  return sp_pass_device_launch(h, gamma, sigma);
}

// passes per host poll: large problems are polled every pass (a pass is milliseconds), small ones every 8 passes
static int poll_batch(const spk_handle *h) { return (h->D.nz + h->D.nv) > (1L << 21) ? 1 : 8; }

// ------------------------------------------------------------------------------------------------ device group

static void worker_main(SpkWorker *w, int device) {
#ifndef SPK_HOST_EMU
  if (device >= 0) cudaSetDevice(device);
#else
  (void)device;
#endif
  std::unique_lock<std::mutex> lk(w->m);
  for (;;) {
    w->cv.wait(lk, [&] { return w->has_job || w->stop; });
    if (w->stop) return;
    std::function<int()> job = std::move(w->job);
    w->has_job = false;
    lk.unlock();
    const int rc = job();
    const std::string err = rc != SPK_OK ? g_last_error : std::string();  // thread-local of THIS worker
    lk.lock();
    w->rc = rc;
    w->err = err;
    w->done = true;
    w->cv.notify_all();
  }
}
static void worker_submit(SpkWorker *w, std::function<int()> job) {
  std::lock_guard<std::mutex> lk(w->m);
  w->job = std::move(job);
  w->has_job = true;
  w->done = false;
  w->cv.notify_all();
}
static int worker_wait(SpkWorker *w, std::string &err) {
  std::unique_lock<std::mutex> lk(w->m);
  w->cv.wait(lk, [&] { return w->done; });
  if (w->rc != SPK_OK) err = w->err;
  return w->rc;
}
// fn(sub-handle, rank) on every device of the group, concurrently, each on its own worker thread
template <class F>
static int group_all(const spk_handle *h, F fn) {
  SpkGroup *G = h->group;
  for (size_t g = 0; g < G->workers.size(); ++g) {
    spk_handle *sub = G->subs[g];
    worker_submit(G->workers[g], [fn, sub, g]() { return fn(sub, (int)g); });
  }
  int rc = SPK_OK;
  std::string err, e;
  for (size_t g = 0; g < G->workers.size(); ++g) {
    const int r = worker_wait(G->workers[g], e);
    if (r != SPK_OK && rc == SPK_OK) { rc = r; err = e; }
  }
  return rc == SPK_OK ? SPK_OK : fail(rc, err);
}
// fn(sub-handle 0) on rank 0's worker thread (queries and the single-GPU operator interface)
template <class F>
static int group_first(const spk_handle *h, F fn) {
  SpkGroup *G = h->group;
  spk_handle *sub = G->subs[0];
  worker_submit(G->workers[0], [fn, sub]() { return fn(sub); });
  std::string err;
  const int rc = worker_wait(G->workers[0], err);
  return rc == SPK_OK ? SPK_OK : fail(rc, err);
}
static void group_shutdown(SpkGroup *G) {
  for (SpkWorker *w : G->workers) {
    {
      std::lock_guard<std::mutex> lk(w->m);
      w->stop = true;
      w->cv.notify_all();
    }
    if (w->th.joinable()) w->th.join();
    delete w;
  }
  delete G;
}

// ------------------------------------------------------------------------------------------------ C ABI

static int check_problem(const spk_problem *p) {
  if (!p) return fail(SPK_ERR_INVALID, "problem is NULL");
  if (p->d <= 1) return fail(SPK_ERR_INVALID, "Branching factor d must be larger than 1, but is " + std::to_string(p->d) + ".");
  if (p->N < 2 || p->N > SPK_MAX_LEVELS) return fail(SPK_ERR_INVALID, "N must be in 2.." + std::to_string(SPK_MAX_LEVELS));
  if (p->nx < 1 || p->nu < 1 || p->ny < 1) return fail(SPK_ERR_INVALID, "nx, nu, ny must be positive");
  if (p->ny + 2 * p->d > SPK_MAX_LS) return fail(SPK_ERR_INVALID, "ny + 2d exceeds " + std::to_string(SPK_MAX_LS));
  if (p->n_cones < 1 || p->n_cones > SPK_MAX_CONES) return fail(SPK_ERR_INVALID, "1.." + std::to_string(SPK_MAX_CONES) + " sub-cones supported");
  if (p->algorithm != SPK_CP && p->algorithm != SPK_SP)
    return fail(SPK_ERR_INVALID, "This combination of solver options is not supported.");  // model.jl:295
  if (!p->A || !p->B || !p->sqrtQ || !p->sqrtR || !p->sqrtQN || !p->E || !p->F || !p->b || !p->cone_type || !p->cone_dim)
    return fail(SPK_ERR_INVALID, "NULL array in spk_problem");
  int sum = 0;
  for (int c = 0; c < p->n_cones; ++c) {
    if (p->cone_type[c] < 0 || p->cone_type[c] > 4 || p->cone_dim[c] < 1) return fail(SPK_ERR_INVALID, "bad sub-cone");
    sum += p->cone_dim[c];
  }
  if (sum != p->ny) return fail(SPK_ERR_INVALID, "sub-cone dimensions must add up to ny");
  if (p->factor_mode != SPK_FACTORS_PER_NODE && p->factor_mode != SPK_FACTORS_PER_STAGE)
    return fail(SPK_ERR_INVALID, "unknown factor_mode");
  if (p->world_size > 1) {
    if (!p->comm_id) return fail(SPK_ERR_INVALID, "world_size > 1 needs comm_id (spk_comm_unique_id)");
    if (p->rank < 0 || p->rank >= p->world_size) return fail(SPK_ERR_INVALID, "rank out of range");
  }
  if (p->n_rms < 0) return fail(SPK_ERR_INVALID, "n_rms must not be negative");
  if (p->mode_of_node) {
    if (p->factor_mode != SPK_FACTORS_PER_NODE) return fail(SPK_ERR_INVALID, "a mode map needs factor_mode = SPK_FACTORS_PER_NODE");
    if (p->world_size > 1 || p->n_devices > 1) return fail(SPK_ERR_INVALID, "a mode map is single-GPU");
  }
  if (p->n_devices > 1) {
    if (!p->devices) return fail(SPK_ERR_INVALID, "n_devices > 1 needs the devices array");
    if (p->world_size > 1) return fail(SPK_ERR_INVALID, "n_devices > 1 (one process) and world_size > 1 (one process per GPU) are exclusive");
    if (p->n_devices > 64) return fail(SPK_ERR_INVALID, "at most 64 devices");
  }
  double dn = 1.0;
  for (int k = 0; k < p->N; ++k) dn *= p->d;
  if (dn > 4.0e9) return fail(SPK_ERR_INVALID, "tree too large");
  return SPK_OK;
}

int spk_tree_maps(int32_t N, int32_t d, int64_t *child_first, int64_t *anc, int32_t *w, int64_t *min_index) {
  if (d <= 1) return fail(SPK_ERR_INVALID, "Branching factor d must be larger than 1, but is " + std::to_string(d) + ".");
  if (N < 1) return fail(SPK_ERR_INVALID, "N must be positive");
  int64_t dN = 1;
  for (int k = 0; k < N; ++k) dN *= d;
  const int64_t n = (dN - 1) / (d - 1), nnl = (dN / d - 1) / (d - 1);
  if (child_first)
    for (int64_t i0 = 0; i0 < nnl; ++i0) child_first[i0] = (int64_t)d * i0 + 2;          // scenario_tree.jl:82-88
  if (anc)
    for (int64_t j0 = 0; j0 < n; ++j0) anc[j0] = j0 == 0 ? 0 : (j0 - 1) / d + 1;         // :90-95
  if (w)
    for (int64_t j0 = 0; j0 < n; ++j0) w[j0] = j0 == 0 ? 0 : (int32_t)((j0 + 1) % d) + 1; // :77
  if (min_index)
    for (int t = 0; t < N; ++t) min_index[t] = spk_level_first0(d, t) + 1;               // :107
  return SPK_OK;
}

int spk_destroy(spk_handle *h) {
  if (!h) return SPK_OK;
  if (h->group) {
    group_all(h, [](spk_handle *s, int) { return spk_destroy(s); });
    group_shutdown(h->group);
    delete h;
    return SPK_OK;
  }
  rt_sync(h->stream);
  h->comm.destroy();
  for (void *p : h->allocs) rt_free(p);
#ifndef SPK_HOST_EMU
  for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
  if (h->sp_graph) cudaGraphExecDestroy(h->sp_graph);
  for (auto &s : h->prof_open) { cudaEventDestroy(s.e0); cudaEventDestroy(s.e1); }
  for (void *q : h->ipc_opened) cudaIpcCloseMemHandle(q);
  if (h->mbox) cudaFree(h->mbox);
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev_a) cudaEventDestroy(h->ev_a);
  if (h->ev_b) cudaEventDestroy(h->ev_b);
  if (h->stream) cudaStreamDestroy(h->stream);
#endif
  delete h;
  return SPK_OK;
}

static int fill_ones(spk_handle *h, double *dst, long count) {
  std::vector<double> ones((size_t)std::min(count, 1L << 20), 1.0);
  for (long off = 0; off < count; off += (long)ones.size()) {
    const long cnt = std::min((long)ones.size(), count - off);
    SPK_TRY(rt_h2d(dst + off, ones.data(), sizeof(double) * cnt, h->stream));
  }
  return SPK_OK;
}
// The iterate-sized vectors outside the fused Chambolle-Pock path, allocated on first use (see create_impl)
static int ensure_scratch(spk_handle *h) {
  if (h->scratch_ready) return SPK_OK;
  const long nr = h->D.nz + h->D.nv;
  SPK_TRY(dev_alloc(h, &h->zv_old, nr));
  SPK_TRY(dev_alloc(h, &h->dzv, nr));
  SPK_TRY(dev_alloc(h, &h->xi12, nr));
  SPK_TRY(dev_alloc(h, &h->zw, h->D.nz));
  SPK_TRY(dev_alloc(h, &h->xi, h->D.nz));
  SPK_TRY(dev_alloc(h, &h->ws_L, nr));
  SPK_TRY(fill_ones(h, h->zv_old, nr));  // z_old, v_old (cpock.jl:98-99)
  h->scratch_ready = true;
  return SPK_OK;
}

#ifndef SPK_HOST_EMU
// Mailbox of the peer-memory exchanges (spk_peer.cuh).  Between processes the ranks swap cudaIpc handles of their
// mailboxes through one NCCL all-gather and map each other's; inside a device group the front handle links the pointers
// after all sub-handles exist (group_link_peers).  Any failure on any rank leaves every rank on the NCCL path.
static int peer_setup(spk_handle *h) {
  const char *np = getenv("SPK_NO_PEER");
  if (np && np[0] == '1') return SPK_OK;
  if (h->world > SPK_PEER_MAX || h->world > 128) return SPK_OK;
  h->lenA = h->m_sub * h->D.nx;
  h->lenB = h->m_sub + 2;
  const size_t bytes = (size_t)spk_peer_mailbox_doubles(h->world, h->lenA, h->lenB) * sizeof(double);
  void *q = nullptr;
  if (cudaMalloc(&q, bytes) != cudaSuccess) { cudaGetLastError(); return SPK_OK; }
  h->mbox = (double *)q;
  h->bytes += (int64_t)bytes;
  RT_CHECK(cudaMemsetAsync(q, 0, bytes, h->stream));
  RT_CHECK(cudaStreamSynchronize(h->stream));
  h->peer_box[h->rank] = h->mbox;
  if (h->in_group) return SPK_OK;  // same process: linked by the group front
  // ---- one process per GPU: exchange IPC handles
  double ok = 1.0;
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, q) != cudaSuccess) { cudaGetLastError(); ok = 0.0; memset(&mine, 0, sizeof(mine)); }
  std::vector<cudaIpcMemHandle_t> all((size_t)h->world);
  void *stage = nullptr;
  RT_CHECK(cudaMalloc(&stage, sizeof(cudaIpcMemHandle_t) * h->world));
  RT_CHECK(cudaMemcpyAsync((char *)stage + sizeof(mine) * h->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  std::string cerr;
  if (!h->comm.allgather_bytes(stage, sizeof(mine), h->stream, cerr)) { cudaFree(stage); return fail(SPK_ERR_COMM, cerr); }
  RT_CHECK(cudaMemcpyAsync(all.data(), stage, sizeof(mine) * h->world, cudaMemcpyDeviceToHost, h->stream));
  RT_CHECK(cudaStreamSynchronize(h->stream));
  cudaFree(stage);
  for (int r = 0; r < h->world && ok > 0; ++r) {
    if (r == h->rank) continue;
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0.0; break; }
    h->ipc_opened.push_back(ptr);
    h->peer_box[r] = (double *)ptr;
  }
  // all ranks or none
  RT_CHECK(cudaMemcpyAsync(h->d_scal + 31, &ok, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (!h->comm.allreduce_sum(h->d_scal + 31, 1, h->stream, cerr)) return fail(SPK_ERR_COMM, cerr);
  double sum = 0.0;
  RT_CHECK(cudaMemcpyAsync(&sum, h->d_scal + 31, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  RT_CHECK(cudaStreamSynchronize(h->stream));
  h->peer_ok = sum > h->world - 0.5;
  return SPK_OK;
}
#endif

static int create_impl(const spk_problem *p, spk_handle *h) {
  SpkDims &D = h->D;
  memset(&D, 0, sizeof(D));
  D.N = p->N; D.d = p->d; D.nx = p->nx; D.nu = p->nu; D.ny = p->ny; D.nls = p->ny + 2 * p->d;
  long dN = 1;
  for (int k = 0; k < p->N; ++k) dN *= p->d;
  D.n = (dN - 1) / (p->d - 1);
  D.nl = dN / p->d;
  D.nnl = (dN / p->d - 1) / (p->d - 1);
  const long n = D.n, nnl = D.nnl, nl = D.nl;
  const int nx = D.nx, nu = D.nu, ny = D.ny, d = D.d;
  D.zo[0] = 0; D.zo[1] = n * nx; D.zo[2] = D.zo[1] + nnl * nu; D.zo[3] = D.zo[2] + n; D.zo[4] = D.zo[3] + (n - 1);
  D.nz = D.zo[4] + nnl * ny;                                                     // implicit_l.jl:5-15
  D.vo[0] = 0; D.vo[1] = nnl * ny; D.vo[2] = D.vo[1] + nnl; D.vo[3] = D.vo[2] + (long)nx * (n - 1);
  D.vo[4] = D.vo[3] + (long)nu * (n - 1); D.vo[5] = D.vo[4] + (n - 1); D.vo[6] = D.vo[5] + (n - 1);
  D.vo[7] = D.vo[6] + nnl * (nx + nu); D.vo[8] = D.vo[7] + nl * nx; D.vo[9] = D.vo[8] + nl; D.vo[10] = D.vo[9] + nl;
  D.nv = D.vo[10] + nl * nx;                                                     // cpock.jl:51-60
  D.n_cones = p->n_cones;
  for (int c = 0; c < p->n_cones; ++c) { D.cone_type[c] = p->cone_type[c]; D.cone_dim[c] = p->cone_dim[c]; }
  D.x_min = p->x_min; D.x_max = p->x_max; D.u_min = p->u_min; D.u_max = p->u_max;
  D.cost_dense = p->cost_dense ? 1 : 0;
  if (p->n_sqrtQ != 1 && p->n_sqrtQ != n - 1) return fail(SPK_ERR_INVALID, "n_sqrtQ must be 1 or n-1");
  if (p->n_sqrtQN != 1 && p->n_sqrtQN != nl) return fail(SPK_ERR_INVALID, "n_sqrtQN must be 1 or n_leaf");
  D.n_sqrtQ = p->n_sqrtQ; D.n_sqrtQN = p->n_sqrtQN;
  D.per_node = (p->factor_mode == SPK_FACTORS_PER_NODE) ? 1 : 0;
  h->alg = p->algorithm;
  h->L_norm = p->L_norm > 0 ? p->L_norm : (p->algorithm == SPK_SP ? 3.28 : 10.0);  // spock.jl:64 / cpock.jl:63

  std::string err;
  // The offline phase (Riccati recursion, kernel projectors) runs on the device (spk_offline_dev.cuh); SPK_HOST_OFFLINE=1
  // selects the host versions (spk_offline.cpp), which the tests use as the cross-check.
  const char *ho = getenv("SPK_HOST_OFFLINE");
  const bool host_offline = ho && ho[0] == '1' && !p->mode_of_node;  // (a mode map: the per-node recursion exists on the device only)
  const long n_rms = p->n_rms > 1 ? p->n_rms : 1;
  if (n_rms != 1 && n_rms != nnl) return fail(SPK_ERR_INVALID, "n_rms must be 1 (uniform risk measures) or n_non_leaf");
  D.rms_per_node = n_rms > 1 ? 1 : 0;
  const size_t nls2 = (size_t)D.nls * D.nls;
  if (host_offline) {
    if (!spk_offline_factors(p->N, d, nx, nu, p->A, p->B, h->table, err)) return fail(SPK_ERR_INVALID, err);
    // kernel projector (cpock.jl:71-81,125-128): one for all nodes, or one per non-leaf node with non-uniform risk measures
    h->ls_host.resize(nls2 * n_rms);
    std::vector<double> one;
    const double *prevE = nullptr, *prevF = nullptr;
    for (long i = 0; i < n_rms; ++i) {
      const double *Ei = p->E + (size_t)i * ny * d, *Fi = p->F + (size_t)i * ny * d;
      // consecutive nodes often share (E, F): reuse the projector instead of recomputing it
      if (!(prevE && memcmp(prevE, Ei, sizeof(double) * ny * d) == 0 && memcmp(prevF, Fi, sizeof(double) * ny * d) == 0))
        if (!spk_kernel_projector(ny, d, Ei, Fi, one, err)) return fail(SPK_ERR_INVALID, err);
      std::copy(one.begin(), one.end(), h->ls_host.begin() + nls2 * i);
      prevE = Ei; prevF = Fi;
    }
  } else {
    spk_table_init(p->N, d, nx, nu, h->table);
  }
  D.abk_stride = h->table.abk_stride; D.k_stride = h->table.k_stride; D.r_stride = h->table.r_stride;

  // ---- problem data
  double *tmp;
  SPK_TRY(upload(h, &tmp, p->A, (long)d * nx * nx)); h->P.A = tmp;
  SPK_TRY(upload(h, &tmp, p->B, (long)d * nx * nu)); h->P.B = tmp;
  h->B_host.assign(p->B, p->B + (size_t)d * nx * nu);
  h->P.w = nullptr;
  if (p->mode_of_node) {  // node_info[j].w (1-based) -> 0-based modes on the device
    std::vector<int> w0((size_t)n, 0);
    for (long j = 1; j < n; ++j) {
      const int w = p->mode_of_node[j];
      if (w < 1 || w > d) return fail(SPK_ERR_INVALID, "mode_of_node entries must be in 1..d");
      w0[(size_t)j] = w - 1;
    }
    void *q = nullptr;
    SPK_TRY(rt_malloc(&q, sizeof(int) * (size_t)n));
    h->allocs.push_back(q);
    h->bytes += (int64_t)(sizeof(int) * (size_t)n);
    SPK_TRY(rt_h2d(q, w0.data(), sizeof(int) * (size_t)n, h->stream));
    h->P.w = (const int *)q;
    h->custom_modes = true;
  }
  std::vector<double> Bt((size_t)d * nx * nu);
  for (int w = 0; w < d; ++w)
    for (int r = 0; r < nx; ++r)
      for (int k = 0; k < nu; ++k) Bt[(size_t)w * nx * nu + k + (size_t)r * nu] = p->B[(size_t)w * nx * nu + r + (size_t)k * nx];
  SPK_TRY(upload(h, &tmp, Bt.data(), (long)Bt.size())); h->P.Bt = tmp;
  const long qsz = D.cost_dense ? (long)nx * nx : nx, rsz = D.cost_dense ? (long)nu * nu : nu;
  SPK_TRY(upload(h, &tmp, p->sqrtQ, p->n_sqrtQ * qsz)); h->P.sqrtQ = tmp;
  SPK_TRY(upload(h, &tmp, p->sqrtR, p->n_sqrtQ * rsz)); h->P.sqrtR = tmp;
  SPK_TRY(upload(h, &tmp, p->sqrtQN, p->n_sqrtQN * qsz)); h->P.sqrtQN = tmp;
  SPK_TRY(upload(h, &tmp, p->b, n_rms * ny)); h->P.b = tmp;
  if (host_offline) {
    SPK_TRY(upload(h, &tmp, h->ls_host.data(), (long)h->ls_host.size())); h->P.ls = tmp;
  } else {  // kernel projectors on the device: one thread per risk measure (spk_k_projector)
    double *dE = nullptr, *dF = nullptr, *rows = nullptr, *dls = nullptr, *flag = nullptr;
    SPK_TRY(dev_alloc(h, &dls, (long)(nls2 * n_rms), false));
    SPK_TRY(dev_alloc(h, &flag, 1));
    void *t1 = nullptr, *t2 = nullptr, *t3 = nullptr;  // temporaries: freed right after the kernel
    const size_t eb = sizeof(double) * (size_t)n_rms * ny * d, rb = sizeof(double) * (size_t)n_rms * 2 * d * D.nls;
    SPK_TRY(rt_malloc(&t1, eb));
    if (rt_malloc(&t2, eb) != SPK_OK) { rt_free(t1); return SPK_ERR_ALLOC; }
    if (rt_malloc(&t3, rb) != SPK_OK) { rt_free(t1); rt_free(t2); return SPK_ERR_ALLOC; }
    dE = (double *)t1; dF = (double *)t2; rows = (double *)t3;
    int rc = rt_h2d(dE, p->E, eb, h->stream);
    if (rc == SPK_OK) rc = rt_h2d(dF, p->F, eb, h->stream);
    if (rc == SPK_OK) {
      const int b = SPK_EMU ? 1 : 128;
      SPK_LAUNCH(spk_k_projector, grid_for(n_rms, b), b, 0, h->stream, ny, d, n_rms, dE, dF, rows, dls, reinterpret_cast<int *>(flag));
      h->launches++;
      rc = rt_launch_check();
    }
    int bad = 0;
    if (rc == SPK_OK) rc = rt_d2h(&bad, flag, sizeof(int), h->stream);
    rt_free(t1); rt_free(t2); rt_free(t3);
    SPK_TRY(rc);
    if (bad) return fail(SPK_ERR_INVALID, "kernel projector: the risk measure has E = F = 0");
    h->P.ls = dls;
    if (n_rms == 1) {  // host copy of the one projector: spk_k_s2t_st takes it as a kernel parameter
      h->ls_host.resize(nls2);
      SPK_TRY(rt_d2h(h->ls_host.data(), dls, sizeof(double) * nls2, h->stream));
    }
  }
  SPK_TRY(dev_alloc(h, &h->x0, nx)); h->P.x0 = h->x0;

  // ---- subtree partition: levels < t_cut replicated, rank g owns m_sub subtrees below the cut
  h->rank = p->rank; h->world = p->world_size > 1 ? p->world_size : 1;
  if (h->world > 1) {
    int tc = -1;
    for (int t = 1; t <= D.N - 2; ++t)
      if (spk_level_count(d, t) % h->world == 0) { tc = t; break; }
    if (tc < 0) return fail(SPK_ERR_INVALID, "no tree level splits evenly over world_size ranks: report this tree at 1 GPU");
    h->t_cut = tc;
    h->m_sub = spk_level_count(d, tc) / h->world;
  }

  // ---- factors: per-stage table; in PER_NODE mode replicated into one slot per node of the families this rank computes
  //      (all of them on a single GPU), packed level by level
  SpkFactorTable &T = h->table;
  double *tABK, *tABKt, *tK, *tKt, *tG, *tR;
  if (h->custom_modes) {
    // ---- ricatti_offline per node (implicit_l.jl:71-95), batched on the device: one CTA per node, level by level from the
    //      leaves up; P is kept for two adjacent levels only.  The factor slots ARE the per-node arrays (slot = node).
    if (T.ABK.empty()) spk_table_init(p->N, d, nx, nu, T);
    D.abk_stride = T.abk_stride; D.k_stride = T.k_stride; D.r_stride = T.r_stride;
    double *nABK, *nABKt, *nK, *nKt, *nG, *nR, *ws = nullptr, *flag = nullptr;
    SPK_TRY(dev_alloc(h, &nABK, (n - 1) * D.abk_stride, false));
    SPK_TRY(dev_alloc(h, &nABKt, (n - 1) * D.abk_stride, false));
    SPK_TRY(dev_alloc(h, &nK, nnl * D.k_stride, false));
    SPK_TRY(dev_alloc(h, &nKt, nnl * D.k_stride, false));
    SPK_TRY(dev_alloc(h, &nG, nnl * D.k_stride, false));
    SPK_TRY(dev_alloc(h, &nR, nnl * D.r_stride, false));
    SPK_TRY(dev_alloc(h, &flag, 1));
    const long ws_stride = (spk_riccati_ws_doubles(d, nx, nu) + 1) & ~1L;
    const long widest = spk_level_count(d, D.N - 2);
    const long ctas = SPK_EMU ? std::min(widest, 4L) : std::min(widest, (long)h->num_sms * 2);
    void *wsv = nullptr, *pa = nullptr, *pb = nullptr;  // temporaries: freed after the recursion
    const size_t pbytes = sizeof(double) * (size_t)std::max(1L, spk_level_count(d, std::max(D.N - 3, 0))) * nx * nx;
    const size_t pbytes2 = sizeof(double) * (size_t)widest * nx * nx;
    SPK_TRY(rt_malloc(&wsv, sizeof(double) * (size_t)(ws_stride * ctas)));
    if (rt_malloc(&pa, pbytes2) != SPK_OK) { rt_free(wsv); return SPK_ERR_ALLOC; }
    if (rt_malloc(&pb, pbytes) != SPK_OK) { rt_free(wsv); rt_free(pa); return SPK_ERR_ALLOC; }
    ws = (double *)wsv;
    SpkRiccatiArgs R;
    R.N = p->N; R.d = d; R.nx = nx; R.nu = nu;
    R.abk_stride = D.abk_stride; R.k_stride = D.k_stride; R.r_stride = D.r_stride;
    R.A = h->P.A; R.B = h->P.B;
    R.P = nullptr; R.ABK = nABK; R.ABKt = nABKt; R.K = nK; R.Kt = nKt; R.G = nG; R.Rinv = nR; R.ws = ws;
    R.fail = reinterpret_cast<int *>(flag);
    int rc = SPK_OK;
    double *Pcur = (double *)pa, *Pnext = (double *)pb;  // Pcur: level t (written), Pnext: level t+1 (read); sizes alternate
    const double *Pchild = nullptr;                       // leaves: P = I
    for (int t = D.N - 2; t >= 0 && rc == SPK_OK; --t) {
      const long cnt = spk_level_count(d, t);
      // level N-2 (widest) goes to the large buffer, then the two buffers alternate (each level fits: counts shrink by d)
      double *Pout = ((D.N - 2 - t) % 2 == 0) ? (double *)pa : (double *)pb;
      SPK_LAUNCH(spk_k_riccati_level, (unsigned)std::max(1L, std::min(cnt, ctas)), SPK_EMU ? 1 : 256, 0, h->stream, R, t,
                 spk_level_first0(d, t), cnt, Pchild, Pout, h->P.w, ws_stride);
      h->launches++;
      rc = rt_launch_check();
      Pchild = Pout;
    }
    (void)Pcur; (void)Pnext;
    int bad = 0;
    if (rc == SPK_OK) rc = rt_d2h(&bad, flag, sizeof(int), h->stream);
    rt_free(wsv); rt_free(pa); rt_free(pb);
    SPK_TRY(rc);
    if (bad) return fail(SPK_ERR_INVALID, "Riccati recursion: I + sum B'PB is not positive definite at stage " + std::to_string(bad - 1));
    h->P.ABK = nABK; h->P.ABKt = nABKt; h->P.K = nK; h->P.Kt = nKt; h->P.G = nG; h->P.Rinv = nR;
    h->D_tab = D; h->P_tab = h->P;  // the modular kernels read the per-node slots as well
    for (int t = 0; t <= D.N - 2; ++t) h->fac_start[t] = spk_level_first0(d, t);
  } else {
  if (host_offline) {
    SPK_TRY(upload(h, &tABK, T.ABK.data(), (long)T.ABK.size()));
    SPK_TRY(upload(h, &tABKt, T.ABKt.data(), (long)T.ABKt.size()));
    SPK_TRY(upload(h, &tK, T.K.data(), (long)T.K.size()));
    SPK_TRY(upload(h, &tKt, T.Kt.data(), (long)T.Kt.size()));
    SPK_TRY(upload(h, &tG, T.G.data(), (long)T.G.size()));
    SPK_TRY(upload(h, &tR, T.Rinv.data(), (long)T.Rinv.size()));
  } else {  // ricatti_offline on the device (spk_k_riccati), then a host copy of the small table for spk_get_factors and the
            // fragment tables of the tensor-pipe kernels
    double *tP = nullptr, *ws = nullptr, *flag = nullptr;
    SPK_TRY(dev_alloc(h, &tABK, (long)T.ABK.size()));
    SPK_TRY(dev_alloc(h, &tABKt, (long)T.ABKt.size()));
    SPK_TRY(dev_alloc(h, &tK, (long)T.K.size()));
    SPK_TRY(dev_alloc(h, &tKt, (long)T.Kt.size()));
    SPK_TRY(dev_alloc(h, &tG, (long)T.G.size()));
    SPK_TRY(dev_alloc(h, &tR, (long)T.Rinv.size()));
    SPK_TRY(dev_alloc(h, &tP, (long)T.P.size()));
    SPK_TRY(dev_alloc(h, &ws, spk_riccati_ws_doubles(d, nx, nu)));
    SPK_TRY(dev_alloc(h, &flag, 1));
    SpkRiccatiArgs R;
    R.N = p->N; R.d = d; R.nx = nx; R.nu = nu;
    R.abk_stride = T.abk_stride; R.k_stride = T.k_stride; R.r_stride = T.r_stride;
    R.A = h->P.A; R.B = h->P.B;
    R.P = tP; R.ABK = tABK; R.ABKt = tABKt; R.K = tK; R.Kt = tKt; R.G = tG; R.Rinv = tR; R.ws = ws;
    R.fail = reinterpret_cast<int *>(flag);
    SPK_LAUNCH(spk_k_riccati, 1, SPK_EMU ? 1 : 256, 0, h->stream, R);
    h->launches++;
    SPK_TRY(rt_launch_check());
    int bad = 0;
    SPK_TRY(rt_d2h(&bad, flag, sizeof(int), h->stream));
    if (bad) return fail(SPK_ERR_INVALID, "Riccati recursion: I + sum B'PB is not positive definite at stage " + std::to_string(bad - 1));
    auto back = [&](std::vector<double> &dst, const double *src) { return dst.empty() ? (int)SPK_OK : rt_d2h(dst.data(), src, sizeof(double) * dst.size(), h->stream); };
    SPK_TRY(back(T.ABK, tABK)); SPK_TRY(back(T.ABKt, tABKt)); SPK_TRY(back(T.K, tK)); SPK_TRY(back(T.Kt, tKt));
    SPK_TRY(back(T.G, tG)); SPK_TRY(back(T.Rinv, tR)); SPK_TRY(back(T.P, tP));
  }
  h->P.ABK = tABK; h->P.ABKt = tABKt; h->P.K = tK; h->P.Kt = tKt; h->P.G = tG; h->P.Rinv = tR;
  h->D_tab = D; h->D_tab.per_node = 0;
  h->P_tab = h->P;
  {
    long packed = 0;
    for (int t = 0; t <= D.N - 2; ++t) {
      long lo, cnt;
      level_range(h, t, lo, cnt);
      h->fac_start[t] = packed;
      packed += cnt;
    }
    if (D.per_node) {
      double *nABK, *nABKt, *nK, *nKt, *nG, *nR;
      SPK_TRY(dev_alloc(h, &nABK, packed * d * D.abk_stride, false));
      SPK_TRY(dev_alloc(h, &nABKt, packed * d * D.abk_stride, false));
      SPK_TRY(dev_alloc(h, &nK, packed * D.k_stride, false));
      SPK_TRY(dev_alloc(h, &nKt, packed * D.k_stride, false));
      SPK_TRY(dev_alloc(h, &nG, packed * D.k_stride, false));
      SPK_TRY(dev_alloc(h, &nR, packed * D.r_stride, false));
      const int b = elem_block();
      for (int t = 0; t <= D.N - 2; ++t) {
        long lo, cnt;
        level_range(h, t, lo, cnt);
        const long f0 = h->fac_start[t];
        const unsigned g = SPK_EMU ? 8 : (unsigned)std::max(1L, std::min((long)h->num_sms * 16, (cnt * d * D.abk_stride + b - 1) / b));
        SPK_LAUNCH(spk_k_replicate_level, g, b, 0, h->stream, d, 0, t, D.abk_stride, cnt * d, tABK, nABK + f0 * d * D.abk_stride);
        SPK_LAUNCH(spk_k_replicate_level, g, b, 0, h->stream, d, 0, t, D.abk_stride, cnt * d, tABKt, nABKt + f0 * d * D.abk_stride);
        SPK_LAUNCH(spk_k_replicate_level, g, b, 0, h->stream, d, 1, t, D.k_stride, cnt, tK, nK + f0 * D.k_stride);
        SPK_LAUNCH(spk_k_replicate_level, g, b, 0, h->stream, d, 1, t, D.k_stride, cnt, tKt, nKt + f0 * D.k_stride);
        SPK_LAUNCH(spk_k_replicate_level, g, b, 0, h->stream, d, 1, t, D.k_stride, cnt, tG, nG + f0 * D.k_stride);
        SPK_LAUNCH(spk_k_replicate_level, g, b, 0, h->stream, d, 1, t, D.r_stride, cnt, tR, nR + f0 * D.r_stride);
        h->launches += 6;
      }
      SPK_TRY(rt_launch_check());
      h->P.ABK = nABK; h->P.ABKt = nABKt; h->P.K = nK; h->P.Kt = nKt; h->P.G = nG; h->P.Rinv = nR;
    }
  }
  }  // (!custom_modes)

  // ---- state (cpock.jl:83-103)
  const long nr = D.nz + D.nv;
  // The fused Chambolle-Pock path touches only (z, v), (zbar, vbar) and the sweep scratch q, d.  The other iterate-sized
  // vectors (z_old, delta, xi_1 / xi_2, xi, the workspaces of the modular operators: 4 (nz+nv) + 2 nz doubles, 8.6 GB at config 3)
  // belong to the LOG mode, the host-driven loop, SuperMann and the operator interface: ensure_scratch allocates them on first use.
  SPK_TRY(dev_alloc(h, &h->zv, nr));
  SPK_TRY(dev_alloc(h, &h->zvbar, nr));
  SPK_TRY(dev_alloc(h, &h->ric_q, n * nx));
  SPK_TRY(dev_alloc(h, &h->ric_d, std::max(nnl, 1L) * nu));
  h->red_blocks = SPK_EMU ? 8 : h->num_sms * 8;
  SPK_TRY(dev_alloc(h, &h->d_partial, h->red_blocks));
  SPK_TRY(dev_alloc(h, &h->d_scal, 32));
  SPK_TRY(dev_alloc(h, &h->d_pnorm, nnl + 64));
  {
    double *ctl_mem = nullptr, *tick = nullptr;
    SPK_TRY(dev_alloc(h, &ctl_mem, (long)(sizeof(SpkCtl) / sizeof(double)) + 2));
    h->d_ctl = reinterpret_cast<SpkCtl *>(ctl_mem);
    SPK_TRY(dev_alloc(h, &tick, 2));
    h->d_ticket = reinterpret_cast<unsigned *>(tick);
    SPK_TRY(dev_alloc(h, &h->d_sp_partial, (long)h->red_blocks * 4));
    h->log_cap = 5000;
    SPK_TRY(dev_alloc(h, &h->d_log, 3L * h->log_cap));
  }
  SPK_TRY(fused_setup(h));
  if (h->world > 1) {
    if (!h->fused_ok) return fail(SPK_ERR_INVALID, "the subtree partition needs the fused kernels (problem too large for shared memory)");
    std::string cerr;
    if (!h->comm.init(h->rank, h->world, p->comm_id, cerr)) return fail(SPK_ERR_COMM, cerr);
    // One eager round of the collectives of an operator evaluation: NCCL sets up its connections lazily on the first
    // call (allocations, IPC), which must not happen inside a stream capture.  The buffers hold zeros at this point.
    if (!h->comm.allgather_inplace(h->ric_q + spk_level_first0(d, h->t_cut) * nx, h->m_sub * nx, h->stream, cerr)) return fail(SPK_ERR_COMM, cerr);
    if (!h->comm.allgather_inplace(h->zvbar + D.nz + D.vo[1] + spk_level_first0(d, h->t_cut), h->m_sub, h->stream, cerr)) return fail(SPK_ERR_COMM, cerr);
    if (!h->comm.allreduce_max(h->d_scal, 2, h->stream, cerr)) return fail(SPK_ERR_COMM, cerr);
    SPK_TRY(rt_sync(h->stream));
    const char *gn = getenv("SPK_GRAPH_NCCL");
    h->graph_nccl = gn && gn[0] == '1';
#ifndef SPK_HOST_EMU
    SPK_TRY(peer_setup(h));
#endif
  }
  if (h->world > 1 && h->alg == SPK_SP) {
    // Entries of (z|v) this rank counts in SuperMann's global sums and maxima: those of its owned families, and of the
    // replicated families above the cut on rank 0 only.  Family i counts x_i, u_i, y_i, s_i, tau of its children, x and s
    // of its leaf children, and its dual blocks (the partition the dual kernel uses for <r, P r>); every operand of L / L'
    // at such an entry is valid on the counting rank.
    std::vector<unsigned char> mask((size_t)nr, 0);
    SpkFusedArgs R;
    fill_ranges(h, R);
    for (int k = 0; k < R.n_ranges; ++k) {
      if (k == 0 && h->t_cut > 0 && h->rank != 0) continue;
      const long lo = R.r_start[k], cnt = R.r_prefix[k + 1] - R.r_prefix[k];
      for (long i0 = lo; i0 < lo + cnt; ++i0) {
        const long c0 = (long)d * i0 + 1;
        auto setz = [&](long a, long len) { std::fill(mask.begin() + a, mask.begin() + a + len, 1); };
        auto setv = [&](long a, long len) { std::fill(mask.begin() + D.nz + a, mask.begin() + D.nz + a + len, 1); };
        setz(i0 * nx, nx); setz(D.zo[1] + i0 * nu, nu); setz(D.zo[4] + i0 * ny, ny); setz(D.zo[2] + i0, 1);
        setz(D.zo[3] + c0 - 1, d);
        setv(i0 * ny, ny); setv(D.vo[1] + i0, 1); setv(D.vo[6] + i0 * (nx + nu), nx + nu);
        setv(D.vo[2] + (c0 - 1) * nx, (long)d * nx); setv(D.vo[3] + (c0 - 1) * nu, (long)d * nu);
        setv(D.vo[4] + c0 - 1, d); setv(D.vo[5] + c0 - 1, d);
        if (c0 >= nnl) {
          const long l0 = c0 - nnl;
          setz(c0 * nx, (long)d * nx); setz(D.zo[2] + c0, d);
          setv(D.vo[7] + l0 * nx, (long)d * nx); setv(D.vo[8] + l0, d); setv(D.vo[9] + l0, d); setv(D.vo[10] + l0 * nx, (long)d * nx);
        }
      }
    }
    void *q = nullptr;
    SPK_TRY(rt_malloc(&q, (size_t)nr));
    h->allocs.push_back(q);
    h->bytes += nr;
    h->d_mask = (unsigned char *)q;
    SPK_TRY(rt_h2d(h->d_mask, mask.data(), (size_t)nr, h->stream));
  }
  h->res_0[0] = h->res_0[1] = -INFINITY;
  {
    const char *nd = getenv("SPK_HOST_LOOP");
    h->dev_loop = h->fused_ok && !(nd && nd[0] == '1') && (h->alg == SPK_CP || h->world == 1);
  }
  if (h->alg == SPK_SP) {  // spock.jl:160-202
    SPK_TRY(dev_alloc(h, &h->rzv, nr));
    SPK_TRY(dev_alloc(h, &h->rzv_old, nr));
    SPK_TRY(dev_alloc(h, &h->drzv, nr));
    SPK_TRY(dev_alloc(h, &h->dir, nr));
    SPK_TRY(dev_alloc(h, &h->wu, nr));
    SPK_TRY(dev_alloc(h, &h->wubar, nr));
    SPK_TRY(dev_alloc(h, &h->rwu, nr));
    SPK_TRY(dev_alloc(h, &h->ws_rho, nr));
    for (int j = 0; j < AA_M; ++j) {
      SPK_TRY(dev_alloc(h, &h->MP[j], nr));
      SPK_TRY(dev_alloc(h, &h->MR[j], nr));
      SPK_TRY(dev_alloc(h, &h->AQ[j], nr));
    }
    for (int k = 0; k < AA_M * AA_M; ++k) h->aaR[k] = 0.0;
    for (int k = 0; k < AA_M; ++k) h->aaR[k + k * AA_M] = 1.0;
  }
  if (h->alg == SPK_SP || !h->fused_ok) SPK_TRY(ensure_scratch(h));
  if (h->alg == SPK_SP) SPK_TRY(fill_ones(h, h->rzv_old, nr));  // rz_old, rv_old (spock.jl:187-188)
  for (int c = 0; c < PC_COUNT; ++c) { h->prof_ms[c] = 0.0; h->prof_launches[c] = 0; }
  return rt_sync(h->stream);
}

static int create_single(const spk_problem *p, spk_handle **out, bool in_group) {
  *out = nullptr;
  spk_handle *h = new spk_handle();
  h->in_group = in_group;
#ifndef SPK_HOST_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    delete h;
    return fail(SPK_ERR_CUDA, "no CUDA device available: libspock_b200 has no CPU fallback");
  }
  if (p->device >= 0) {
    if (cudaSetDevice(p->device) != cudaSuccess) { delete h; return fail(SPK_ERR_CUDA, "cudaSetDevice failed"); }
  }
  cudaGetDevice(&h->device);
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device) == cudaSuccess && sms > 0) h->num_sms = sms;
  }
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return fail(SPK_ERR_CUDA, "cudaStreamCreate failed");
  }
  cudaEventCreate(&h->ev_a);
  cudaEventCreate(&h->ev_b);
#endif
  const int rc = create_impl(p, h);
  if (rc != SPK_OK) {
    const std::string keep = g_last_error;
    spk_destroy(h);
    g_last_error = keep;
    return rc;
  }
  *out = h;
  return SPK_OK;
}

int spk_create(const spk_problem *p, spk_handle **out) {
  if (!out) return fail(SPK_ERR_INVALID, "out is NULL");
  *out = nullptr;
  SPK_TRY(check_problem(p));
  if (p->n_devices <= 1) {
    if (p->n_devices == 1 && p->devices) {
      spk_problem q = *p;
      q.device = p->devices[0];
      q.n_devices = 0;
      return create_single(&q, out, false);
    }
    return create_single(p, out, false);
  }
  // ---- device group: one sub-handle (rank of the subtree partition) and one worker thread per device
  const int G = p->n_devices;
  uint8_t id[128];
  SPK_TRY(spk_comm_unique_id(id));
  spk_handle *h = new spk_handle();
  h->group = new SpkGroup();
  h->group->subs.assign((size_t)G, nullptr);
  for (int g = 0; g < G; ++g) {
    SpkWorker *w = new SpkWorker();
    w->th = std::thread(worker_main, w, (int)p->devices[g]);
    h->group->workers.push_back(w);
  }
  const spk_problem base = *p;
  SpkGroup *grp = h->group;
  const uint8_t *idp = id;
  for (int g = 0; g < G; ++g)  // NCCL initialisation is a collective: all ranks enter create concurrently
    worker_submit(grp->workers[g], [base, grp, g, G, idp]() {
      spk_problem q = base;
      q.n_devices = 0; q.devices = nullptr;
      q.device = base.devices[g];
      q.rank = g; q.world_size = G; q.comm_id = idp;
      return create_single(&q, &grp->subs[g], true);
    });
  int rc = SPK_OK;
  std::string err, e;
  for (int g = 0; g < G; ++g) {
    const int r = worker_wait(grp->workers[g], e);
    if (r != SPK_OK && rc == SPK_OK) { rc = r; err = e; }
  }
  if (rc != SPK_OK) {
    for (int g = 0; g < G; ++g)
      if (grp->subs[g]) { spk_handle *sub = grp->subs[g]; worker_submit(grp->workers[g], [sub]() { return spk_destroy(sub); }); worker_wait(grp->workers[g], e); }
    group_shutdown(grp);
    delete h;
    return fail(rc, err);
  }
  h->D = grp->subs[0]->D;
  h->alg = grp->subs[0]->alg;
  h->world = G;
#ifndef SPK_HOST_EMU
  {  // peer-memory exchanges inside the process: enable peer access both ways and hand every sub-handle the mailboxes
    bool all = true;
    for (int g = 0; g < G; ++g) all = all && grp->subs[g]->mbox != nullptr;
    if (all) {
      std::vector<int> okv((size_t)G, 0);
      int *okp = okv.data();
      group_all(h, [grp, G, okp](spk_handle *s, int g) {
        int ok = 1;
        for (int r = 0; r < G && ok; ++r) {
          if (r == g) continue;
          int can = 0;
          if (cudaDeviceCanAccessPeer(&can, s->device, grp->subs[r]->device) != cudaSuccess || !can) { ok = 0; break; }
          const cudaError_t e = cudaDeviceEnablePeerAccess(grp->subs[r]->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
          cudaGetLastError();
        }
        okp[g] = ok;
        return (int)SPK_OK;
      });
      for (int g = 0; g < G; ++g) all = all && okv[g];
      if (all)
        for (int g = 0; g < G; ++g) {
          for (int r = 0; r < G; ++r) grp->subs[g]->peer_box[r] = grp->subs[r]->mbox;
          grp->subs[g]->peer_ok = true;
        }
    }
  }
#endif
  *out = h;
  return SPK_OK;
}

int spk_sizes(const spk_handle *h, int64_t s[5]) {
  if (!h || !s) return fail(SPK_ERR_INVALID, "NULL argument");
  s[0] = h->D.n; s[1] = h->D.nnl; s[2] = h->D.nl; s[3] = h->D.nz; s[4] = h->D.nv;
  return SPK_OK;
}
int spk_offsets(const spk_handle *h, int64_t zo[5], int64_t vo[11]) {
  if (!h) return fail(SPK_ERR_INVALID, "NULL handle");
  if (zo) for (int k = 0; k < 5; ++k) zo[k] = h->D.zo[k];
  if (vo) for (int k = 0; k < 11; ++k) vo[k] = h->D.vo[k];
  return SPK_OK;
}
int spk_device_bytes(const spk_handle *h, int64_t *b) {
  if (!h || !b) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group) {
    *b = 0;
    for (const spk_handle *s : h->group->subs) *b += s->bytes;
    return SPK_OK;
  }
  *b = h->bytes;
  return SPK_OK;
}
int spk_launch_count(const spk_handle *h, int64_t *l) {
  if (!h || !l) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group) {
    *l = 0;
    for (const spk_handle *s : h->group->subs) *l += s->launches;
    return SPK_OK;
  }
  *l = h->launches;
  return SPK_OK;
}

int spk_set_x0(spk_handle *h, const double *x0) {
  if (!h || !x0) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group) return group_all(h, [=](spk_handle *s, int) { return spk_set_x0(s, x0); });
  return rt_h2d(h->x0, x0, sizeof(double) * h->D.nx, h->stream);
}
int spk_set_iterate(spk_handle *h, const double *z, const double *v) {
  if (!h) return fail(SPK_ERR_INVALID, "NULL handle");
  if (h->group) return group_all(h, [=](spk_handle *s, int) { return spk_set_iterate(s, z, v); });
  if (z) SPK_TRY(rt_h2d(h->z(), z, sizeof(double) * h->D.nz, h->stream));
  if (v) SPK_TRY(rt_h2d(h->v(), v, sizeof(double) * h->D.nv, h->stream));
  h->gathered = z && v;  // complete host vectors make the device iterate complete on this rank
  return SPK_OK;
}
// The contiguous slices [off, off+len) of z and of v that the families of this rank write: for a
// range [a, b) of owned families the entries indexed by the family's node (u, y, v1, v2, v7), by its children (x, s, tau,
// v3..v6) and, for leaf children, by the leaf ordinal (v11..v14); plus x_1 and s_1 with the root family.
template <class F>
static void for_each_owned_slice(const spk_handle *h, F &&fn) {  // fn(bool in_v, long off, long len)
  const SpkDims &D = h->D;
  SpkFusedArgs R;
  R.gate = nullptr; R.vec2 = 0;
  fill_ranges(h, R);
  for (int k = 0; k < R.n_ranges; ++k) {
    // device group: all sub-handles copy into the SAME host vectors; the replicated top is rank 0's to write
    if (h->in_group && h->world > 1 && h->t_cut > 0 && k == 0 && h->rank != 0) continue;
    const long a = R.r_start[k], b = a + (R.r_prefix[k + 1] - R.r_prefix[k]);
    if (b <= a) continue;
    const long c0 = (long)D.d * a + 1, c1 = (long)D.d * b + 1;  // children nodes [c0, c1)
    if (a == 0) { fn(false, 0, D.nx); fn(false, D.zo[2], 1); }
    fn(false, D.zo[1] + a * D.nu, (b - a) * D.nu);
    fn(false, D.zo[4] + a * D.ny, (b - a) * D.ny);
    fn(false, c0 * D.nx, (c1 - c0) * D.nx);
    fn(false, D.zo[2] + c0, c1 - c0);
    fn(false, D.zo[3] + c0 - 1, c1 - c0);
    fn(true, D.vo[0] + a * D.ny, (b - a) * D.ny);
    fn(true, D.vo[1] + a, b - a);
    fn(true, D.vo[6] + a * (D.nx + D.nu), (b - a) * (D.nx + D.nu));
    fn(true, D.vo[2] + (c0 - 1) * D.nx, (c1 - c0) * D.nx);
    fn(true, D.vo[3] + (c0 - 1) * D.nu, (c1 - c0) * D.nu);
    fn(true, D.vo[4] + c0 - 1, c1 - c0);
    fn(true, D.vo[5] + c0 - 1, c1 - c0);
    if (c1 - 1 >= D.nnl) {  // (part of) the children are leaves; a level is either all leaves or none
      const long l0 = std::max(c0, D.nnl) - D.nnl, l1 = c1 - D.nnl;
      fn(true, D.vo[7] + l0 * D.nx, (l1 - l0) * D.nx);
      fn(true, D.vo[8] + l0, l1 - l0);
      fn(true, D.vo[9] + l0, l1 - l0);
      fn(true, D.vo[10] + l0 * D.nx, (l1 - l0) * D.nx);
    }
  }
}

int spk_get_iterate(spk_handle *h, double *z, double *v) {
  if (!h) return fail(SPK_ERR_INVALID, "NULL handle");
  // device group: every device copies the entries its families own into the SAME host vectors -- together the complete
  // iterate (solver_state.z / .v of the reference), without any device-side gather
  if (h->group) return group_all(h, [=](spk_handle *s, int) { return spk_get_iterate(s, z, v); });
  if (h->world > 1 && (!h->gathered || h->in_group)) {
    // subtree partition: only the entries this rank's families write are valid here, so only they are copied (1/world of
    // the iterate plus the replicated top); the other entries of the host vectors are left untouched.  After
    // spk_gather_iterate the whole device iterate is valid and is copied in one piece.
    int rc = SPK_OK;
#ifndef SPK_HOST_EMU
    for_each_owned_slice(h, [&](bool in_v, long off, long len) {
      double *dst = in_v ? v : z;
      if (!dst || len <= 0 || rc != SPK_OK) return;
      const cudaError_t e = cudaMemcpyAsync(dst + off, (in_v ? h->v() : h->z()) + off, sizeof(double) * len, cudaMemcpyDeviceToHost, h->stream);
      if (e != cudaSuccess) rc = fail(SPK_ERR_CUDA, std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e));
    });
#else
    for_each_owned_slice(h, [&](bool in_v, long off, long len) {
      double *dst = in_v ? v : z;
      if (dst && len > 0) memcpy(dst + off, (in_v ? h->v() : h->z()) + off, sizeof(double) * len);
    });
#endif
    SPK_TRY(rc);
    return rt_sync(h->stream);
  }
  if (z) SPK_TRY(rt_d2h(z, h->z(), sizeof(double) * h->D.nz, h->stream));
  if (v) SPK_TRY(rt_d2h(v, h->v(), sizeof(double) * h->D.nv, h->stream));
  return SPK_OK;
}

int spk_apply(spk_handle *h, int32_t op, const double *in, double *out, double scalar) {
  if (!h || !in || !out) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group) return group_first(h, [=](spk_handle *s) { return spk_apply(s, op, in, out, scalar); });
  SPK_TRY(ensure_scratch(h));
  const SpkDims &D = h->D;
  const long nz = D.nz, nv = D.nv;
  // scratch that carries nothing across passes or solves: zvbar (rewritten by every operator evaluation) and xi12
  // (dzv is NOT scratch: SuperMann's next Anderson step reads delta_z, delta_v of the previous pass)
  double *tz = h->zvbar, *tv = h->zvbar + nz, *oz = h->xi12, *ov = h->xi12 + nz;
  switch (op) {
    case SPK_OP_L:
      SPK_TRY(rt_h2d(tz, in, sizeof(double) * nz, h->stream));
      SPK_TRY(op_L(h, tz, 1.0, 0.0, nullptr, ov));
      return rt_d2h(out, ov, sizeof(double) * nv, h->stream);
    case SPK_OP_LT:
      SPK_TRY(rt_h2d(tv, in, sizeof(double) * nv, h->stream));
      SPK_TRY(op_Lt(h, tv, 1.0, 0.0, nullptr, oz));
      return rt_d2h(out, oz, sizeof(double) * nz, h->stream);
    case SPK_OP_PROX_F:
      SPK_TRY(rt_h2d(tz, in, sizeof(double) * nz, h->stream));
      SPK_TRY(op_prox_f(h, tz, scalar));
      return rt_d2h(out, tz, sizeof(double) * nz, h->stream);
    case SPK_OP_PROX_HCONJ:
      if (!(scalar > 0)) return fail(SPK_ERR_INVALID, "sigma must be positive");
      SPK_TRY(rt_h2d(tv, in, sizeof(double) * nv, h->stream));
      SPK_TRY(op_prox_hconj(h, tv, scalar, ov));
      return rt_d2h(out, ov, sizeof(double) * nv, h->stream);
    case SPK_OP_PROJ_S1:
      SPK_TRY(rt_h2d(tz, in, sizeof(double) * D.zo[2], h->stream));
      SPK_TRY(op_proj_S1(h, tz));
      return rt_d2h(out, tz, sizeof(double) * D.zo[2], h->stream);
    case SPK_OP_PROJ_S2:  // z2 = view(arg, s_inds[1]+1 : nz), implicit_l.jl:747
      SPK_TRY(rt_h2d(tz + D.zo[2] + 1, in, sizeof(double) * (nz - D.zo[2] - 1), h->stream));
      SPK_TRY(op_proj_S2(h, tz));
      return rt_d2h(out, tz + D.zo[2] + 1, sizeof(double) * (nz - D.zo[2] - 1), h->stream);
    default:
      return fail(SPK_ERR_INVALID, "unknown operator");
  }
}

// spock_mul!(model, C, trans, B, alpha, beta): C <- alpha * op(L) B + beta * C (implicit_l.jl:451-481)
int spk_mul(spk_handle *h, int32_t trans, const double *B, double alpha, double beta, double *C) {
  if (!h || !B || !C) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group) return group_first(h, [=](spk_handle *s) { return spk_mul(s, trans, B, alpha, beta, C); });
  SPK_TRY(ensure_scratch(h));
  const long nz = h->D.nz, nv = h->D.nv;
  double *tz = h->zvbar, *tv = h->zvbar + nz, *cz = h->xi12, *cv = h->xi12 + nz;  // scratch (see spk_apply)
  if (trans) {  // C (nz) <- alpha L'B + beta C, B (nv)
    SPK_TRY(rt_h2d(tv, B, sizeof(double) * nv, h->stream));
    SPK_TRY(rt_h2d(cz, C, sizeof(double) * nz, h->stream));
    SPK_TRY(op_Lt(h, tv, alpha, beta, cz, cz));
    return rt_d2h(C, cz, sizeof(double) * nz, h->stream);
  }
  SPK_TRY(rt_h2d(tz, B, sizeof(double) * nz, h->stream));
  SPK_TRY(rt_h2d(cv, C, sizeof(double) * nv, h->stream));
  SPK_TRY(op_L(h, tz, alpha, beta, cv, cv));
  return rt_d2h(C, cv, sizeof(double) * nv, h->stream);
}

int spk_dot(spk_handle *h, const double *a_z, const double *a_v, const double *b_z, const double *b_v, double gamma,
            double sigma, double *out) {
  if (!h || !a_z || !a_v || !b_z || !b_v || !out) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group) {  // every rank takes part (the scalar fetch is a collective under the partition); rank 0's value is returned
    std::vector<double> loc(h->group->subs.size(), 0.0);
    double *lp = loc.data();
    SPK_TRY(group_all(h, [=](spk_handle *s, int g) { return spk_dot(s, a_z, a_v, b_z, b_v, gamma, sigma, lp + g); }));
    *out = loc[0];
    return SPK_OK;
  }
  SPK_TRY(ensure_scratch(h));
  const long nz = h->D.nz, nv = h->D.nv;
  SPK_TRY(rt_h2d(h->zvbar, a_z, sizeof(double) * nz, h->stream));
  SPK_TRY(rt_h2d(h->zvbar + nz, a_v, sizeof(double) * nv, h->stream));
  SPK_TRY(rt_h2d(h->xi12, b_z, sizeof(double) * nz, h->stream));
  SPK_TRY(rt_h2d(h->xi12 + nz, b_v, sizeof(double) * nv, h->stream));
  SPK_TRY(op_pdot(h, h->zvbar, h->xi12, gamma, sigma, 4));
  SPK_TRY(fetch_scalars(h, 5));
  *out = h->h_scal[4];
  return SPK_OK;
}

int spk_get_factors(const spk_handle *h, int64_t i, double *Pm, double *ABK, double *K, double *Rinv, double *ls) {
  if (!h) return fail(SPK_ERR_INVALID, "NULL handle");
  if (h->group) return spk_get_factors(h->group->subs[0], i, Pm, ABK, K, Rinv, ls);  // host tables only
  const SpkDims &D = h->D;
  if (i < 1 || i > D.n) return fail(SPK_ERR_INVALID, "node index out of range");
  const long i0 = i - 1;
  int t = 0;
  while (t + 1 < D.N && i0 >= spk_level_first0(D.d, t + 1)) ++t;
  const SpkFactorTable &T = h->table;
  const size_t nx2 = (size_t)D.nx * D.nx;
  if (h->custom_modes) {  // per-node recursion: the slots of node i live on the device; P is not kept
    if (Pm) return fail(SPK_ERR_INVALID, "P is not kept with a mode map (only two levels of it exist during the recursion)");
    if (ABK && i0 >= 1) SPK_TRY(rt_d2h(ABK, h->P.ABK + (i0 - 1) * D.abk_stride, sizeof(double) * nx2, h->stream));
    if (i0 < D.nnl) {
      if (K) SPK_TRY(rt_d2h(K, h->P.K + i0 * D.k_stride, sizeof(double) * D.nu * D.nx, h->stream));
      if (Rinv) SPK_TRY(rt_d2h(Rinv, h->P.Rinv + i0 * D.r_stride, sizeof(double) * D.nu * D.nu, h->stream));
    }
  } else {
  if (Pm) std::copy(T.P.begin() + t * nx2, T.P.begin() + (t + 1) * nx2, Pm);
  if (ABK && i0 >= 1) {
    const int c = (int)((i0 - 1) % D.d);
    const double *src = T.ABK.data() + ((size_t)(t - 1) * D.d + c) * T.abk_stride;
    std::copy(src, src + nx2, ABK);
  }
  if (i0 < D.nnl) {
    if (K) std::copy(T.K.begin() + t * T.k_stride, T.K.begin() + t * T.k_stride + (size_t)D.nu * D.nx, K);
    if (Rinv) std::copy(T.Rinv.begin() + t * T.r_stride, T.Rinv.begin() + t * T.r_stride + (size_t)D.nu * D.nu, Rinv);
  }
  }
  if (ls) {
    const size_t nls2 = (size_t)D.nls * D.nls, off = (D.rms_per_node && i0 < D.nnl) ? nls2 * (size_t)i0 : 0;
    if (!h->ls_host.empty()) {
      std::copy(h->ls_host.begin() + off, h->ls_host.begin() + off + nls2, ls);
    } else {  // computed on the device (spk_k_projector): fetch this node's projector
      SPK_TRY(rt_d2h(ls, h->P.ls + off, sizeof(double) * nls2, h->stream));
    }
  }
  return SPK_OK;
}

// a peer-memory exchange gave up waiting for a peer (spk_peer.cuh): report it instead of returning a wrong iterate
static int peer_check(spk_handle *h) {
#ifndef SPK_HOST_EMU
  if (h->peer_ok) {
    unsigned long long err = 0;
    const double *w = h->mbox + 2L * h->world * (h->lenA + h->lenB) + 2L * h->world + 2;
    SPK_TRY(rt_d2h(&err, w, sizeof(err), h->stream));
    if (err) return fail(SPK_ERR_COMM, "peer-memory exchange timed out waiting for another rank");
  }
#endif
  return SPK_OK;
}

struct LoopTimer {
  spk_handle *h;
  double t0 = 0;
  explicit LoopTimer(spk_handle *h_) : h(h_) {
#ifndef SPK_HOST_EMU
    cudaEventRecord(h->ev_a, h->stream);
#else
    t0 = wall_ms();
#endif
  }
  float stop() {
#ifndef SPK_HOST_EMU
    cudaEventRecord(h->ev_b, h->stream);
    cudaEventSynchronize(h->ev_b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
    return ms;
#else
    return (float)(wall_ms() - t0);
#endif
  }
};

int spk_step(spk_handle *h, int32_t K, double gamma, double sigma, float *ms) {
  if (!h) return fail(SPK_ERR_INVALID, "NULL handle");
  if (h->group) {
    std::vector<float> loc(h->group->subs.size(), 0.f);
    float *lp = loc.data();
    SPK_TRY(group_all(h, [=](spk_handle *s, int g) { return spk_step(s, K, gamma, sigma, lp + g); }));
    if (ms) *ms = *std::max_element(loc.begin(), loc.end());
    return SPK_OK;
  }
  LoopState S;
  h->gathered = false;
  default_steps(h, gamma, sigma);
  S.gamma = gamma; S.sigma = sigma; S.tol = 0.0; S.verbose_log = 0;
  SPK_TRY(ctl_begin(h, 0.0, 0, 1));  // also clears the device-side `done` flag that gates the fused kernels
  LoopTimer timer(h);
  if (h->dev_loop && h->alg == SPK_SP && !h->sp_options) {  // SuperMann without host decisions (spk_sp.cuh)
    for (int k = 0; k < K; ++k) {
      SPK_TRY(sp_pass_device(h, gamma, sigma));
      if (h->prof_on && h->prof_open.size() > 8192) prof_harvest(h);
    }
    SPK_TRY(ctl_poll(h));
    const double r0a = h->res_0[0], r0b = h->res_0[1];
    ctl_end(h);
    h->res_0[0] = r0a; h->res_0[1] = r0b;  // spk_step leaves solve_model!'s res_0 alone
    const float t = timer.stop();
    if (ms) *ms = t;
    return peer_check(h);
  }
  for (int k = 0; k < K; ++k) {
    S.iter = k;
    int bc = 0;
    if (h->alg == SPK_SP) SPK_TRY(sp_pass_launch(h, S, &bc));
    else SPK_TRY(cp_pass_launch(h, S));
    if (h->prof_on && h->prof_open.size() > 8192) prof_harvest(h);
  }
  SPK_TRY(fetch_scalars(h, 3));
  h->last_xi1 = h->h_scal[0];
  h->last_xi2 = h->h_scal[1];
  const float t = timer.stop();
  if (ms) *ms = t;
  return peer_check(h);
}

// ---- persistent solver memory (checkpoint / resume), layout documented in include/spock_b200.h
int spk_state_size(const spk_handle *h, int64_t *doubles) {
  if (!h || !doubles) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group) return spk_state_size(h->group->subs[0], doubles);
  const int64_t nr = h->D.nz + h->D.nv;
  *doubles = h->alg == SPK_SP ? 12 * nr + AA_M * AA_M : nr;
  return SPK_OK;
}
static void state_blocks(spk_handle *h, std::vector<double *> &blocks) {
  blocks.push_back(h->zv_old);
  if (h->alg == SPK_SP) {
    blocks.push_back(h->rzv_old);
    blocks.push_back(h->dzv);
    // logical column j lives in slot (j + 2 rot) % 3 (device loop); the host-driven loop rotates the pointers, rot = 0
    for (int j = 0; j < AA_M; ++j) blocks.push_back(h->MP[(j + 2 * h->sp_rot) % 3]);
    for (int j = 0; j < AA_M; ++j) blocks.push_back(h->MR[(j + 2 * h->sp_rot) % 3]);
    for (int j = 0; j < AA_M; ++j) blocks.push_back(h->AQ[(j + 2 * h->sp_rot) % 3]);
  }
}
int spk_get_state(spk_handle *h, double *state) {
  if (!h || !state) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->world > 1) return fail(SPK_ERR_INVALID, "spk_get_state is single-GPU");
  SPK_TRY(ensure_scratch(h));
  const long nr = h->D.nz + h->D.nv;
  std::vector<double *> blocks;
  state_blocks(h, blocks);
  for (size_t k = 0; k < blocks.size(); ++k) SPK_TRY(rt_d2h(state + k * nr, blocks[k], sizeof(double) * nr, h->stream));
  if (h->alg == SPK_SP)
    for (int k = 0; k < AA_M * AA_M; ++k) state[blocks.size() * nr + k] = h->aaR[k];
  return SPK_OK;
}
int spk_set_state(spk_handle *h, const double *state) {
  if (!h || !state) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->world > 1) return fail(SPK_ERR_INVALID, "spk_set_state is single-GPU");
  SPK_TRY(ensure_scratch(h));
  const long nr = h->D.nz + h->D.nv;
  std::vector<double *> blocks;
  state_blocks(h, blocks);
  for (size_t k = 0; k < blocks.size(); ++k) SPK_TRY(rt_h2d(blocks[k], state + k * nr, sizeof(double) * nr, h->stream));
  if (h->alg == SPK_SP)
    for (int k = 0; k < AA_M * AA_M; ++k) h->aaR[k] = state[blocks.size() * nr + k];
  return SPK_OK;
}

// ---- ||L||^2: power iteration on L'L (largest eigenvalue) with the modular L / L' kernels
int spk_estimate_L_norm(spk_handle *h, int32_t iters, double *L_norm) {
  if (!h || !L_norm) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->world > 1) return fail(SPK_ERR_INVALID, "spk_estimate_L_norm is single-GPU");
  SPK_TRY(ensure_scratch(h));
  const long nz = h->D.nz;
  if (iters <= 0) iters = 50;
  // scratch that carries nothing across passes or solves
  double *w = h->zw, *Lw = h->ws_L + nz, *LtLw = h->xi;
  {  // deterministic start vector with a component along every coordinate
    std::vector<double> init((size_t)nz);
    for (long k = 0; k < nz; ++k) init[(size_t)k] = 1.0 + 0.5 * std::sin(0.7 * (double)k + 0.3);
    SPK_TRY(rt_h2d(w, init.data(), sizeof(double) * nz, h->stream));
  }
  double lam = 0.0;
  for (int it = 0; it < iters; ++it) {
    SPK_TRY(op_reduce(h, nz, 0, w, w, 8, 0));
    SPK_TRY(fetch_scalars(h, 9));
    const double nrm = std::sqrt(h->h_scal[8]);
    if (!(nrm > 0)) return fail(SPK_ERR_INVALID, "power iteration collapsed to zero");
    SPK_TRY(op_axpby(h, nz, 1.0 / nrm, w, 0.0, nullptr, w));
    SPK_TRY(op_L(h, w, 1.0, 0.0, nullptr, Lw));
    SPK_TRY(op_Lt(h, Lw, 1.0, 0.0, nullptr, LtLw));
    SPK_TRY(op_reduce(h, nz, 0, w, LtLw, 8, 0));  // Rayleigh quotient w'L'Lw with ||w|| = 1
    SPK_TRY(fetch_scalars(h, 9));
    const double next = h->h_scal[8];
    SPK_TRY(rt_d2d(w, LtLw, sizeof(double) * nz, h->stream));
    const bool done = it > 3 && std::fabs(next - lam) <= 1e-10 * next;
    lam = next;
    if (done) break;
  }
  SPK_TRY(rt_sync(h->stream));
  *L_norm = lam;
  return SPK_OK;
}
int spk_set_L_norm(spk_handle *h, double L_norm) {
  if (!h || !(L_norm > 0)) return fail(SPK_ERR_INVALID, "L_norm must be positive");
  if (h->group)
    for (spk_handle *s : h->group->subs) s->L_norm = L_norm;
  h->L_norm = L_norm;
  return SPK_OK;
}

int spk_shift_iterate(spk_handle *h, int32_t child) {
  if (!h) return fail(SPK_ERR_INVALID, "NULL handle");
  if (h->world > 1 || h->group) return fail(SPK_ERR_INVALID, "spk_shift_iterate is single-GPU");
  if (child < 0 || child >= h->D.d) return fail(SPK_ERR_INVALID, "child ordinal out of range");
  const long nr = h->D.nz + h->D.nv;
  SPK_TRY(rt_d2d(h->zvbar, h->zv, sizeof(double) * nr, h->stream));  // zvbar is scratch between evaluations
  const int b = elem_block();
  SPK_LAUNCH(spk_k_shift, grid_for(h->D.n, b), b, 0, h->stream, h->D, (int)child, h->zvbar, h->zv);
  COUNT_LAUNCH(h, PC_VEC);
  SPK_TRY(rt_launch_check());
  return rt_sync(h->stream);
}

// solve_model! for several independent models at once (MPC serving: many small problems share one GPU).  Every handle
// runs its device-resident loop on its own stream; the host launches a few passes of each, then polls each, so the
// latency-bound kernels of the different problems overlap on the device.  Handles that cannot run the device-resident
// loop (LOG mode, SuperMann options, several GPUs) are solved one after the other.
int spk_solve_batch(spk_handle **hs, int32_t n, const double *x0s, double tol, int32_t max_iter, spk_result *results) {
  if (!hs || n < 0 || !x0s) return fail(SPK_ERR_INVALID, "NULL argument");
  std::vector<int> launched((size_t)n, 0), maxit((size_t)n, 0);
  std::vector<char> active((size_t)n, 0);
  std::vector<double> gam((size_t)n, 0.0), sig((size_t)n, 0.0);
  size_t x0_off = 0;
  for (int k = 0; k < n; ++k) {
    spk_handle *h = hs[k];
    if (!h) return fail(SPK_ERR_INVALID, "NULL handle in batch");
    const double *x0 = x0s + x0_off;  // back to back: the models may have different nx
    x0_off += (size_t)h->D.nx;
    const bool dev = !h->group && h->world == 1 && h->dev_loop && !(h->alg == SPK_SP && h->sp_options);
    if (!dev) {
      SPK_TRY(spk_solve(h, x0, tol, max_iter, 0.0, 0.0, 0, nullptr, nullptr, results ? results + k : nullptr, nullptr, nullptr,
                        nullptr, nullptr, 0));
      continue;
    }
    h->gathered = false;
    SPK_TRY(spk_set_x0(h, x0));
    h->res_0[0] = h->res_0[1] = -INFINITY;
    default_steps(h, gam[k], sig[k]);
    maxit[k] = max_iter > 0 ? max_iter : (h->alg == SPK_SP ? 1000 : 5000);
    SPK_TRY(ctl_begin(h, tol, 0, 0));
#ifndef SPK_HOST_EMU
    cudaEventRecord(h->ev_a, h->stream);
#endif
    active[k] = 1;
  }
  bool any = true;
  while (any) {
    any = false;
    for (int k = 0; k < n; ++k) {  // queue a few passes of every unfinished problem ...
      if (!active[k]) continue;
      spk_handle *h = hs[k];
      LoopState S;
      S.gamma = gam[k]; S.sigma = sig[k]; S.tol = tol; S.verbose_log = 0;
      const int nb = std::min(poll_batch(h), maxit[k] - launched[k]);
      for (int b = 0; b < nb; ++b, ++launched[k]) {
        if (h->alg == SPK_SP) {
          SPK_TRY(sp_pass_device(h, gam[k], sig[k]));
        } else {
          SPK_TRY(cp_pass_launch(h, S));
          SPK_TRY(launch_terminate(h, 0, 0));
        }
      }
    }
    for (int k = 0; k < n; ++k) {  // ... then look at each one's control block
      if (!active[k]) continue;
      spk_handle *h = hs[k];
      SPK_TRY(ctl_poll(h));
      if (!h->h_ctl.done && launched[k] < maxit[k]) { any = true; continue; }
      const SpkCtl &c = h->h_ctl;
      const int passes = c.passes, converged = c.done ? 1 : 0;
      if (h->alg == SPK_CP && ((launched[k] - passes) & 1)) std::swap(h->zv, h->zvbar);  // gated passes did not move the iterate
      ctl_end(h);
      float ms = 0.f;
#ifndef SPK_HOST_EMU
      cudaEventRecord(h->ev_b, h->stream);
      cudaEventSynchronize(h->ev_b);
      cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
#endif
      if (results) {
        spk_result &r = results[k];
        r.passes = passes; r.printed_iter = converged ? passes - 1 : passes; r.converged = converged;
        r.xi1 = c.xi1; r.xi2 = c.xi2; r.xi = NAN;
        r.k1_count = c.k1; r.k2_count = c.k2; r.fallback_count = c.fb;
        r.solve_ms = ms;
      }
      active[k] = 0;
    }
  }
  return SPK_OK;
}

int spk_set_sp_options(spk_handle *h, int32_t options) {
  if (!h || (options & ~3)) return fail(SPK_ERR_INVALID, "unknown SuperMann option");
  if (h->group)
    for (spk_handle *s : h->group->subs) spk_set_sp_options(s, options);
  h->sp_options = options;
  return SPK_OK;
}

int spk_get_residuals(spk_handle *h, double out[2]) {
  if (!h || !out) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group) h = h->group->subs[0];  // the residual pair is global on every rank
  out[0] = h->last_xi1;
  out[1] = h->last_xi2;
  return SPK_OK;
}

int spk_solve(spk_handle *h, const double *x0, double tol, int32_t max_iter, double gamma, double sigma,
              int32_t verbose_log, const double *z0, const double *v0, spk_result *result, double *log_xi,
              double *log_xi1, double *log_xi2, int32_t *log_bc, int32_t log_cap) {
  if (!h || !x0) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group)  // every rank runs the same loop and takes identical decisions; rank 0 reports
    return group_all(h, [=](spk_handle *s, int g) {
      const bool r0 = g == 0;
      return spk_solve(s, x0, tol, max_iter, gamma, sigma, verbose_log, z0, v0, r0 ? result : nullptr, r0 ? log_xi : nullptr,
                       r0 ? log_xi1 : nullptr, r0 ? log_xi2 : nullptr, r0 ? log_bc : nullptr, r0 ? log_cap : 0);
    });
  if (z0 != nullptr || v0 != nullptr) {  // cpock.jl:180-183: logical-or guard, then both are copied
    if (z0 == nullptr || v0 == nullptr) return fail(SPK_ERR_INVALID, "z0 and v0 must be given together");
    SPK_TRY(spk_set_iterate(h, z0, v0));
  }
  if (h->world > 1 && verbose_log) return fail(SPK_ERR_INVALID, "verbose=LOG is single-GPU only");
  h->gathered = false;
  SPK_TRY(spk_set_x0(h, x0));                      // :185
  h->res_0[0] = h->res_0[1] = -INFINITY;           // :186
  LoopState S;
  default_steps(h, gamma, sigma);
  S.gamma = gamma; S.sigma = sigma; S.tol = tol; S.verbose_log = verbose_log;
  if (max_iter <= 0) max_iter = (h->alg == SPK_SP) ? 1000 : 5000;  // sp.jl:360 / cp.jl:190
  int iter = 0, passes = 0, converged = 0;
  double xi = NAN, xi1 = NAN, xi2 = NAN;
  SPK_TRY(ctl_begin(h, tol, verbose_log, 0));
  LoopTimer timer(h);
  // Device-resident loop: CP (SILENT; LOG needs the xi vector, cp.jl:82-86) and SP (both modes, sp.jl:304)
  if (h->dev_loop && (h->alg == SPK_SP || !verbose_log) && !(h->alg == SPK_SP && h->sp_options)) {
    const int batch = poll_batch(h);
    int launched = 0;
    while (launched < max_iter) {
      const int nb = std::min(batch, max_iter - launched);
      for (int b = 0; b < nb; ++b, ++launched) {
        if (h->alg == SPK_SP) {
          SPK_TRY(sp_pass_device(h, gamma, sigma));
        } else {
          SPK_TRY(cp_pass_launch(h, S));
          SPK_TRY(launch_terminate(h, 0, 0));
        }
        if (h->prof_on && h->prof_open.size() > 8192) prof_harvest(h);
      }
      SPK_TRY(ctl_poll(h));
      if (h->h_ctl.done) break;
    }
    const SpkCtl &c = h->h_ctl;
    passes = c.passes;
    converged = c.done ? 1 : 0;
    iter = converged ? passes - 1 : passes;
    if (h->alg == SPK_CP && ((launched - passes) & 1)) std::swap(h->zv, h->zvbar);  // gated passes did not move the iterate
    ctl_end(h);
    xi1 = c.xi1; xi2 = c.xi2;
    xi = verbose_log ? spk_max_nn(xi1, xi2) : NAN;
    S.k1 = c.k1; S.k2 = c.k2; S.fb = c.fb;
    const float ms = timer.stop();
    if ((log_xi || log_xi1 || log_xi2 || log_bc) && passes > 0) {
      const int cnt = std::min(std::min(passes, (int)log_cap), h->log_cap);
      std::vector<double> lg((size_t)3 * cnt);
      if (cnt > 0) SPK_TRY(rt_d2h(lg.data(), h->d_log, sizeof(double) * 3 * cnt, h->stream));
      for (int k = 0; k < cnt; ++k) {
        if (log_xi) log_xi[k] = verbose_log ? spk_max_nn(lg[3 * k], lg[3 * k + 1]) : NAN;
        if (log_xi1) log_xi1[k] = lg[3 * k];
        if (log_xi2) log_xi2[k] = lg[3 * k + 1];
        if (log_bc) log_bc[k] = (int32_t)lg[3 * k + 2];
      }
    }
    if (result) {
      result->passes = passes; result->printed_iter = iter; result->converged = converged;
      result->xi1 = xi1; result->xi2 = xi2; result->xi = xi;
      result->k1_count = S.k1; result->k2_count = S.k2; result->fallback_count = S.fb;
      result->solve_ms = ms;
    }
    return peer_check(h);
  }
  while (iter < max_iter) {
    int bc = 0;
    S.iter = iter;
    if (h->alg == SPK_SP) SPK_TRY(sp_pass_launch(h, S, &bc));
    else SPK_TRY(cp_pass_launch(h, S));
    SPK_TRY(fetch_scalars(h, 3));
    const bool stop = decide_termination(h, S, &xi, &xi1, &xi2);
    if (passes < log_cap) {
      if (log_xi) log_xi[passes] = xi;
      if (log_xi1) log_xi1[passes] = xi1;
      if (log_xi2) log_xi2[passes] = xi2;
      if (log_bc) log_bc[passes] = bc;
    }
    passes++;
    if (h->prof_on && h->prof_open.size() > 8192) prof_harvest(h);
    if (stop) { converged = 1; break; }
    iter += 1;
  }
  const float ms = timer.stop();
  if (result) {
    result->passes = passes; result->printed_iter = iter; result->converged = converged;
    result->xi1 = xi1; result->xi2 = xi2; result->xi = xi;
    result->k1_count = S.k1; result->k2_count = S.k2; result->fallback_count = S.fb;
    result->solve_ms = ms;
  }
  return peer_check(h);
}

int spk_profile(spk_handle *h, int32_t enable, int32_t reset) {
  if (!h) return fail(SPK_ERR_INVALID, "NULL handle");
  if (h->group) return group_all(h, [=](spk_handle *s, int) { return spk_profile(s, enable, reset); });
  prof_harvest(h);
  h->prof_on = enable != 0;
  if (reset) for (int c = 0; c < PC_COUNT; ++c) { h->prof_ms[c] = 0.0; h->prof_launches[c] = 0; }
  return SPK_OK;
}
int spk_profile_get(spk_handle *h, int32_t cls, double *ms, int64_t *launches) {
  if (!h || cls < 0 || cls >= PC_COUNT) return fail(SPK_ERR_INVALID, "bad profile class");
  if (h->group) return group_first(h, [=](spk_handle *s) { return spk_profile_get(s, cls, ms, launches); });
  prof_harvest(h);
  if (ms) *ms = h->prof_ms[cls];
  if (launches) *launches = h->prof_launches[cls];
  return SPK_OK;
}

int spk_comm_unique_id(uint8_t id[128]) {
  if (!id) return fail(SPK_ERR_INVALID, "NULL argument");
  SpkComm c;
  std::string cerr;
  if (!c.make_id(id, cerr)) return fail(SPK_ERR_COMM, cerr);
  return SPK_OK;
}

int spk_owned_ranges(const spk_handle *h, int32_t *n_ranges, int64_t *start, int64_t *count) {
  if (!h || !n_ranges) return fail(SPK_ERR_INVALID, "NULL argument");
  if (h->group) {  // the group as a whole computes every family
    *n_ranges = 1;
    if (start) start[0] = 1;
    if (count) count[0] = h->D.nnl;
    return SPK_OK;
  }
  SpkFusedArgs A;
  A.gate = nullptr; A.vec2 = 0;
  fill_ranges(h, A);
  *n_ranges = A.n_ranges;
  for (int k = 0; k < A.n_ranges; ++k) {
    if (start) start[k] = A.r_start[k] + 1;  // 1-based node numbers like the reference
    if (count) count[k] = A.r_prefix[k + 1] - A.r_prefix[k];
  }
  return SPK_OK;
}

int spk_gather_iterate(spk_handle *h) {
  if (!h) return fail(SPK_ERR_INVALID, "NULL handle");
  if (h->group) return group_all(h, [](spk_handle *s, int) { return spk_gather_iterate(s); });
  if (h->world <= 1) return SPK_OK;
  const SpkDims &D = h->D;
  double *z = h->z(), *v = h->v();
  std::string cerr;
  auto gather = [&](double *base, long first, long per_rank, long stride) -> int {  // entries [first, first + world*per_rank) x stride
    if (!h->comm.allgather_inplace(base + first * stride, per_rank * stride, h->stream, cerr)) return comm_fail(cerr);
    return SPK_OK;
  };
  for (int t = h->t_cut; t <= D.N - 2; ++t) {
    long lo, cnt;
    level_range(h, t, lo, cnt);  // this rank's families of level t: positions [rank*cnt, (rank+1)*cnt)
    const long f0 = spk_level_first0(D.d, t), c0 = spk_level_first0(D.d, t + 1), ccnt = cnt * D.d;
    // indexed by the family's node: u_i, y_i, v1_i, v2_i, v7_i
    SPK_TRY(gather(z + D.zo[1], f0, cnt, D.nu));
    SPK_TRY(gather(z + D.zo[4], f0, cnt, D.ny));
    SPK_TRY(gather(v + D.vo[0], f0, cnt, D.ny));
    SPK_TRY(gather(v + D.vo[1], f0, cnt, 1));
    SPK_TRY(gather(v + D.vo[6], f0, cnt, D.nx + D.nu));
    // indexed by the children: x_c, s_c, tau_c and (v3..v6)_c (entry c-1); leaf children: (v11..v14) by leaf ordinal
    SPK_TRY(gather(z, c0, ccnt, D.nx));
    SPK_TRY(gather(z + D.zo[2], c0, ccnt, 1));
    SPK_TRY(gather(z + D.zo[3], c0 - 1, ccnt, 1));
    SPK_TRY(gather(v + D.vo[2], c0 - 1, ccnt, D.nx));
    SPK_TRY(gather(v + D.vo[3], c0 - 1, ccnt, D.nu));
    SPK_TRY(gather(v + D.vo[4], c0 - 1, ccnt, 1));
    SPK_TRY(gather(v + D.vo[5], c0 - 1, ccnt, 1));
    if (t == D.N - 2) {
      const long l0 = c0 - D.nnl;
      SPK_TRY(gather(v + D.vo[7], l0, ccnt, D.nx));
      SPK_TRY(gather(v + D.vo[8], l0, ccnt, 1));
      SPK_TRY(gather(v + D.vo[9], l0, ccnt, 1));
      SPK_TRY(gather(v + D.vo[10], l0, ccnt, D.nx));
    }
  }
  h->gathered = true;
  return rt_sync(h->stream);
}

#ifdef SPK_HOST_EMU
// Test hook of the host-emulation build only (tests/test_mm_fragments.py): the fragment-ordered weight tables that the
// tensor-pipe S1 kernels of the CUDA build read (spk_mm_layout.h), so that the CPU test-suite can check the host-side
// fragment / permutation logic by emulating the DMMA chain.  counts = {bwd fragments per level, fwd fragments per level,
// and the layout offsets}.  Either table pointer may be NULL (size query).
#include "spk_mm_layout.h"
extern "C" int spk_emu_mm_tables(spk_handle *h, double *bwd, double *fwd, int64_t counts[16]) {
  if (!h || !counts) return fail(SPK_ERR_INVALID, "NULL argument");
  const SpkMmLayout L = spk_mm_layout(h->D.nx, h->D.nu, h->D.d);
  const int vals[16] = {L.bwd_frags, L.fwd_frags, L.KX, L.KU, L.KUP, L.NBX, L.NBU, L.bwd_sweep, L.bwd_s2, L.bwd_s3b,
                        L.bwd_s3c, L.fwd_k, L.fwd_abk, L.fwd_b, h->D.N, 0};
  for (int k = 0; k < 16; ++k) counts[k] = vals[k];
  if (bwd || fwd) {
    std::vector<double> fb, ff;
    spk_mm_fragment_tables(h->table, h->B_host.data(), fb, ff);
    if (bwd) memcpy(bwd, fb.data(), fb.size() * sizeof(double));
    if (fwd) memcpy(fwd, ff.data(), ff.size() * sizeof(double));
  }
  return SPK_OK;
}
#endif

