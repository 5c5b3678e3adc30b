// C-ABI implementation (include/dad_b200.h): layer plan, weight re-packing, TMA descriptors,
// CUDA-graph capture of one diffusion step, and the sampling loops.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "dad_b200.h"
#include "common.cuh"
#include "conv_tc.cuh"
#include "conv_t3.cuh"
#include "conv_chain.cuh"
#include "chain_host.h"
#include "launch.cuh"
#include "conv_small.cuh"
#include "kernels_f32.cuh"
#include "step_kernel.cuh"
#include "projector_build.cuh"

#define STEP_FOR_EACH_SPT(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#define SMALL_FOR_EACH(X) X(1, 2) X(2, 2) X(1, 4) X(2, 4)

using namespace dad;

namespace {

thread_local std::string g_create_error;
// Live handles, so that destroying an fp32 companion detaches it from every handle that still points at it
// (dad_set_fp32_steps) instead of leaving a dangling pointer.
static std::mutex g_handles_mutex;
static std::vector<dad_handle *> g_handles;


struct Act {
  size_t off;     // byte offset per sample
  int L, C;       // rows per sample, stored channels
};

struct ConvOp {
  std::string wname, bname, gname;  // state_dict stems: weight/bias of the conv, GroupNorm stem ("" = none)
  int ksize = 1;
  bool transposed = false;          // ConvTranspose1d weight layout (Cin, Cout, k)
  TapSel sel{};                     // raw kernel index per packed tap
  ConvGeom g{};                     // geometry with STORED channel counts
  int Cin_real = 0, Cin_store = 0, Cout_pad = 0;
  int in1 = -1, in2 = -1, out = -1, res = -1;
  int tblock = -1;                  // index into time tables
  bool head = false;
  // device data
  float *w_f32 = nullptr;
  __nv_bfloat16 *w_b16 = nullptr;
  // fp32 handles: K-major fp32 weights (TF32-rounded) for the tcgen05 kind::tf32 path (BN / GW / tcp / tmA1 / tmA2 / tmW
  // below then describe THAT launch), see setup_tc32_op
  float *w_k32 = nullptr;
  bool tc32 = false;
  float *bias = nullptr, *gamma = nullptr, *beta = nullptr;
  // bf16 path
  int BN = 0, GW = 0;
  CUtensorMap tmA1, tmA2, tmW;
  ConvTcParams tcp{};
  // v3 path (conv_t3.cuh): stride-1 convs with position-major tiles
  bool t3 = false;
  int t3_MH = 1, t3_mode = 0, t3_NS = 1, t3_smem = 0;
  CUtensorMap t3A1, t3A2, t3W, t3W2, t3R, t3O;
  ConvT3Params t3p{};
  // chain path (conv_chain.cuh): index of the launch unit this conv belongs to, or -1
  int chain = -1;
  // small-batch latency path (conv_small.cuh)
  bool small = false;
  int sm_mt = 1, sm_nt = 2, sm_cluster = 1, sm_smem = 0;
  ConvSmallParams smp{};
};

struct TimeBlock {
  std::string stem;   // "<block>.time_mlp.1"
  int C = 0;
  float *tab = nullptr;  // [S][C]
};

// One launch of conv_chain_kernel: consecutive stride-1 convolutions of one U-Net level (same length, width and
// GroupNorm shape), or a single one.
struct ChainUnit {
  std::vector<int> ops;             // indices into dad_handle::ops, in execution order
  int GW = 0, MH = 1, NS = 1, L = 0, S_t = 0;
  int smem = 0, max_clusters = 0;
  // launches with few work items (small batches) run 128-wide items instead of 256-wide ones: twice the entries, half
  // the MMA time on each one's critical path (conv_t3's "half entries").  Ring depths / shared memory of that variant:
  bool has_alt = false;
  int alt_smem = 0, alt_max_clusters = 0, alt_n_a = 0, alt_nb = 0;
  int main_n_a = 0, main_nb = 0;
  ChainArgs args{};
};

// What enqueue_unet walks: a chain (index into dad_handle::chains) or a single op on the generic kernels.
struct LaunchUnit {
  int chain = -1;
  int op = -1;
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  long long kernels = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

}  // namespace

// Batch from which the dynamics projector runs on the tensor cores (pointwise kernel + bf16x3 tcgen05 GEMM, K8) even
// though the fused SIMT kernel fits (D*D*4 B in shared memory, PointMaze D = 192).  Measured cross-over on a B200
// (tools/projector_paths.py, isolated launches): B = 4096: 19.2 vs 17.9 us, 8192: 34.4 vs 20.0, 65536: 343 vs 108,
// 262144: 2722 vs 387; below 4096 the single fused launch wins (9.6 vs 16.1 us at B = 512).
constexpr int kProjTcMinBatch = 8192;

struct dad_handle {
  dad_config cfg{};
  std::string err;
  int sm_count = 0;
  int step_ctas_per_sm = 8;   // resident step_pointwise_kernel CTAs per SM (occupancy query)
  int step_ctas_per_sm_lean = 8;
  int max_smem_optin = 0;
  bool bf16 = false;
  int time_dim = 0, D = 0, Cpad_in = 0;
  size_t elt = 4;                    // activation element size
  std::vector<Act> acts;
  size_t act_bytes_per_sample = 0;
  std::vector<ConvOp> ops;
  std::vector<TimeBlock> tblocks;
  std::vector<void *> allocs;        // everything freed in destroy
  char *arena = nullptr;
  float *d_eps = nullptr, *d_xtmp = nullptr;
  LoopState *d_ls = nullptr;
  float *d_sched[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool have_weights = false, have_sched = false;
  bool rows_t = false;               // the forward being enqueued has per-row timesteps
  long long epoch = 0;               // bumped whenever captured pointers / kernel choices go stale
  long long loop_kernels = 0;        // kernels of the last dad_loop_unet + dad_loop_step pair
  int small_max_b = 24;              // batches up to this size take the latency kernels (DAD_SMALL_MAX_B, 0 = never)
  // projector
  float *d_Nt = nullptr, *d_Nrow = nullptr, *d_q = nullptr, *d_alpha = nullptr;
  int projD = 0;
  // tensor-core projector (bf16 engines): 3-term bf16 split GEMM through conv_tc_kernel<128,0>
  __nv_bfloat16 *d_projW = nullptr, *d_split = nullptr;
  float *d_qpad = nullptr;
  int projKp = 0, projNp = 0;
  CUtensorMap tmProjA, tmProjW;
  bool proj_tc = false;
  // The fused SIMT projector (projector in shared memory) wins while the batch is small enough to be latency-bound; from
  // this batch on the tensor-core path (pointwise + bf16x3 GEMM) is used even when the fused kernel fits (measured
  // cross-over, tools/step_times.py).  force_proj_tc: measurement switch of dad_time_step_kernel (flag 0x200).
  bool step_lean = false;            // the step being enqueued has Philox noise, no gradient, no trace (step_pointwise_kernel<true>)
  int f32_math = 0;                  // dad_set_fp32_math: 0 IEEE fp32 SIMT, 1 TF32 tensor cores, 2 3xTF32
  // dad_set_fp32_steps: reverse steps with index >= fp32_min_step take eps from this fp32 handle
  dad_handle *companion = nullptr;
  int fp32_min_step = INT_MAX;
  int proj_tc_min_batch = tuning_env("DAD_PROJ_TC_MIN_B", kProjTcMinBatch);
  bool force_proj_tc = false;
  // conditions
  int n_cond = 0, cond_per_batch = 0, cond_B = 0;
  int cond_h[kMaxCond] = {0};
  float *d_cond = nullptr;
  size_t cond_cap = 0;
  // host-buffer path
  float *d_hostx = nullptr, *d_hostnoise = nullptr;
  size_t hostx_cap = 0, hostnoise_cap = 0;
  cudaStream_t cap_stream = nullptr, own_stream = nullptr, side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::map<long long, GraphEntry> graphs;
  long long launches = 0;            // total kernels launched
  long long counting = 0;            // kernels enqueued since last reset (capture accounting)
  EncodeTiledFn encode = nullptr;
  int64_t conv_flops = 0;
  // conv chains: 0 = one launch per conv (round-1 kernels), 1 = chain kernel without fusion, 2 = one launch per
  // ResidualTemporalBlock, 3 = one launch per run of blocks of a level
  int fusion = 3;
  std::vector<ChainUnit> chains;
  std::vector<LaunchUnit> units;
  unsigned *d_flags = nullptr;       // [ops][tiles_cap] tile-completion counters, zeroed by stage_x_kernel every pass
  int tiles_cap = 0;
  unsigned *d_err = nullptr;
};

namespace {

#define DAD_FAIL(h, code, ...)                              \
  do {                                                      \
    char _b[512];                                           \
    snprintf(_b, sizeof(_b), __VA_ARGS__);                  \
    (h)->err = _b;                                          \
    return (code);                                          \
  } while (0)

#define CK(h, call)                                                                              \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      DAD_FAIL(h, DAD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

template <typename T>
int dev_alloc(dad_handle *h, T **p, size_t n) {
  void *q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(n * sizeof(T), 16));
  if (e != cudaSuccess) DAD_FAIL(h, DAD_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
  h->allocs.push_back(q);
  *p = reinterpret_cast<T *>(q);
  return DAD_OK;
}

// Replace *p by a larger allocation; the old buffer is released (the caller has made sure nothing reads it any more).
template <typename T>
int dev_regrow(dad_handle *h, T **p, size_t n) {
  T *old = *p;
  int rc = dev_alloc(h, p, n);
  if (rc) return rc;
  if (old) {
    auto it = std::find(h->allocs.begin(), h->allocs.end(), (void *)old);
    if (it != h->allocs.end()) h->allocs.erase(it);
    cudaFree(old);
  }
  return DAD_OK;
}

// Set-up entry points rewrite device tables (weights, schedule, projector, conditions) that sampling work still in
// flight on ANY stream may be reading: they first wait for the device to go idle.
#define DAD_QUIESCE(h) CK(h, cudaDeviceSynchronize())

// Temporaries of a set-up call, released on every return path.
struct DevTemps {
  std::vector<void *> v;
  ~DevTemps() { for (void *p : v) cudaFree(p); }
  template <typename T>
  cudaError_t alloc(T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n * sizeof(T), 16));
    if (e == cudaSuccess) { v.push_back(q); *p = reinterpret_cast<T *>(q); }
    return e;
  }
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// NVTX range around the host-side enqueue of a library call (header-only NVTX3: free when no profiler is attached)
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// Set-up copy (host or device source) that has fully landed when the call returns.  A plain cudaMemcpy from
// pageable host memory may return while the DMA is still in flight on the legacy stream, and the library's
// kernels run on non-blocking streams that do not wait for it.
// Same for fills: cudaMemset runs asynchronously on the legacy stream, which the library's streams do not wait on.
inline cudaError_t fill_now(void *dst, int value, size_t bytes, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(dst, value, bytes, st);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(st);
}

inline cudaError_t copy_now(void *dst, const void *src, size_t bytes, cudaStream_t st) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(st);
}

int new_act(dad_handle *h, int L, int C) {
  Act a;
  a.off = h->act_bytes_per_sample;
  a.L = L;
  a.C = C;
  size_t bytes = (size_t)L * C * h->elt;
  bytes = (bytes + 255) / 256 * 256;
  h->act_bytes_per_sample += bytes;
  h->acts.push_back(a);
  return (int)h->acts.size() - 1;
}

void *act_ptr(const dad_handle *h, int id) {
  // activations are laid out act-major: [act][sample][L][C]; offset scales with the capacity
  return h->arena + h->acts[id].off * (size_t)h->cfg.max_batch;
}

// ---- plan ------------------------------------------------------------------------------------
// conv over `in1 (+ in2)` with `ksize` taps, stride 1, 'same' padding
int add_conv(dad_handle *h, const std::string &stem, int in1, int in2, int Cin_real, int Cout, int L, int ksize,
             const std::string &gn_stem, int tblock, int res) {
  ConvOp op;
  op.wname = stem + ".weight";
  op.bname = stem + ".bias";
  op.gname = gn_stem;
  op.ksize = ksize;
  op.g.C1 = h->acts[in1].C;
  op.g.C2 = in2 >= 0 ? h->acts[in2].C : 0;
  op.g.Cout = Cout;
  op.g.taps = ksize;
  for (int t = 0; t < ksize; ++t) {
    op.g.tap_off[t] = t - ksize / 2;
    op.sel.k[t] = t;
  }
  op.g.in_stride = 1;
  op.g.L_in = L;
  op.g.L_out = L;
  op.g.out_mul = 1;
  op.g.out_phase = 0;
  op.Cin_real = Cin_real;
  op.Cin_store = op.g.C1 + op.g.C2;
  op.in1 = in1;
  op.in2 = in2;
  op.out = new_act(h, L, Cout);
  op.res = res;
  op.tblock = tblock;
  h->ops.push_back(op);
  return op.out;
}

int add_time_block(dad_handle *h, const std::string &stem, int C) {
  TimeBlock tb;
  tb.stem = stem;
  tb.C = C;
  h->tblocks.push_back(tb);
  return (int)h->tblocks.size() - 1;
}

// ResidualTemporalBlock (temporal_unet.py:79-122)
int add_res_block(dad_handle *h, const std::string &name, int in1, int in2, int Cin_real, int Cout, int L) {
  const int ks = h->cfg.kernel_size;
  const int tb = add_time_block(h, name + ".time_mlp.1", Cout);
  const int h1 = add_conv(h, name + ".blocks.0.block.0", in1, in2, Cin_real, Cout, L, ks, name + ".blocks.0.block.1", tb, -1);
  int r = in1;
  if (Cin_real != Cout) r = add_conv(h, name + ".residual_conv", in1, in2, Cin_real, Cout, L, 1, "", -1, -1);
  return add_conv(h, name + ".blocks.1.block.0", h1, -1, Cout, Cout, L, ks, name + ".blocks.1.block.1", -1, r);
}

int build_plan(dad_handle *h) {
  const dad_config &c = h->cfg;
  const int T = c.transition_dim;
  int L = c.horizon;
  h->Cpad_in = (T + 63) / 64 * 64;
  const int x_act = new_act(h, L, h->bf16 ? h->Cpad_in : T);
  int cur = x_act, curC = T;
  std::vector<int> skips, skipC;
  char nm[128];
  for (int lvl = 0; lvl < c.n_levels; ++lvl) {
    const int Co = c.dim * c.dim_mults[lvl];
    snprintf(nm, sizeof(nm), "downs.%d.0", lvl);
    cur = add_res_block(h, nm, cur, -1, curC, Co, L);
    snprintf(nm, sizeof(nm), "downs.%d.1", lvl);
    cur = add_res_block(h, nm, cur, -1, Co, Co, L);
    curC = Co;
    skips.push_back(cur);
    skipC.push_back(Co);
    if (lvl < c.n_levels - 1) {
      // Downsample1d: Conv1d k3 s2 p1 (temporal_unet.py:40)
      ConvOp op;
      snprintf(nm, sizeof(nm), "downs.%d.2.conv", lvl);
      op.wname = std::string(nm) + ".weight";
      op.bname = std::string(nm) + ".bias";
      op.ksize = 3;
      op.g.C1 = h->acts[cur].C; op.g.C2 = 0; op.g.Cout = Co; op.g.taps = 3;
      for (int t = 0; t < 3; ++t) { op.g.tap_off[t] = t - 1; op.sel.k[t] = t; }
      op.g.in_stride = 2; op.g.L_in = L; op.g.L_out = L / 2; op.g.out_mul = 1; op.g.out_phase = 0;
      op.Cin_real = Co; op.Cin_store = Co; op.in1 = cur; op.out = new_act(h, L / 2, Co);
      h->ops.push_back(op);
      cur = op.out;
      L /= 2;
    }
  }
  cur = add_res_block(h, "mid_block1", cur, -1, curC, curC, L);
  cur = add_res_block(h, "mid_block2", cur, -1, curC, curC, L);
  for (int u = 0; u < c.n_levels - 1; ++u) {
    // ups[u] built from reversed(in_out[1:]) (temporal_unet.py:184-192): (dim_in, dim_out) = level n-2-u -> n-1-u
    const int lvl_out = c.n_levels - 1 - u;
    const int d_out = c.dim * c.dim_mults[lvl_out], d_in = c.dim * c.dim_mults[lvl_out - 1];
    const int skip = skips.back();
    skips.pop_back();
    snprintf(nm, sizeof(nm), "ups.%d.0", u);
    cur = add_res_block(h, nm, cur, skip, 2 * d_out, d_in, L);       // cat([x, h.pop()]) (:230)
    snprintf(nm, sizeof(nm), "ups.%d.1", u);
    cur = add_res_block(h, nm, cur, -1, d_in, d_in, L);
    // Upsample1d: ConvTranspose1d k4 s2 p1 (:51) as two 2-tap phases.
    //   even o=2j  : x[j-1] W[:,:,3] + x[j]   W[:,:,1]
    //   odd  o=2j+1: x[j]   W[:,:,2] + x[j+1] W[:,:,0]
    const int up_out = new_act(h, 2 * L, d_in);
    for (int ph = 0; ph < 2; ++ph) {
      ConvOp op;
      snprintf(nm, sizeof(nm), "ups.%d.2.conv", u);
      op.wname = std::string(nm) + ".weight";
      op.bname = std::string(nm) + ".bias";
      op.ksize = 4;
      op.transposed = true;
      op.g.C1 = h->acts[cur].C; op.g.C2 = 0; op.g.Cout = d_in; op.g.taps = 2;
      if (ph == 0) { op.g.tap_off[0] = -1; op.sel.k[0] = 3; op.g.tap_off[1] = 0; op.sel.k[1] = 1; }
      else         { op.g.tap_off[0] = 0;  op.sel.k[0] = 2; op.g.tap_off[1] = 1; op.sel.k[1] = 0; }
      op.g.in_stride = 1; op.g.L_in = L; op.g.L_out = L; op.g.out_mul = 2; op.g.out_phase = ph;
      op.Cin_real = d_in; op.Cin_store = d_in; op.in1 = cur; op.out = up_out;
      h->ops.push_back(op);
    }
    cur = up_out;
    curC = d_in;
    L *= 2;
  }
  (void)skipC;
  // final_conv: Conv1dBlock(dim, dim, k) -> Conv1d(dim, T, 1) (:194-197)
  cur = add_conv(h, "final_conv.0.block.0", cur, -1, c.dim, c.dim, L, c.kernel_size, "final_conv.0.block.1", -1, -1);
  {
    ConvOp op;
    op.wname = "final_conv.1.weight";
    op.bname = "final_conv.1.bias";
    op.ksize = 1;
    op.g.C1 = h->acts[cur].C; op.g.C2 = 0; op.g.Cout = T; op.g.taps = 1; op.g.tap_off[0] = 0; op.sel.k[0] = 0;
    op.g.in_stride = 1; op.g.L_in = L; op.g.L_out = L; op.g.out_mul = 1; op.g.out_phase = 0;
    op.Cin_real = c.dim; op.Cin_store = c.dim; op.in1 = cur; op.out = -1; op.head = true;
    h->ops.push_back(op);
  }
  if (L != c.horizon) DAD_FAIL(h, DAD_ERR_INVALID, "internal: decoder length %d != horizon %d", L, c.horizon);
  h->conv_flops = 0;
  for (const ConvOp &op : h->ops)
    h->conv_flops += 2LL * op.g.L_out * op.g.taps * op.Cin_real * op.g.Cout;
  return DAD_OK;
}

// ---- bf16 path set-up --------------------------------------------------------------------------
int make_tmap(dad_handle *h, CUtensorMap *m, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
              const cuuint32_t *box) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DAD_FAIL(h, DAD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DAD_OK;
}

int make_act_tmap(dad_handle *h, CUtensorMap *m, int act, const ConvGeom &g) {
  const Act &a = h->acts[act];
  const int P = g.in_stride, J = a.L / P;
  cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)P, (cuuint64_t)J, (cuuint64_t)h->cfg.max_batch};
  cuuint64_t strides[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)P * a.C * 2, (cuuint64_t)a.L * a.C * 2};
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, 1, (cuuint32_t)g.L_out, (cuuint32_t)(TC_BM / g.L_out)};
  return make_tmap(h, m, act_ptr(h, act), 4, dims, strides, box);
}

int setup_tc_op(dad_handle *h, ConvOp &op) {
  const ConvGeom &g = op.g;
  if (g.L_out > TC_BM || TC_BM % g.L_out != 0)
    DAD_FAIL(h, DAD_ERR_INVALID, "bf16 mode needs every level length to divide 128 (got %d); use fp32 mode", g.L_out);
  if (g.C1 % TC_BK || g.C2 % TC_BK)
    DAD_FAIL(h, DAD_ERR_INVALID, "bf16 mode needs channel counts that are multiples of 64 (got %d,%d); use fp32 mode", g.C1, g.C2);
  if (op.head) {
    op.GW = 0;
    op.BN = g.Cout <= 16 ? 16 : g.Cout <= 32 ? 32 : g.Cout <= 64 ? 64 : 128;
    if (g.Cout > 128) DAD_FAIL(h, DAD_ERR_INVALID, "bf16 mode supports transition_dim <= 128 (got %d)", g.Cout);
  } else if (!op.gname.empty()) {
    const int gw = g.Cout / kGroups;
    if (g.Cout % kGroups) DAD_FAIL(h, DAD_ERR_INVALID, "channels %d not divisible by 8 groups", g.Cout);
    if (gw == 8 && g.Cout == 64) { op.BN = 64; op.GW = 8; }
    else if ((gw == 16 || gw == 32 || gw == 64 || gw == 128) && g.Cout % 128 == 0) { op.BN = 128; op.GW = gw; }
    else if (gw == 256) { op.BN = 256; op.GW = 256; }
    else DAD_FAIL(h, DAD_ERR_INVALID, "bf16 mode: unsupported GroupNorm width %d for %d channels; use fp32 mode", gw, g.Cout);
  } else {
    op.GW = 0;
    if (g.Cout % 128 == 0) op.BN = 128;
    else if (g.Cout % 64 == 0) op.BN = 64;
    else DAD_FAIL(h, DAD_ERR_INVALID, "bf16 mode: unsupported channel count %d; use fp32 mode", g.Cout);
  }
  op.Cout_pad = cdiv(g.Cout, op.BN) * op.BN;
  ConvTcParams &p = op.tcp;
  p.L_out = g.L_out;
  p.out_mul = g.out_mul;
  p.out_phase = g.out_phase;
  p.Cout = g.Cout;
  p.n_tiles_n = op.Cout_pad / op.BN;
  p.kch1 = g.C1 / TC_BK;
  p.kch2 = g.C2 / TC_BK;
  p.taps = g.taps;
  for (int t = 0; t < g.taps; ++t) {
    const int s = g.in_stride, off = g.tap_off[t];
    const int ph = ((off % s) + s) % s;
    p.tap_p[t] = ph;
    p.tap_j[t] = (off - ph) / s;
  }
  p.out_f32 = op.head ? 1 : 0;
  p.debug = tuning_env("DAD_TC_DEBUG", 0);
  p.prof = nullptr;
  p.ls = h->d_ls;
  return DAD_OK;
}

int finish_tc_op(dad_handle *h, ConvOp &op) {
  // tensor maps need the final arena / weight addresses
  int rc = make_act_tmap(h, &op.tmA1, op.in1, op.g);
  if (rc) return rc;
  rc = make_act_tmap(h, &op.tmA2, op.in2 >= 0 ? op.in2 : op.in1, op.g);
  if (rc) return rc;
  const cuuint64_t K = (cuuint64_t)op.g.taps * op.Cin_store;
  cuuint64_t dims[2] = {K, (cuuint64_t)op.Cout_pad};
  cuuint64_t strides[1] = {K * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)op.BN};
  rc = make_tmap(h, &op.tmW, op.w_b16, 2, dims, strides, box);
  if (rc) return rc;
  ConvTcParams &p = op.tcp;
  p.bias = op.bias;
  p.gamma = op.gamma;
  p.beta = op.beta;
  p.ttab = op.tblock >= 0 ? h->tblocks[op.tblock].tab : nullptr;
  p.residual = op.res >= 0 ? reinterpret_cast<const __nv_bfloat16 *>(act_ptr(h, op.res)) : nullptr;
  p.out = op.head ? (void *)h->d_eps : act_ptr(h, op.out);
  return DAD_OK;
}

// ---- TF32 tensor-core path of an fp32-precision handle (dad_set_fp32_math(h, 1)) ------------------------------
// The fp32 sibling of a bf16 model evaluates ONE U-Net pass per plan (the ill-conditioned leading reverse step); with
// fp32 activations and TF32 operands on tcgen05 (conv_tc_kernel<BN, GW, true>) that pass costs a few bf16 steps' worth.
// Layers whose channel counts are not multiples of 64 (the first conv, the head) stay on the SIMT kernel.
int make_tmap_f32(dad_handle *h, CUtensorMap *m, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                  const cuuint32_t *box) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, base, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DAD_FAIL(h, DAD_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
  return DAD_OK;
}

bool setup_tc32_op(dad_handle *h, ConvOp &op) {
  const ConvGeom &g = op.g;
  if (op.head || g.C1 % 64 || g.C2 % 64 || g.L_out > TC_BM || TC_BM % g.L_out) return false;
  if (!op.gname.empty()) {
    const int gw = g.Cout / kGroups;
    if (g.Cout % kGroups) return false;
    if (gw == 8 && g.Cout == 64) { op.BN = 64; op.GW = 8; }
    else if ((gw == 16 || gw == 32 || gw == 64 || gw == 128) && g.Cout % 128 == 0) { op.BN = 128; op.GW = gw; }
    else if (gw == 256) { op.BN = 256; op.GW = 256; }
    else return false;
  } else {
    op.GW = 0;
    if (g.Cout % 128 == 0) op.BN = 128;
    else if (g.Cout % 64 == 0) op.BN = 64;
    else return false;
  }
  if (g.Cout % op.BN) return false;
  ConvTcParams &p = op.tcp;
  p = ConvTcParams{};
  p.L_out = g.L_out;
  p.out_mul = g.out_mul;
  p.out_phase = g.out_phase;
  p.Cout = g.Cout;
  p.n_tiles_n = g.Cout / op.BN;
  p.kch1 = g.C1 / 32;          // k blocks of 32 fp32 channels (one 128-byte swizzle row)
  p.kch2 = g.C2 / 32;
  p.taps = g.taps;
  for (int t = 0; t < g.taps; ++t) {
    const int s = g.in_stride, off = g.tap_off[t];
    const int ph = ((off % s) + s) % s;
    p.tap_p[t] = ph;
    p.tap_j[t] = (off - ph) / s;
  }
  p.ls = h->d_ls;
  return true;
}

int finish_tc32_op(dad_handle *h, ConvOp &op) {
  const ConvGeom &g = op.g;
  auto act_map = [&](CUtensorMap *m, int act) {
    const Act &a = h->acts[act];
    const int P = g.in_stride, J = a.L / P;
    cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)P, (cuuint64_t)J, (cuuint64_t)h->cfg.max_batch};
    cuuint64_t strides[3] = {(cuuint64_t)a.C * 4, (cuuint64_t)P * a.C * 4, (cuuint64_t)a.L * a.C * 4};
    cuuint32_t box[4] = {32, 1, (cuuint32_t)g.L_out, (cuuint32_t)(TC_BM / g.L_out)};
    return make_tmap_f32(h, m, act_ptr(h, act), 4, dims, strides, box);
  };
  int rc = act_map(&op.tmA1, op.in1);
  if (rc) return rc;
  if ((rc = act_map(&op.tmA2, op.in2 >= 0 ? op.in2 : op.in1))) return rc;
  const cuuint64_t K = (cuuint64_t)g.taps * (g.C1 + g.C2);
  cuuint64_t dims[2] = {K, (cuuint64_t)g.Cout};
  cuuint64_t strides[1] = {K * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)op.BN};
  if ((rc = make_tmap_f32(h, &op.tmW, op.w_k32, 2, dims, strides, box))) return rc;
  ConvTcParams &p = op.tcp;
  p.bias = op.bias;
  p.gamma = op.gamma;
  p.beta = op.beta;
  p.ttab = op.tblock >= 0 ? h->tblocks[op.tblock].tab : nullptr;
  p.residual = op.res >= 0 ? reinterpret_cast<const __nv_bfloat16 *>(act_ptr(h, op.res)) : nullptr;   // fp32 data (TF32 kernel)
  p.out = act_ptr(h, op.out);
  return DAD_OK;
}

// ---- v3 path ---------------------------------------------------------------------------------------
int make_tmap_raw(dad_handle *h, CUtensorMap *m, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                  const cuuint32_t *box, const char *what) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DAD_FAIL(h, DAD_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return DAD_OK;
}

// (channel, sample, position) view of a channels-last activation: position-major boxes
int make_t3_act_tmap(dad_handle *h, CUtensorMap *m, int act, int box_s, int box_l, const char *what) {
  const Act &a = h->acts[act];
  cuuint64_t dims[3] = {(cuuint64_t)a.C, (cuuint64_t)h->cfg.max_batch, (cuuint64_t)a.L};
  cuuint64_t strides[2] = {(cuuint64_t)a.L * a.C * 2, (cuuint64_t)a.C * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_s, (cuuint32_t)box_l};
  return make_tmap_raw(h, m, act_ptr(h, act), 3, dims, strides, box, what);
}

bool t3_eligible(const dad_handle *h, const ConvOp &op) {
  const ConvGeom &g = op.g;
  if (tuning_env("DAD_T3", 1) == 0) return false;
  if (op.head || op.transposed || g.in_stride != 1 || g.out_mul != 1) return false;
  if (g.Cout % T3_BN || g.C1 % 64 || g.C2 % 64) return false;
  if (!(g.L_out == 4 || g.L_out == 8 || g.L_out == 16 || g.L_out == 32)) return false;
  if (!op.gname.empty()) {
    const int gw = g.Cout / kGroups;
    if (!(gw == 16 || gw == 32 || gw == 64 || gw == 128)) return false;
  }
  (void)h;
  return true;
}

int setup_t3_op(dad_handle *h, ConvOp &op) {
  const ConvGeom &g = op.g;
  ConvT3Params &p = op.t3p;
  const int L = g.L_out;
  op.t3_MH = (L == 32) ? 2 : 1;
  const int S_t = 128 * op.t3_MH / L;
  int lo = 0, hi = 0;
  for (int t = 0; t < g.taps; ++t) { lo = std::min(lo, g.tap_off[t]); hi = std::max(hi, g.tap_off[t]); }
  p.halo_lo = -lo;
  const int box_l = L + hi - lo;
  p.a_tx_bytes = box_l * S_t * 128;
  p.a_stage_bytes = (p.a_tx_bytes + 1023) / 1024 * 1024;
  for (int t = 0; t < g.taps; ++t) p.tap_row[t] = (g.tap_off[t] - lo) * S_t;
  p.L = L;
  p.S_t = S_t;
  p.Cout = g.Cout;
  p.kch1 = g.C1 / 64;
  p.kch2 = g.C2 / 64;
  p.taps = g.taps;
  p.has_res = op.res >= 0 ? 1 : 0;
  p.debug = tuning_env("DAD_TC_DEBUG", 0);
  p.prof = nullptr;
  p.ls = h->d_ls;
  // cooperation mode: CTA pairs (cta_group::2) by default, 256-wide items where the layer allows it
  op.t3_mode = tuning_env("DAD_T3_MODE", T3_PAIR);
  if (op.t3_mode < 0 || op.t3_mode > 2) op.t3_mode = T3_PAIR;
  const int want_ns = tuning_env("DAD_T3_NS", 2);
  op.t3_NS = (op.t3_mode == T3_PAIR && op.t3_MH == 1 && g.Cout % 256 == 0 && want_ns == 2) ? 2 : 1;
  p.n_tiles_n = g.Cout / (T3_BN * op.t3_NS);
  p.b_stage_bytes = op.t3_mode == T3_PAIR ? op.t3_NS * 8192 : 16384;
  int na = 4;
  for (; na >= 2; --na) {
    op.t3_smem = t3_smem_layout(p.a_stage_bytes, na, p.b_stage_bytes, g.Cout, S_t, op.GW).total;
    if (op.t3_smem <= h->max_smem_optin) break;
  }
  if (na < 2) return DAD_ERR_INVALID;     // caller falls back to the generic path
  p.n_a_stages = na;
  return DAD_OK;
}

int finish_t3_op(dad_handle *h, ConvOp &op) {
  const ConvGeom &g = op.g;
  ConvT3Params &p = op.t3p;
  const int box_l = p.a_tx_bytes / (p.S_t * 128);
  int rc = make_t3_act_tmap(h, &op.t3A1, op.in1, p.S_t, box_l, "activation");
  if (rc) return rc;
  if ((rc = make_t3_act_tmap(h, &op.t3A2, op.in2 >= 0 ? op.in2 : op.in1, p.S_t, box_l, "activation 2"))) return rc;
  const cuuint64_t K = (cuuint64_t)g.taps * op.Cin_store;
  cuuint64_t dims[2] = {K, (cuuint64_t)op.Cout_pad};
  cuuint64_t strides[1] = {K * 2};
  const int wrows = op.t3_mode == T3_SINGLE ? 128 : op.t3_mode == T3_MCAST ? 64 : 64 * op.t3_NS;
  cuuint32_t box[2] = {64, (cuuint32_t)wrows};
  if ((rc = make_tmap_raw(h, &op.t3W, op.w_b16, 2, dims, strides, box, "weights"))) return rc;
  cuuint32_t box2[2] = {64, 64};          // half entries of 256-wide items: 64 rows per CTA
  if ((rc = make_tmap_raw(h, &op.t3W2, op.w_b16, 2, dims, strides, box2, "weights (half)"))) return rc;
  const int pph = 128 / p.S_t;
  if ((rc = make_t3_act_tmap(h, &op.t3O, op.out, p.S_t, pph, "output"))) return rc;
  if ((rc = make_t3_act_tmap(h, &op.t3R, op.res >= 0 ? op.res : op.out, p.S_t, pph, "residual"))) return rc;
  p.bias = op.bias;
  p.gamma = op.gamma;
  p.beta = op.beta;
  p.ttab = op.tblock >= 0 ? h->tblocks[op.tblock].tab : nullptr;
  return DAD_OK;
}

template <int GW, int MH, int MODE, int NS>
cudaError_t set_t3_attr(int max_optin) {
  return cudaFuncSetAttribute(conv_t3_kernel<GW, MH, MODE, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
}

template <int GW, int MH, int MODE, int NS>
int launch_t3(dad_handle *h, const ConvOp &op, const ConvT3Params &p, int grid, cudaStream_t st) {
  cudaError_t e = launch_k(conv_t3_kernel<GW, MH, MODE, NS>, dim3((unsigned)grid), dim3(t3_threads(GW)), (size_t)op.t3_smem, st,
                           MODE == T3_SINGLE ? 1 : 2, op.t3A1, op.t3A2, op.t3W, op.t3W2, op.t3R, op.t3O, p);
  if (e != cudaSuccess) DAD_FAIL(h, DAD_ERR_CUDA, "conv_t3 launch failed: %s", cudaGetErrorString(e));
  return DAD_OK;
}

#define T3_FOR_EACH_SHAPE(X, gw) X(gw, 1, 0, 1) X(gw, 2, 0, 1) X(gw, 1, 1, 1) X(gw, 2, 1, 1) X(gw, 1, 2, 1) X(gw, 2, 2, 1) X(gw, 1, 2, 2)
#define T3_FOR_EACH(X) T3_FOR_EACH_SHAPE(X, 0) T3_FOR_EACH_SHAPE(X, 16) T3_FOR_EACH_SHAPE(X, 32) T3_FOR_EACH_SHAPE(X, 64) T3_FOR_EACH_SHAPE(X, 128)

int enqueue_t3(dad_handle *h, const ConvOp &op, int B, cudaStream_t st) {
  ConvT3Params p = op.t3p;
  p.B = B;
  p.n_mst = cdiv(B, p.S_t);
  const int CL = op.t3_mode == T3_SINGLE ? 1 : 2;
  const int items = cdiv(p.n_mst, CL) * p.n_tiles_n;
  // 256-wide items: when the last round would leave at most half of the clusters busy, its items are split
  // into two 128-wide half entries (a half entry costs ~0.75 of a whole one, so only then does it pay)
  const int n_cl = h->sm_count / CL;
  const int rem = items % n_cl;
  p.split_tail = (op.t3_NS == 2 && rem > 0 && 2 * rem <= n_cl && !(p.debug & 128)) ? 1 : 0;
  const int grid = CL * std::min(p.split_tail && items < n_cl ? 2 * items : items, n_cl);
  int rc = DAD_ERR_INVALID;
#define T3_CASE(gw, mh, mode, ns) if (op.GW == gw && op.t3_MH == mh && op.t3_mode == mode && op.t3_NS == ns) rc = launch_t3<gw, mh, mode, ns>(h, op, p, grid, st);
  T3_FOR_EACH(T3_CASE)
#undef T3_CASE
  if (rc == DAD_ERR_INVALID) DAD_FAIL(h, DAD_ERR_INVALID, "internal: no conv_t3 instantiation for GW=%d MH=%d mode=%d NS=%d", op.GW, op.t3_MH, op.t3_mode, op.t3_NS);
  h->counting += 1;
  return rc;
}

// ---- conv chains (conv_chain.cuh) -------------------------------------------------------------------
// A stride-1 conv qualifies when the CTA-pair kernel can run it: whole 128-column tiles, 64-channel K blocks, a
// sequence length that packs whole samples into 128-row half tiles, and a GroupNorm width the epilogue implements
// (a plain 1x1 residual conv borrows the width of its block: same C_out).
struct ChainShape {
  int GW, MH, NS, S_t;
};

bool chain_shape(const dad_handle *h, const ConvOp &op, ChainShape *out) {
  const ConvGeom &g = op.g;
  if (!h->bf16 || op.head || op.transposed || g.in_stride != 1 || g.out_mul != 1) return false;
  if (g.Cout % CH_BN || g.C1 % 64 || g.C2 % 64 || g.Cout % kGroups) return false;
  if (!(g.L_out == 4 || g.L_out == 8 || g.L_out == 16 || g.L_out == 32)) return false;
  const int gw = g.Cout / kGroups;
  if (!(gw == 16 || gw == 32 || gw == 64 || gw == 128 || gw == 256)) return false;
  ChainShape s;
  s.GW = gw;
  s.MH = g.L_out == 32 ? 2 : 1;
  s.NS = (s.MH == 1 && g.Cout % 256 == 0) ? 2 : 1;
  s.S_t = 128 * s.MH / g.L_out;
  if (gw == 256 && s.NS != 2) return false;       // a 256-column group needs 256-wide items (L <= 16)
  *out = s;
  return true;
}

// "downs.0.1.blocks.0.block.0.weight" / "downs.0.1.residual_conv.weight" -> "downs.0.1": the ResidualTemporalBlock
std::string block_stem(const ConvOp &op) {
  size_t pos = op.wname.find(".blocks.");
  if (pos == std::string::npos) pos = op.wname.find(".residual_conv");
  if (pos == std::string::npos) pos = op.wname.find(".block.");
  return pos == std::string::npos ? op.wname : op.wname.substr(0, pos);
}

// Shared memory and ring depths of a chain run with `ns`-wide items; false when even the smallest configuration does
// not fit.
bool chain_config_ns(const dad_handle *h, const ChainUnit &cu, int ns, int *a_stage_out, int *n_a, int *nb, int *smem) {
  int a_stage = 0;
  for (int oi : cu.ops) {
    const ConvGeom &g = h->ops[oi].g;
    int lo = 0, hi = 0;
    for (int t = 0; t < g.taps; ++t) { lo = std::min(lo, g.tap_off[t]); hi = std::max(hi, g.tap_off[t]); }
    const int bytes = (cu.L + hi - lo) * cu.S_t * 128;
    a_stage = std::max(a_stage, (bytes + 1023) / 1024 * 1024);
  }
  const int b_stage = ns * 8192;
  // deepest weight ring first (the MMA issue rate depends on it), then as many activation stages as still fit
  static const int combos[][2] = {{4, 6}, {3, 6}, {2, 6}, {3, 5}, {2, 5}, {3, 4}, {2, 4}, {2, 3}};
  for (auto &c : combos) {
    const ChSmem lay = ch_smem_layout(a_stage, c[0], b_stage, c[1], cu.S_t, cu.GW);
    if (lay.total <= h->max_smem_optin) {
      *a_stage_out = a_stage;
      *n_a = c[0];
      *nb = c[1];
      *smem = lay.total;
      return true;
    }
  }
  return false;
}

bool chain_config(const dad_handle *h, ChainUnit &cu) {
  int a_stage, n_a, nb, smem;
  if (!chain_config_ns(h, cu, cu.NS, &a_stage, &n_a, &nb, &smem)) return false;
  ChainParams &p = cu.args.p;
  p.a_stage_bytes = a_stage;
  cu.main_n_a = n_a;
  cu.main_nb = nb;
  cu.smem = smem;
  cu.has_alt = false;
  if (cu.NS == 2 && cu.GW != 256 && chain_config_ns(h, cu, 1, &a_stage, &cu.alt_n_a, &cu.alt_nb, &cu.alt_smem)) cu.has_alt = true;
  return true;
}

// Tensor maps, scalars and dependency counters of every conv of the chain.
int finish_chain(dad_handle *h, ChainUnit &cu, int unit_index) {
  const int UC = ch_unit_cols(cu.GW);
  ChainParams &p = cu.args.p;
  p.ls = h->d_ls;
  p.err = h->d_err;
  p.n_convs = (int)cu.ops.size();
  p.flag_epoch = 1;
  p.L = cu.L;
  p.S_t = cu.S_t;
  p.n_tiles_n = h->ops[cu.ops[0]].g.Cout / (CH_BN * cu.NS);
  p.w_alt = 0;
  p.debug = 0;
  // which act is produced by which conv of this chain
  std::map<int, int> producer;
  for (size_t k = 0; k < cu.ops.size(); ++k) producer[h->ops[cu.ops[k]].out] = (int)k;
  std::vector<bool> consumed(cu.ops.size(), false);
  for (size_t k = 0; k < cu.ops.size(); ++k) {
    ConvOp &op = h->ops[cu.ops[k]];
    const ConvGeom &g = op.g;
    ChainConv &cv = cu.args.convs[k];
    ChainConvMeta &m = cv.m;
    int lo = 0, hi = 0;
    for (int t = 0; t < g.taps; ++t) { lo = std::min(lo, g.tap_off[t]); hi = std::max(hi, g.tap_off[t]); }
    const int box_l = cu.L + hi - lo;
    int rc;
    if ((rc = make_t3_act_tmap(h, &cv.tmA1, op.in1, cu.S_t, box_l, "chain activation"))) return rc;
    if ((rc = make_t3_act_tmap(h, &cv.tmA2, op.in2 >= 0 ? op.in2 : op.in1, cu.S_t, box_l, "chain activation 2"))) return rc;
    const cuuint64_t K = (cuuint64_t)g.taps * op.Cin_store;
    cuuint64_t dims[2] = {K, (cuuint64_t)op.Cout_pad};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(64 * cu.NS)};         // each CTA of the pair keeps half of the item's rows
    if ((rc = make_tmap_raw(h, &cv.tmW, op.w_b16, 2, dims, strides, box, "chain weights"))) return rc;
    cuuint32_t box1[2] = {64, 64};
    if ((rc = make_tmap_raw(h, &cv.tmW1, op.w_b16, 2, dims, strides, box1, "chain weights (128-wide items)"))) return rc;
    const int pph = 128 / cu.S_t;
    if ((rc = make_t3_act_tmap(h, &cv.tmO, op.out, cu.S_t, pph, "chain output"))) return rc;
    if ((rc = make_t3_act_tmap(h, &cv.tmR, op.res >= 0 ? op.res : op.out, cu.S_t, pph, "chain residual"))) return rc;
    m.bias = op.bias;
    m.gamma = op.gamma;
    m.beta = op.beta;
    m.ttab = op.tblock >= 0 ? h->tblocks[op.tblock].tab : nullptr;
    m.Cout = g.Cout;
    m.kch1 = g.C1 / 64;
    m.kch2 = g.C2 / 64;
    m.taps = g.taps;
    m.halo_lo = -lo;
    m.tap_first = (g.tap_off[0] - lo) * cu.S_t * 8;              // rows * 128 B in 16-byte units
    m.tap_step = g.taps > 1 ? (g.tap_off[1] - g.tap_off[0]) * cu.S_t * 8 : 0;
    for (int t = 2; t < g.taps; ++t)
      if (g.tap_off[t] - g.tap_off[t - 1] != g.tap_off[1] - g.tap_off[0])
        DAD_FAIL(h, DAD_ERR_INVALID, "internal: chain conv %s has unevenly spaced taps", op.wname.c_str());
    m.a_tx_bytes = box_l * cu.S_t * 128;
    m.has_res = op.res >= 0 ? 1 : 0;
    m.plain = op.gname.empty() ? 1 : 0;
    m.flag_out = nullptr;
    m.flag_a = nullptr;
    m.flag_r = nullptr;
    m.need_a = m.need_r = 0;
    auto dep = [&](int act, const unsigned **flag, int *need) {
      auto it = producer.find(act);
      if (act < 0 || it == producer.end() || it->second >= (int)k) return;
      const ConvOp &src = h->ops[cu.ops[it->second]];
      *flag = h->d_flags + (size_t)cu.ops[it->second] * h->tiles_cap;
      *need = (src.g.Cout / UC) * cu.MH;                          // stores that complete one tile of the producer
      consumed[it->second] = true;
    };
    dep(op.in1, &m.flag_a, &m.need_a);
    dep(op.res, &m.flag_r, &m.need_r);
    if (op.in2 >= 0 && producer.count(op.in2) && producer[op.in2] < (int)k)
      DAD_FAIL(h, DAD_ERR_INVALID, "internal: chain conv %s takes its second source from the same chain", op.wname.c_str());
    op.chain = unit_index;
  }
  for (size_t k = 0; k < cu.ops.size(); ++k)
    if (consumed[k]) cu.args.convs[k].m.flag_out = h->d_flags + (size_t)cu.ops[k] * h->tiles_cap;
  const ChainOps *ops = chain_ops(cu.GW);
  CK(h, ops->prepare(cu.MH, cu.NS, h->max_smem_optin));
  CK(h, ops->max_clusters(cu.MH, cu.NS, cu.smem, &cu.max_clusters));
  cu.max_clusters = std::min(cu.max_clusters, h->sm_count / 2);
  if (cu.max_clusters < 1) DAD_FAIL(h, DAD_ERR_CUDA, "conv_chain_kernel<GW=%d> cannot be resident with %d bytes of shared memory", cu.GW, cu.smem);
  if (cu.has_alt) {
    CK(h, ops->prepare(cu.MH, 1, h->max_smem_optin));
    CK(h, ops->max_clusters(cu.MH, 1, cu.alt_smem, &cu.alt_max_clusters));
    cu.alt_max_clusters = std::min(cu.alt_max_clusters, h->sm_count / 2);
    if (cu.alt_max_clusters < 1) cu.has_alt = false;
  }
  return DAD_OK;
}

// Partition the plan into launch units for the current fusion level.
int build_units(dad_handle *h) {
  h->chains.clear();
  h->units.clear();
  for (ConvOp &op : h->ops) op.chain = -1;
  ChainUnit cur;
  std::string cur_block;
  auto flush = [&]() {
    if (cur.ops.empty()) return;
    LaunchUnit lu;
    lu.chain = (int)h->chains.size();
    h->chains.push_back(cur);
    h->units.push_back(lu);
    cur = ChainUnit();
  };
  for (size_t i = 0; i < h->ops.size(); ++i) {
    const ConvOp &op = h->ops[i];
    ChainShape sh;
    if (h->fusion >= 1 && chain_shape(h, op, &sh)) {
      const std::string blk = block_stem(op);
      bool join = !cur.ops.empty() && h->fusion >= 2 && (int)cur.ops.size() < CH_MAX_CONVS && cur.L == op.g.L_out &&
                  cur.GW == sh.GW && cur.MH == sh.MH && cur.NS == sh.NS &&
                  h->ops[cur.ops[0]].g.Cout == op.g.Cout && (h->fusion >= 3 || blk == cur_block);
      if (join) {
        // the second source (a skip tensor) must come from an earlier launch
        for (int oi : cur.ops) if (h->ops[oi].out == op.in2) join = false;
      }
      if (join) {
        ChainUnit trial = cur;
        trial.ops.push_back((int)i);
        if (!chain_config(h, trial)) join = false; else cur = trial;
      }
      if (!join) {
        flush();
        cur.ops.push_back((int)i);
        cur.GW = sh.GW; cur.MH = sh.MH; cur.NS = sh.NS; cur.L = op.g.L_out; cur.S_t = sh.S_t;
        if (!chain_config(h, cur)) {        // does not fit at all: generic kernel
          cur = ChainUnit();
          LaunchUnit lu;
          lu.op = (int)i;
          h->units.push_back(lu);
          continue;
        }
      }
      cur_block = blk;
      continue;
    }
    flush();
    LaunchUnit lu;
    lu.op = (int)i;
    h->units.push_back(lu);
  }
  flush();
  for (size_t u = 0; u < h->units.size(); ++u)
    if (h->units[u].chain >= 0) {
      int rc = finish_chain(h, h->chains[h->units[u].chain], (int)u);
      if (rc) return rc;
    }
  return DAD_OK;
}

// Measured (tools/fusion_sweep.py, PointMaze): with 2 (= 128-wide items whenever a conv's 256-wide items fit in ONE round
// of the clusters) B = 1024 steps in 0.46 instead of 0.52 ms, B = 2048 / 4096 are unchanged (their convs have 1.7 / 3.5
// rounds of 256-wide items and keep them); with 4 the B = 2048 / 4096 chains lose 2-5 %.
constexpr int kChainAltHalfRoundsNarrow = 2, kChainAltHalfRoundsWide = 2;

int enqueue_chain(dad_handle *h, ChainUnit &cu, int B, cudaStream_t st, int flag_epoch = 1) {
  ChainParams &p = cu.args.p;
  p.B = B;
  p.n_mst = cdiv(B, cu.S_t);
  p.flag_epoch = flag_epoch;
  const int cout = h->ops[cu.ops[0]].g.Cout;
  // 256-wide items are the efficient shape (1,610 vs ~1,100 TFLOP/s of MMA issue) when there are rounds of them; when
  // ALL of a conv's 256-wide items fit in one round of the clusters, 128-wide items put twice the clusters to work (or
  // double the distance, in work-list rounds, between a conv and the conv that consumes it) and halve the MMA time on
  // every item's critical path
  const int items2 = cdiv(p.n_mst, 2) * (cout / (CH_BN * cu.NS));
  // kChainAltHalfRounds = how many HALF rounds of 256-wide items per conv still take 128-wide items (1 = the rule above)
  const int alt_r = cu.GW <= 32 ? tuning_env("DAD_CH_ALT_R_NARROW", kChainAltHalfRoundsNarrow) : tuning_env("DAD_CH_ALT_R_WIDE", kChainAltHalfRoundsWide);
  const bool alt = cu.has_alt && 2 * items2 <= cu.alt_max_clusters * alt_r;
  const int ns = alt ? 1 : cu.NS;
  p.n_tiles_n = cout / (CH_BN * ns);
  p.w_alt = alt ? 1 : 0;
  p.b_stage_bytes = ns * 8192;
  p.n_a_stages = alt ? cu.alt_n_a : cu.main_n_a;
  p.nb_stages = alt ? cu.alt_nb : cu.main_nb;
  const int smem = alt ? cu.alt_smem : cu.smem;
  const int entries = cdiv(p.n_mst, 2) * p.n_tiles_n * p.n_convs;
  // persistent, and never more clusters than can be co-resident: a cluster spins on tiles that other clusters of
  // this launch produce
  const int n_cl = std::min(entries, alt ? cu.alt_max_clusters : cu.max_clusters);
  cudaError_t e = chain_ops(cu.GW)->launch(cu.MH, ns, 2 * n_cl, smem, st, cu.args);
  if (e != cudaSuccess) DAD_FAIL(h, DAD_ERR_CUDA, "conv_chain launch failed (%s ...): %s", h->ops[cu.ops[0]].wname.c_str(), cudaGetErrorString(e));
  h->counting += 1;
  return DAD_OK;
}

template <int BN, int GW>
int launch_tc(dad_handle *h, const ConvOp &op, const ConvTcParams &p, int grid, cudaStream_t st) {
  auto kern = conv_tc_kernel<BN, GW>;
  const int num_kb = p.taps * (p.kch1 + p.kch2);
  const size_t smem = p.ws ? (size_t)TcCfg<BN>::smem_bytes_ws(op.Cout_pad, num_kb) : (size_t)TcCfg<BN>::smem_bytes(op.Cout_pad);
  launch_k(kern, dim3((unsigned)grid), dim3(TC_THREADS), smem, st, 1, op.tmA1, op.tmA2, op.tmW, p);
  return DAD_OK;
}

template <int BN, int GW>
cudaError_t set_tc_attr(int max_optin) {
  // the dynamic size depends on the layer's channel count (per-column epilogue parameters); allow the maximum
  // (the opt-in limit covers static + dynamic shared memory)
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, conv_tc_kernel<BN, GW>);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_tc_kernel<BN, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              max_optin - (int)fa.sharedSizeBytes);
}

template <int BN, int GW>
cudaError_t set_tc32_attr(int max_optin) {
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, conv_tc_kernel<BN, GW, true>);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_tc_kernel<BN, GW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              max_optin - (int)fa.sharedSizeBytes);
}

// Opt-in shared memory sizes are set once at create (never inside a stream capture).
int set_kernel_attrs(dad_handle *h) {
  CK(h, (set_tc_attr<64, 8>(h->max_smem_optin)));
  CK(h, (set_tc_attr<128, 16>(h->max_smem_optin)));
  CK(h, (set_tc_attr<128, 32>(h->max_smem_optin)));
  CK(h, (set_tc_attr<128, 64>(h->max_smem_optin)));
  CK(h, (set_tc_attr<128, 128>(h->max_smem_optin)));
  CK(h, (set_tc_attr<256, 256>(h->max_smem_optin)));
  CK(h, (set_tc_attr<16, 0>(h->max_smem_optin)));
  CK(h, (set_tc_attr<32, 0>(h->max_smem_optin)));
  CK(h, (set_tc_attr<64, 0>(h->max_smem_optin)));
  CK(h, (set_tc_attr<128, 0>(h->max_smem_optin)));
  CK(h, (set_tc32_attr<64, 8>(h->max_smem_optin)));
  CK(h, (set_tc32_attr<128, 16>(h->max_smem_optin)));
  CK(h, (set_tc32_attr<128, 32>(h->max_smem_optin)));
  CK(h, (set_tc32_attr<128, 64>(h->max_smem_optin)));
  CK(h, (set_tc32_attr<128, 128>(h->max_smem_optin)));
  CK(h, (set_tc32_attr<256, 256>(h->max_smem_optin)));
  CK(h, (set_tc32_attr<64, 0>(h->max_smem_optin)));
  CK(h, (set_tc32_attr<128, 0>(h->max_smem_optin)));
#define T3_ATTR(gw, mh, mode, ns) CK(h, (set_t3_attr<gw, mh, mode, ns>(h->max_smem_optin)));
  T3_FOR_EACH(T3_ATTR)
#undef T3_ATTR
#define STEP_ATTR(spt)                                                                                                        \
  CK(h, cudaFuncSetAttribute(step_project_fused_kernel<spt, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin)); \
  CK(h, cudaFuncSetAttribute(step_project_fused_kernel<spt, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  STEP_FOR_EACH_SPT(STEP_ATTR)
#undef STEP_ATTR
#define SMALL_ATTR(mt, nt) CK(h, cudaFuncSetAttribute(conv_small_kernel<mt, nt>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  SMALL_FOR_EACH(SMALL_ATTR)
#undef SMALL_ATTR
  CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->step_ctas_per_sm, step_pointwise_kernel<false>, 256, 0));
  if (h->step_ctas_per_sm < 1) h->step_ctas_per_sm = 1;
  CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->step_ctas_per_sm_lean, step_pointwise_kernel<true>, 256, 0));
  if (h->step_ctas_per_sm_lean < 1) h->step_ctas_per_sm_lean = 1;
  return DAD_OK;
}

// ---- small-batch latency path ---------------------------------------------------------------------
// Eligible: every bf16-mode layer of at most 32 rows per sample whose weight slab (16 output channels) and haloed
// activation tile fit shared memory, with GroupNorm groups of 16..128 channels (a cluster of <= 8 CTAs) or none.
void setup_small_op(dad_handle *h, ConvOp &op) {
  const ConvGeom &g = op.g;
  op.small = false;
  if (!h->bf16 || g.L_out > SM_MAX_L) return;
  const int gw = op.gname.empty() ? 0 : g.Cout / kGroups;
  // channels per CTA: 16, or 32 when a GroupNorm group would otherwise span more than 8 CTAs (the portable cluster limit)
  const int nch = (gw && gw / 16 > 8) ? 32 : 16;
  if (op.Cout_pad % nch || (gw && (gw % nch || gw / nch > 8))) return;
  int lo = 0, hi = 0;
  for (int t = 0; t < g.taps; ++t) { lo = std::min(lo, g.tap_off[t]); hi = std::max(hi, g.tap_off[t]); }
  ConvSmallParams &p = op.smp;
  p = ConvSmallParams{};
  p.halo = -lo;
  p.rows = p.halo + g.L_in + std::max(0, (g.L_out - 1) * g.in_stride + hi - (g.L_in - 1));
  op.sm_mt = g.L_out <= 16 ? 1 : 2;
  op.sm_nt = nch / 8;
  // weight ring: the whole K when it fits (every chunk then streams in ahead of the dependency wait), else 3 stages
  const int Cin = g.C1 + g.C2, K = g.taps * Cin;
  p.kc = K;
  p.n_stages = 1;
  if ((int)small_layout(Cin, p.rows, op.sm_mt, nch, p.kc, 1).total > h->max_smem_optin) {
    p.n_stages = 3;
    const int fixed = (int)small_layout(Cin, p.rows, op.sm_mt, nch, 0, 0).total;
    const int per_row = (h->max_smem_optin - fixed) / (p.n_stages * nch) - 16;      // bytes of K per weight row and stage
    p.kc = per_row / 2 / 128 * 128;             // whole k16 steps for each of the 8 warps
    if (p.kc < 256) return;
  }
  op.sm_smem = (int)small_layout(Cin, p.rows, op.sm_mt, nch, p.kc, p.n_stages).total;
  if (op.sm_smem > h->max_smem_optin) return;
  op.sm_cluster = gw ? gw / nch : 1;
  p.in1 = reinterpret_cast<const __nv_bfloat16 *>(act_ptr(h, op.in1));
  p.in2 = op.in2 >= 0 ? reinterpret_cast<const __nv_bfloat16 *>(act_ptr(h, op.in2)) : nullptr;
  p.w = op.w_b16;
  p.residual = op.tcp.residual;
  p.bias = op.bias;
  p.gamma = op.gamma;
  p.beta = op.beta;
  p.ttab = op.tcp.ttab;
  p.out = op.tcp.out;
  p.ls = h->d_ls;
  p.C1 = g.C1;
  p.C2 = g.C2;
  p.Cout = g.Cout;
  p.taps = g.taps;
  for (int t = 0; t < g.taps; ++t) p.tap_off[t] = g.tap_off[t];
  p.in_stride = g.in_stride;
  p.L_in = g.L_in;
  p.L_out = g.L_out;
  p.out_mul = g.out_mul;
  p.out_phase = g.out_phase;
  p.gw = gw;
  p.out_f32 = op.head ? 1 : 0;
  op.small = true;
}

int enqueue_small(dad_handle *h, const ConvOp &op, int B, cudaStream_t st) {
  const dim3 grid((unsigned)(op.Cout_pad / (8 * op.sm_nt)), (unsigned)B);
  cudaError_t e = cudaErrorInvalidValue;
#define SMALL_CASE(mt, nt) if (op.sm_mt == mt && op.sm_nt == nt) e = launch_k(conv_small_kernel<mt, nt>, grid, dim3(SM_THREADS), (size_t)op.sm_smem, st, op.sm_cluster, op.smp);
  SMALL_FOR_EACH(SMALL_CASE)
#undef SMALL_CASE
  if (e != cudaSuccess) DAD_FAIL(h, DAD_ERR_CUDA, "conv_small launch failed for %s: %s", op.wname.c_str(), cudaGetErrorString(e));
  h->counting += 1;
  return DAD_OK;
}

static bool step_fused_fits(const dad_handle *h) {
  return step_fused_smem(h->D) <= (size_t)h->max_smem_optin && step_fused_threads(h->D) <= STEP_FUSED_MAX_THREADS &&
         h->D % 4 == 0;
}

bool use_small(const dad_handle *h, const ConvOp &op, int B);

int enqueue_tc(dad_handle *h, const ConvOp &op, int B, cudaStream_t st) {
  if (use_small(h, op, B)) return enqueue_small(h, op, B, st);
  if (op.t3 && !h->rows_t) return enqueue_t3(h, op, B, st);      // conv_t3 assumes one timestep for the whole batch
  ConvTcParams p = op.tcp;
  p.B = B;
  p.n_tiles_m = cdiv((long long)B * op.g.L_out, TC_BM);
  const int tiles = p.n_tiles_m * p.n_tiles_n;
  int grid = std::min(tiles, h->sm_count);
  // Weight-stationary (conv_tc.cuh): when the whole K x 128 weight tile fits beside the activation ring and every CTA
  // sees several M tiles, keep it resident.  Needs a fixed N tile per CTA: a grid that is a multiple of n_tiles_n.
  {
    const int num_kb = p.taps * (p.kch1 + p.kch2);
    const int g_ws = std::min(tiles, h->sm_count / p.n_tiles_n * p.n_tiles_n);
    // 2 KB of static shared memory (GroupNorm scratch) also counts against the opt-in limit
    p.ws = (op.BN == 128 && tuning_env("DAD_TC_WS", 1) != 0 && g_ws > 0 && tiles >= 2 * g_ws &&
            TcCfg<128>::smem_bytes_ws(op.Cout_pad, num_kb) + 2048 <= h->max_smem_optin) ? 1 : 0;
    if (p.ws) grid = g_ws;
  }
  int rc = DAD_ERR_INVALID;
#define TC_CASE(bn, gw) if (op.BN == bn && op.GW == gw) rc = launch_tc<bn, gw>(h, op, p, grid, st);
  TC_CASE(64, 8) TC_CASE(128, 16) TC_CASE(128, 32) TC_CASE(128, 64) TC_CASE(128, 128) TC_CASE(256, 256)
  TC_CASE(16, 0) TC_CASE(32, 0) TC_CASE(64, 0) TC_CASE(128, 0)
#undef TC_CASE
  if (rc == DAD_ERR_INVALID) DAD_FAIL(h, DAD_ERR_INVALID, "internal: no tcgen05 instantiation for BN=%d GW=%d", op.BN, op.GW);
  h->counting += 1;
  return rc;
}

// ---- fp32 path ---------------------------------------------------------------------------------
// One conv (+ GroupNorm + Mish + time bias + residual) of an fp32 handle on tcgen05 kind::tf32.
int enqueue_tc32(dad_handle *h, const ConvOp &op, int B, cudaStream_t st) {
  ConvTcParams p = op.tcp;
  p.B = B;
  p.n_tiles_m = cdiv((long long)B * op.g.L_out, TC_BM);
  const int tiles = p.n_tiles_m * p.n_tiles_n;
  const int grid = std::min(tiles, h->sm_count);
  bool ok = false;
#define TC32_CASE(bn, gw)                                                                                                   \
  if (op.BN == bn && op.GW == gw) {                                                                                         \
    launch_k(conv_tc_kernel<bn, gw, true>, dim3((unsigned)grid), dim3(TC_THREADS), (size_t)TcCfg<bn>::smem_bytes(op.Cout_pad), \
             st, 1, op.tmA1, op.tmA2, op.tmW, p);                                                                           \
    ok = true;                                                                                                              \
  }
  TC32_CASE(64, 8) TC32_CASE(128, 16) TC32_CASE(128, 32) TC32_CASE(128, 64) TC32_CASE(128, 128) TC32_CASE(256, 256)
  TC32_CASE(64, 0) TC32_CASE(128, 0)
#undef TC32_CASE
  if (!ok) DAD_FAIL(h, DAD_ERR_INVALID, "internal: no tf32 instantiation for BN=%d GW=%d", op.BN, op.GW);
  h->counting += 1;
  CK(h, cudaGetLastError());
  return DAD_OK;
}

int enqueue_f32(dad_handle *h, const ConvOp &op, int B, cudaStream_t st) {
  if (h->f32_math == 1 && op.tc32 && !h->rows_t) return enqueue_tc32(h, op, B, st);
  ConvF32Params p{};
  p.in1 = reinterpret_cast<const float *>(act_ptr(h, op.in1));
  p.in2 = op.in2 >= 0 ? reinterpret_cast<const float *>(act_ptr(h, op.in2)) : nullptr;
  p.w = op.w_f32;
  p.bias = op.bias;
  const bool gn = !op.gname.empty();
  // with GroupNorm the residual is added after Mish (in the GN kernel), otherwise here
  p.residual = (!gn && op.res >= 0) ? reinterpret_cast<const float *>(act_ptr(h, op.res)) : nullptr;
  p.out = op.head ? h->d_eps : reinterpret_cast<float *>(act_ptr(h, op.out));
  p.g = op.g;
  p.B = B;
  const bool vec = (op.g.C1 % 4 == 0) && (op.g.C2 % 4 == 0) && (op.g.Cout % 4 == 0);
  dim3 grid(cdiv((long long)B * op.g.L_out, F32_BM), cdiv(op.g.Cout, F32_BN));
  // dad_set_fp32_math: the tensor-core variants take the layers whose channel counts they are written for
  const bool tf = h->f32_math != 0 && op.g.C1 % 16 == 0 && op.g.C2 % 16 == 0 && op.g.Cout % 8 == 0;
  if (tf && h->f32_math == 2) launch_k(conv_tf32_kernel<true>, grid, dim3(256), 0, st, 1, p);
  else if (tf) launch_k(conv_tf32_kernel<false>, grid, dim3(256), 0, st, 1, p);
  else if (vec) launch_k(conv_f32_kernel<EPI_BIAS, true>, grid, dim3(256), 0, st, 1, p);
  else launch_k(conv_f32_kernel<EPI_BIAS, false>, grid, dim3(256), 0, st, 1, p);
  h->counting += 1;
  if (gn) {
    GnF32Params q{};
    q.in = p.out;
    q.out = p.out;
    q.gamma = op.gamma;
    q.beta = op.beta;
    q.ttab = op.tblock >= 0 ? h->tblocks[op.tblock].tab : nullptr;
    q.residual = op.res >= 0 ? reinterpret_cast<const float *>(act_ptr(h, op.res)) : nullptr;
    q.ls = h->d_ls;
    q.L = op.g.L_out;
    q.C = op.g.Cout;
    // one warp per (sample, group) with the group in registers when it fits, else one block per (sample, group)
    const int gw = q.C / kGroups, n4 = q.L * gw / 4, pairs = B * kGroups;
    if (gw % 4 == 0 && n4 <= 32 * 4) launch_k(gn_mish_f32_warp_kernel<4>, dim3(cdiv(pairs, 8)), dim3(256), 0, st, 1, q, pairs);
    else if (gw % 4 == 0 && n4 <= 32 * 8) launch_k(gn_mish_f32_warp_kernel<8>, dim3(cdiv(pairs, 8)), dim3(256), 0, st, 1, q, pairs);
    else if (gw % 4 == 0 && n4 <= 32 * 16) launch_k(gn_mish_f32_warp_kernel<16>, dim3(cdiv(pairs, 8)), dim3(256), 0, st, 1, q, pairs);
    else launch_k(gn_mish_f32_kernel, dim3(B, kGroups), dim3(128), 0, st, 1, q);
    h->counting += 1;
  }
  return DAD_OK;
}

// One U-Net forward of the staged trajectories -> d_eps.  (TemporalUnet.forward, temporal_unet.py:199-241)
// In the captured step, kernels that do not depend on each other are put on a forked branch of the graph: a
// block's 1x1 residual conv beside its first k5 conv (both read the block input), and the two phases of a
// ConvTranspose.  Every kernel is persistent over all SMs, so the branch kernel's CTAs move in exactly while the
// main kernel's tail drains its SMs.
bool forks_with_next(const dad_handle *h, size_t i) {
  if (!h->bf16 || i + 1 >= h->ops.size()) return false;
  const ConvOp &a = h->ops[i], &b = h->ops[i + 1];
  if (b.wname.find("residual_conv") != std::string::npos && b.in1 == a.in1 && b.in2 == a.in2) return true;
  if (a.transposed && b.transposed && a.out == b.out && a.g.out_phase != b.g.out_phase) return true;
  return false;
}

// Whether a chain runs as one conv_chain launch for this batch: not when its convs take the latency kernels
// (small batches) or carry per-row timesteps (stand-alone forward; the generic kernel indexes the tables per row).
bool use_small(const dad_handle *h, const ConvOp &op, int B) {
  // latency kernels: every sample's CTAs fetch the layer's weights again, so past a re-read budget (256 MB per
  // layer and launch: never reached by PointMaze below 24 samples, 6 samples of HalfCheetah's largest layer) the throughput kernels win
  return op.small && B <= h->small_max_b && !h->rows_t &&
         (size_t)B * op.Cout_pad * op.g.taps * op.Cin_store * 2 <= ((size_t)256 << 20);
}

bool use_chain(const dad_handle *h, const ChainUnit &cu, int B) {
  if (h->rows_t) return false;
  for (int oi : cu.ops)
    if (use_small(h, h->ops[oi], B)) return false;
  return true;
}

int enqueue_unet(dad_handle *h, int B, cudaStream_t st, bool advance = false, cudaStream_t side = nullptr) {
  const dad_config &c = h->cfg;
  const size_t rows = (size_t)B * c.horizon;
  const int n_flags = h->d_flags ? (int)h->ops.size() * h->tiles_cap : 0;
  if (h->bf16) {
    const size_t n = rows * (h->Cpad_in / 8);
    launch_k(stage_x_kernel, dim3(cdiv(n, 256)), dim3(256), 0, st, 1, h->d_ls, (float *)nullptr,
             reinterpret_cast<__nv_bfloat16 *>(act_ptr(h, 0)), rows, c.transition_dim, h->Cpad_in, advance ? 1 : 0,
             h->d_flags, n_flags);
  } else {
    const size_t n = rows * c.transition_dim;
    launch_k(stage_x_kernel, dim3(cdiv(n, 256)), dim3(256), 0, st, 1, h->d_ls, reinterpret_cast<float *>(act_ptr(h, 0)),
             (__nv_bfloat16 *)nullptr, rows, c.transition_dim, 0, advance ? 1 : 0, (unsigned *)nullptr, 0);
  }
  h->counting += 1;
  // ops in execution order; a chain unit contributes ONE launch when it runs fused
  std::vector<int> seq;
  for (const LaunchUnit &lu : h->units) {
    if (lu.chain >= 0 && h->bf16 && use_chain(h, h->chains[lu.chain], B)) {
      seq.push_back(-1 - lu.chain);
    } else if (lu.chain >= 0) {
      for (int oi : h->chains[lu.chain].ops) seq.push_back(oi);
    } else {
      seq.push_back(lu.op);
    }
  }
  for (size_t k = 0; k < seq.size(); ++k) {
    if (seq[k] < 0) {
      int rc = enqueue_chain(h, h->chains[-1 - seq[k]], B, st);
      if (rc) return rc;
      continue;
    }
    const size_t i = (size_t)seq[k];
    const ConvOp &op = h->ops[i];
    if (side && k + 1 < seq.size() && seq[k + 1] == (int)i + 1 && forks_with_next(h, i)) {
      CK(h, cudaEventRecord(h->ev_fork, st));
      CK(h, cudaStreamWaitEvent(side, h->ev_fork, 0));
      int rc = enqueue_tc(h, op, B, st);
      if (!rc) rc = enqueue_tc(h, h->ops[i + 1], B, side);
      if (rc) return rc;
      CK(h, cudaEventRecord(h->ev_join, side));
      CK(h, cudaStreamWaitEvent(st, h->ev_join, 0));
      ++k;
      continue;
    }
    int rc = h->bf16 ? enqueue_tc(h, op, B, st) : enqueue_f32(h, op, B, st);
    if (rc) return rc;
  }
  CK(h, cudaGetLastError());
  return DAD_OK;
}

// The fused remainder of the step (K7 [+ K8 for large D]).
int enqueue_step(dad_handle *h, const float *model_out, int B, bool project, bool advance, cudaStream_t st) {
  const dad_config &c = h->cfg;
  StepParams p{};
  p.ls = h->d_ls;
  p.model_out = model_out;
  p.xtmp = h->d_xtmp;
  p.sqrt_recip = h->d_sched[0];
  p.sqrt_recipm1 = h->d_sched[1];
  p.coef1 = h->d_sched[2];
  p.coef2 = h->d_sched[3];
  p.logvar = h->d_sched[4];
  p.alpha_tab = h->d_alpha;
  p.Nt = h->d_Nt;
  p.q = h->d_q;
  p.cond_vals = h->d_cond;
  p.B = B;
  p.D = h->D;
  p.T = c.transition_dim;
  p.predict_epsilon = c.predict_epsilon;
  p.clip_denoised = c.clip_denoised;
  const size_t total4 = (size_t)B * h->D / 4;
  if (total4 >= (size_t)1 << 31) DAD_FAIL(h, DAD_ERR_INVALID, "step kernel: B*H*T/4 must be below 2^31");
  if (!project) {
    p.to_tmp = 0;
    const bool lean = h->step_lean && c.predict_epsilon && c.clip_denoised;
    const int grid = (int)std::min<size_t>(cdiv(total4, 256), (size_t)h->sm_count * (lean ? h->step_ctas_per_sm_lean : h->step_ctas_per_sm));
    if (lean) launch_k(step_pointwise_kernel<true>, dim3(grid), dim3(256), 0, st, 1, p);
    else launch_k(step_pointwise_kernel<false>, dim3(grid), dim3(256), 0, st, 1, p);
    h->counting += 1;
  } else if (h->proj_tc && (!step_fused_fits(h) || B >= h->proj_tc_min_batch || h->force_proj_tc)) {
    // large D (the projector does not fit shared memory): pointwise part -> x' (fp32) + its bf16 (hi | lo | hi) split; then one tcgen05 GEMM against (N_hi | N_hi | N_lo)
    // whose epilogue blends, inpaints and writes x (K8)
    p.to_tmp = 1;
    p.split = h->d_split;
    p.Kp = h->projKp;
    const int grid = (int)std::min<size_t>(cdiv(total4, 256), (size_t)h->sm_count * h->step_ctas_per_sm);
    launch_k(step_pointwise_kernel<false>, dim3(grid), dim3(256), 0, st, 1, p);
    ConvTcParams t{};
    t.bias = h->d_qpad;
    t.ls = h->d_ls;
    t.B = B;
    t.L_out = 1;
    t.out_mul = 1;
    t.out_phase = 0;
    t.Cout = h->D;
    t.n_tiles_m = cdiv(B, TC_BM);
    t.n_tiles_n = h->projNp / 128;
    t.kch1 = 3 * h->projKp / TC_BK;
    t.kch2 = 0;
    t.taps = 1;
    t.out_f32 = 1;
    t.proj_x = h->d_xtmp;
    t.alpha_tab = h->d_alpha;
    t.cond_vals = h->d_cond;
    t.T = c.transition_dim;
    const int tiles = t.n_tiles_m * t.n_tiles_n;
    launch_k(conv_tc_kernel<128, 0>, dim3((unsigned)std::min(tiles, h->sm_count)), dim3(TC_THREADS),
             (size_t)TcCfg<128>::smem_bytes(h->projNp), st, 1, h->tmProjA, h->tmProjA, h->tmProjW, t);
    h->counting += 2;
  } else {
    if (step_fused_fits(h)) {
      // samples per thread: whole rounds of 4*SPT-sample groups over the SMs, least (rounds x SPT); ties -> larger SPT
      int spt = 8, best = INT_MAX;
      for (int c = 8; c >= 1; --c) {
        const int cost = cdiv(cdiv(B, 4 * c), h->sm_count) * c;
        if (cost < best) { best = cost; spt = c; }
      }
      const int grid = std::min(cdiv(B, 4 * spt), h->sm_count);
      const dim3 blk(step_fused_threads(h->D));
      const size_t smem = step_fused_smem(h->D);
      const bool lean = h->step_lean && h->cfg.predict_epsilon && h->cfg.clip_denoised;      // see step_pointwise_kernel<true>
      switch (spt) {
#define STEP_CASE(n)                                                                                \
  case n:                                                                                           \
    if (lean) launch_k(step_project_fused_kernel<n, true>, dim3(grid), blk, smem, st, 1, p);        \
    else launch_k(step_project_fused_kernel<n, false>, dim3(grid), blk, smem, st, 1, p);            \
    break;
        STEP_FOR_EACH_SPT(STEP_CASE)
#undef STEP_CASE
      }
      h->counting += 1;
    } else {
      // large D: pointwise part to scratch, then the projector as a tiled GEMM whose epilogue
      // blends, inpaints and writes x (K8).
      p.to_tmp = 1;
      const int grid = (int)std::min<size_t>(cdiv(total4, 256), (size_t)h->sm_count * h->step_ctas_per_sm);
      launch_k(step_pointwise_kernel<false>, dim3(grid), dim3(256), 0, st, 1, p);
      ConvF32Params g{};
      g.in1 = h->d_xtmp;
      g.w = h->d_Nt;            // Nt[k][d] is exactly the [c][n] weight layout
      g.bias = h->d_q;
      g.g.C1 = h->D; g.g.C2 = 0; g.g.Cout = h->D; g.g.taps = 1; g.g.tap_off[0] = 0;
      g.g.in_stride = 1; g.g.L_in = 1; g.g.L_out = 1; g.g.out_mul = 1; g.g.out_phase = 0;
      g.B = B;
      g.ls = h->d_ls;
      g.alpha_tab = h->d_alpha;
      g.cond_vals = h->d_cond;
      g.T = c.transition_dim;
      dim3 grid2(cdiv(B, F32_BM), cdiv(h->D, F32_BN));
      launch_k(conv_f32_kernel<EPI_PROJECT, true>, grid2, dim3(256), 0, st, 1, g);
      h->counting += 2;
    }
  }
  (void)advance;    // the step index is advanced by the first kernel of the captured step (stage_x_kernel)
  CK(h, cudaGetLastError());
  return DAD_OK;
}

int set_loop_state(dad_handle *h, const LoopState &v, cudaStream_t st) {
  set_loop_state_kernel<<<1, 1, 0, st>>>(h->d_ls, v);
  h->launches += 1;
  CK(h, cudaGetLastError());
  return DAD_OK;
}

void fill_cond(const dad_handle *h, LoopState &ls, int row0) {
  ls.n_cond = h->n_cond;
  ls.cond_per_batch = h->cond_per_batch;
  ls.cond_B = h->cond_B;
  ls.cond_row0 = row0;
  for (int i = 0; i < kMaxCond; ++i) ls.cond_h[i] = h->cond_h[i];
}

void drop_graphs(dad_handle *h) {
  h->epoch += 1;            // caller-captured steps (dad_loop_*) point at the same buffers
  for (auto &kv : h->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->graphs.clear();
}

// One captured graph = `reps` consecutive diffusion steps (each: [stage x, step index -= 1] -> U-Net -> fused step
// kernel).  Several steps per graph keep the programmatic (PDL) edges ACROSS step boundaries and cut the number of
// graph launches per plan (500 -> 25 for the B = 1 plan of get_action).
constexpr int kStepsPerGraph = 20;

int get_graph(dad_handle *h, int B, bool project, int reps, GraphEntry **out) {
  const long long key = (((long long)B * 2 + (project ? 1 : 0)) * 64 + reps) * 2 + (h->step_lean ? 1 : 0);
  auto it = h->graphs.find(key);
  if (it != h->graphs.end()) { *out = &it->second; return DAD_OK; }
  cudaGraph_t graph = nullptr;
  h->counting = 0;
  CK(h, cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
  // the loop starts one index high: every captured step first decrements it
  const bool fork = tuning_env("DAD_FORK", 1) != 0;
  int rc = DAD_OK;
  for (int r = 0; r < reps && !rc; ++r) {
    rc = enqueue_unet(h, B, h->cap_stream, true, fork ? h->side_stream : nullptr);
    if (!rc) rc = enqueue_step(h, h->d_eps, B, project, false, h->cap_stream);
  }
  cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (e != cudaSuccess) DAD_FAIL(h, DAD_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  GraphEntry ge;
  ge.kernels = h->counting;
  e = cudaGraphInstantiate(&ge.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) DAD_FAIL(h, DAD_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  h->graphs[key] = ge;
  *out = &h->graphs[key];
  return DAD_OK;
}

struct NamedTensors {
  std::unordered_map<std::string, const dad_tensor *> map;
  const dad_tensor *get(const std::string &n) const {
    auto it = map.find(n);
    return it == map.end() ? nullptr : it->second;
  }
};

// Returns a DEVICE pointer to the tensor's data: the caller's pointer if it already is device
// memory, otherwise a staged copy in `stage`.
int device_view(dad_handle *h, const dad_tensor *t, float *stage, const float **out) {
  cudaPointerAttributes at{};
  cudaError_t e = cudaPointerGetAttributes(&at, t->data);
  if (e == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
    *out = t->data;
    return DAD_OK;
  }
  cudaGetLastError();
  CK(h, copy_now(stage, t->data, t->numel * sizeof(float), h->own_stream));
  *out = stage;
  return DAD_OK;
}

int check_dims(dad_handle *h, const dad_tensor *t, const std::string &name, long long want) {
  if (!t) DAD_FAIL(h, DAD_ERR_INVALID, "missing tensor '%s' in state_dict", name.c_str());
  if (t->numel != want)
    DAD_FAIL(h, DAD_ERR_INVALID, "tensor '%s' has %lld elements, expected %lld", name.c_str(), (long long)t->numel, want);
  return DAD_OK;
}

}  // namespace

namespace {
template <typename F>
int time_launches(dad_handle *h, cudaStream_t st, int iters, float *ms, F &&launch) {
  cudaEvent_t a, b;
  CK(h, cudaEventCreate(&a));
  CK(h, cudaEventCreate(&b));
  int rc = launch();
  if (rc) return rc;
  CK(h, cudaEventRecord(a, st));
  for (int i = 0; i < iters; ++i)
    if ((rc = launch())) return rc;
  CK(h, cudaEventRecord(b, st));
  CK(h, cudaStreamSynchronize(st));
  CK(h, cudaGetLastError());
  float t = 0.f;
  CK(h, cudaEventElapsedTime(&t, a, b));
  *ms = t / (float)iters;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return DAD_OK;
}
}  // namespace

// =================================================================================================
extern "C" {

int dad_abi_version(void) { return DAD_ABI_VERSION; }

const char *dad_last_error(const dad_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int dad_create(const dad_config *cfg, dad_handle **out) {
  if (!cfg || !out) { g_create_error = "null argument"; return DAD_ERR_INVALID; }
  *out = nullptr;
  dad_handle *h = new dad_handle();
  h->cfg = *cfg;
  auto fail = [&](int code) { g_create_error = h->err; dad_destroy(h); return code; };
  const dad_config &c = h->cfg;
  if (c.abi_version != DAD_ABI_VERSION) { h->err = "ABI version mismatch"; return fail(DAD_ERR_INVALID); }
  if (c.n_levels < 1 || c.n_levels > DAD_MAX_LEVELS || c.transition_dim < 1 || c.dim < 8 || c.dim % 8 ||
      c.horizon < 1 || c.n_timesteps < 1 || c.max_batch < 1 || c.kernel_size < 1 || c.kernel_size > 7 ||
      c.kernel_size % 2 == 0) {
    h->err = "invalid configuration (levels/dims/horizon/kernel_size)";
    return fail(DAD_ERR_INVALID);
  }
  if (c.horizon % (1 << (c.n_levels - 1))) {
    h->err = "horizon must be divisible by 2^(n_levels-1) (the reference U-Net cannot concatenate its skips otherwise)";
    return fail(DAD_ERR_INVALID);
  }
  if ((c.horizon * c.transition_dim) % 4) { h->err = "horizon * transition_dim must be a multiple of 4"; return fail(DAD_ERR_INVALID); }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    h->err = "no CUDA device: this library has no CPU fallback";
    return fail(DAD_ERR_DEVICE);
  }
  if (c.device < 0 || c.device >= ndev) { h->err = "device ordinal out of range"; return fail(DAD_ERR_DEVICE); }
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, c.device) != cudaSuccess) { h->err = "cudaGetDeviceProperties failed"; return fail(DAD_ERR_CUDA); }
  if (prop.major != 10) {
    char b[160];
    snprintf(b, sizeof(b), "device %d is sm_%d%d; this library is built for sm_100a (B200) only", c.device, prop.major, prop.minor);
    h->err = b;
    return fail(DAD_ERR_DEVICE);
  }
  if (cudaSetDevice(c.device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return fail(DAD_ERR_CUDA); }
  h->sm_count = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  h->bf16 = c.precision == DAD_PRECISION_BF16;
  h->elt = h->bf16 ? 2 : 4;
  h->time_dim = c.time_dim > 0 ? c.time_dim : c.dim;
  h->D = c.horizon * c.transition_dim;
  {   // both precisions build tensor maps (fp32 handles for their TF32 tensor-core path)
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
      h->err = "cuTensorMapEncodeTiled not available from the driver";
      return fail(DAD_ERR_CUDA);
    }
    h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  int rc = build_plan(h);
  if (rc) return fail(rc);
  if ((rc = set_kernel_attrs(h))) return fail(rc);
  if (cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    h->err = "cudaStreamCreate failed";
    return fail(DAD_ERR_CUDA);
  }
  // workspaces
  const size_t arena_bytes = h->act_bytes_per_sample * (size_t)c.max_batch;
  if ((rc = dev_alloc(h, &h->arena, arena_bytes))) return fail(rc);
  fill_now(h->arena, 0, arena_bytes, h->own_stream);
  if ((rc = dev_alloc(h, &h->d_eps, (size_t)c.max_batch * h->D))) return fail(rc);
  if ((rc = dev_alloc(h, &h->d_xtmp, (size_t)c.max_batch * h->D))) return fail(rc);
  if ((rc = dev_alloc(h, &h->d_ls, 1))) return fail(rc);
  fill_now(h->d_ls, 0, sizeof(LoopState), h->own_stream);
  for (int i = 0; i < 5; ++i)
    if ((rc = dev_alloc(h, &h->d_sched[i], (size_t)c.n_timesteps))) return fail(rc);
  if ((rc = dev_alloc(h, &h->d_alpha, (size_t)c.n_timesteps))) return fail(rc);
  fill_now(h->d_alpha, 0, sizeof(float) * c.n_timesteps, h->own_stream);
  h->cond_cap = (size_t)kMaxCond * c.transition_dim;
  if ((rc = dev_alloc(h, &h->d_cond, h->cond_cap))) return fail(rc);
  // per-op parameter storage
  for (ConvOp &op : h->ops) {
    if (h->bf16) {
      if ((rc = setup_tc_op(h, op))) return fail(rc);
      if ((rc = dev_alloc(h, &op.w_b16, (size_t)op.Cout_pad * op.g.taps * op.Cin_store))) return fail(rc);
    } else {
      op.Cout_pad = op.g.Cout;
      if ((rc = dev_alloc(h, &op.w_f32, (size_t)op.g.taps * op.Cin_store * op.g.Cout))) return fail(rc);
      op.tc32 = setup_tc32_op(h, op);
      if (op.tc32 && (rc = dev_alloc(h, &op.w_k32, (size_t)op.g.taps * op.Cin_store * op.g.Cout))) return fail(rc);
    }
    if ((rc = dev_alloc(h, &op.bias, (size_t)op.Cout_pad))) return fail(rc);
    fill_now(op.bias, 0, sizeof(float) * op.Cout_pad, h->own_stream);
    if (!op.gname.empty()) {
      if ((rc = dev_alloc(h, &op.gamma, (size_t)op.g.Cout))) return fail(rc);
      if ((rc = dev_alloc(h, &op.beta, (size_t)op.g.Cout))) return fail(rc);
    }
  }
  for (TimeBlock &tb : h->tblocks)
    if ((rc = dev_alloc(h, &tb.tab, (size_t)c.n_timesteps * tb.C))) return fail(rc);
  if (h->bf16)
    for (ConvOp &op : h->ops) {
      if ((rc = finish_tc_op(h, op))) return fail(rc);
      if (t3_eligible(h, op) && setup_t3_op(h, op) == DAD_OK) {
        if ((rc = finish_t3_op(h, op))) return fail(rc);
        op.t3 = true;
      }
      setup_small_op(h, op);
    }
  if (!h->bf16)
    for (ConvOp &op : h->ops)
      if (op.tc32 && (rc = finish_tc32_op(h, op))) return fail(rc);
  h->small_max_b = tuning_env("DAD_SMALL_MAX_B", h->small_max_b);
  if (h->bf16) {
    // tile-completion counters of the conv chains: one row per conv, one word per sample tile (>= 8 samples each)
    h->tiles_cap = cdiv(c.max_batch, 8);
    const size_t words = h->ops.size() * (size_t)h->tiles_cap;
    if ((rc = dev_alloc(h, &h->d_flags, words))) return fail(rc);
    fill_now(h->d_flags, 0, words * sizeof(unsigned), h->own_stream);
    if ((rc = dev_alloc(h, &h->d_err, 4))) return fail(rc);      // [0] error code; [1..3] stall counters of -DDAD_TUNING builds
    fill_now(h->d_err, 0, 4 * sizeof(unsigned), h->own_stream);
    h->fusion = tuning_env("DAD_FUSION", h->fusion);
  } else {
    h->fusion = 0;
  }
  if ((rc = build_units(h))) return fail(rc);
  if (cudaDeviceSynchronize() != cudaSuccess) { h->err = "device error during create"; return fail(DAD_ERR_CUDA); }
  {
    std::lock_guard<std::mutex> lock(g_handles_mutex);
    g_handles.push_back(h);
  }
  *out = h;
  return DAD_OK;
}

int dad_destroy(dad_handle *h) {
  if (!h) return DAD_OK;
  {
    std::lock_guard<std::mutex> lock(g_handles_mutex);
    g_handles.erase(std::remove(g_handles.begin(), g_handles.end(), h), g_handles.end());
    for (dad_handle *o : g_handles)
      if (o->companion == h) { o->companion = nullptr; o->fp32_min_step = INT_MAX; }
  }
  cudaDeviceSynchronize();
  drop_graphs(h);
  for (void *p : h->allocs) cudaFree(p);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return DAD_OK;
}

int dad_load_weights(dad_handle *h, const dad_tensor *tensors, int32_t n) {
  if (!h || !tensors) return DAD_ERR_INVALID;
  NvtxRange nvtx("dad_load_weights");
  const dad_config &c = h->cfg;
  CK(h, cudaSetDevice(c.device));
  DAD_QUIESCE(h);
  NamedTensors nt;
  size_t max_numel = 0;
  for (int i = 0; i < n; ++i) {
    if (!tensors[i].name || !tensors[i].data) DAD_FAIL(h, DAD_ERR_INVALID, "tensor %d has a null name or pointer", i);
    nt.map[tensors[i].name] = &tensors[i];
    max_numel = std::max<size_t>(max_numel, (size_t)tensors[i].numel);
  }
  DevTemps tmp;                 // freed on every return path
  float *stage = nullptr;
  CK(h, tmp.alloc(&stage, max_numel));
  int rc = DAD_OK;
  cudaStream_t st = h->own_stream;
  // ---- convolutions
  for (ConvOp &op : h->ops) {
    const dad_tensor *w = nt.get(op.wname), *b = nt.get(op.bname);
    if ((rc = check_dims(h, w, op.wname, (long long)op.g.Cout * op.Cin_real * op.ksize))) return rc;
    if ((rc = check_dims(h, b, op.bname, op.g.Cout))) return rc;
    const float *wd = nullptr;
    if ((rc = device_view(h, w, stage, &wd))) return rc;
    if (h->bf16) {
      const size_t total = (size_t)op.Cout_pad * op.g.taps * op.Cin_store;
      pack_w_bf16_kernel<<<cdiv(total, 256), 256, 0, st>>>(wd, op.w_b16, op.g.Cout, op.Cin_real, op.Cout_pad,
                                                           op.Cin_store, op.ksize, op.g.taps, op.sel, op.transposed);
    } else {
      const size_t total = (size_t)op.g.taps * op.Cin_real * op.g.Cout;
      pack_w_f32_kernel<<<cdiv(total, 256), 256, 0, st>>>(wd, op.w_f32, op.g.Cout, op.Cin_real, op.ksize, op.g.taps,
                                                          op.sel, op.transposed);
      if (op.tc32)
        pack_w_tf32_kmajor_kernel<<<cdiv(total, 256), 256, 0, st>>>(wd, op.w_k32, op.g.Cout, op.Cin_real, op.ksize, op.g.taps,
                                                                    op.sel, op.transposed);
    }
    CK(h, cudaStreamSynchronize(st));   // `stage` is reused by the next tensor
    CK(h, copy_now(op.bias, b->data, sizeof(float) * op.g.Cout, h->own_stream));
    if (!op.gname.empty()) {
      const dad_tensor *gw = nt.get(op.gname + ".weight"), *gb = nt.get(op.gname + ".bias");
      if ((rc = check_dims(h, gw, op.gname + ".weight", op.g.Cout))) return rc;
      if ((rc = check_dims(h, gb, op.gname + ".bias", op.g.Cout))) return rc;
      CK(h, copy_now(op.gamma, gw->data, sizeof(float) * op.g.Cout, h->own_stream));
      CK(h, copy_now(op.beta, gb->data, sizeof(float) * op.g.Cout, h->own_stream));
    }
  }
  // ---- time tables: every step index, once (K6)
  const int S = c.n_timesteps, dim = c.dim, td = h->time_dim;
  if (dim / 2 < 2) DAD_FAIL(h, DAD_ERR_INVALID, "dim too small for the sinusoidal embedding");
  float *emb = nullptr, *h1 = nullptr, *temb = nullptr, *bd = nullptr;
  int max_c = td * 4;
  for (const TimeBlock &tb : h->tblocks) max_c = std::max(max_c, tb.C);
  CK(h, tmp.alloc(&emb, (size_t)S * dim));
  CK(h, tmp.alloc(&h1, (size_t)S * td * 4));
  CK(h, tmp.alloc(&temb, (size_t)S * td));
  CK(h, tmp.alloc(&bd, (size_t)max_c));
  {
    const dad_tensor *w1 = nt.get("time_mlp.1.weight"), *b1 = nt.get("time_mlp.1.bias");
    const dad_tensor *w3 = nt.get("time_mlp.3.weight"), *b3 = nt.get("time_mlp.3.bias");
    if ((rc = check_dims(h, w1, "time_mlp.1.weight", (long long)td * 4 * dim))) return rc;
    if ((rc = check_dims(h, b1, "time_mlp.1.bias", td * 4))) return rc;
    if ((rc = check_dims(h, w3, "time_mlp.3.weight", (long long)td * td * 4))) return rc;
    if ((rc = check_dims(h, b3, "time_mlp.3.bias", td))) return rc;
    sinusoid_table_kernel<<<cdiv((long long)S * (dim / 2), 256), 256, 0, st>>>(emb, S, dim);
    const float *wd = nullptr;
    if ((rc = device_view(h, w1, stage, &wd))) return rc;
    CK(h, copy_now(bd, b1->data, sizeof(float) * td * 4, h->own_stream));
    linear_rows_kernel<<<cdiv((long long)S * td * 4, 256), 256, 0, st>>>(emb, wd, bd, h1, S, dim, td * 4, 0, 1);
    CK(h, cudaStreamSynchronize(st));
    if ((rc = device_view(h, w3, stage, &wd))) return rc;
    CK(h, copy_now(bd, b3->data, sizeof(float) * td, h->own_stream));
    linear_rows_kernel<<<cdiv((long long)S * td, 256), 256, 0, st>>>(h1, wd, bd, temb, S, td * 4, td, 0, 0);
    CK(h, cudaStreamSynchronize(st));
  }
  for (TimeBlock &tb : h->tblocks) {
    const dad_tensor *w = nt.get(tb.stem + ".weight"), *b = nt.get(tb.stem + ".bias");
    if ((rc = check_dims(h, w, tb.stem + ".weight", (long long)tb.C * td))) return rc;
    if ((rc = check_dims(h, b, tb.stem + ".bias", tb.C))) return rc;
    const float *wd = nullptr;
    if ((rc = device_view(h, w, stage, &wd))) return rc;
    CK(h, copy_now(bd, b->data, sizeof(float) * tb.C, h->own_stream));
    linear_rows_kernel<<<cdiv((long long)S * tb.C, 256), 256, 0, st>>>(temb, wd, bd, tb.tab, S, td, tb.C, 1, 0);
    CK(h, cudaStreamSynchronize(st));
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("weight packing failed: ") + cudaGetErrorString(e); return DAD_ERR_CUDA; }
  h->have_weights = true;
  return DAD_OK;
}

int dad_set_schedule(dad_handle *h, const float *sr, const float *srm1, const float *c1, const float *c2,
                     const float *lv, int32_t n) {
  if (!h || !sr || !srm1 || !c1 || !c2 || !lv) return DAD_ERR_INVALID;
  if (n != h->cfg.n_timesteps) DAD_FAIL(h, DAD_ERR_INVALID, "schedule length %d != n_timesteps %d", n, h->cfg.n_timesteps);
  CK(h, cudaSetDevice(h->cfg.device));
  DAD_QUIESCE(h);
  const float *src[5] = {sr, srm1, c1, c2, lv};
  for (int i = 0; i < 5; ++i) CK(h, copy_now(h->d_sched[i], src[i], sizeof(float) * n, h->own_stream));
  h->have_sched = true;
  return DAD_OK;
}

int dad_set_projector(dad_handle *h, const float *Nmat, const float *q, const float *alpha, int32_t D, int32_t n) {
  if (!h) return DAD_ERR_INVALID;
  CK(h, cudaSetDevice(h->cfg.device));
  DAD_QUIESCE(h);
  if (!Nmat) { h->projD = 0; return DAD_OK; }
  if (!q || !alpha) return DAD_ERR_INVALID;
  if (D != h->D) DAD_FAIL(h, DAD_ERR_INVALID, "projector dimension %d != horizon*transition_dim %d", D, h->D);
  if (n != h->cfg.n_timesteps) DAD_FAIL(h, DAD_ERR_INVALID, "alpha length %d != n_timesteps %d", n, h->cfg.n_timesteps);
  if (!h->d_Nt) {
    int rc;
    if ((rc = dev_alloc(h, &h->d_Nt, (size_t)D * D))) return rc;
    if ((rc = dev_alloc(h, &h->d_Nrow, (size_t)D * D))) return rc;
    if ((rc = dev_alloc(h, &h->d_q, (size_t)D))) return rc;
    drop_graphs(h);   // new buffers -> captured pointers are stale
  }
  CK(h, copy_now(h->d_Nrow, Nmat, sizeof(float) * D * D, h->own_stream));
  CK(h, copy_now(h->d_q, q, sizeof(float) * D, h->own_stream));
  CK(h, copy_now(h->d_alpha, alpha, sizeof(float) * n, h->own_stream));
  // transpose on device via the fp32 packer: treat Nmat as a (Cout=D, Cin=D, k=1) conv weight -> [c][n]
  TapSel sel{};
  pack_w_f32_kernel<<<cdiv((long long)D * D, 256), 256, 0, h->own_stream>>>(h->d_Nrow, h->d_Nt, D, D, 1, 1, sel, 0);
  CK(h, cudaStreamSynchronize(h->own_stream));
  h->projD = D;
  h->proj_tc = false;
  if (h->bf16 && tuning_env("DAD_PROJ_TC", 1) != 0) {
    const int Kp = (D + 63) / 64 * 64, Np = (D + 127) / 128 * 128;
    if (!h->d_projW) {
      int rc;
      if ((rc = dev_alloc(h, &h->d_projW, (size_t)Np * 3 * Kp))) return rc;
      if ((rc = dev_alloc(h, &h->d_split, (size_t)h->cfg.max_batch * 3 * Kp))) return rc;
      if ((rc = dev_alloc(h, &h->d_qpad, (size_t)Np))) return rc;
      CK(h, fill_now(h->d_split, 0, sizeof(__nv_bfloat16) * (size_t)h->cfg.max_batch * 3 * Kp, h->own_stream));
      h->projKp = Kp;
      h->projNp = Np;
      // A: (3Kp, 1, 1, max_batch) boxes of 64 x 128 rows; W: (3Kp, Np) boxes of 64 x 128
      cuuint64_t ad[4] = {(cuuint64_t)3 * Kp, 1, 1, (cuuint64_t)h->cfg.max_batch};
      cuuint64_t as[3] = {(cuuint64_t)3 * Kp * 2, (cuuint64_t)3 * Kp * 2, (cuuint64_t)3 * Kp * 2};
      cuuint32_t ab[4] = {64, 1, 1, 128};
      int rc2 = make_tmap(h, &h->tmProjA, h->d_split, 4, ad, as, ab);
      if (rc2) return rc2;
      cuuint64_t wd[2] = {(cuuint64_t)3 * Kp, (cuuint64_t)Np};
      cuuint64_t ws[1] = {(cuuint64_t)3 * Kp * 2};
      cuuint32_t wb[2] = {64, 128};
      if ((rc2 = make_tmap(h, &h->tmProjW, h->d_projW, 2, wd, ws, wb))) return rc2;
    }
    CK(h, fill_now(h->d_qpad, 0, sizeof(float) * h->projNp, h->own_stream));
    CK(h, copy_now(h->d_qpad, h->d_q, sizeof(float) * D, h->own_stream));
    pack_projector_bf16_kernel<<<cdiv((long long)h->projNp * h->projKp, 256), 256, 0, h->own_stream>>>(
        h->d_Nrow, h->d_projW, D, h->projKp, h->projNp);
    CK(h, cudaStreamSynchronize(h->own_stream));
    h->proj_tc = true;
  }
  return DAD_OK;
}

int dad_set_conditions(dad_handle *h, const int32_t *h_idx, const float *vals, int32_t n_cond, int32_t per_batch,
                       int32_t B) {
  if (!h) return DAD_ERR_INVALID;
  CK(h, cudaSetDevice(h->cfg.device));
  DAD_QUIESCE(h);
  if (n_cond <= 0) { h->n_cond = 0; return DAD_OK; }
  if (!h_idx || !vals) return DAD_ERR_INVALID;
  if (n_cond > kMaxCond) DAD_FAIL(h, DAD_ERR_INVALID, "at most %d conditions are supported (got %d)", kMaxCond, n_cond);
  for (int i = 0; i < n_cond; ++i) {
    int hh = h_idx[i];
    if (hh < 0) hh += h->cfg.horizon;       // python-style negative index (e.g. {-1: goal})
    if (hh < 0 || hh >= h->cfg.horizon) DAD_FAIL(h, DAD_ERR_INVALID, "condition index %d outside the horizon", h_idx[i]);
    h->cond_h[i] = hh;
  }
  const size_t need = (size_t)n_cond * (per_batch ? B : 1) * h->cfg.transition_dim;
  if (need > h->cond_cap) {
    int rc;
    if ((rc = dev_regrow(h, &h->d_cond, need))) return rc;
    h->cond_cap = need;
    drop_graphs(h);
  }
  CK(h, copy_now(h->d_cond, vals, sizeof(float) * need, h->own_stream));
  h->n_cond = n_cond;
  h->cond_per_batch = per_batch ? 1 : 0;
  h->cond_B = per_batch ? B : 1;
  return DAD_OK;
}

int dad_unet_forward(dad_handle *h, const float *x, const int64_t *t, int32_t step, float *eps, int32_t B, void *stream) {
  if (!h || !x || !eps || B < 1) return DAD_ERR_INVALID;
  NvtxRange nvtx("dad_unet_forward");
  if (!h->have_weights) DAD_FAIL(h, DAD_ERR_STATE, "dad_unet_forward before dad_load_weights");
  if (!t && (step < 0 || step >= h->cfg.n_timesteps)) DAD_FAIL(h, DAD_ERR_INVALID, "step %d outside [0, %d)", step, h->cfg.n_timesteps);
  CK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int c0 = 0; c0 < B; c0 += h->cfg.max_batch) {
    const int Bc = std::min(h->cfg.max_batch, B - c0);
    LoopState ls{};
    ls.step = step;
    ls.n_steps = step + 1;
    ls.x = const_cast<float *>(x) + (size_t)c0 * h->D;
    ls.t_rows = t ? reinterpret_cast<const long long *>(t) + c0 : nullptr;
    ls.n_table = h->cfg.n_timesteps;
    int rc = set_loop_state(h, ls, st);
    if (rc) return rc;
    h->counting = 0;
    h->rows_t = t != nullptr;
    rc = enqueue_unet(h, Bc, st);
    h->rows_t = false;
    if (rc) return rc;
    h->launches += h->counting;
    CK(h, cudaMemcpyAsync(eps + (size_t)c0 * h->D, h->d_eps, sizeof(float) * (size_t)Bc * h->D, cudaMemcpyDeviceToDevice, st));
  }
  return DAD_OK;
}

int dad_step(dad_handle *h, float *x, const float *model_out, const float *noise, const float *grad, float guide_w,
             int32_t step, uint32_t flags, uint64_t seed, uint64_t sample_offset, int32_t B, void *stream) {
  if (!h || !x || !model_out || B < 1) return DAD_ERR_INVALID;
  NvtxRange nvtx("dad_step");
  if (!h->have_sched) DAD_FAIL(h, DAD_ERR_STATE, "dad_step before dad_set_schedule");
  if (step < 0 || step >= h->cfg.n_timesteps) DAD_FAIL(h, DAD_ERR_INVALID, "step %d outside [0, %d)", step, h->cfg.n_timesteps);
  const bool project = (flags & DAD_FLAG_PROJECT) != 0;
  if (project && !h->projD) DAD_FAIL(h, DAD_ERR_STATE, "DAD_FLAG_PROJECT without dad_set_projector");
  CK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int c0 = 0; c0 < B; c0 += h->cfg.max_batch) {
    const int Bc = std::min(h->cfg.max_batch, B - c0);
    const size_t o = (size_t)c0 * h->D;
    LoopState ls{};
    ls.step = step;
    ls.n_steps = step + 1;          // noise slot 0 == the tensor passed in
    ls.flags = flags;
    ls.guide_w = guide_w;
    ls.x = x + o;
    ls.noise = noise ? noise + o : nullptr;
    ls.noise_stride = 0;
    ls.grad = grad ? grad + o : nullptr;
    ls.seed = seed;
    ls.sample_offset = sample_offset + c0;
    fill_cond(h, ls, c0);
    int rc = set_loop_state(h, ls, st);
    if (rc) return rc;
    h->counting = 0;
    if ((rc = enqueue_step(h, model_out + o, Bc, project, false, st))) return rc;
    h->launches += h->counting;
  }
  return DAD_OK;
}

int dad_project(dad_handle *h, float *x, int32_t step, int32_t B, void *stream) {
  if (!h || !x || B < 1) return DAD_ERR_INVALID;
  if (!h->projD) DAD_FAIL(h, DAD_ERR_STATE, "dad_project without dad_set_projector");
  if (step < 0 || step >= h->cfg.n_timesteps) DAD_FAIL(h, DAD_ERR_INVALID, "step %d outside [0, %d)", step, h->cfg.n_timesteps);
  CK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int c0 = 0; c0 < B; c0 += h->cfg.max_batch) {
    const int Bc = std::min(h->cfg.max_batch, B - c0);
    float *xc = x + (size_t)c0 * h->D;
    // the GEMM reads whole rows while other CTAs write them: go through the scratch copy
    CK(h, cudaMemcpyAsync(h->d_xtmp, xc, sizeof(float) * (size_t)Bc * h->D, cudaMemcpyDeviceToDevice, st));
    LoopState ls{};
    ls.step = step;
    ls.n_steps = step + 1;
    ls.x = xc;
    int rc = set_loop_state(h, ls, st);
    if (rc) return rc;
    ConvF32Params g{};
    g.in1 = h->d_xtmp;
    g.w = h->d_Nt;
    g.bias = h->d_q;
    g.g.C1 = h->D; g.g.C2 = 0; g.g.Cout = h->D; g.g.taps = 1; g.g.tap_off[0] = 0;
    g.g.in_stride = 1; g.g.L_in = 1; g.g.L_out = 1; g.g.out_mul = 1; g.g.out_phase = 0;
    g.B = Bc;
    g.ls = h->d_ls;
    g.alpha_tab = h->d_alpha;
    g.cond_vals = h->d_cond;
    g.T = h->cfg.transition_dim;
    dim3 grid(cdiv(Bc, F32_BM), cdiv(h->D, F32_BN));
    conv_f32_kernel<EPI_PROJECT, true><<<grid, 256, 0, st>>>(g);
    h->launches += 1;
    CK(h, cudaGetLastError());
  }
  return DAD_OK;
}

int dad_sample(dad_handle *h, float *x, const float *noise_seq, uint64_t seed, uint64_t sample_offset, int32_t B,
               int32_t n_steps, uint32_t flags, float *trace, void *stream) {
  if (!h || !x || B < 1) return DAD_ERR_INVALID;
  NvtxRange nvtx("dad_sample");
  if (!h->have_weights || !h->have_sched) DAD_FAIL(h, DAD_ERR_STATE, "dad_sample before weights and schedule are set");
  if (n_steps < 1 || n_steps > h->cfg.n_timesteps)
    DAD_FAIL(h, DAD_ERR_INVALID, "n_steps %d outside [1, %d] (the schedule tables have n_timesteps entries)", n_steps, h->cfg.n_timesteps);
  const bool project = (flags & DAD_FLAG_PROJECT) != 0;
  if (project && !h->projD) DAD_FAIL(h, DAD_ERR_STATE, "DAD_FLAG_PROJECT without dad_set_projector");
  if ((flags & DAD_FLAG_CONDITIONS) && h->n_cond && h->cond_per_batch && h->cond_B != B)
    DAD_FAIL(h, DAD_ERR_INVALID, "per-batch conditions were registered for B=%d, sampling B=%d", h->cond_B, B);
  CK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t D = h->D;
  // the loop state this call installs has Philox noise, no gradient and no trace: the lean step kernel applies
  h->step_lean = !noise_seq && !trace;
  struct LeanReset { dad_handle *h; ~LeanReset() { h->step_lean = false; } } lean_reset{h};
  for (int c0 = 0; c0 < B; c0 += h->cfg.max_batch) {
    const int Bc = std::min(h->cfg.max_batch, B - c0);
    LoopState ls{};
    ls.step = n_steps;          // every graph replay first decrements it
    ls.n_steps = n_steps;
    ls.flags = flags;
    ls.x = x + (size_t)c0 * D;
    ls.noise = noise_seq ? noise_seq + (size_t)c0 * D : nullptr;
    ls.noise_stride = (long long)B * D;
    ls.trace = trace ? trace + (size_t)c0 * D : nullptr;
    ls.trace_stride = (long long)B * D;
    ls.seed = seed;
    ls.sample_offset = sample_offset + c0;
    fill_cond(h, ls, c0);
    int rc = set_loop_state(h, ls, st);
    if (rc) return rc;
    const bool draw = (flags & DAD_FLAG_PHILOX_INIT) != 0;
    if (draw || ((flags & DAD_FLAG_CONDITIONS) && h->n_cond)) {
      init_x_kernel<<<cdiv((size_t)Bc * D / 4, 256), 256, 0, st>>>(h->d_ls, h->d_cond, Bc, (int)D, h->cfg.transition_dim, draw ? 1 : 0);
      h->launches += 1;
    }
    // leading steps routed through the fp32 companion (dad_set_fp32_steps): its U-Net reads x in place and leaves eps
    // in OUR eps buffer, the rest of the step is ours.  Noise / trace slots and Philox slots are indexed by the step, so
    // the split does not change the draws.
    int done = 0;
    if (h->companion && h->fp32_min_step < n_steps) {
      const int lead = n_steps - std::max(h->fp32_min_step, 0);
      for (; done < lead; ++done) {
        const int i = n_steps - 1 - done;
        const long long before = h->companion->launches;
        if ((rc = dad_unet_forward(h->companion, ls.x, nullptr, i, h->d_eps, Bc, stream)))
          DAD_FAIL(h, rc, "fp32 companion: %s", dad_last_error(h->companion));
        ls.step = i;
        if ((rc = set_loop_state(h, ls, st))) return rc;
        h->counting = 0;
        if ((rc = enqueue_step(h, h->d_eps, Bc, project, false, st))) return rc;
        h->launches += h->counting + (h->companion->launches - before);
      }
      if (done == n_steps) continue;
      ls.step = n_steps - done;       // every graph replay first decrements it
      if ((rc = set_loop_state(h, ls, st))) return rc;
    }
    // whole multiples of kStepsPerGraph through the multi-step graph, the remainder step by step
    const int left = n_steps - done;
    const int reps = std::min(left, kStepsPerGraph);
    GraphEntry *ge = nullptr, *ge1 = nullptr;
    if ((rc = get_graph(h, Bc, project, reps, &ge))) return rc;
    const int done0 = done;
    for (; done + reps <= n_steps; done += reps) CK(h, cudaGraphLaunch(ge->exec, st));
    h->launches += ge->kernels * ((done - done0) / reps);
    if (done < n_steps) {
      if ((rc = get_graph(h, Bc, project, 1, &ge1))) return rc;
      for (; done < n_steps; ++done) { CK(h, cudaGraphLaunch(ge1->exec, st)); h->launches += ge1->kernels; }
    }
  }
  return DAD_OK;
}

static bool stream_is_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
}

int dad_loop_begin(dad_handle *h, float *x, const float *noise, int32_t noise_single, const float *grad, float guide_w,
                   uint64_t seed, uint64_t sample_offset, int32_t B, int32_t n_steps, uint32_t flags, float *trace,
                   void *stream) {
  if (!h || !x || B < 1) return DAD_ERR_INVALID;
  if (!h->have_weights || !h->have_sched) DAD_FAIL(h, DAD_ERR_STATE, "dad_loop_begin before weights and schedule are set");
  if (B > h->cfg.max_batch) DAD_FAIL(h, DAD_ERR_INVALID, "dad_loop_begin needs B <= max_batch (%d)", h->cfg.max_batch);
  if (n_steps < 1 || n_steps > h->cfg.n_timesteps)
    DAD_FAIL(h, DAD_ERR_INVALID, "n_steps %d outside [1, %d] (the schedule tables have n_timesteps entries)", n_steps, h->cfg.n_timesteps);
  if ((flags & DAD_FLAG_PROJECT) && !h->projD) DAD_FAIL(h, DAD_ERR_STATE, "DAD_FLAG_PROJECT without dad_set_projector");
  if ((flags & DAD_FLAG_CONDITIONS) && h->n_cond && h->cond_per_batch && h->cond_B != B)
    DAD_FAIL(h, DAD_ERR_INVALID, "per-batch conditions were registered for B=%d, sampling B=%d", h->cond_B, B);
  CK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LoopState ls{};
  ls.step = n_steps;            // dad_loop_unet decrements it first
  ls.n_steps = n_steps;
  ls.flags = flags;
  ls.guide_w = guide_w;
  ls.x = x;
  ls.noise = noise;
  ls.noise_stride = noise_single ? 0 : (long long)B * h->D;
  ls.grad = grad;
  ls.trace = trace;
  ls.trace_stride = (long long)B * h->D;
  ls.seed = seed;
  ls.sample_offset = sample_offset;
  fill_cond(h, ls, 0);
  int rc = set_loop_state(h, ls, st);
  if (rc) return rc;
  const bool draw = (flags & DAD_FLAG_PHILOX_INIT) != 0;
  if (draw || ((flags & DAD_FLAG_CONDITIONS) && h->n_cond)) {
    init_x_kernel<<<cdiv((size_t)B * h->D / 4, 256), 256, 0, st>>>(h->d_ls, h->d_cond, B, h->D, h->cfg.transition_dim, draw ? 1 : 0);
    h->launches += 1;
    CK(h, cudaGetLastError());
  }
  return DAD_OK;
}

int dad_loop_unet(dad_handle *h, int32_t B, void *stream) {
  if (!h || B < 1 || B > h->cfg.max_batch) return DAD_ERR_INVALID;
  if (!h->have_weights) DAD_FAIL(h, DAD_ERR_STATE, "dad_loop_unet before dad_load_weights");
  CK(h, cudaSetDevice(h->cfg.device));
  h->counting = 0;
  const int rc = enqueue_unet(h, B, reinterpret_cast<cudaStream_t>(stream), true, nullptr);
  h->loop_kernels = h->counting;
  if (!stream_is_capturing(reinterpret_cast<cudaStream_t>(stream))) h->launches += h->counting;
  return rc;
}

int dad_loop_step(dad_handle *h, int32_t B, uint32_t flags, void *stream) {
  if (!h || B < 1 || B > h->cfg.max_batch) return DAD_ERR_INVALID;
  if (!h->have_sched) DAD_FAIL(h, DAD_ERR_STATE, "dad_loop_step before dad_set_schedule");
  const bool project = (flags & DAD_FLAG_PROJECT) != 0;
  if (project && !h->projD) DAD_FAIL(h, DAD_ERR_STATE, "DAD_FLAG_PROJECT without dad_set_projector");
  CK(h, cudaSetDevice(h->cfg.device));
  h->counting = 0;
  const int rc = enqueue_step(h, h->d_eps, B, project, false, reinterpret_cast<cudaStream_t>(stream));
  h->loop_kernels += h->counting;
  if (!stream_is_capturing(reinterpret_cast<cudaStream_t>(stream))) h->launches += h->counting;
  return rc;
}

int64_t dad_graph_epoch(const dad_handle *h) { return h ? h->epoch : -1; }

int dad_build_projection_matrix(int32_t device, const double *F, int32_t rows, int32_t cols, float *P) {
  char buf[256];
  if (!F || !P || rows < 1 || cols < 1 || cols > rows) { g_create_error = "dad_build_projection_matrix: bad arguments"; return DAD_ERR_INVALID; }
  cudaError_t e = cudaSetDevice(device);
  int bad = 0;
  if (e == cudaSuccess) e = build_projection_matrix(F, rows, cols, P, &bad);
  if (e != cudaSuccess) {
    snprintf(buf, sizeof(buf), "dad_build_projection_matrix: %s", cudaGetErrorString(e));
    g_create_error = buf;
    return DAD_ERR_CUDA;
  }
  if (bad) {
    snprintf(buf, sizeof(buf), "dad_build_projection_matrix: F is not of full column rank (pivot %d of F^T F is not positive)", bad);
    g_create_error = buf;
    return DAD_ERR_INVALID;
  }
  return DAD_OK;
}

int dad_fit_linear_dynamics(int32_t device, const double *X, const double *U, const double *Xn, int64_t N, int32_t n,
                            int32_t m, double *A, double *B) {
  char buf[256];
  if (!X || !U || !Xn || !A || !B || N < 1 || n < 1 || m < 1 || N < n + m) {
    g_create_error = "dad_fit_linear_dynamics: bad arguments (need N >= n + m transitions)";
    return DAD_ERR_INVALID;
  }
  cudaError_t e = cudaSetDevice(device);
  int bad = 0;
  if (e == cudaSuccess) e = fit_linear_dynamics_device(X, U, Xn, N, n, m, A, B, &bad);
  if (e != cudaSuccess) {
    snprintf(buf, sizeof(buf), "dad_fit_linear_dynamics: %s", cudaGetErrorString(e));
    g_create_error = buf;
    return DAD_ERR_CUDA;
  }
  if (bad) {
    snprintf(buf, sizeof(buf), "dad_fit_linear_dynamics: [X U] is not of full column rank (pivot %d): the transitions do not identify (A, B)", bad);
    g_create_error = buf;
    return DAD_ERR_INVALID;
  }
  return DAD_OK;
}

int dad_dynamics_residual(int32_t device, const float *x, int32_t B, int32_t H, int32_t n, int32_t m, const float *P,
                          const float *obs_mean, const float *obs_std, const float *act_mean, const float *act_std,
                          double *out, void *stream) {
  char buf[256];
  if (!x || !P || !obs_mean || !obs_std || !act_mean || !act_std || !out || B < 1 || H < 1 || n < 1 || m < 1) {
    g_create_error = "dad_dynamics_residual: bad arguments";
    return DAD_ERR_INVALID;
  }
  cudaError_t e = cudaSetDevice(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float *d_stats = nullptr;
  double *d_out = nullptr;
  std::vector<float> stats;
  stats.insert(stats.end(), obs_mean, obs_mean + n);
  stats.insert(stats.end(), obs_std, obs_std + n);
  stats.insert(stats.end(), act_mean, act_mean + m);
  stats.insert(stats.end(), act_std, act_std + m);
  double sum = 0.0;
  if (e == cudaSuccess) e = cudaMalloc(&d_stats, stats.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_stats, stats.data(), stats.size() * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0, sizeof(double), st);
  if (e == cudaSuccess) {
    ResidualParams p{};
    p.x = x; p.P = P; p.stats = d_stats; p.out = d_out;
    p.B = B; p.H = H; p.n = n; p.m = m; p.T = n + m; p.Dc = (H + 1) * n + H * m;
    residual_kernel<<<dim3((p.Dc + RS_N - 1) / RS_N, (B + RS_S - 1) / RS_S), RS_N, 0, st>>>(p);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&sum, d_out, sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    *out = sum / ((double)B * (double)p.Dc);
  }
  cudaFree(d_stats);
  cudaFree(d_out);
  if (e != cudaSuccess) {
    snprintf(buf, sizeof(buf), "dad_dynamics_residual: %s", cudaGetErrorString(e));
    g_create_error = buf;
    return DAD_ERR_CUDA;
  }
  return DAD_OK;
}

int dad_loop_replayed(dad_handle *h, int32_t n) {
  if (!h || n < 0) return DAD_ERR_INVALID;
  h->launches += h->loop_kernels * n;
  return DAD_OK;
}

int dad_sample_host(dad_handle *h, float *x_host, const float *noise_seq_host, uint64_t seed, uint64_t sample_offset,
                    int32_t B, int32_t n_steps, uint32_t flags) {
  if (!h || !x_host || B < 1) return DAD_ERR_INVALID;
  CK(h, cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)B * h->D;
  if (n > h->hostx_cap) {
    int rc;
    CK(h, cudaStreamSynchronize(h->own_stream));
    if ((rc = dev_regrow(h, &h->d_hostx, n))) return rc;
    h->hostx_cap = n;
  }
  const float *d_noise = nullptr;
  if (noise_seq_host) {
    const size_t nn = n * (size_t)n_steps;
    if (nn > h->hostnoise_cap) {
      int rc;
      CK(h, cudaStreamSynchronize(h->own_stream));
      if ((rc = dev_regrow(h, &h->d_hostnoise, nn))) return rc;
      h->hostnoise_cap = nn;
    }
    CK(h, cudaMemcpyAsync(h->d_hostnoise, noise_seq_host, sizeof(float) * nn, cudaMemcpyHostToDevice, h->own_stream));
    d_noise = h->d_hostnoise;
  }
  if (!(flags & DAD_FLAG_PHILOX_INIT))
    CK(h, cudaMemcpyAsync(h->d_hostx, x_host, sizeof(float) * n, cudaMemcpyHostToDevice, h->own_stream));
  int rc = dad_sample(h, h->d_hostx, d_noise, seed, sample_offset, B, n_steps, flags, nullptr, h->own_stream);
  if (rc) return rc;
  CK(h, cudaMemcpyAsync(x_host, h->d_hostx, sizeof(float) * n, cudaMemcpyDeviceToHost, h->own_stream));
  CK(h, cudaStreamSynchronize(h->own_stream));
  return DAD_OK;
}

int dad_get_info(const dad_handle *h, dad_info *out) {
  if (!h || !out) return DAD_ERR_INVALID;
  out->conv_flops_per_sample = h->conv_flops;
  long long per_step = 1;  // stage_x
  for (const LaunchUnit &lu : h->units) {
    if (lu.chain >= 0) per_step += 1;
    else per_step += (h->bf16 || h->ops[lu.op].gname.empty()) ? 1 : 2;
  }
  per_step += 1;           // fused step kernel (the projector GEMM adds one for large D)
  out->launches_per_step = per_step;
  out->workspace_bytes = (int64_t)(h->act_bytes_per_sample * (size_t)h->cfg.max_batch);
  out->n_conv_layers = (int32_t)h->ops.size();
  out->sm_count = h->sm_count;
  return DAD_OK;
}

int64_t dad_launch_count(const dad_handle *h) { return h ? h->launches : 0; }

int dad_sample_profile(dad_handle *h, float *x, uint64_t seed, uint64_t sample_offset, int32_t B, int32_t n_steps,
                       uint32_t flags, float *step_ms, void *stream) {
  if (!h || !x || !step_ms || B < 1) return DAD_ERR_INVALID;
  if (!h->have_weights || !h->have_sched) DAD_FAIL(h, DAD_ERR_STATE, "dad_sample_profile before weights and schedule are set");
  if (B > h->cfg.max_batch) DAD_FAIL(h, DAD_ERR_INVALID, "dad_sample_profile needs B <= max_batch");
  if (n_steps < 1 || n_steps > h->cfg.n_timesteps) DAD_FAIL(h, DAD_ERR_INVALID, "n_steps outside the schedule");
  const bool project = (flags & DAD_FLAG_PROJECT) != 0;
  if (project && !h->projD) DAD_FAIL(h, DAD_ERR_STATE, "DAD_FLAG_PROJECT without dad_set_projector");
  CK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LoopState ls{};
  ls.step = n_steps;            // every graph replay first decrements it
  ls.n_steps = n_steps;
  ls.flags = flags;
  ls.x = x;
  ls.seed = seed;
  ls.sample_offset = sample_offset;
  fill_cond(h, ls, 0);
  int rc = set_loop_state(h, ls, st);
  if (rc) return rc;
  const bool draw = (flags & DAD_FLAG_PHILOX_INIT) != 0;
  if (draw || ((flags & DAD_FLAG_CONDITIONS) && h->n_cond)) {
    init_x_kernel<<<cdiv((size_t)B * h->D / 4, 256), 256, 0, st>>>(h->d_ls, h->d_cond, B, h->D, h->cfg.transition_dim, draw ? 1 : 0);
    h->launches += 1;
  }
  GraphEntry *ge = nullptr;
  h->step_lean = true;           // Philox noise, no trace: the kernels dad_sample runs for such a loop
  rc = get_graph(h, B, project, 1, &ge);
  h->step_lean = false;
  if (rc) return rc;
  std::vector<cudaEvent_t> ev(n_steps + 1);
  for (auto &e : ev) CK(h, cudaEventCreate(&e));
  CK(h, cudaEventRecord(ev[0], st));
  for (int s = 0; s < n_steps; ++s) {
    CK(h, cudaGraphLaunch(ge->exec, st));
    CK(h, cudaEventRecord(ev[s + 1], st));
  }
  h->launches += ge->kernels * n_steps;
  CK(h, cudaStreamSynchronize(st));
  for (int s = 0; s < n_steps; ++s) CK(h, cudaEventElapsedTime(&step_ms[s], ev[s], ev[s + 1]));
  for (auto &e : ev) cudaEventDestroy(e);
  return DAD_OK;
}

int dad_set_fp32_steps(dad_handle *h, dad_handle *companion, int32_t min_step) {
  if (!h) return DAD_ERR_INVALID;
  if (!companion) { h->companion = nullptr; h->fp32_min_step = INT_MAX; return DAD_OK; }
  if (companion == h || companion->bf16) DAD_FAIL(h, DAD_ERR_INVALID, "dad_set_fp32_steps: the companion must be another handle of fp32 precision");
  if (!h->bf16) DAD_FAIL(h, DAD_ERR_INVALID, "dad_set_fp32_steps: this handle already computes in fp32");
  const dad_config &a = h->cfg, &b = companion->cfg;
  bool same = a.device == b.device && a.transition_dim == b.transition_dim && a.dim == b.dim && a.n_levels == b.n_levels &&
              a.kernel_size == b.kernel_size && a.horizon == b.horizon && a.n_timesteps == b.n_timesteps;
  for (int l = 0; same && l < a.n_levels; ++l) same = a.dim_mults[l] == b.dim_mults[l];
  if (!same) DAD_FAIL(h, DAD_ERR_INVALID, "dad_set_fp32_steps: the companion was created for a different architecture / device");
  if (min_step < 0) DAD_FAIL(h, DAD_ERR_INVALID, "dad_set_fp32_steps: min_step %d < 0", min_step);
  h->companion = companion;
  h->fp32_min_step = min_step;
  return DAD_OK;
}

int dad_set_fp32_math(dad_handle *h, int32_t mode) {
  if (!h || mode < 0 || mode > 2) return DAD_ERR_INVALID;
  if (h->bf16) DAD_FAIL(h, DAD_ERR_INVALID, "dad_set_fp32_math applies to fp32-precision handles");
  h->f32_math = mode;
  return DAD_OK;
}

int dad_set_latency_batch(dad_handle *h, int32_t max_b) {
  if (!h || max_b < 0) return DAD_ERR_INVALID;
  if (max_b != h->small_max_b) {
    cudaDeviceSynchronize();
    drop_graphs(h);           // captured steps chose their kernels by batch size
    h->small_max_b = max_b;
  }
  return DAD_OK;
}

int dad_layer_count(const dad_handle *h) { return h ? (int)h->ops.size() : 0; }

int dad_layer_info(const dad_handle *h, int32_t index, dad_layer_desc *out) {
  if (!h || !out || index < 0 || index >= (int)h->ops.size()) return DAD_ERR_INVALID;
  const ConvOp &op = h->ops[index];
  memset(out, 0, sizeof(*out));
  std::string nm = op.wname.substr(0, op.wname.size() - 7);   // strip ".weight"
  if (op.transposed) nm += op.g.out_phase ? "[odd]" : "[even]";
  snprintf(out->name, sizeof(out->name), "%s", nm.c_str());
  out->L_out = op.g.L_out;
  out->C_in = op.Cin_real;
  out->C_out = op.g.Cout;
  out->taps = op.g.taps;
  out->tile_n = op.t3 ? 128 * op.t3_NS : op.BN;
  out->group_width = op.GW;
  if (!h->bf16) snprintf(out->kernel, sizeof(out->kernel), "conv_f32_kernel%s", op.gname.empty() ? "" : "+gn_mish_f32_kernel");
  else if (op.chain >= 0) {
    const ChainUnit &cu = h->chains[h->units[op.chain].chain];
    snprintf(out->kernel, sizeof(out->kernel), "conv_chain_kernel<GW=%d,MH=%d,N=%d>", cu.GW, cu.MH, 128 * cu.NS);
    out->tile_n = 128 * cu.NS;
  } else if (op.t3) {
    static const char *modes[3] = {"single", "mcast", "pair"};
    snprintf(out->kernel, sizeof(out->kernel), "conv_t3_kernel<GW=%d,MH=%d,%s,N=%d>", op.GW, op.t3_MH, modes[op.t3_mode], 128 * op.t3_NS);
  } else snprintf(out->kernel, sizeof(out->kernel), "conv_tc_kernel<%d,%d>", op.BN, op.GW);
  out->flops_per_sample = 2LL * op.g.L_out * op.g.taps * op.Cin_real * op.g.Cout;
  return DAD_OK;
}


int dad_time_layer(dad_handle *h, int32_t index, int32_t B, int32_t iters, float *ms, void *stream) {
  if (!h || !ms || index < 0 || index >= (int)h->ops.size() || iters < 1 || B < 1 || B > h->cfg.max_batch) return DAD_ERR_INVALID;
  if (!h->have_weights) DAD_FAIL(h, DAD_ERR_STATE, "dad_time_layer before dad_load_weights");
  CK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LoopState ls{};
  ls.step = 0;
  ls.n_steps = 1;
  ls.x = h->d_xtmp;
  int rc = set_loop_state(h, ls, st);
  if (rc) return rc;
  ConvOp &op = h->ops[index];
  unsigned long long *prof = nullptr;
  if (h->bf16 && tuning_env("DAD_TC_PROF", 0)) {
    CK(h, cudaMalloc(&prof, 4 * sizeof(unsigned long long)));
    CK(h, fill_now(prof, 0, 4 * sizeof(unsigned long long), h->own_stream));
    op.tcp.prof = prof;
    op.t3p.prof = prof;
  }
  h->counting = 0;
  rc = time_launches(h, st, iters, ms, [&]() { return h->bf16 ? enqueue_tc(h, op, B, st) : enqueue_f32(h, op, B, st); });
  h->launches += h->counting;
  if (prof) {
    unsigned long long v[4];
    cudaMemcpy(v, prof, sizeof(v), cudaMemcpyDeviceToHost);
    const double n = v[3] ? (double)v[3] : 1.0;     // warp-tiles
    fprintf(stderr, "[prof] layer %d %s: per warp-tile cycles: wait %.0f  pass1 %.0f  pass2 %.0f  (warp-tiles %llu over %d launches)\n",
            index, op.wname.c_str(), v[0] / n, v[1] / n, v[2] / n, v[3], iters + 1);
    op.tcp.prof = nullptr;
    op.t3p.prof = nullptr;
    cudaFree(prof);
  }
  return rc;
}

int dad_set_fusion(dad_handle *h, int32_t level) {
  if (!h || level < 0 || level > 3) return DAD_ERR_INVALID;
  if (!h->bf16) return DAD_OK;                 // fp32 mode has no chains
  if (level == h->fusion) return DAD_OK;
  CK(h, cudaSetDevice(h->cfg.device));
  CK(h, cudaDeviceSynchronize());
  drop_graphs(h);
  h->fusion = level;
  return build_units(h);
}

int dad_debug_counters(dad_handle *h, uint32_t *out4, int32_t reset) {
  if (!h || !out4) return DAD_ERR_INVALID;
  for (int i = 0; i < 4; ++i) out4[i] = 0;
  if (!h->d_err) return DAD_OK;
  CK(h, cudaSetDevice(h->cfg.device));
  CK(h, cudaDeviceSynchronize());
  CK(h, cudaMemcpy(out4, h->d_err, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost));
  if (reset) CK(h, cudaMemset(h->d_err, 0, 4 * sizeof(unsigned)));
  return DAD_OK;
}

int dad_unit_count(const dad_handle *h) { return h ? (int)h->units.size() : 0; }

int dad_unit_info(const dad_handle *h, int32_t index, dad_unit_desc *out) {
  if (!h || !out || index < 0 || index >= (int)h->units.size()) return DAD_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  const LaunchUnit &lu = h->units[index];
  std::vector<int> ops;
  if (lu.chain >= 0) ops = h->chains[lu.chain].ops; else ops.push_back(lu.op);
  out->first_layer = ops.front();
  out->n_layers = (int)ops.size();
  out->is_chain = lu.chain >= 0 ? 1 : 0;
  const ConvOp &last = h->ops[ops.back()];
  out->L_out = last.g.L_out;
  out->C_out = last.g.Cout;
  for (int oi : ops) out->flops_per_sample += 2LL * h->ops[oi].g.L_out * h->ops[oi].g.taps * h->ops[oi].Cin_real * h->ops[oi].g.Cout;
  dad_layer_desc d;
  dad_layer_info(h, ops.front(), &d);
  snprintf(out->kernel, sizeof(out->kernel), "%s", d.kernel);
  return DAD_OK;
}

int dad_time_unit(dad_handle *h, int32_t index, int32_t B, int32_t iters, float *ms, void *stream) {
  if (!h || !ms || index < 0 || index >= (int)h->units.size() || iters < 1 || B < 1 || B > h->cfg.max_batch) return DAD_ERR_INVALID;
  if (!h->have_weights) DAD_FAIL(h, DAD_ERR_STATE, "dad_time_unit before dad_load_weights");
  const LaunchUnit &lu = h->units[index];
  if (lu.chain < 0) return dad_time_layer(h, lu.op, B, iters, ms, stream);
  CK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LoopState ls{};
  ls.step = 0;
  ls.n_steps = 1;
  ls.x = h->d_xtmp;
  int rc = set_loop_state(h, ls, st);
  if (rc) return rc;
  ChainUnit &cu = h->chains[lu.chain];
  if (!use_chain(h, cu, B)) DAD_FAIL(h, DAD_ERR_INVALID, "unit %d does not run as a chain at B=%d (latency kernels)", index, B);
  CK(h, cudaMemsetAsync(h->d_flags, 0, h->ops.size() * (size_t)h->tiles_cap * sizeof(unsigned), st));
  int epoch = 0;
  h->counting = 0;
  rc = time_launches(h, st, iters, ms, [&]() { return enqueue_chain(h, cu, B, st, ++epoch); });
  h->launches += h->counting;
  // leave the counters as a U-Net pass expects to find them
  CK(h, cudaMemsetAsync(h->d_flags, 0, h->ops.size() * (size_t)h->tiles_cap * sizeof(unsigned), st));
  CK(h, cudaStreamSynchronize(st));
  return rc;
}

int dad_time_step_kernel(dad_handle *h, int32_t B, int32_t step, uint32_t flags, int32_t iters, float *ms, void *stream) {
  if (!h || !ms || iters < 1 || B < 1 || B > h->cfg.max_batch) return DAD_ERR_INVALID;
  if (!h->have_sched) DAD_FAIL(h, DAD_ERR_STATE, "dad_time_step_kernel before dad_set_schedule");
  if (step < 0 || step >= h->cfg.n_timesteps) return DAD_ERR_INVALID;
  const bool project = (flags & DAD_FLAG_PROJECT) != 0;
  if (project && !h->projD) DAD_FAIL(h, DAD_ERR_STATE, "DAD_FLAG_PROJECT without dad_set_projector");
  CK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = (size_t)B * h->D;
  if (n > h->hostx_cap) {
    int rc;
    CK(h, cudaDeviceSynchronize());
    if ((rc = dev_regrow(h, &h->d_hostx, n))) return rc;
    h->hostx_cap = n;
  }
  CK(h, cudaMemsetAsync(h->d_hostx, 0, n * sizeof(float), st));
  const bool injected = (flags & 0x100u) != 0;      // measurement only: read the noise from a buffer instead of Philox
  h->force_proj_tc = (flags & 0x200u) != 0;         // measurement only: projector on the tensor cores whatever the batch
  const bool fused_simt = (flags & 0x400u) != 0;    // measurement only: the fused SIMT projector whatever the batch
  flags &= ~0x700u;
  const int saved_min_batch = h->proj_tc_min_batch;
  if (fused_simt) h->proj_tc_min_batch = INT_MAX;
  if (injected) {
    if (n > h->hostnoise_cap) {
      int rc;
      CK(h, cudaDeviceSynchronize());
      if ((rc = dev_regrow(h, &h->d_hostnoise, n))) return rc;
      h->hostnoise_cap = n;
    }
    CK(h, cudaMemsetAsync(h->d_hostnoise, 0, n * sizeof(float), st));
  }
  LoopState ls{};
  ls.step = step;
  ls.n_steps = step + 1;
  ls.flags = flags;
  ls.x = h->d_hostx;
  ls.noise = injected ? h->d_hostnoise : nullptr;
  ls.seed = 1;
  fill_cond(h, ls, 0);
  int rc = set_loop_state(h, ls, st);
  if (rc) return rc;
  h->counting = 0;
  h->step_lean = !injected;          // what dad_sample selects for a loop without injected noise / trace
  rc = time_launches(h, st, iters, ms, [&]() { return enqueue_step(h, h->d_eps, B, project, false, st); });
  h->step_lean = false;
  h->launches += h->counting;
  h->force_proj_tc = false;
  h->proj_tc_min_batch = saved_min_batch;
  return rc;
}

}  // extern "C"
