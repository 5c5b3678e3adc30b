// Shared definitions for the sampler kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace dad {

// Kernel-side ablation bits (skip the epilogue arithmetic, the MMAs, ...) and cycle counters are compiled in only
// with -DDAD_TUNING; in the shipping build the expressions are constants and the branches disappear.
#ifdef DAD_TUNING
#define DAD_DEBUG_BITS(p) ((p).debug)
#define DAD_PROF_PTR(p) ((p).prof)
#else
#define DAD_DEBUG_BITS(p) 0
#define DAD_PROF_PTR(p) (static_cast<unsigned long long *>(nullptr))
#endif

constexpr int kMaxTaps = 8;
constexpr int kMaxCond = 8;
constexpr int kGroups = 8;          // nn.GroupNorm(8, C), temporal_unet.py:67
constexpr float kGnEps = 1e-5f;     // torch default

// Per-loop state that lives in device memory so that ONE captured graph can be replayed for
// every step, every chunk and every plan: kernels read it instead of taking by-value params.
struct LoopState {
  int step;                 // current step index i (decremented by advance_step_kernel)
  int n_steps;              // loop length; noise slot k = n_steps - 1 - step
  unsigned flags;           // DAD_FLAG_*
  float guide_w;
  float *x;                 // (B,H,T) trajectories of the current chunk
  const float *noise;       // slot 0 of the injected noise for this chunk, or nullptr -> Philox
  long long noise_stride;   // elements between slots
  const float *grad;        // guide gradient (B,H,T) or nullptr
  float *trace;             // slot 0 of the per-step trace for this chunk, or nullptr
  long long trace_stride;
  unsigned long long seed;
  unsigned long long sample_offset;  // global index of row 0 of this chunk (Philox subsequence)
  const long long *t_rows;  // per-row timesteps for a stand-alone forward, or nullptr (uniform `step`)
  int n_table;              // rows of the time tables: per-row timesteps are clamped to [0, n_table) on the device
  // conditions (GuidedPolicy.apply_conditions)
  int n_cond;
  int cond_per_batch;
  int cond_B;               // batch stride of cond_vals when per_batch (the FULL batch, not the chunk)
  int cond_row0;            // row of the full batch at which this chunk starts
  int cond_h[kMaxCond];
};

// One implicit-GEMM convolution over channels-last activations.
//   out[b, lo*out_mul + out_phase, n] = bias[n] + sum_{tap, c} in[b, lo*in_stride + tap_off[tap], c] * W[tap, c, n]
// with `in` the channel-concatenation of in1 (C1 channels) and in2 (C2 channels, may be 0).
struct ConvGeom {
  int C1, C2, Cout;
  int taps;
  int tap_off[kMaxTaps];
  int in_stride;            // 1, or 2 for Downsample1d
  int L_in;                 // rows per sample of the input
  int L_out;                // rows per sample this launch produces
  int out_mul, out_phase;   // 2/phase for the two ConvTranspose1d phases, else 1/0
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; `red` is >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float *red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

// nn.Mish, accurate form for the fp32 mode: x * tanh(softplus(x)).
__device__ __forceinline__ float mish_precise(float x) {
  if (x > 20.f) return x;
  const float sp = log1pf(expf(x));
  return x * tanhf(sp);
}

// nn.Mish, fast form for the bf16 mode: tanh(softplus(x)) = n / (n + 2), n = e^x (e^x + 2).
// One ex2 + one rcp on the SFU; relative error ~1e-6, far below bf16 resolution.
__device__ __forceinline__ float mish_fast(float x) {
  const float e = __expf(fminf(x, 20.f));
  const float n = e * (e + 2.f);
  return x * __fdividef(n, n + 2.f);
}

}  // namespace dad
