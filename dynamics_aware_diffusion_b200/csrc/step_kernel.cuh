// The memory-bound remainder of one reverse-diffusion step, fused (K7):
//   x0 = sqrt_recip[i] x - sqrt_recipm1[i] eps      predict_start_from_noise   diffusion.py:159-166
//   x0 = clamp(x0, -1, 1)                            clip_denoised              diffusion.py:199-200
//   mu = c1[i] x0 + c2[i] x                          q_posterior                diffusion.py:168-180
//   mu += w * exp(logvar[i]) * grad                  guidance                   policies.py:87-97
//   x' = mu + [i != 0] exp(0.5 logvar[i]) z          noise                      diffusion.py:217-223
//   x' = x' + alpha[i] (N x' + q)                    dynamics projector         policies.py:409-485
//   x'[:, h, :] = cond                               inpainting                 policies.py:48-63
// 128-bit loads/stores over the flattened (B, H*T) batch; the small-D projector variant keeps the
// projector in shared memory and reduces with warp shuffles.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace dad {

struct StepParams {
  const LoopState *ls;
  const float *model_out;       // eps (or x0 when predict_epsilon == 0), (B, H*T)
  float *xtmp;                  // destination when a separate projector GEMM follows, else unused
  const float *sqrt_recip, *sqrt_recipm1, *coef1, *coef2, *logvar;
  const float *alpha_tab;       // per-step projector strength (fused variant)
  const float *Nt;              // projector, TRANSPOSED: Nt[k*D + d] = N[d][k] (fused variant)
  const float *q;
  const float *cond_vals;
  int B, D, T;
  int predict_epsilon, clip_denoised;
  int to_tmp;                   // 1: write x' to xtmp and skip inpainting (projector GEMM follows)
  __nv_bfloat16 *split;         // tensor-core projector operand [B][3*Kp]: (hi | lo | hi) bf16 split of x', or nullptr
  int Kp;
};

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: (quad index, step index, sample lo, sample hi)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// Four standard normals from one Philox block (Box-Muller on two pairs).  The transcendental part uses the SFU
// approximations (lg2 / sqrt / sin / cos .approx, |error| ~1e-6): the generator sits inside a kernel that must
// stream at HBM speed, and the library-precision versions cost ~4x the instructions.
__device__ __forceinline__ float4 philox_normal4(unsigned quad, unsigned slot, unsigned long long sample,
                                                 unsigned long long seed) {
  const uint4 r = philox4x32_10(make_uint4(quad, slot, (unsigned)sample, (unsigned)(sample >> 32)),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)r.x + 0.5f) * k, u1 = ((float)r.y + 0.5f) * k;
  const float u2 = ((float)r.z + 0.5f) * k, u3 = ((float)r.w + 0.5f) * k;
  float r0, r1;
  // sqrt(-2 ln u) = sqrt(-2 ln2 * lg2 u)
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(-1.3862943611198906f * __log2f(u0)));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(-1.3862943611198906f * __log2f(u2)));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

constexpr unsigned kInitSlot = 0xFFFFFFFFu;  // Philox slot of the initial x_S draw

// x_S ~ N(0, I) drawn in-kernel (torch.randn at diffusion.py:241 / policies.py:134), + initial inpainting
// (policies.py:137-138).  One thread per float4.
__global__ void __launch_bounds__(256) init_x_kernel(const LoopState *lsp, const float *cond_vals, int B,
                                                      int D, int T, int draw) {
  const LoopState ls = *lsp;
  const size_t q4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total4 = (size_t)B * D / 4;
  if (q4 >= total4) return;
  const size_t e = q4 * 4;
  const int b = (int)(e / D);
  const int d0 = (int)(e - (size_t)b * D);
  float4 v;
  if (draw) v = philox_normal4((unsigned)(d0 >> 2), kInitSlot, ls.sample_offset + b, ls.seed);
  else v = *reinterpret_cast<const float4 *>(ls.x + e);
  float vv[4] = {v.x, v.y, v.z, v.w};
  if (ls.flags & 1u) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + j, hh = d / T, tt = d - hh * T;
      for (int c = 0; c < ls.n_cond; ++c)
        if (ls.cond_h[c] == hh)
          vv[j] = cond_vals[((size_t)c * (ls.cond_per_batch ? ls.cond_B : 1) +
                             (ls.cond_per_batch ? (ls.cond_row0 + b) : 0)) * T + tt];
    }
  }
  *reinterpret_cast<float4 *>(ls.x + e) = make_float4(vv[0], vv[1], vv[2], vv[3]);
}

// The pointwise part for 4 consecutive elements of sample b starting at d0.
__device__ __forceinline__ void step_pointwise4(const StepParams &p, const LoopState &ls, int b, int d0,
                                                float cr, float crm1, float c1, float c2, float sig,
                                                float gvar, float out[4]) {
  const size_t e = (size_t)b * p.D + d0;
  const float4 x4 = *reinterpret_cast<const float4 *>(ls.x + e);
  const float4 m4 = __ldg(reinterpret_cast<const float4 *>(p.model_out + e));
  const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w};
  float zv[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
  if (sig != 0.f) {
    if (ls.noise) {
      const unsigned slot = (unsigned)(ls.n_steps - 1 - ls.step);
      const float4 z4 = __ldg(reinterpret_cast<const float4 *>(ls.noise + (size_t)slot * ls.noise_stride + e));
      zv[0] = z4.x; zv[1] = z4.y; zv[2] = z4.z; zv[3] = z4.w;
    } else {
      const float4 z4 = philox_normal4((unsigned)(d0 >> 2), (unsigned)ls.step, ls.sample_offset + b, ls.seed);
      zv[0] = z4.x; zv[1] = z4.y; zv[2] = z4.z; zv[3] = z4.w;
    }
  }
  if (ls.grad && gvar != 0.f) {
    const float4 g4 = __ldg(reinterpret_cast<const float4 *>(ls.grad + e));
    gv[0] = g4.x; gv[1] = g4.y; gv[2] = g4.z; gv[3] = g4.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x0 = p.predict_epsilon ? (cr * xv[j] - crm1 * mv[j]) : mv[j];
    if (p.clip_denoised) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float mu = c1 * x0 + c2 * xv[j];
    mu += gvar * gv[j];
    out[j] = mu + sig * zv[j];
  }
}

// x[:, h] = val for every registered (h, val): element d of the flattened (H*T) sample belongs to condition c
// iff 0 <= d - h_c*T < T (no division: this runs per element inside a bandwidth-bound kernel).
__device__ __forceinline__ float cond_override(const LoopState &ls, const float *cond_vals, int n_cond, int b,
                                               int d, int T, float v) {
  for (int c = 0; c < n_cond; ++c) {
    const unsigned tt = (unsigned)(d - ls.cond_h[c] * T);
    if (tt < (unsigned)T)
      v = cond_vals[((size_t)c * (ls.cond_per_batch ? ls.cond_B : 1) +
                     (ls.cond_per_batch ? (ls.cond_row0 + b) : 0)) * T + tt];
  }
  return v;
}

// Variant A: no projector in this kernel (guided / plain policies, or a projector GEMM follows).
__global__ void __launch_bounds__(256) step_pointwise_kernel(const StepParams p) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  const LoopState ls = *p.ls;
  const int i = ls.step;
  const float cr = p.sqrt_recip[i], crm1 = p.sqrt_recipm1[i], c1 = p.coef1[i], c2 = p.coef2[i];
  const float lv = p.logvar[i];
  const float sig = (i != 0) ? expf(0.5f * lv) : 0.f;
  const float gvar = ls.grad ? ls.guide_w * expf(lv) : 0.f;
  // when a projector GEMM follows, inpaint here only in the inpaint -> project order
  const int n_cond = ((ls.flags & 1u) && (!p.to_tmp || (ls.flags & 4u))) ? ls.n_cond : 0;
  float *dst = p.to_tmp ? p.xtmp : ls.x;
  float *tr = (!p.to_tmp && ls.trace) ? ls.trace + (size_t)(ls.n_steps - 1 - i) * ls.trace_stride : nullptr;
  const size_t total4 = (size_t)p.B * p.D / 4;
  // (sample, offset) of this thread's float4 advance by a fixed stride: one division up front, none in the loop
  const unsigned D4 = (unsigned)p.D / 4u;
  const size_t stride4 = (size_t)gridDim.x * blockDim.x;
  const unsigned sb = (unsigned)(stride4 / D4), sd = (unsigned)(stride4 - (size_t)sb * D4);
  size_t q4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned b_u = (unsigned)(q4 / D4), d4 = (unsigned)(q4 - (size_t)b_u * D4);
  for (; q4 < total4; q4 += stride4, b_u += sb, d4 += sd) {
    if (d4 >= D4) { d4 -= D4; b_u += 1; }
    const size_t e = q4 * 4;
    const int b = (int)b_u;
    const int d0 = (int)(d4 * 4u);
    float o[4];
    step_pointwise4(p, ls, b, d0, cr, crm1, c1, c2, sig, gvar, o);
    if (n_cond) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = cond_override(ls, p.cond_vals, n_cond, b, d0 + j, p.T, o[j]);
    }
    const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4 *>(dst + e) = o4;
    if (tr) *reinterpret_cast<float4 *>(tr + e) = o4;
    if (p.split) {
      // x' = hi + lo with hi, lo in bf16: three bf16 products recover fp32-level accuracy on the tensor cores
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[j] = __float2bfloat16_rn(o[j]);
        lo[j] = __float2bfloat16_rn(o[j] - __bfloat162float(hi[j]));
      }
      __nv_bfloat16 *row = p.split + (size_t)b * 3 * p.Kp + d0;
      const uint2 h2 = *reinterpret_cast<const uint2 *>(hi), l2 = *reinterpret_cast<const uint2 *>(lo);
      *reinterpret_cast<uint2 *>(row) = h2;
      *reinterpret_cast<uint2 *>(row + p.Kp) = l2;
      *reinterpret_cast<uint2 *>(row + 2 * p.Kp) = h2;
    }
  }
}

// Variant B: projector fused, for D*D*4 bytes that fit shared memory (PointMaze H=32: D=192).
// Block = 256 threads; the transposed projector Nt[k][d] is staged in shared memory once per block
// and reused for every sample group the (persistent) block processes.  Per group of SB samples:
// phase 1 writes x' (pointwise part) to shared memory with float4 accesses; phase 2 has each thread
// own one output column d for SB/2 samples and accumulate over k with conflict-free Nt reads and
// broadcast x' reads.
constexpr int STEP_SB = 16;

__global__ void __launch_bounds__(256) step_project_fused_kernel(const StepParams p) {
  extern __shared__ __align__(16) float smem[];
  const int D = p.D;
  float *Nt = smem;                 // D * D
  float *xs = Nt + (size_t)D * D;   // D * STEP_SB, layout xs[k][s]
  float *qs = xs + (size_t)D * STEP_SB;
  ptx::griddep_launch();
  // the projector itself is constant: stage it while the U-Net's last kernel drains
  for (int idx = threadIdx.x * 4; idx < D * D; idx += blockDim.x * 4)
    *reinterpret_cast<float4 *>(Nt + idx) = __ldg(reinterpret_cast<const float4 *>(p.Nt + idx));
  for (int d = threadIdx.x; d < D; d += blockDim.x) qs[d] = p.q[d];
  ptx::griddep_wait();
  const LoopState ls = *p.ls;
  const int i = ls.step;
  const float cr = p.sqrt_recip[i], crm1 = p.sqrt_recipm1[i], c1 = p.coef1[i], c2 = p.coef2[i];
  const float lv = p.logvar[i];
  const float sig = (i != 0) ? expf(0.5f * lv) : 0.f;
  const float gvar = ls.grad ? ls.guide_w * expf(lv) : 0.f;
  const float alpha = p.alpha_tab[i];
  const bool inpaint_first = (ls.flags & 4u) != 0;
  const int n_cond = (ls.flags & 1u) ? ls.n_cond : 0;
  float *tr = ls.trace ? ls.trace + (size_t)(ls.n_steps - 1 - i) * ls.trace_stride : nullptr;

  const int n_groups = (p.B + STEP_SB - 1) / STEP_SB;
  const int D4 = D / 4;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int b0 = grp * STEP_SB;
    __syncthreads();   // previous group's phase 2 is done with xs (and Nt/qs are loaded)
    // phase 1: pointwise part -> xs[k][s]
    for (int w = threadIdx.x; w < STEP_SB * D4; w += blockDim.x) {
      const int s = w / D4, d0 = (w - s * D4) * 4;
      const int b = b0 + s;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if (b < p.B) {
        step_pointwise4(p, ls, b, d0, cr, crm1, c1, c2, sig, gvar, o);
        if (inpaint_first && n_cond) {
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = cond_override(ls, p.cond_vals, n_cond, b, d0 + j, p.T, o[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) xs[(d0 + j) * STEP_SB + s] = o[j];
    }
    __syncthreads();
    // phase 2: y[s][d] = x'[s][d] + alpha * (sum_k Nt[k][d] x'[s][k] + q[d])
    for (int w = threadIdx.x; w < 2 * D; w += blockDim.x) {
      const int d = w % D, half = w / D;     // half selects samples [8*half, 8*half + 8)
      float acc[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) acc[s] = 0.f;
      const float *xk = xs + half * 8;
#pragma unroll 4
      for (int k = 0; k < D; ++k) {
        const float m = Nt[k * D + d];
        const float4 xa = *reinterpret_cast<const float4 *>(xk + k * STEP_SB);
        const float4 xb = *reinterpret_cast<const float4 *>(xk + k * STEP_SB + 4);
        acc[0] = fmaf(m, xa.x, acc[0]); acc[1] = fmaf(m, xa.y, acc[1]);
        acc[2] = fmaf(m, xa.z, acc[2]); acc[3] = fmaf(m, xa.w, acc[3]);
        acc[4] = fmaf(m, xb.x, acc[4]); acc[5] = fmaf(m, xb.y, acc[5]);
        acc[6] = fmaf(m, xb.z, acc[6]); acc[7] = fmaf(m, xb.w, acc[7]);
      }
      const float qd = qs[d];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int b = b0 + half * 8 + s;
        if (b < p.B) {
          float v = xs[d * STEP_SB + half * 8 + s] + alpha * (acc[s] + qd);
          if (!inpaint_first && n_cond) v = cond_override(ls, p.cond_vals, n_cond, b, d, p.T, v);
          ls.x[(size_t)b * D + d] = v;
          if (tr) tr[(size_t)b * D + d] = v;
        }
      }
    }
  }
}

// ---- loop-state plumbing ------------------------------------------------------------------
__global__ void set_loop_state_kernel(LoopState *dst, const LoopState v) { *dst = v; }

// Stage the current trajectories as the U-Net's first operand (fixed address -> graph-replayable):
// fp32 copy (fp32 mode) or bf16 with zero-padded channels (bf16 mode; 8 channels = one 16-byte store per thread).
// With `advance` the kernel also moves the loop to its next step index: it is the FIRST kernel of the captured
// step, does not read the index itself, and every later kernel of the step sees the new value.
__global__ void __launch_bounds__(256) stage_x_kernel(LoopState *lsp, float *out_f32, __nv_bfloat16 *out_bf16,
                                                       size_t rows, int T, int Cpad, int advance) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  const float *x = lsp->x;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (advance && idx == 0) lsp->step -= 1;
  if (out_bf16) {
    const int vec = Cpad >> 3;
    if (idx >= rows * (size_t)vec) return;
    const size_t r = idx / vec;
    const int c0 = (int)(idx - r * vec) << 3;
    uint4 o;
    __nv_bfloat162 *o2 = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + 2 * j;
      const float a = c < T ? x[r * T + c] : 0.f, b = c + 1 < T ? x[r * T + c + 1] : 0.f;
      o2[j] = __floats2bfloat162_rn(a, b);
    }
    *reinterpret_cast<uint4 *>(out_bf16 + r * Cpad + c0) = o;
  } else {
    if (idx >= rows * (size_t)T) return;
    out_f32[idx] = x[idx];
  }
}

}  // namespace dad
