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
#include "f32x2.cuh"

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
  // u in (0, 1): (r + 0.5) 2^-32 as one FMA per draw; lg2 through the raw SFU op (u is never denormal)
  const float k = 2.3283064365386963e-10f, h = 1.1641532182693481e-10f;   // 2^-32, 2^-33
  const float u0 = fmaf((float)r.x, k, h), u2 = fmaf((float)r.z, k, h);
  const float a1 = fmaf((float)r.y, 6.283185307179586f * k, 6.283185307179586f * h);
  const float a3 = fmaf((float)r.w, 6.283185307179586f * k, 6.283185307179586f * h);
  float l0, l2, r0, r1;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(u0));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u2));
  // sqrt(-2 ln u) = sqrt(-2 ln2 * lg2 u)
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(-1.3862943611198906f * l0));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(-1.3862943611198906f * l2));
  float s0, c0, s1, c1;
  __sincosf(a1, &s0, &c0);
  __sincosf(a3, &s1, &c1);
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

// The scalar part of the loop state a step kernel needs, read once into registers (the condition table stays
// in global memory: indexing a by-value copy of LoopState would put the whole struct in local memory).
struct StepView {
  float *x;
  const float *noise;           // this step's slot of the injected noise, or nullptr -> Philox
  const float *grad;
  float *trace;                 // this step's slot of the trace, or nullptr
  unsigned long long seed, sample_offset;
  unsigned flags;
  int step, n_cond, cond_mul, cond_row;
  float cr, crm1, c1, c2, sig, gvar;
};

// Values every thread loads from the same address: the shuffle tells the compiler they are warp-uniform, so the
// Philox key schedule and the step coefficients live in uniform registers instead of being recomputed per thread.
__device__ __forceinline__ unsigned uniform32(unsigned v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ unsigned long long uniform64(unsigned long long v) {
  return ((unsigned long long)uniform32((unsigned)(v >> 32)) << 32) | uniform32((unsigned)v);
}

__device__ __forceinline__ StepView step_view(const StepParams &p) {
  const LoopState *ls = p.ls;
  StepView v;
  const int i = (int)uniform32((unsigned)ls->step);
  v.step = i;
  v.x = ls->x;
  const float *nz = ls->noise;
  v.noise = nz ? nz + (size_t)(ls->n_steps - 1 - i) * ls->noise_stride : nullptr;
  v.grad = ls->grad;
  float *tr = ls->trace;
  v.trace = tr ? tr + (size_t)(ls->n_steps - 1 - i) * ls->trace_stride : nullptr;
  v.seed = uniform64(ls->seed);
  v.sample_offset = uniform64(ls->sample_offset);
  v.flags = ls->flags;
  v.n_cond = (v.flags & 1u) ? ls->n_cond : 0;
  const int per_batch = ls->cond_per_batch;
  v.cond_mul = per_batch ? ls->cond_B : 1;
  v.cond_row = per_batch ? ls->cond_row0 : -1;      // -1: one value row shared by the whole batch
  v.cr = p.sqrt_recip[i];
  v.crm1 = p.sqrt_recipm1[i];
  v.c1 = p.coef1[i];
  v.c2 = p.coef2[i];
  const float lv = p.logvar[i];
  v.sig = (i != 0) ? expf(0.5f * lv) : 0.f;
  v.gvar = v.grad ? ls->guide_w * expf(lv) : 0.f;
  return v;
}

// The loads of one float4 of the pointwise part, issued together so that several quads can be in flight per thread.
struct StepQuad {
  float4 x, m, z, g;
};
// LEAN: the common loop configuration fixed at compile time -- epsilon prediction, clipping, in-kernel Philox noise, no
// guidance gradient (see step_pointwise_kernel<true>); the generic code is predicated on all of these per element.
template <bool LEAN = false>
__device__ __forceinline__ void step_load4(const StepParams &p, const StepView &v, size_t e, StepQuad &q) {
  q.x = *reinterpret_cast<const float4 *>(v.x + e);
  q.m = __ldg(reinterpret_cast<const float4 *>(p.model_out + e));
  if constexpr (!LEAN) {
    if (v.sig != 0.f && v.noise) q.z = __ldg(reinterpret_cast<const float4 *>(v.noise + e));
    if (v.gvar != 0.f) q.g = __ldg(reinterpret_cast<const float4 *>(v.grad + e));
  }
}

// The pointwise part for 4 consecutive elements of sample b starting at d0 (loads already issued).
template <bool LEAN = false>
__device__ __forceinline__ void step_math4(const StepParams &p, const StepView &v, int b, int d0, const StepQuad &q,
                                           float out[4]) {
  const float xv[4] = {q.x.x, q.x.y, q.x.z, q.x.w}, mv[4] = {q.m.x, q.m.y, q.m.z, q.m.w};
  float zv[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
  if (v.sig != 0.f) {
    const float4 z4 = (!LEAN && v.noise) ? q.z : philox_normal4((unsigned)(d0 >> 2), (unsigned)v.step, v.sample_offset + b, v.seed);
    zv[0] = z4.x; zv[1] = z4.y; zv[2] = z4.z; zv[3] = z4.w;
  }
  if (!LEAN && v.gvar != 0.f) { gv[0] = q.g.x; gv[1] = q.g.y; gv[2] = q.g.z; gv[3] = q.g.w; }
  if (LEAN || (p.predict_epsilon && p.clip_denoised && v.gvar == 0.f)) {
    // the usual configuration (epsilon prediction, clipping, no guidance gradient) on packed fp32x2; the same
    // operations in the same order as the general path below, so the results are bit-identical
    const f32x2 cr2 = pk2(v.cr, v.cr), ncrm1 = pk2(-v.crm1, -v.crm1), c12 = pk2(v.c1, v.c1), c22 = pk2(v.c2, v.c2);
    const f32x2 sg2 = pk2(v.sig, v.sig);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const f32x2 x2 = pk2(xv[2 * h], xv[2 * h + 1]);
      float a, b2;
      upk2(ffma2(cr2, x2, fmul2(ncrm1, pk2(mv[2 * h], mv[2 * h + 1]))), a, b2);
      a = fminf(fmaxf(a, -1.f), 1.f);
      b2 = fminf(fmaxf(b2, -1.f), 1.f);
      const f32x2 mu = ffma2(c22, x2, fmul2(c12, pk2(a, b2)));
      upk2(ffma2(sg2, pk2(zv[2 * h], zv[2 * h + 1]), mu), out[2 * h], out[2 * h + 1]);
    }
    return;
  }
  if constexpr (!LEAN) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x0 = p.predict_epsilon ? (v.cr * xv[j] - v.crm1 * mv[j]) : mv[j];
    if (p.clip_denoised) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float mu = v.c1 * x0 + v.c2 * xv[j];
    mu += v.gvar * gv[j];
    out[j] = mu + v.sig * zv[j];
  }
  }
}

// x[:, h] = val for every registered (h, val): element d of the flattened (H*T) sample belongs to condition c
// iff 0 <= d - h_c*T < T (no division: this runs per element inside a bandwidth-bound kernel).
template <int N>
__device__ __forceinline__ void cond_override(const StepParams &p, const StepView &v, int n_cond, int b, int d0,
                                              float (&o)[N]) {
  for (int c = 0; c < n_cond; ++c) {
    const int base = d0 - __ldg(&p.ls->cond_h[c]) * p.T;
    if (base + N <= 0 || base >= p.T) continue;
    const float *row = p.cond_vals + ((size_t)c * v.cond_mul + (v.cond_row >= 0 ? v.cond_row + b : 0)) * p.T;
#pragma unroll
    for (int j = 0; j < N; ++j)
      if ((unsigned)(base + j) < (unsigned)p.T) o[j] = row[base + j];
  }
}

// Variant A: no projector in this kernel (guided / plain policies, or a projector GEMM follows).  Grid-stride over
// float4 quads, STEP_U quads per thread per iteration with all their loads issued before the first use; the grid
// is the resident CTA count (occupancy query on the host).  A shared-memory ring fed by 1-D bulk copies was
// measured too (round 1): not faster -- with in-kernel Philox the kernel is bound by instruction issue, with
// injected noise this version already streams at the measured HBM peak.
#ifndef DAD_STEP_U
#define DAD_STEP_U 2
#endif
constexpr int STEP_U = DAD_STEP_U;

// LEAN = true: epsilon prediction + clipping, Philox noise, no guidance gradient, no trace, no projector GEMM behind it
// (to_tmp / split) -- the host selects it when the loop it enqueues is exactly that (dad_sample without noise_seq /
// trace; part of the captured graph's key).  Same arithmetic, ~1/4 fewer instructions per float4.
template <bool LEAN>
__global__ void __launch_bounds__(256) step_pointwise_kernel(const StepParams p) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  const StepView v = step_view(p);
  // when a projector GEMM follows, inpaint here only in the inpaint -> project order
  const int n_cond = (LEAN || !p.to_tmp || (v.flags & 4u)) ? v.n_cond : 0;
  float *dst = (!LEAN && p.to_tmp) ? p.xtmp : v.x;
  float *tr = (LEAN || p.to_tmp) ? nullptr : v.trace;
  // 32-bit quad indices (the host rejects B*D/4 >= 2^31); (sample, offset) of this thread's float4 advance by a
  // fixed stride: one division up front, none in the loop
  const unsigned total4 = (unsigned)((size_t)p.B * p.D / 4);
  const unsigned D4 = (unsigned)p.D / 4u;
  const unsigned stride4 = gridDim.x * blockDim.x;
  const unsigned sb = stride4 / D4, sd = stride4 - sb * D4;
  unsigned q4 = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned b_u = q4 / D4, d4 = q4 - b_u * D4;
  // quick reject for inpainting: the first two conditions' element ranges live in registers
  const unsigned span = (unsigned)p.T + 3u;
  const int lo0 = n_cond > 0 ? __ldg(&p.ls->cond_h[0]) * p.T : 0, lo1 = n_cond > 1 ? __ldg(&p.ls->cond_h[1]) * p.T : lo0;
  // LEAN: software-pipelined -- the loads of the NEXT pair of quads are issued before the Philox / Box-Muller work of
  // the current pair (~140 instructions per quad), so every warp has memory requests in flight while it computes
  // (without it the Philox variant ran at 0.64-0.70 of the HBM peak with issue slots and pipes half idle).
  StepQuad q[STEP_U], qn[STEP_U];
  if constexpr (LEAN) {
#pragma unroll
    for (int u = 0; u < STEP_U; ++u)
      if (q4 + u * stride4 < total4) step_load4<true>(p, v, (size_t)(q4 + u * stride4) * 4, q[u]);
  }
  while (q4 < total4) {
    if constexpr (LEAN) {
      const unsigned nx = q4 + STEP_U * stride4;
#pragma unroll
      for (int u = 0; u < STEP_U; ++u)
        if (nx + u * stride4 < total4) step_load4<true>(p, v, (size_t)(nx + u * stride4) * 4, qn[u]);
    } else {
#pragma unroll
      for (int u = 0; u < STEP_U; ++u)
        if (q4 + u * stride4 < total4) step_load4<false>(p, v, (size_t)(q4 + u * stride4) * 4, q[u]);
    }
#pragma unroll
    for (int u = 0; u < STEP_U; ++u) {
      if (q4 < total4) {
        const size_t e = (size_t)q4 * 4;
        const int b = (int)b_u;
        const int d0 = (int)(d4 * 4u);
        float o[4];
        step_math4<LEAN>(p, v, b, d0, q[u], o);
        if (n_cond && ((unsigned)(d0 - lo0 + 3) < span || (unsigned)(d0 - lo1 + 3) < span || n_cond > 2))
          cond_override<4>(p, v, n_cond, b, d0, o);
        const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4 *>(dst + e) = o4;
        if (tr) *reinterpret_cast<float4 *>(tr + e) = o4;
        if (!LEAN && p.split) {
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
      q4 += stride4;
      b_u += sb;
      d4 += sd;
      if (d4 >= D4) { d4 -= D4; b_u += 1; }
    }
    if constexpr (LEAN) {
#pragma unroll
      for (int u = 0; u < STEP_U; ++u) q[u] = qn[u];
    }
  }
}

// Variant B: projector fused, for D*D*4 bytes that fit shared memory (PointMaze H=32: D=192).
// Block = 4*D threads (thread = one output column d for one quarter of the CTA's samples); the transposed
// projector Nt[k][d] is staged in shared memory once per block and reused for every sample group the (persistent)
// block processes.  A group is 4*SPT samples, SPT chosen by the host so that the groups fill the SMs in whole rounds
// (B=4096 on 148 SMs: SPT=7 -> 147 groups).  Phase 1 writes x' (pointwise part) to shared memory; phase 2
// accumulates over k with conflict-free Nt reads, broadcast x' reads and packed fp32x2 FMAs.
//   xs layout: quarter q, row k -> 8 floats (SPT used) at q*9*D + k*8 + (k>>2)*4  (the 4-float pad every 4 rows
//   spreads phase 1's stores, which walk k in steps of 4 across a warp, over 8 bank groups)
constexpr int STEP_FUSED_MAX_THREADS = 896;
__host__ __device__ constexpr size_t step_fused_smem(int D) {
  return ((size_t)D * D + (size_t)36 * D + D) * sizeof(float);
}
__host__ __device__ constexpr int step_fused_threads(int D) { return (4 * D + 31) / 32 * 32; }

template <int SPT, bool LEAN = false>
__global__ void __launch_bounds__(STEP_FUSED_MAX_THREADS) step_project_fused_kernel(const StepParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int SB = 4 * SPT, NP = (SPT + 1) / 2;
  const int D = p.D;
  float *Nt = smem;                 // D * D
  float *xs = Nt + (size_t)D * D;   // 4 quarters * 9 * D
  float *qs = xs + (size_t)36 * D;
  ptx::griddep_launch();
  // the projector itself is constant: stage it while the U-Net's last kernel drains
  for (int idx = threadIdx.x * 4; idx < D * D; idx += blockDim.x * 4)
    *reinterpret_cast<float4 *>(Nt + idx) = __ldg(reinterpret_cast<const float4 *>(p.Nt + idx));
  for (int d = threadIdx.x; d < D; d += blockDim.x) qs[d] = p.q[d];
  for (int idx = threadIdx.x; idx < 36 * D; idx += blockDim.x) xs[idx] = 0.f;
  ptx::griddep_wait();
  const StepView v = step_view(p);
  const float alpha = p.alpha_tab[v.step];
  const bool inpaint_first = (v.flags & 4u) != 0;
  const int n_cond = v.n_cond;

  const int n_groups = (p.B + SB - 1) / SB;
  const int D4 = D / 4;
  const int pd = (int)threadIdx.x % D, pq = (int)threadIdx.x / D;   // phase-2 role: column pd, quarter pq (< 4 iff active)
  // inpainting after the projection: whether column pd is overridden (and by which condition) does not depend on the sample
  int pc = -1, ptt = 0;
  if (!inpaint_first)
    for (int c = 0; c < n_cond; ++c) {
      const int tt = pd - __ldg(&p.ls->cond_h[c]) * p.T;
      if ((unsigned)tt < (unsigned)p.T) { pc = c; ptt = tt; }
    }
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int b0 = grp * SB;
    __syncthreads();   // previous group's phase 2 is done with xs (and Nt/qs are loaded)
    // phase 1: pointwise part -> xs
    for (int w0 = threadIdx.x; w0 < SB * D4; w0 += 2 * blockDim.x) {
      StepQuad q[2];
      int s[2], d0[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int w = w0 + u * blockDim.x;
        s[u] = w / D4;
        d0[u] = (w - s[u] * D4) * 4;
        if (w < SB * D4 && b0 + s[u] < p.B) step_load4<LEAN>(p, v, (size_t)(b0 + s[u]) * D + d0[u], q[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int w = w0 + u * blockDim.x;
        if (w >= SB * D4) break;
        const int b = b0 + s[u];
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (b < p.B) {
          step_math4<LEAN>(p, v, b, d0[u], q[u], o);
          if (inpaint_first && n_cond) cond_override<4>(p, v, n_cond, b, d0[u], o);
        }
        const int qq = s[u] / SPT, sl = s[u] - qq * SPT;
        float *dstq = xs + qq * 9 * D + d0[u] * 9 + sl;     // row(d0 + j) = (d0 + j)*8 + d0
#pragma unroll
        for (int j = 0; j < 4; ++j) dstq[j * 8] = o[j];
      }
    }
    __syncthreads();
    // phase 2: y[s][d] = x'[s][d] + alpha * (sum_k Nt[k][d] x'[s][k] + q[d])
    if (pq < 4) {
      f32x2 acc[NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) acc[j] = pk2(0.f, 0.f);
      const float *xq = xs + pq * 9 * D;
      const float *np = Nt + pd;
#pragma unroll 2
      for (int k4 = 0; k4 < D4; ++k4) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = k4 * 4 + kk;
          const float m = np[(size_t)k * D];
          const f32x2 m2 = pk2(m, m);
          const float *xr = xq + k4 * 36 + kk * 8;
          const float4 xa = *reinterpret_cast<const float4 *>(xr);
          acc[0] = ffma2(m2, pk2(xa.x, xa.y), acc[0]);
          if (NP > 1) acc[1 % NP] = ffma2(m2, pk2(xa.z, xa.w), acc[1 % NP]);
          if (NP > 2) {
            const float4 xb = *reinterpret_cast<const float4 *>(xr + 4);
            acc[2 % NP] = ffma2(m2, pk2(xb.x, xb.y), acc[2 % NP]);
            if (NP > 3) acc[3 % NP] = ffma2(m2, pk2(xb.z, xb.w), acc[3 % NP]);
          }
        }
      }
      float r[2 * NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) upk2(acc[j], r[2 * j], r[2 * j + 1]);
      const float qd = qs[pd];
      const float *xd = xq + pd * 8 + (pd >> 2) * 4;
#pragma unroll
      for (int s = 0; s < SPT; ++s) {
        const int b = b0 + pq * SPT + s;
        if (b < p.B) {
          float y = xd[s] + alpha * (r[s] + qd);
          if (pc >= 0) y = p.cond_vals[((size_t)pc * v.cond_mul + (v.cond_row >= 0 ? v.cond_row + b : 0)) * p.T + ptt];
          v.x[(size_t)b * D + pd] = y;
          if (!LEAN && v.trace) v.trace[(size_t)b * D + pd] = y;
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
// It also zeroes the tile-completion counters of the conv chains (conv_chain.cuh): every kernel of the previous pass
// has completed when this one passes its dependency wait, and every kernel of this pass waits for this one.
__global__ void __launch_bounds__(256) stage_x_kernel(LoopState *lsp, float *out_f32, __nv_bfloat16 *out_bf16,
                                                       size_t rows, int T, int Cpad, int advance, unsigned *flags,
                                                       int n_flags) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  const float *x = lsp->x;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (advance && idx == 0) lsp->step -= 1;
  for (size_t i = idx; i < (size_t)n_flags; i += (size_t)gridDim.x * blockDim.x) flags[i] = 0u;
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
