// K1-K5: Conv1d-over-horizon as an implicit GEMM on the 5th-gen tensor cores (tcgen05), bf16 operands,
// fp32 accumulation in TMEM, with the whole Conv1dBlock tail fused into the epilogue.
//
//   replaces  Conv1dBlock            nn.Conv1d(k,pad k//2) -> GroupNorm(8) -> Mish   temporal_unet.py:69-73
//             + time-embedding add   out + time_mlp(t)[:, :, None]                   temporal_unet.py:117
//             + residual add         out + residual_conv(x)                          temporal_unet.py:122
//             Downsample1d / Upsample1d (two 2-tap phases) / 1x1 residual + head     temporal_unet.py:40,51,103,196
//
// GEMM view: rows m = (sample, position), M-tile = 128 rows = 128/L WHOLE samples, so GroupNorm
// statistics never leave the CTA and the conv halo never crosses a tile; columns n = output
// channels, N-tile a multiple of the GroupNorm group width; K = taps x input channels.
//
// A operand: a 4-D TMA tensor map (channel, stride-phase, position, sample) over the channels-last
// bf16 activation.  For tap t the box is fetched at position offset tap_j[t]; positions outside
// [0, L) are zero-filled by the TMA unit, which IS the convolution's zero padding.  The box lands in
// shared memory as a 128 x 64 K-major tile with the 128-byte swizzle tcgen05 expects.
// B operand: pre-packed bf16 weights Wb[n][tap * Cin + c] (K-major), 2-D TMA, same swizzle.
//
// Weight-stationary mode (p.ws, chosen by the host when K x BN weights + the activation ring fit shared memory: the
// Downsample / ConvTranspose convs of the narrow levels): a persistent CTA always works on the same N tile (the grid is a
// multiple of n_tiles_n), so its weight boxes are fetched ONCE -- before griddepcontrol.wait, they are constants -- and
// only activation boxes stream through the ring.  Without it every M tile re-fetched the whole weight tile: L2->SM
// traffic was 3-6x the DRAM traffic (profiles/r01_ncu_full_pointmaze.md) and bounded these kernels.
//
// Warp roles (320 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread MMA issuer, warps 2-9 = two epilogue warpgroups (TMEM -> registers -> global) that take
// alternate tiles.  Up to four TMEM accumulator stages let the MMAs run ahead of the epilogues.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace dad {

constexpr int TC_BM = 128;       // UMMA M (cta_group::1)
constexpr int TC_BK = 64;        // bf16 elements per k-block = one 128 B swizzle row
constexpr int TC_UMMA_K = 16;
#ifndef DAD_TC_EPI_WG
#define DAD_TC_EPI_WG 2
#endif
#ifndef DAD_TC_CW
#define DAD_TC_CW 16
#endif
constexpr int TC_EPI_WG = DAD_TC_EPI_WG;     // epilogue warpgroups (4 warps each), alternating tiles
constexpr int TC_THREADS = 64 + 128 * TC_EPI_WG;

struct ConvTcParams {
  const float *bias;             // [Cout_pad]
  const float *gamma, *beta;     // GroupNorm affine (GN variants)
  const float *ttab;             // [n_timesteps][Cout] time-bias table, or nullptr
  const __nv_bfloat16 *residual; // (B, L_total, Cout) or nullptr
  void *out;                     // bf16 (B, L_total, Cout), or fp32 (B, L_total, Cout) when out_f32
  const LoopState *ls;
  int B;                         // samples in this launch
  int L_out;                     // rows per sample in the GEMM (divides 128)
  int out_mul, out_phase;        // output row = l * out_mul + out_phase
  int Cout;                      // real output channels (stores are clipped to it)
  int n_tiles_m, n_tiles_n;
  int kch1, kch2;                // 64-channel chunks of source 1 / source 2 (channel concat)
  int taps;
  int tap_j[kMaxTaps], tap_p[kMaxTaps];
  int out_f32;
  // projector epilogue (K8): out = x' + alpha[step] * (acc + bias), then inpainting / trace, written to ls->x
  const float *proj_x;           // x' fp32 (B, D), or nullptr for a plain convolution
  const float *alpha_tab;
  const float *cond_vals;
  int T;
  int ws;                        // 1: weight-stationary (the CTA's whole weight tile stays in shared memory, see below)
  unsigned long long *prof;      // optional cycle counters (DAD_TC_PROF): [wait, pass1, pass2, tiles] summed over epilogue warps
  int debug;                     // 0 normal; 1 = skip the epilogue arithmetic; 2 = skip TMA + MMA (profiling only, DAD_TC_DEBUG)
};

template <int BN>
struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;                 // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int B_ALLOC = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = A_BYTES + B_ALLOC;
  static constexpr int STAGES = (BN >= 256) ? 3 : 5;
  static constexpr int ACC_STAGES = (BN >= 256) ? 2 : 4;
  static constexpr int TMEM_COLS = (ACC_STAGES * BN <= 32) ? 32 : (ACC_STAGES * BN <= 64) ? 64
                                   : (ACC_STAGES * BN <= 128) ? 128 : (ACC_STAGES * BN <= 256) ? 256 : 512;
  static constexpr int BAR_BYTES = 256;
  // + per-column epilogue parameters (bias, gamma, beta, time bias) for every output channel: 16 B per channel
  static constexpr int smem_bytes(int cout_pad) { return STAGES * STAGE_BYTES + 1024 /*align slack*/ + BAR_BYTES + 16 * cout_pad; }
  // Weight-stationary layout: the ring holds activation boxes only, the num_kb weight boxes of the CTA's N tile follow it.
  __host__ __device__ static constexpr int ring_bytes(bool ws, int num_kb) { return ws ? STAGES * A_BYTES + num_kb * B_ALLOC : STAGES * STAGE_BYTES; }
  static constexpr int smem_bytes_ws(int cout_pad, int num_kb) {
    return ring_bytes(true, num_kb) + 1024 + BAR_BYTES + 16 * cout_pad;
  }
};

// tanh(softplus(y)) = (w - 1) / (w + 1) with w = (1 + e^y)^2, i.e. 1 - 2 / (w + 1): one ex2, one rcp, no clamp
// (e^y = inf gives rcp(inf) = 0 -> 1).  Absolute error < 1e-6; the output is rounded to bf16 afterwards.
__device__ __forceinline__ float mish_tc(float y) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y * 1.4426950408889634f));
  const float u = 1.f + e;
  const float w1 = fmaf(u, u, 1.f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w1));
  return y * fmaf(-2.f, r, 1.f);
}

// GW = GroupNorm group width in columns (0: no GroupNorm/Mish, plain bias epilogue).
// TF32 = true: the fp32 SIBLING of a bf16 model (dad_set_fp32_steps / dad_set_fp32_math 1) -- activations and weights are
// fp32 words in global and shared memory (a 128-byte swizzle row holds 32 of them: k blocks of 32 channels), the MMAs are
// kind::tf32 (K = 8 per instruction, the same 32-byte descriptor advance), the residual is read and the output written
// as fp32, rounded to TF32 with round-to-nearest (the tensor core would otherwise truncate the next layer's operands).
template <int BN, int GW, bool TF32 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const ConvTcParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int BKE = TF32 ? 32 : TC_BK;         // elements per k block = one 128 B swizzle row
  constexpr int CW = (BN < 32) ? 16 : DAD_TC_CW;       // columns per TMEM load
  constexpr int NCHUNK = BN / CW;
  constexpr int NG = (GW > 0) ? BN / GW : 1;    // groups per N-tile

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kch = p.kch1 + p.kch2;
  const int num_kb = p.taps * kch;
  const bool ws = p.ws != 0;
  const int ring_bytes = Cfg::ring_bytes(ws, num_kb);
  const int a_stride = ws ? Cfg::A_BYTES : Cfg::STAGE_BYTES;     // distance between activation slots of the ring
  uint8_t *w_res = smem + Cfg::STAGES * Cfg::A_BYTES;            // resident weight boxes (ws only)
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + ring_bytes);
  uint64_t *empty_bar = full_bar + Cfg::STAGES;
  uint64_t *tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t *tempty_bar = tfull_bar + Cfg::ACC_STAGES;
  uint64_t *wfull_bar = tempty_bar + Cfg::ACC_STAGES;            // resident weights have landed (ws only)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wfull_bar + 1);
  const int cout_pad = p.n_tiles_n * BN;
  // per-column epilogue parameters, interleaved: s_par[col] = {gamma, beta, bias, time bias} (one LDS.128 per column)
  const uint32_t s_par = ptx::smem_u32(smem + ring_bytes + Cfg::BAR_BYTES);
  __shared__ float gn_scratch[TC_EPI_WG][4][2 * 8];   // [warpgroup][epilogue warp][sum, sumsq per group]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles_m * p.n_tiles_n;
  const int spt = TC_BM / p.L_out;              // samples per M-tile

  ptx::griddep_launch();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA1);
    ptx::prefetch_tmap(&tmA2);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < Cfg::ACC_STAGES; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], 4);        // one arrive per warp of the warpgroup that drained it
    }
    ptx::mbar_init(wfull_bar, 1);
    ptx::fence_barrier_init();
    if (ws && (int)blockIdx.x < total_tiles) {
      // the weights are constants: fetch this CTA's whole N tile now, while the previous kernel is still running
      const int n0 = ((int)blockIdx.x % p.n_tiles_n) * BN;
      ptx::mbar_arrive_expect_tx(wfull_bar, (uint32_t)num_kb * Cfg::B_BYTES);
      for (int kb = 0; kb < num_kb; ++kb) ptx::tma_load_2d(w_res + kb * Cfg::B_ALLOC, &tmW, wfull_bar, kb * BKE, n0);
    }
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::griddep_wait();                            // no global data was touched above
  if (warp >= 2) {
    // per-column epilogue parameters -> shared memory, once per CTA (all tiles of this CTA reuse them)
    const int step = p.ls->step;
    const float *tt = (p.ttab && !p.ls->t_rows) ? p.ttab + (size_t)step * p.Cout : nullptr;
    for (int n = threadIdx.x - 64; n < cout_pad; n += TC_THREADS - 64) {
      const bool in = n < p.Cout;
      float g = 0.f, e = 0.f;
      if constexpr (GW > 0) {
        if (in) { g = p.gamma[n]; e = p.beta[n]; }
      }
      ptx::sts32(s_par + n * 16 + 0, g);
      ptx::sts32(s_par + n * 16 + 4, e);
      ptx::sts32(s_par + n * 16 + 8, p.bias[n]);   // allocated and zero-filled up to Cout_pad
      ptx::sts32(s_par + n * 16 + 12, (tt && in) ? tt[n] : 0.f);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0 && !(DAD_DEBUG_BITS(p) & 2)) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tm = tile / p.n_tiles_n, tn = tile - tm * p.n_tiles_n;
        const int b0 = tm * spt, n0 = tn * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / kch, ch = kb - tap * kch;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t *sa = smem + stage * a_stride;
          uint8_t *sb = sa + Cfg::A_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], ws ? Cfg::A_BYTES : Cfg::A_BYTES + Cfg::B_BYTES);
          if (ch < p.kch1) ptx::tma_load_4d(sa, &tmA1, &full_bar[stage], ch * BKE, p.tap_p[tap], p.tap_j[tap], b0);
          else ptx::tma_load_4d(sa, &tmA2, &full_bar[stage], (ch - p.kch1) * BKE, p.tap_p[tap], p.tap_j[tap], b0);
          if (!ws) ptx::tma_load_2d(sb, &tmW, &full_bar[stage], kb * BKE, n0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc = TF32 ? ptx::make_idesc_tf32(TC_BM, BN) : ptx::make_idesc_bf16(TC_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (ws && (int)blockIdx.x < total_tiles) ptx::mbar_wait(wfull_bar, 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it % Cfg::ACC_STAGES;
        const uint32_t aphase = (it / Cfg::ACC_STAGES) & 1;
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);   // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < ((DAD_DEBUG_BITS(p) & 2) ? 0 : num_kb); ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);      // TMA bytes have landed
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * a_stride);
          const uint64_t da = ptx::make_smem_desc_sw128(sa);
          const uint64_t db = ptx::make_smem_desc_sw128(ws ? ptx::smem_u32(w_res + kb * Cfg::B_ALLOC) : sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            // advance 16 bf16 (8 tf32) = 32 B along K inside the swizzle row: +2 in 16 B units
            if constexpr (TF32) ptx::umma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            else ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&empty_bar[stage]);          // frees the smem slot when the MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[as]);               // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue ====================================
    // TC_EPI_WG warpgroups take alternate tiles so that the TMEM / shared / global latencies of one hide
    // behind the arithmetic of the others; each warp owns the TMEM lane quarter (warp % 4).  The column
    // loops are deliberately NOT unrolled across chunks: the body is ~1K instructions and must stay
    // resident in the instruction cache while 8-16 warps run it at different phases.
    const int wg = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                  // row of the tile
    const int s_in_tile = r / p.L_out;
    const int l = r - s_in_tile * p.L_out;
    const int L_total = p.L_out * p.out_mul;
    const bool rows_t = (p.ttab != nullptr) && (p.ls->t_rows != nullptr);   // stand-alone forward with per-row timesteps
    const long long *t_rows = rows_t ? p.ls->t_rows : nullptr;
    const int lanes = p.L_out < 32 ? p.L_out : 32;
    const float inv_n = 1.0f / (float)(p.L_out * (GW > 0 ? GW : 1));
    constexpr int GPC = (GW > 0 && GW < CW) ? CW / GW : 1;     // GroupNorm groups per column chunk
    constexpr int CPG = (GW >= CW) ? GW / CW : 1;              // column chunks per GroupNorm group
    int it = wg;
    long long pc_wait = 0, pc_p1 = 0, pc_p2 = 0, pc_n = 0, pc_t0 = 0, pc_t1 = 0;
    // projector epilogue (K8): the loop state is read once per thread, not per element
    float pj_alpha = 0.f;
    float *pj_x = nullptr, *pj_trace = nullptr;
    int pj_n_cond = 0, pj_cond_mul = 1, pj_cond_row = -1;
    if (p.proj_x) {
      const LoopState *ls = p.ls;
      const int step = ls->step;
      pj_alpha = p.alpha_tab[step];
      pj_x = ls->x;
      float *trc = ls->trace;
      pj_trace = trc ? trc + (size_t)(ls->n_steps - 1 - step) * ls->trace_stride : nullptr;
      const unsigned fl = ls->flags;
      pj_n_cond = ((fl & 1u) && !(fl & 4u)) ? ls->n_cond : 0;
      const int per_batch = ls->cond_per_batch;
      pj_cond_mul = per_batch ? ls->cond_B : 1;
      pj_cond_row = per_batch ? ls->cond_row0 : -1;
    }
    for (int tile = blockIdx.x + wg * gridDim.x; tile < total_tiles; tile += TC_EPI_WG * gridDim.x, it += TC_EPI_WG) {
      if (DAD_PROF_PTR(p)) pc_t0 = clock64();
      const int tm = tile / p.n_tiles_n, tn = tile - tm * p.n_tiles_n;
      const int n0 = tn * BN;
      const int b = tm * spt + s_in_tile;
      const bool valid = b < p.B;
      const int as = it % Cfg::ACC_STAGES;
      const uint32_t aphase = (it / Cfg::ACC_STAGES) & 1;
      const uint32_t t_addr = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
      const size_t orow = ((size_t)b * L_total + (size_t)l * p.out_mul + p.out_phase) * p.Cout + n0;
      // residual of the first column chunk: requested before the accumulator is even ready
      const bool has_res = !TF32 && (p.residual != nullptr) && valid && !p.out_f32 && !(DAD_DEBUG_BITS(p) & 8);
      uint4 rcur[CW / 8];
      if (has_res) {
        const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + orow);
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) rcur[j] = __ldg(rp + j);
      }
      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tc_fence_after();
      if (DAD_PROF_PTR(p)) { pc_t1 = clock64(); pc_wait += pc_t1 - pc_t0; pc_t0 = pc_t1; }
      if (DAD_DEBUG_BITS(p) == 1) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
        continue;
      }

      float mean_a[NG], rstd_a[NG];               // dynamically indexed -> thread-local memory (L1 resident)
      if (DAD_DEBUG_BITS(p) & 16) {
#pragma unroll
        for (int g = 0; g < NG; ++g) { mean_a[g] = 0.f; rstd_a[g] = 1.f; }
      } else if constexpr (GW > 0) {
        // ---- pass 1: GroupNorm statistics of (conv + bias) over (L rows) x (GW columns)
        float run1 = 0.f, run2 = 0.f;             // carried across the chunks of a wide group
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          uint32_t v[32];
          if constexpr (CW == 32) ptx::tmem_ld32(t_addr + c * CW, v); else ptx::tmem_ld16(t_addr + c * CW, v);
          float s1[GPC], s2[GPC];
#pragma unroll
          for (int g = 0; g < GPC; ++g) { s1[g] = 0.f; s2[g] = 0.f; }
          // parameter loads are issued BEFORE waiting for the TMEM load so that the two latencies overlap
          const uint32_t sp = s_par + (uint32_t)(n0 + c * CW) * 16u;
          float bb[CW];
#pragma unroll
          for (int j = 0; j < CW; ++j) bb[j] = ptx::lds128(sp + j * 16).z;
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const float x = __uint_as_float(v[j]) + bb[j];
            const int g = (GW < CW) ? j / GW : 0;   // compile-time
            s1[g] += x;
            s2[g] = fmaf(x, x, s2[g]);
          }
          if constexpr (GW >= CW) {
            run1 += s1[0];
            run2 += s2[0];
            if ((c + 1) % CPG != 0) continue;
            s1[0] = run1; s2[0] = run2;
            run1 = 0.f; run2 = 0.f;
          }
          // reduce across the L rows of this sample (contiguous, aligned lanes)
#pragma unroll
          for (int g = 0; g < GPC; ++g) {
            for (int o = lanes >> 1; o > 0; o >>= 1) {
              s1[g] += __shfl_xor_sync(0xffffffffu, s1[g], o);
              s2[g] += __shfl_xor_sync(0xffffffffu, s2[g], o);
            }
          }
          if (p.L_out > 32) {
            // a sample spans several epilogue warps: combine through shared memory (per-warpgroup barrier)
            float *scr = &gn_scratch[wg][0][0];
            ptx::named_bar_sync(1 + wg, 128);           // previous readers are done with scr
            if (lane == 0) {
#pragma unroll
              for (int g = 0; g < GPC; ++g) { scr[q * 16 + g] = s1[g]; scr[q * 16 + 8 + g] = s2[g]; }
            }
            ptx::named_bar_sync(1 + wg, 128);
            const int wps = p.L_out / 32;
            const int w0 = (q / wps) * wps;
#pragma unroll
            for (int g = 0; g < GPC; ++g) {
              float a = 0.f, c2 = 0.f;
              for (int w = 0; w < wps; ++w) { a += scr[(w0 + w) * 16 + g]; c2 += scr[(w0 + w) * 16 + 8 + g]; }
              s1[g] = a; s2[g] = c2;
            }
          }
          const int g0 = (GW >= CW) ? c / CPG : c * GPC;
#pragma unroll
          for (int g = 0; g < GPC; ++g) {
            const float m = s1[g] * inv_n;
            const float var = fmaxf(s2[g] * inv_n - m * m, 0.f);
            mean_a[g0 + g] = m;
            rstd_a[g0 + g] = rsqrtf(var + kGnEps);
          }
        }
      }

      if (DAD_PROF_PTR(p)) { pc_t1 = clock64(); pc_p1 += pc_t1 - pc_t0; pc_t0 = pc_t1; }
      // ---- pass 2: normalise, Mish, (+ time bias | + residual), convert, store
      const float *trow = nullptr;
      if (rows_t) {
        // clamped: an out-of-range timestep must not read outside the table (the Python wrapper rejects it with an error)
        const long long tr = valid ? t_rows[b] : 0;
        trow = p.ttab + (size_t)min(max(tr, 0ll), (long long)p.ls->n_table - 1) * p.Cout + n0;
      }
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        uint32_t v[32];
        if constexpr (CW == 32) ptx::tmem_ld32(t_addr + c * CW, v); else ptx::tmem_ld16(t_addr + c * CW, v);
        // residual of the NEXT chunk is requested while this one is processed
        uint4 rnext[CW / 8];
        if (has_res && c + 1 < NCHUNK) {
          const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + orow + (c + 1) * CW);
#pragma unroll
          for (int j = 0; j < CW / 8; ++j) rnext[j] = __ldg(rp + j);
        }
        float mg[GPC], rg[GPC];
        if constexpr (GW > 0) {
          const int g0 = (GW >= CW) ? c / CPG : c * GPC;
#pragma unroll
          for (int g = 0; g < GPC; ++g) { mg[g] = mean_a[g0 + g]; rg[g] = rstd_a[g0 + g]; }
        }
        const int nc = n0 + c * CW;
        const uint32_t sp = s_par + (uint32_t)nc * 16u;
        float y[CW];
        if constexpr (GW > 0) {
          // fold (bias, mean, rstd, gamma, beta) into one multiply-add per element while the TMEM load flies
          float a[CW], bsh[CW], tt[CW];
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const float4 pr = ptx::lds128(sp + j * 16);       // {gamma, beta, bias, time bias}
            const int g = (GW < CW) ? j / GW : 0;
            a[j] = rg[g] * pr.x;
            bsh[j] = fmaf(pr.z - mg[g], a[j], pr.y);
            tt[j] = pr.w;
          }
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const float xn = fmaf(__uint_as_float(v[j]), a[j], bsh[j]);
            y[j] = ((DAD_DEBUG_BITS(p) & 32) ? xn : mish_tc(xn)) + tt[j];
          }
          if (trow) {
#pragma unroll
            for (int j = 0; j < CW; j += 4) {
              const float4 r4 = __ldg(reinterpret_cast<const float4 *>(trow + c * CW + j));
              y[j] += r4.x; y[j + 1] += r4.y; y[j + 2] += r4.z; y[j + 3] += r4.w;
            }
          }
        } else {
          float bb[CW];
#pragma unroll
          for (int j = 0; j < CW; ++j) bb[j] = ptx::lds128(sp + j * 16).z;
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CW; ++j) y[j] = __uint_as_float(v[j]) + bb[j];
        }
        if (valid && (!(DAD_DEBUG_BITS(p) & 4) || y[0] == 123.456f)) {
          if (p.proj_x) {
            // dynamics projector: y = x' + alpha (N x' + q), Diffuser-style inpainting, optional trace (policies.py:409-485,61-62)
            // The chunk's x' values are fetched with 128-bit loads, inpainting is decided per (chunk, condition) from the
            // column range [h_c T, h_c T + T) of the condition (no per-element division), results leave as 128-bit stores.
            const float *xr = p.proj_x + orow + c * CW;
            float *o = pj_x + orow + c * CW;
            float *tr = pj_trace ? pj_trace + orow + c * CW : nullptr;
            const bool full = nc + CW <= p.Cout;      // Cout = D is a multiple of 4 (the step kernels' float4 contract)
            float xv[CW];
            if (full) {
#pragma unroll
              for (int j = 0; j < CW; j += 4) {
                const float4 x4 = __ldg(reinterpret_cast<const float4 *>(xr + j));
                xv[j] = x4.x; xv[j + 1] = x4.y; xv[j + 2] = x4.z; xv[j + 3] = x4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < CW; ++j) xv[j] = (nc + j < p.Cout) ? xr[j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) y[j] = fmaf(pj_alpha, y[j], xv[j]);
            for (int cc = 0; cc < pj_n_cond; ++cc) {
              const int base = nc - __ldg(&p.ls->cond_h[cc]) * p.T;      // column nc + j belongs to the condition iff 0 <= base + j < T
              if (base + CW <= 0 || base >= p.T) continue;
              const float *row = p.cond_vals + ((size_t)cc * pj_cond_mul + (pj_cond_row >= 0 ? pj_cond_row + b : 0)) * p.T;
#pragma unroll
              for (int j = 0; j < CW; ++j)
                if ((unsigned)(base + j) < (unsigned)p.T) y[j] = row[base + j];
            }
            if (full) {
#pragma unroll
              for (int j = 0; j < CW; j += 4) {
                const float4 o4 = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                *reinterpret_cast<float4 *>(o + j) = o4;
                if (tr) *reinterpret_cast<float4 *>(tr + j) = o4;
              }
            } else {
#pragma unroll
              for (int j = 0; j < CW; ++j)
                if (nc + j < p.Cout) {
                  o[j] = y[j];
                  if (tr) tr[j] = y[j];
                }
            }
          } else if constexpr (TF32) {
            // fp32 activations (channel counts are multiples of 64 here): + fp32 residual, round to TF32, 128-bit stores
            float *o = reinterpret_cast<float *>(p.out) + orow + c * CW;
            const float *rs = p.residual ? reinterpret_cast<const float *>(p.residual) + orow + c * CW : nullptr;
#pragma unroll
            for (int j = 0; j < CW; j += 4) {
              float4 v = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
              if (rs) {
                const float4 r4 = __ldg(reinterpret_cast<const float4 *>(rs + j));
                v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
              }
              uint32_t t0, t1, t2, t3;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t0) : "f"(v.x));
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t1) : "f"(v.y));
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t2) : "f"(v.z));
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t3) : "f"(v.w));
              *reinterpret_cast<uint4 *>(o + j) = make_uint4(t0, t1, t2, t3);
            }
          } else if (p.out_f32) {
            float *o = reinterpret_cast<float *>(p.out) + orow + c * CW;
#pragma unroll
            for (int j = 0; j < CW; ++j)
              if (nc + j < p.Cout) o[j] = y[j];
          } else {
            if (has_res) {
#pragma unroll
              for (int j = 0; j < CW; j += 8) {
                const __nv_bfloat162 *r2 = reinterpret_cast<const __nv_bfloat162 *>(&rcur[j / 8]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  const float2 f = __bfloat1622float2(r2[jj]);
                  y[j + 2 * jj] += f.x;
                  y[j + 2 * jj + 1] += f.y;
                }
              }
            }
            uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + orow + c * CW);
#pragma unroll
            for (int j = 0; j < CW; j += 8) {
              uint4 ov;
              __nv_bfloat162 *o2 = reinterpret_cast<__nv_bfloat162 *>(&ov);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) o2[jj] = __floats2bfloat162_rn(y[j + 2 * jj], y[j + 2 * jj + 1]);
              op[j / 8] = ov;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) rcur[j] = rnext[j];
      }
      // release the accumulator stage back to the MMA issuer
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
      if (DAD_PROF_PTR(p)) { pc_p2 += clock64() - pc_t0; pc_n += 1; }
    }
    if (DAD_PROF_PTR(p) && lane == 0) {
      atomicAdd(DAD_PROF_PTR(p) + 0, (unsigned long long)pc_wait);
      atomicAdd(DAD_PROF_PTR(p) + 1, (unsigned long long)pc_p1);
      atomicAdd(DAD_PROF_PTR(p) + 2, (unsigned long long)pc_p2);
      atomicAdd(DAD_PROF_PTR(p) + 3, (unsigned long long)pc_n);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dad
