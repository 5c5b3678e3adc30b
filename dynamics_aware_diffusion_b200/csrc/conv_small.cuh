// Small-batch (latency) variant of every U-Net convolution: GuidedPolicy.get_action plans ONE trajectory per call
// (policies.py:193-223), so a layer is a 8..32-row GEMM against up to 2.6 MB of weights and the step is a chain of
// 39 dependent launches.  The throughput kernels (conv_t3 / conv_tc) put such a layer on 2-4 CTAs and pay their
// fixed cluster / TMEM / pipeline costs per launch; here a layer is spread over Cout/16 CTAs per sample:
//   * CTA = 16 output channels of one sample (32 when a GroupNorm group is 256 wide, so that its cluster stays
//     within 8 CTAs).  Its weight slab ([16][taps*Cin] bf16, K-major rows as packed for the tensor-map path) is
//     fetched by a producer warp with 1-D bulk copies issued BEFORE griddepcontrol.wait -- weights do not depend
//     on the previous layer, so under programmatic dependent launch they stream in while that layer still runs.
//     A slab larger than shared memory (HalfCheetah / Door widths: up to 20 KB per weight row) goes through a
//     3-stage ring of K chunks instead (full / empty mbarriers).
//   * After the wait the haloed activation tile of the sample ((L_in + halo) rows x Cin) follows the same way.
//   * 8 consumer warps split K; each runs mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with ldmatrix fragments straight
//     from the padded rows (pitch = row bytes + 16 -> conflict-free), then the partial tiles are summed through
//     shared memory.  tcgen05 would add TMEM allocation and a commit round trip to a GEMM of < 1 MFLOP per CTA.
//   * Epilogue as in the throughput kernels: + bias, GroupNorm over (L x group) -- the group's CTAs form a
//     cluster and exchange (sum, sum of squares) through distributed shared memory --, Mish, + time bias,
//     + residual, bf16 (or fp32 for the head) store.
// Reference ops: Conv1dBlock (temporal_unet.py:57-76), ResidualTemporalBlock (:106-122), Downsample1d (:35-43),
// Upsample1d (:46-54), final_conv (:194-197).
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"   // mish_tc
#include "ptx.cuh"

namespace dad {

constexpr int SM_CONSUMERS = 256;   // 8 MMA / epilogue warps
constexpr int SM_WARPS = SM_CONSUMERS / 32;
constexpr int SM_THREADS = SM_CONSUMERS + 32;      // + one producer warp (bulk copies)
constexpr int SM_MAX_L = 32;        // rows per sample the kernel handles (two m16 tiles)
constexpr int SM_MAX_STAGES = 4;

struct ConvSmallParams {
  const __nv_bfloat16 *in1, *in2;   // (B, L_in, C1) / (B, L_in, C2) channels-last; in2 = nullptr when C2 == 0
  const __nv_bfloat16 *w;           // [Cout_pad][taps * (C1 + C2)]
  const __nv_bfloat16 *residual;    // (B, L_out * out_mul, Cout) or nullptr
  const float *bias, *gamma, *beta; // [Cout_pad], [Cout], [Cout]
  const float *ttab;                // [n_timesteps][Cout] or nullptr
  void *out;                        // bf16 (B, L_out * out_mul, Cout) or fp32 when out_f32
  const LoopState *ls;
  int C1, C2, Cout;
  int taps, tap_off[kMaxTaps];
  int in_stride, L_in, L_out, out_mul, out_phase;
  int gw;                           // GroupNorm width, 0 = plain convolution
  int out_f32;
  int halo, rows;                   // tile row of input row r is halo + r; rows = tile height
  int kc, n_stages;                 // weight ring: K elements per chunk (multiple of 16), stages (1 = whole K resident)
};

struct SmallLayout {
  uint32_t pitch_a, pitch_w, off_w, off_red, off_misc, total;
};
// mt: m16 tiles per sample, nch: output channels per CTA
__host__ __device__ inline SmallLayout small_layout(int Cin, int rows, int mt, int nch, int kc, int n_stages) {
  SmallLayout s;
  s.pitch_a = (uint32_t)Cin * 2u + 16u;
  s.pitch_w = (uint32_t)kc * 2u + 16u;
  s.off_w = (uint32_t)rows * s.pitch_a;
  s.off_red = s.off_w + (uint32_t)n_stages * nch * s.pitch_w;
  s.off_red = (s.off_red + 15u) & ~15u;
  s.off_misc = s.off_red + (uint32_t)SM_WARPS * mt * nch * 16u * 4u;      // per warp: mt x (16 x nch) fp32 partials
  s.total = s.off_misc + 256u;
  return s;
}

namespace ptx {
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_saddr, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst_saddr), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t saddr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float2 ld_cluster_f32x2(uint32_t cluster_saddr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(cluster_saddr));
  return v;
}
}  // namespace ptx

// grid = (Cout_pad / NCH, B); cluster = (max(1, gw / NCH), 1, 1); MT = m16 tiles per sample (1: L_out <= 16,
// 2: <= 32); NT = n8 tiles per CTA (NCH = 8 NT: 16, or 32 for 256-wide GroupNorm groups so the cluster stays <= 8).
template <int MT, int NT>
__global__ void __launch_bounds__(SM_THREADS, 1) conv_small_kernel(const ConvSmallParams p) {
  extern __shared__ __align__(128) unsigned char sm_raw[];
  constexpr int NCH = 8 * NT;
  constexpr int NQ = MT * NT * 2;                     // (m tile, n tile, row half) output slices of 32 column pairs
  constexpr int QI = (NQ + SM_WARPS - 1) / SM_WARPS;  // slices per consumer warp
  const int Cin = p.C1 + p.C2;
  const int K = p.taps * Cin;
  const SmallLayout lay = small_layout(Cin, p.rows, MT, NCH, p.kc, p.n_stages);
  const uint32_t sA = ptx::smem_u32(sm_raw), sW = sA + lay.off_w;
  float *red = reinterpret_cast<float *>(sm_raw + lay.off_red);
  uint64_t *full = reinterpret_cast<uint64_t *>(sm_raw + lay.off_misc);     // [stage] weights landed
  uint64_t *empty = full + SM_MAX_STAGES;                                    // [stage] all consumer warps done
  uint64_t *abar = empty + SM_MAX_STAGES;                                    // activations landed
  float *stat = reinterpret_cast<float *>(sm_raw + lay.off_misc + 96);       // [0..1] this CTA's (sum, sumsq); [2..17] warp partials
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_base = blockIdx.x * NCH, b = blockIdx.y;
  const int n_chunks = (K + p.kc - 1) / p.kc;

  if (tid == 0) {
    for (int s = 0; s < SM_MAX_STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], SM_WARPS);
    }
    ptx::mbar_init(abar, 1);
    ptx::fence_barrier_init();
  }
  ptx::griddep_launch();
  __syncthreads();

  if (warp == SM_WARPS) {
    // ---- producer warp.  Weights do not depend on the previous layer: the first n_stages chunks are requested
    // ahead of the dependency wait; the activations of sample b (one bulk copy per row and source) after it.
    const int ahead = min(p.n_stages, n_chunks);
    for (int j = 0; j < ahead; ++j) {
      const uint32_t bytes = (uint32_t)min(p.kc, K - j * p.kc) * 2u;
      if (lane == 0) ptx::mbar_arrive_expect_tx(&full[j], (uint32_t)NCH * bytes);
      __syncwarp();
      if (lane < NCH)
        ptx::bulk_load_1d(sW + (uint32_t)(j * NCH + lane) * lay.pitch_w, p.w + (size_t)(n_base + lane) * K + (size_t)j * p.kc,
                          bytes, &full[j]);
    }
    ptx::griddep_wait();
    {
      const uint32_t row_bytes = (uint32_t)Cin * 2u;
      if (lane == 0) ptx::mbar_arrive_expect_tx(abar, (uint32_t)p.L_in * row_bytes);
      __syncwarp();
      for (int r = lane; r < p.L_in; r += 32) {
        const uint32_t dst = sA + (uint32_t)(p.halo + r) * lay.pitch_a;
        ptx::bulk_load_1d(dst, p.in1 + ((size_t)b * p.L_in + r) * p.C1, (uint32_t)p.C1 * 2u, abar);
        if (p.C2) ptx::bulk_load_1d(dst + (uint32_t)p.C1 * 2u, p.in2 + ((size_t)b * p.L_in + r) * p.C2, (uint32_t)p.C2 * 2u, abar);
      }
    }
  } else {
    // zero the halo rows (generic stores; nothing else writes them)
    const int halo_hi = p.rows - p.halo - p.L_in;
    const int vec_per_row = (int)(lay.pitch_a / 16u);
    for (int i = tid; i < (p.halo + halo_hi) * vec_per_row; i += SM_CONSUMERS) {
      const int r = i / vec_per_row, v = i - r * vec_per_row;
      const int row = r < p.halo ? r : p.halo + p.L_in + (r - p.halo);
      *reinterpret_cast<uint4 *>(sm_raw + (size_t)row * lay.pitch_a + v * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::griddep_wait();
  }
  const int step = p.ls->step;
  __syncthreads();                 // halo rows visible to every warp
  if (warp == SM_WARPS) {
    // refill the ring behind the consumers
    for (int j = min(p.n_stages, n_chunks); j < n_chunks; ++j) {
      ptx::mbar_wait(&empty[j % p.n_stages], (uint32_t)((j / p.n_stages - 1) & 1));
      const int st = j % p.n_stages;
      const uint32_t bytes = (uint32_t)min(p.kc, K - j * p.kc) * 2u;
      if (lane == 0) ptx::mbar_arrive_expect_tx(&full[st], (uint32_t)NCH * bytes);
      __syncwarp();
      if (lane < NCH)
        ptx::bulk_load_1d(sW + (uint32_t)(st * NCH + lane) * lay.pitch_w, p.w + (size_t)(n_base + lane) * K + (size_t)j * p.kc,
                          bytes, &full[st]);
    }
    if (p.gw / NCH <= 1) return;      // the producer warp only stays for the cluster barriers
  }

  float v0[QI], v1[QI];
  bool valid[QI], valid1[QI];
  int row_o[QI], col_o[QI];
  if (warp < SM_WARPS) {
    ptx::mbar_wait(abar, 0);
    // ---- K split over the warps: k16 step s covers tap s / (Cin/16), channels 16 (s % (Cin/16)) ..
    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[mt][nt][j] = 0.f;
    const int cps = Cin >> 4, spc = p.kc >> 4;
    // ldmatrix row roles of this lane: A matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15); B matrices (n 0-7 | 8-15) x (k 0-7 | 8-15)
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kofs = (lane >> 4) * 8;
    const int b_row = (lane & 7) + (lane >> 4) * 8, b_kofs = ((lane >> 3) & 1) * 8;
    int l_of[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) l_of[mt] = min(mt * 16 + a_row, p.L_out - 1);      // padding rows re-read a valid row
    for (int j = 0; j < n_chunks; ++j) {
      const int st = j % p.n_stages;
      ptx::mbar_wait(&full[st], (uint32_t)((j / p.n_stages) & 1));
      const uint32_t w_lane = sW + (uint32_t)(st * NCH + b_row) * lay.pitch_w + (uint32_t)b_kofs * 2u;
      const int steps = min(p.kc, K - j * p.kc) >> 4;
      for (int sl = warp; sl < steps; sl += SM_WARPS) {
        const int s = j * spc + sl;
        const int t = s / cps, c = (s - t * cps) << 4;
        uint32_t bf[NT / 2][4];
#pragma unroll
        for (int h = 0; h < NT / 2; ++h)
          ptx::ldmatrix_x4(w_lane + (uint32_t)h * 16u * lay.pitch_w + (uint32_t)sl * 32u, bf[h][0], bf[h][1], bf[h][2], bf[h][3]);
        const int off = p.tap_off[t] + p.halo;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t af[4];
          ptx::ldmatrix_x4(sA + (uint32_t)(l_of[mt] * p.in_stride + off) * lay.pitch_a + (uint32_t)(c + a_kofs) * 2u, af[0], af[1],
                           af[2], af[3]);
#pragma unroll
          for (int h = 0; h < NT / 2; ++h) {
            ptx::mma_bf16_16816(acc[mt][2 * h], af, bf[h][0], bf[h][1]);
            ptx::mma_bf16_16816(acc[mt][2 * h + 1], af, bf[h][2], bf[h][3]);
          }
        }
      }
      if (p.n_stages < n_chunks) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[st]);
      }
    }
    // ---- sum the warps' partial tiles: red[warp][q][lane][2], q = (mt, nt, row half)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
          *reinterpret_cast<float2 *>(red + ((warp * NQ + (mt * NT + nt) * 2 + hf) * 32 + lane) * 2) =
              make_float2(acc[mt][nt][2 * hf], acc[mt][nt][2 * hf + 1]);
    ptx::named_bar_sync(1, SM_CONSUMERS);
    // consumer warp -> slices q = warp, warp + 8, ...; thread -> one (row, column pair) of the slice
#pragma unroll
    for (int i = 0; i < QI; ++i) {
      const int q = warp + i * SM_WARPS;
      const bool has_out = q < NQ;
      const int mt_o = q / (NT * 2), nt_o = (q >> 1) % NT, hf_o = q & 1;
      row_o[i] = mt_o * 16 + (lane >> 2) + hf_o * 8;
      col_o[i] = n_base + nt_o * 8 + (lane & 3) * 2;
      valid[i] = has_out && row_o[i] < p.L_out && col_o[i] < p.Cout;
      valid1[i] = valid[i] && col_o[i] + 1 < p.Cout;
      v0[i] = 0.f;
      v1[i] = 0.f;
      if (has_out) {
#pragma unroll
        for (int w = 0; w < SM_WARPS; ++w) {
          const float2 pr = *reinterpret_cast<const float2 *>(red + ((w * NQ + q) * 32 + lane) * 2);
          v0[i] += pr.x;
          v1[i] += pr.y;
        }
        v0[i] += p.bias[col_o[i]];
        v1[i] += p.bias[col_o[i] + 1];
      }
    }
  }
  if (p.gw > 0) {
    // ---- GroupNorm statistics over (L_out x gw): CTA partial, then the cluster's CTAs exchange theirs
    const int csize = p.gw / NCH;
    float t1 = 0.f, t2 = 0.f;
    if (warp < SM_WARPS) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < QI; ++i) {
        if (valid[i]) { s1 += v0[i]; s2 += v0[i] * v0[i]; }
        if (valid1[i]) { s1 += v1[i]; s2 += v1[i] * v1[i]; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0) { stat[2 + warp * 2] = s1; stat[3 + warp * 2] = s2; }
      ptx::named_bar_sync(1, SM_CONSUMERS);
      if (tid == 0) {
        float a = 0.f, c2 = 0.f;
        for (int w = 0; w < SM_WARPS; ++w) { a += stat[2 + w * 2]; c2 += stat[3 + w * 2]; }
        stat[0] = a;
        stat[1] = c2;
      }
    }
    if (csize > 1) {
      ptx::cluster_sync();            // all 288 threads of every CTA of the cluster
      const uint32_t mine = ptx::smem_u32(stat);
      if (warp < SM_WARPS)
        for (int r = 0; r < csize; ++r) {
          const float2 pr = ptx::ld_cluster_f32x2(ptx::mapa(mine, (uint32_t)r));
          t1 += pr.x;
          t2 += pr.y;
        }
      ptx::cluster_sync();            // nobody leaves while a peer may still read its statistics
      if (warp == SM_WARPS) return;
    } else {
      ptx::named_bar_sync(1, SM_CONSUMERS);
      t1 = stat[0];
      t2 = stat[1];
    }
    const float inv_n = 1.f / (float)(p.L_out * p.gw);
    const float m = t1 * inv_n;
    const float var = fmaxf(t2 * inv_n - m * m, 0.f);
    const float rs = rsqrtf(var + kGnEps);
#pragma unroll
    for (int i = 0; i < QI; ++i)
      if (valid[i]) {
        const int col = col_o[i];
        v0[i] = mish_tc((v0[i] - m) * rs * p.gamma[col] + p.beta[col]);
        v1[i] = mish_tc((v1[i] - m) * rs * p.gamma[col + 1] + p.beta[col + 1]);
        if (p.ttab) {
          v0[i] += p.ttab[(size_t)step * p.Cout + col];
          v1[i] += p.ttab[(size_t)step * p.Cout + col + 1];
        }
      }
  }
#pragma unroll
  for (int i = 0; i < QI; ++i)
    if (valid[i]) {
      const size_t o = ((size_t)b * p.L_out * p.out_mul + (size_t)row_o[i] * p.out_mul + p.out_phase) * p.Cout + col_o[i];
      float a0 = v0[i], a1 = v1[i];
      if (p.residual) {
        const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162 *>(p.residual + o);
        a0 += __low2float(r2);
        a1 += __high2float(r2);
      }
      if (p.out_f32) {
        float *op = reinterpret_cast<float *>(p.out) + o;
        op[0] = a0;
        if (valid1[i]) op[1] = a1;
      } else {
        *reinterpret_cast<__nv_bfloat162 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + o) = __floats2bfloat162_rn(a0, a1);
      }
    }
}

}  // namespace dad
