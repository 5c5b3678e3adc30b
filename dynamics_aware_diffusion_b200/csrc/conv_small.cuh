// Small-batch (latency) variant of every U-Net convolution: GuidedPolicy.get_action plans ONE trajectory per call
// (policies.py:193-223), so a layer is a 8..32-row GEMM against up to 2.6 MB of weights and the step is a chain of
// 39 dependent launches.  The throughput kernels (conv_t3 / conv_tc) put such a layer on 2-4 CTAs and pay their
// fixed cluster / TMEM / pipeline costs per launch; here a layer is spread over Cout/16 CTAs per sample:
//   * CTA = 16 output channels of one sample.  Its weight slab ([16][taps*Cin] bf16, K-major rows as packed for the
//     tensor-map path) is fetched by 1-D bulk copies issued BEFORE griddepcontrol.wait -- weights do not depend on
//     the previous layer, so under programmatic dependent launch they stream in while that layer still runs.
//   * After the wait the haloed activation tile of the sample ((L_in + halo) rows x Cin) follows the same way.
//   * 8 warps split K; each runs mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with ldmatrix fragments straight
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

constexpr int SM_NCH = 16;          // output channels per CTA
constexpr int SM_THREADS = 256;
constexpr int SM_WARPS = SM_THREADS / 32;
constexpr int SM_MAX_L = 32;        // rows per sample the kernel handles (two m16 tiles)

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
};

struct SmallLayout {
  uint32_t pitch_a, pitch_w, off_w, off_red, off_misc, total;
};
__host__ __device__ inline SmallLayout small_layout(int Cin, int taps, int rows, int mt) {
  SmallLayout s;
  s.pitch_a = (uint32_t)Cin * 2u + 16u;
  s.pitch_w = (uint32_t)taps * Cin * 2u + 16u;
  s.off_w = (uint32_t)rows * s.pitch_a;
  s.off_red = s.off_w + SM_NCH * s.pitch_w;
  s.off_red = (s.off_red + 15u) & ~15u;
  s.off_misc = s.off_red + (uint32_t)SM_WARPS * mt * 256u * 4u;      // per warp: mt x (16 x 16) fp32 partials
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

// grid = (Cout_pad / 16, B); cluster = (max(1, gw / 16), 1, 1); MT = m16 tiles per sample (1: L_out <= 16, 2: <= 32)
template <int MT>
__global__ void __launch_bounds__(SM_THREADS, 1) conv_small_kernel(const ConvSmallParams p) {
  extern __shared__ __align__(128) unsigned char sm_raw[];
  const int Cin = p.C1 + p.C2;
  const int K = p.taps * Cin;
  const SmallLayout lay = small_layout(Cin, p.taps, p.rows, MT);
  const uint32_t sA = ptx::smem_u32(sm_raw), sW = sA + lay.off_w;
  float *red = reinterpret_cast<float *>(sm_raw + lay.off_red);
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm_raw + lay.off_misc);      // [0] weights, [1] activations
  float *stat = reinterpret_cast<float *>(sm_raw + lay.off_misc + 32);       // [0..1] this CTA's (sum, sumsq); [2..17] warp partials
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_base = blockIdx.x * SM_NCH, b = blockIdx.y;

  if (tid == 0) {
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  ptx::griddep_launch();
  __syncthreads();
  // ---- weights: independent of the previous layer, fetched ahead of the dependency wait
  if (warp == 0) {
    if (lane == 0) ptx::mbar_arrive_expect_tx(&bar[0], (uint32_t)SM_NCH * K * 2u);
    __syncwarp();
    if (lane < SM_NCH)
      ptx::bulk_load_1d(sW + lane * lay.pitch_w, p.w + (size_t)(n_base + lane) * K, (uint32_t)K * 2u, &bar[0]);
  }
  // zero the halo rows (generic stores; nothing else writes them)
  {
    const int halo_hi = p.rows - p.halo - p.L_in;
    const int vec_per_row = (int)(lay.pitch_a / 16u);
    for (int i = tid; i < (p.halo + halo_hi) * vec_per_row; i += SM_THREADS) {
      const int r = i / vec_per_row, v = i - r * vec_per_row;
      const int row = r < p.halo ? r : p.halo + p.L_in + (r - p.halo);
      *reinterpret_cast<uint4 *>(sm_raw + (size_t)row * lay.pitch_a + v * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  ptx::griddep_wait();
  // ---- activations of sample b: one bulk copy per (row, source)
  if (warp == 1) {
    const uint32_t row_bytes = (uint32_t)Cin * 2u;
    if (lane == 0) ptx::mbar_arrive_expect_tx(&bar[1], (uint32_t)p.L_in * row_bytes);
    __syncwarp();
    for (int r = lane; r < p.L_in; r += 32) {
      const uint32_t dst = sA + (uint32_t)(p.halo + r) * lay.pitch_a;
      ptx::bulk_load_1d(dst, p.in1 + ((size_t)b * p.L_in + r) * p.C1, (uint32_t)p.C1 * 2u, &bar[1]);
      if (p.C2) ptx::bulk_load_1d(dst + (uint32_t)p.C1 * 2u, p.in2 + ((size_t)b * p.L_in + r) * p.C2, (uint32_t)p.C2 * 2u, &bar[1]);
    }
  }
  const int step = p.ls->step;
  __syncthreads();                 // halo rows visible to every warp
  ptx::mbar_wait(&bar[0], 0);
  ptx::mbar_wait(&bar[1], 0);

  // ---- K split over the warps: k16 step s covers tap s / (Cin/16), channels 16 (s % (Cin/16)) ..
  float acc[MT][2][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[mt][nt][j] = 0.f;
  const int cps = Cin >> 4, n_steps = K >> 4;
  // ldmatrix row roles of this lane: A matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15); B matrices (n 0-7 | 8-15) x (k 0-7 | 8-15)
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kofs = (lane >> 4) * 8;
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_kofs = ((lane >> 3) & 1) * 8;
  int l_of[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) l_of[mt] = min(mt * 16 + a_row, p.L_out - 1);      // padding rows re-read a valid row
  const uint32_t w_lane = sW + (uint32_t)b_row * lay.pitch_w + (uint32_t)b_kofs * 2u;
  for (int s = warp; s < n_steps; s += SM_WARPS) {
    const int t = s / cps, c = (s - t * cps) << 4;
    uint32_t bf[4];
    ptx::ldmatrix_x4(w_lane + (uint32_t)s * 32u, bf[0], bf[1], bf[2], bf[3]);
    const int off = p.tap_off[t] + p.halo;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      uint32_t af[4];
      ptx::ldmatrix_x4(sA + (uint32_t)(l_of[mt] * p.in_stride + off) * lay.pitch_a + (uint32_t)(c + a_kofs) * 2u, af[0], af[1],
                       af[2], af[3]);
      ptx::mma_bf16_16816(acc[mt][0], af, bf[0], bf[1]);
      ptx::mma_bf16_16816(acc[mt][1], af, bf[2], bf[3]);
    }
  }
  // ---- sum the warps' partial tiles: red[warp][q][lane][2], q = (mt, nt, row half)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
        *reinterpret_cast<float2 *>(red + ((warp * MT * 4 + (mt * 2 + nt) * 2 + hf) * 32 + lane) * 2) =
            make_float2(acc[mt][nt][2 * hf], acc[mt][nt][2 * hf + 1]);
  __syncthreads();
  // thread -> one (row, column pair): q = tid / 32 enumerates (mt, nt, hf); MT = 1 leaves warps 4..7 without outputs
  const int q = tid >> 5;
  const bool has_out = q < MT * 4;
  const int mt_o = q >> 2, nt_o = (q >> 1) & 1, hf_o = q & 1;
  const int row = mt_o * 16 + (lane >> 2) + hf_o * 8;
  const int col = n_base + nt_o * 8 + (lane & 3) * 2;
  const bool valid = has_out && row < p.L_out && col < p.Cout;        // Cout is even whenever it is not the head
  float v0 = 0.f, v1 = 0.f;
  if (has_out) {
#pragma unroll
    for (int w = 0; w < SM_WARPS; ++w) {
      const float2 pr = *reinterpret_cast<const float2 *>(red + ((w * MT * 4 + q) * 32 + lane) * 2);
      v0 += pr.x;
      v1 += pr.y;
    }
    v0 += p.bias[col];
    v1 += p.bias[col + 1];
  }
  const bool valid1 = valid && col + 1 < p.Cout;
  if (p.gw > 0) {
    // ---- GroupNorm statistics over (L_out x gw): CTA partial, then the cluster's CTAs exchange theirs
    float s1 = valid ? v0 + (valid1 ? v1 : 0.f) : 0.f;
    float s2 = valid ? v0 * v0 + (valid1 ? v1 * v1 : 0.f) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) { stat[2 + warp * 2] = s1; stat[3 + warp * 2] = s2; }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f, c2 = 0.f;
      for (int w = 0; w < SM_WARPS; ++w) { a += stat[2 + w * 2]; c2 += stat[3 + w * 2]; }
      stat[0] = a;
      stat[1] = c2;
    }
    const int csize = p.gw / SM_NCH;
    float t1 = 0.f, t2 = 0.f;
    if (csize > 1) {
      ptx::cluster_sync();
      const uint32_t mine = ptx::smem_u32(stat);
      for (int r = 0; r < csize; ++r) {
        const float2 pr = ptx::ld_cluster_f32x2(ptx::mapa(mine, (uint32_t)r));
        t1 += pr.x;
        t2 += pr.y;
      }
      ptx::cluster_sync();          // nobody leaves while a peer may still read its statistics
    } else {
      __syncthreads();
      t1 = stat[0];
      t2 = stat[1];
    }
    const float inv_n = 1.f / (float)(p.L_out * p.gw);
    const float m = t1 * inv_n;
    const float var = fmaxf(t2 * inv_n - m * m, 0.f);
    const float rs = rsqrtf(var + kGnEps);
    if (valid) {
      v0 = mish_tc((v0 - m) * rs * p.gamma[col] + p.beta[col]);
      v1 = mish_tc((v1 - m) * rs * p.gamma[col + 1] + p.beta[col + 1]);
      if (p.ttab) {
        v0 += p.ttab[(size_t)step * p.Cout + col];
        v1 += p.ttab[(size_t)step * p.Cout + col + 1];
      }
    }
  }
  if (valid) {
    const size_t o = ((size_t)b * p.L_out * p.out_mul + (size_t)row * p.out_mul + p.out_phase) * p.Cout + col;
    if (p.residual) {
      const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162 *>(p.residual + o);
      v0 += __low2float(r2);
      v1 += __high2float(r2);
    }
    if (p.out_f32) {
      float *op = reinterpret_cast<float *>(p.out) + o;
      op[0] = v0;
      if (valid1) op[1] = v1;
    } else {
      *reinterpret_cast<__nv_bfloat162 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + o) = __floats2bfloat162_rn(v0, v1);
    }
  }
}

}  // namespace dad
