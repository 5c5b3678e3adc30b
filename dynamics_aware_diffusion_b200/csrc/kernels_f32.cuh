// fp32 SIMT kernels: the 1e-5 parity mode of the U-Net, the time-embedding tables and the
// weight re-packers.  Everything is channels-last: activations (B, L, C).
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "f32x2.cuh"

namespace dad {

// ------------------------------------------------------------------------------------------
// Implicit-GEMM Conv1d / ConvTranspose1d-phase / 1x1 conv, fp32 in, fp32 accumulate.
//   replaces nn.Conv1d / nn.ConvTranspose1d calls at temporal_unet.py:40,51,70,103,196
//   weights: Wp[tap][c][n]  (n contiguous).
// Tile 128 rows x 64 cols x 16 k, 256 threads, 8x4 outputs per thread as 4 row pairs x 4 columns of packed fp32x2
// accumulators (one FFMA2 = two IEEE fmas: the per-output accumulation order -- k ascending, one fma per k -- is the
// same as a scalar loop's, so results do not depend on the tiling).  Shared-memory tiles are double-buffered and the
// next tile's global loads are issued before the current tile is multiplied (one __syncthreads per k block).
// This kernel is the 1e-5 parity mode AND the fp32 sibling that evaluates the ill-conditioned leading reverse step of
// a bf16 model (dad_set_fp32_steps), i.e. one U-Net pass per plan sits in the timed region of the benchmark.
// ------------------------------------------------------------------------------------------
constexpr int F32_BM = 128, F32_BN = 64, F32_BK = 16;

enum { EPI_BIAS = 0, EPI_PROJECT = 1 };

struct ConvF32Params {
  const float *in1, *in2;
  const float *w;
  const float *bias;
  const float *residual;     // same layout as out, or nullptr
  float *out;
  ConvGeom g;
  int B;
  // EPI_PROJECT only
  const LoopState *ls;
  const float *alpha_tab;
  const float *cond_vals;
  int T;
};

template <int EPI, bool VEC>
__global__ void __launch_bounds__(256) conv_f32_kernel(const ConvF32Params p) {
  __shared__ __align__(16) float As[2][F32_BK][F32_BM + 4];
  __shared__ __align__(16) float Bs[2][F32_BK][F32_BN];
  ptx::griddep_launch();
  ptx::griddep_wait();
  const ConvGeom &g = p.g;
  const int Cin = g.C1 + g.C2;
  const int K = g.taps * Cin;
  const int M = p.B * g.L_out;
  const int m0 = blockIdx.x * F32_BM, n0 = blockIdx.y * F32_BN;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;        // rows ty*8 .. +7, columns tx*4 .. +3

  // A loader: one row, two groups of 4 consecutive k per thread.
  const int a_row = tid >> 1, a_k = (tid & 1) * 8;
  const int am = m0 + a_row;
  const bool a_ok = am < M;
  const int ab = a_ok ? am / g.L_out : 0;
  const int alo = a_ok ? am - ab * g.L_out : 0;
  // B loader: one k, 4 consecutive n per thread.
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  f32x2 acc[4][4];                                 // [row pair][column]: rows (ty*8 + 2i, ty*8 + 2i + 1)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = pk2(0.f, 0.f);

  float av[8], bv[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float *a4 = av + 4 * q;
      a4[0] = a4[1] = a4[2] = a4[3] = 0.f;
      if (!a_ok) continue;
      const int kk = k0 + a_k + 4 * q;
      if (VEC) {
        if (kk < K) {
          const int tap = kk / Cin, c = kk - tap * Cin;
          const int li = alo * g.in_stride + g.tap_off[tap];
          if (li >= 0 && li < g.L_in) {
            const float *src = (c < g.C1) ? p.in1 + ((size_t)(ab * g.L_in + li) * g.C1 + c)
                                          : p.in2 + ((size_t)(ab * g.L_in + li) * g.C2 + (c - g.C1));
            const float4 v = *reinterpret_cast<const float4 *>(src);
            a4[0] = v.x; a4[1] = v.y; a4[2] = v.z; a4[3] = v.w;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kj = kk + j;
          if (kj < K) {
            const int tap = kj / Cin, c = kj - tap * Cin;
            const int li = alo * g.in_stride + g.tap_off[tap];
            if (li >= 0 && li < g.L_in)
              a4[j] = (c < g.C1) ? p.in1[(size_t)(ab * g.L_in + li) * g.C1 + c]
                                 : p.in2[(size_t)(ab * g.L_in + li) * g.C2 + (c - g.C1)];
          }
        }
      }
    }
    bv[0] = bv[1] = bv[2] = bv[3] = 0.f;
    const int kk = k0 + b_k, n = n0 + b_n;
    if (kk < K) {
      const float *src = p.w + (size_t)kk * g.Cout + n;
      if (VEC && n + 3 < g.Cout) {
        const float4 v = *reinterpret_cast<const float4 *>(src);
        bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < g.Cout) bv[j] = src[j];
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[buf][a_k + j][a_row] = av[j];
    *reinterpret_cast<float4 *>(&Bs[buf][b_k][b_n]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
  };

  load_tile(0);
  store_tile(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += F32_BK) {
    const bool more = k0 + F32_BK < K;
    if (more) load_tile(k0 + F32_BK);              // global loads in flight while this tile is multiplied
#pragma unroll
    for (int k = 0; k < F32_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
      const f32x2 ar[4] = {pk2(a0.x, a0.y), pk2(a0.z, a0.w), pk2(a1.x, a1.y), pk2(a1.z, a1.w)};
      const f32x2 br[4] = {pk2(b.x, b.x), pk2(b.y, b.y), pk2(b.z, b.z), pk2(b.w, b.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = ffma2(ar[i], br[j], acc[i][j]);
    }
    if (more) {
      store_tile(buf ^ 1);                         // the other buffer: its readers finished before the last barrier
      __syncthreads();
      buf ^= 1;
    }
  }

  // ---- epilogue
  float alpha = 0.f;
  int n_cond = 0;
  LoopState ls;
  if (EPI == EPI_PROJECT) {
    ls = *p.ls;
    alpha = p.alpha_tab[ls.step];
    n_cond = ((ls.flags & 1u) && !(ls.flags & 4u)) ? ls.n_cond : 0;  // project -> inpaint order
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    const int b = m / g.L_out, lo = m - b * g.L_out;
    const size_t orow = ((size_t)b * (g.L_out * g.out_mul) + lo * g.out_mul + g.out_phase) * g.Cout;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.Cout) continue;
      float r0, r1;
      upk2(acc[i >> 1][j], r0, r1);
      float v = ((i & 1) ? r1 : r0) + p.bias[n];
      if (EPI == EPI_BIAS) {
        if (p.residual) v += p.residual[orow + n];
      } else {
        // y = x + alpha * (N x + q), then Diffuser-style inpainting (policies.py:48-63).
        v = p.in1[(size_t)m * g.C1 + n] + alpha * v;
        const int hh = n / p.T, tt = n - hh * p.T;
        for (int c = 0; c < n_cond; ++c)
          if (ls.cond_h[c] == hh)
            v = p.cond_vals[((size_t)c * (ls.cond_per_batch ? ls.cond_B : 1) +
                             (ls.cond_per_batch ? (ls.cond_row0 + m) : 0)) * p.T + tt];
        if (ls.trace) ls.trace[(size_t)(ls.n_steps - 1 - ls.step) * ls.trace_stride + orow + n] = v;
      }
      (EPI == EPI_PROJECT ? ls.x : p.out)[orow + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// The same implicit GEMM on the legacy tensor path: mma.sync.m16n8k8 TF32 operands, fp32 accumulate, fp32 activations
// in and out.  Used by the fp32 SIBLING of a bf16 model for the ill-conditioned leading reverse step
// (dad_set_fp32_steps + dad_set_fp32_math): that step needs eps several times more accurate than bf16 delivers, not
// IEEE fp32, and the SIMT kernel above makes it cost ~5 % of a 500-step plan.
//   X3 = false: operands rounded to TF32 with round-to-nearest when they enter shared memory (10-bit mantissa vs
//               bf16's 7, and activations stay fp32 between layers);
//   X3 = true : 3xTF32 error compensation (a_lo b_hi + a_hi b_lo + a_hi b_hi): fp32-level accuracy at 3 MMAs per step.
// Tile 128 x 64 x 16, 8 warps as 4 (M) x 2 (N), each 32 x 32 = 2 x 4 MMA tiles.  Shared-memory strides 136 / 72 floats
// make every fragment load bank-conflict-free.  Layers with channel counts that are not multiples of 4 / 8 (the first
// conv, the head) stay on the SIMT kernel.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cvt_tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool X3>
__global__ void __launch_bounds__(256) conv_tf32_kernel(const ConvF32Params p) {
  constexpr int SA = F32_BM + 8, SB = F32_BN + 8;
  __shared__ __align__(16) float As[2][F32_BK][SA];
  __shared__ __align__(16) float Bs[2][F32_BK][SB];
  ptx::griddep_launch();
  ptx::griddep_wait();
  const ConvGeom &g = p.g;
  const int Cin = g.C1 + g.C2;
  const int K = g.taps * Cin;
  const int M = p.B * g.L_out;
  const int m0 = blockIdx.x * F32_BM, n0 = blockIdx.y * F32_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  const int gid = lane >> 2, tig = lane & 3;

  const int a_row = tid >> 1, a_k = (tid & 1) * 8;
  const int am = m0 + a_row;
  const bool a_ok = am < M;
  const int ab = a_ok ? am / g.L_out : 0;
  const int alo = a_ok ? am - ab * g.L_out : 0;
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  float av[8], bv[4];
  // Channel counts are multiples of 16 here (the host routes the others to the SIMT kernel), so a 16-wide k block lies
  // inside one tap and one source: (tap, channel) of the block advance incrementally, no division in the loop.
  int l_tap = 0, l_c = 0;                 // tap and first channel of the k block load_tile fetches next
  const float *b_src = p.w + (size_t)b_k * g.Cout + n0 + b_n;
  const bool b_ok = n0 + b_n + 3 < g.Cout;
  auto load_tile = [&](int k0) {
    const int li = alo * g.in_stride + g.tap_off[l_tap];
    const bool ok = a_ok && k0 < K && li >= 0 && li < g.L_in;
    const int c = l_c + a_k;
    const float *src = (c < g.C1) ? p.in1 + ((size_t)(ab * g.L_in + li) * g.C1 + c)
                                  : p.in2 + ((size_t)(ab * g.L_in + li) * g.C2 + (c - g.C1));
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, vb4 = v0;
    if (ok) {
      v0 = *reinterpret_cast<const float4 *>(src);
      v1 = *reinterpret_cast<const float4 *>(src + 4);
    }
    if (b_ok && k0 < K) vb4 = *reinterpret_cast<const float4 *>(b_src + (size_t)k0 * g.Cout);
    av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
    av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
    bv[0] = vb4.x; bv[1] = vb4.y; bv[2] = vb4.z; bv[3] = vb4.w;
    l_c += F32_BK;
    if (l_c >= Cin) { l_c = 0; ++l_tap; }
  };
  auto rnd = [](float x) -> float { return X3 ? x : __uint_as_float(cvt_tf32_rn(x)); };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[buf][a_k + j][a_row] = rnd(av[j]);
    *reinterpret_cast<float4 *>(&Bs[buf][b_k][b_n]) = make_float4(rnd(bv[0]), rnd(bv[1]), rnd(bv[2]), rnd(bv[3]));
  };

  load_tile(0);
  store_tile(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += F32_BK) {
    const bool more = k0 + F32_BK < K;
    if (more) load_tile(k0 + F32_BK);
#pragma unroll
    for (int kk = 0; kk < F32_BK; kk += 8) {
      float af[2][4], bf[4][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r = wm + mi * 16 + gid;
        af[mi][0] = As[buf][kk + tig][r];
        af[mi][1] = As[buf][kk + tig][r + 8];
        af[mi][2] = As[buf][kk + tig + 4][r];
        af[mi][3] = As[buf][kk + tig + 4][r + 8];
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int c = wn + ni * 8 + gid;
        bf[ni][0] = Bs[buf][kk + tig][c];
        bf[ni][1] = Bs[buf][kk + tig + 4][c];
      }
      if constexpr (X3) {
        uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ah[mi][e] = cvt_tf32_rn(af[mi][e]);
            al[mi][e] = cvt_tf32_rn(af[mi][e] - __uint_as_float(ah[mi][e]));
          }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            bh[ni][e] = cvt_tf32_rn(bf[ni][e]);
            bl[ni][e] = cvt_tf32_rn(bf[ni][e] - __uint_as_float(bh[ni][e]));
          }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            mma_tf32_1688(acc[mi][ni], al[mi], bh[ni][0], bh[ni][1]);      // small terms first
            mma_tf32_1688(acc[mi][ni], ah[mi], bl[ni][0], bl[ni][1]);
            mma_tf32_1688(acc[mi][ni], ah[mi], bh[ni][0], bh[ni][1]);
          }
      } else {
        uint32_t au[2][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int e = 0; e < 4; ++e) au[mi][e] = __float_as_uint(af[mi][e]);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
            mma_tf32_1688(acc[mi][ni], au[mi], __float_as_uint(bf[ni][0]), __float_as_uint(bf[ni][1]));
      }
    }
    if (more) {
      store_tile(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // ---- epilogue: bias (+ residual); c0/c1 = (row gid, cols 2 tig, 2 tig + 1), c2/c3 = row gid + 8
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int m = m0 + wm + mi * 16 + gid + hh * 8;
      if (m >= M) continue;
      const int b = m / g.L_out, lo = m - b * g.L_out;
      const size_t orow = ((size_t)b * (g.L_out * g.out_mul) + lo * g.out_mul + g.out_phase) * g.Cout;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int n = n0 + wn + ni * 8 + 2 * tig;
        if (n + 1 >= g.Cout + 1) continue;               // Cout is even here
        float2 v = make_float2(acc[mi][ni][2 * hh] + p.bias[n], acc[mi][ni][2 * hh + 1] + p.bias[n + 1]);
        if (p.residual) {
          const float2 r = *reinterpret_cast<const float2 *>(p.residual + orow + n);
          v.x += r.x;
          v.y += r.y;
        }
        *reinterpret_cast<float2 *>(p.out + orow + n) = v;
      }
    }
}

// ------------------------------------------------------------------------------------------
// GroupNorm(8) + Mish (+ time bias | + residual), fp32.  One block per (sample, group).
//   temporal_unet.py:71-72 (GN, Mish), :117 (time add), :122 (residual add)
// ------------------------------------------------------------------------------------------
struct GnF32Params {
  const float *in;
  float *out;
  const float *gamma, *beta;
  const float *ttab;          // [n_timesteps][C] time-bias table or nullptr
  const float *residual;      // (B, L, C) or nullptr
  const LoopState *ls;
  int L, C;
};

__global__ void __launch_bounds__(128) gn_mish_f32_kernel(const GnF32Params p) {
  __shared__ float red[32];
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int b = blockIdx.x, grp = blockIdx.y;
  const int gw = p.C / kGroups;
  const int n = p.L * gw;
  const float *src = p.in + (size_t)b * p.L * p.C + grp * gw;
  float s = 0.f;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int l = e / gw, j = e - l * gw;
    s += src[(size_t)l * p.C + j];
  }
  const float mean = block_sum(s, red) / (float)n;
  float q = 0.f;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int l = e / gw, j = e - l * gw;
    const float d = src[(size_t)l * p.C + j] - mean;
    q += d * d;
  }
  const float var = block_sum(q, red) / (float)n;
  const float rstd = 1.0f / sqrtf(var + kGnEps);
  long long t = 0;
  if (p.ttab) t = p.ls->t_rows ? min(max(p.ls->t_rows[b], 0ll), (long long)p.ls->n_table - 1) : (long long)p.ls->step;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int l = e / gw, j = e - l * gw;
    const int c = grp * gw + j;
    const size_t idx = ((size_t)b * p.L + l) * p.C + c;
    float v = (p.in[idx] - mean) * rstd * p.gamma[c] + p.beta[c];
    v = mish_precise(v);
    if (p.ttab) v += p.ttab[(size_t)t * p.C + c];
    if (p.residual) v += p.residual[idx];
    p.out[idx] = v;
  }
}

// The same operation with ONE WARP per (sample, group) and the group's L x gw elements held in registers (NV float4 per
// lane): one read, one write, no block barriers, no per-element division in the passes.  Used whenever the group fits
// (L * gw <= 128 * NV; every layer of the PointMaze / HalfCheetah / Door U-Nets at H = 32); two-pass variance as above.
template <int NV>
__global__ void __launch_bounds__(256) gn_mish_f32_warp_kernel(const GnF32Params p, int n_pairs) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);        // (sample, group) index
  if (pair >= n_pairs) return;
  const int b = pair / kGroups, grp = pair - b * kGroups;
  const int gw = p.C / kGroups, g4 = gw >> 2;                  // gw is a multiple of 4 here
  const int n4 = p.L * g4;
  const size_t base = (size_t)b * p.L * p.C + (size_t)grp * gw;
  float4 v[NV];
  size_t off[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e4 = lane + 32 * i;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    off[i] = 0;
    if (e4 < n4) {
      const int l = e4 / g4, j = (e4 - l * g4) * 4;
      off[i] = base + (size_t)l * p.C + j;
      v[i] = *reinterpret_cast<const float4 *>(p.in + off[i]);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float n = (float)(p.L * gw);
  const float mean = warp_sum(s) / n;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (lane + 32 * i < n4) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / n + kGnEps);
  long long t = 0;
  if (p.ttab) t = p.ls->t_rows ? min(max(p.ls->t_rows[b], 0ll), (long long)p.ls->n_table - 1) : (long long)p.ls->step;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e4 = lane + 32 * i;
    if (e4 >= n4) continue;
    const int l = e4 / g4, c = grp * gw + (e4 - l * g4) * 4;
    const float4 ga = *reinterpret_cast<const float4 *>(p.gamma + c), be = *reinterpret_cast<const float4 *>(p.beta + c);
    float4 o;
    o.x = mish_precise((v[i].x - mean) * rstd * ga.x + be.x);
    o.y = mish_precise((v[i].y - mean) * rstd * ga.y + be.y);
    o.z = mish_precise((v[i].z - mean) * rstd * ga.z + be.z);
    o.w = mish_precise((v[i].w - mean) * rstd * ga.w + be.w);
    if (p.ttab) {
      const float4 tt = *reinterpret_cast<const float4 *>(p.ttab + (size_t)t * p.C + c);
      o.x += tt.x; o.y += tt.y; o.z += tt.z; o.w += tt.w;
    }
    if (p.residual) {
      const float4 r = *reinterpret_cast<const float4 *>(p.residual + off[i]);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    *reinterpret_cast<float4 *>(p.out + off[i]) = o;
  }
}

// ------------------------------------------------------------------------------------------
// Time-embedding tables, computed once at load for every step index (K6).
//   temporal_unet.py:26-32 (sinusoid), :155-160 (global MLP), :97-100 (per-block Mish + Linear)
// ------------------------------------------------------------------------------------------
__global__ void sinusoid_table_kernel(float *out, int S, int dim) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * half) return;
  const int s = idx / half, j = idx - s * half;
  const float scale = logf(10000.f) / (float)(half - 1);
  const float f = expf((float)j * -scale);
  const float a = (float)s * f;
  out[(size_t)s * dim + j] = sinf(a);
  out[(size_t)s * dim + half + j] = cosf(a);
}

// out[s][n] = act_out( sum_k act_in(in[s][k]) * W[n][k] + b[n] ), W in torch Linear layout (N, K).
// fp64 accumulation: these tables are built once and feed every step.
__global__ void linear_rows_kernel(const float *__restrict__ in, const float *__restrict__ W,
                                   const float *__restrict__ b, float *__restrict__ out, int S, int K,
                                   int N, int mish_in, int mish_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * N) return;
  const int s = idx / N, n = idx - s * N;
  double acc = 0.0;
  for (int k = 0; k < K; ++k) {
    float a = in[(size_t)s * K + k];
    if (mish_in) a = mish_precise(a);
    acc += (double)a * (double)W[(size_t)n * K + k];
  }
  float v = (float)(acc + (double)b[n]);
  if (mish_out) v = mish_precise(v);
  out[idx] = v;
}

// ------------------------------------------------------------------------------------------
// Weight re-packers (run once per load_state_dict).
// ------------------------------------------------------------------------------------------
// torch Conv1d weight (Cout, Cin, k) or ConvTranspose1d weight (Cin, Cout, k)  ->  Wp[t][c][n] fp32,
// taking kernel index ksel[t] for packed tap t.
struct TapSel { int k[kMaxTaps]; };

__global__ void pack_w_f32_kernel(const float *__restrict__ w, float *__restrict__ out, int Cout, int Cin,
                                  int ksize, int taps, TapSel sel, int transposed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= taps * Cin * Cout) return;
  const int n = idx % Cout, c = (idx / Cout) % Cin, t = idx / (Cout * Cin);
  const int kk = sel.k[t];
  out[idx] = transposed ? w[((size_t)c * Cout + n) * ksize + kk] : w[((size_t)n * Cin + c) * ksize + kk];
}

// Same source layouts -> fp32 K-major rows rounded to TF32 (round to nearest): Wk[n][t * Cin + c], the B operand of
// conv_tc_kernel<BN, GW, true>.
__global__ void pack_w_tf32_kmajor_kernel(const float *__restrict__ w, float *__restrict__ out, int Cout, int Cin,
                                          int ksize, int taps, TapSel sel, int transposed) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)taps * Cin * Cout) return;
  const int c = (int)(idx % Cin), t = (int)((idx / Cin) % taps), n = (int)(idx / ((size_t)Cin * taps));
  const int kk = sel.k[t];
  const float v = transposed ? w[((size_t)c * Cout + n) * ksize + kk] : w[((size_t)n * Cin + c) * ksize + kk];
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  out[idx] = __uint_as_float(r);
}

// Same source layouts -> bf16 K-major rows: Wb[n][t * Cin_pad + c], zero padded to (Cout_pad, Cin_pad).
__global__ void pack_w_bf16_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ out, int Cout,
                                   int Cin, int Cout_pad, int Cin_pad, int ksize, int taps, TapSel sel,
                                   int transposed) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)Cout_pad * taps * Cin_pad;
  if (idx >= total) return;
  const int c = (int)(idx % Cin_pad);
  const int t = (int)((idx / Cin_pad) % taps);
  const int n = (int)(idx / ((size_t)Cin_pad * taps));
  float v = 0.f;
  if (n < Cout && c < Cin) {
    const int kk = sel.k[t];
    v = transposed ? w[((size_t)c * Cout + n) * ksize + kk] : w[((size_t)n * Cin + c) * ksize + kk];
  }
  out[idx] = __float2bfloat16_rn(v);
}

// Projector N (D x D fp32, row-major: y_n = sum_k N[n][k] x_k) -> bf16 K-major rows for the 3-term split GEMM:
// out[n][0:Kp) = hi(N[n][:]), out[n][Kp:2Kp) = hi(N[n][:]), out[n][2Kp:3Kp) = lo(N[n][:]); zero padded to (Np, 3 Kp).
__global__ void pack_projector_bf16_kernel(const float *__restrict__ Nm, __nv_bfloat16 *__restrict__ out, int D, int Kp,
                                           int Np) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Np * Kp) return;
  const int n = (int)(idx / Kp), k = (int)(idx - (size_t)n * Kp);
  float v = (n < D && k < D) ? Nm[(size_t)n * D + k] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16 *row = out + (size_t)n * 3 * Kp;
  row[k] = hi;
  row[Kp + k] = hi;
  row[2 * Kp + k] = lo;
}

// x fp32 (B, H, T) -> bf16 (B, H, Cpad), zero padded channels: first-layer operand of the bf16 path.
__global__ void pack_x_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ out, size_t rows,
                                   int T, int Cpad) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * (size_t)Cpad) return;
  const size_t r = idx / Cpad;
  const int c = (int)(idx - r * Cpad);
  out[idx] = __float2bfloat16_rn(c < T ? x[r * T + c] : 0.f);
}

}  // namespace dad
