// K1/K2 (v3): stride-1 Conv1d-over-horizon on tcgen05 with ONE activation load per 64-channel block.
//
//   replaces  Conv1dBlock (Conv1d k -> GroupNorm(8) -> Mish) + time add + residual add   temporal_unet.py:69-73,117,122
//             and the 1x1 residual convs                                               temporal_unet.py:103
//
// Differences from conv_tc.cuh (the generic path, still used for strided / transposed / head convs):
//  * POSITION-MAJOR tile rows.  A super-tile is S_t whole samples x L positions (S_t a multiple of 8,
//    L * S_t = 128 * MH rows).  The activation box for one 64-channel block is fetched ONCE with its halo,
//    as (L + halo) positions x S_t samples rows of 128 B, ordered position-major by a tensor map whose
//    dimensions are (channel, sample, position).  Tap t of the convolution is then the SAME shared-memory
//    tile viewed S_t rows further down: S_t * 128 B is a multiple of the 1024-byte swizzle atom, so the UMMA
//    descriptor just moves its start address.  The halo rows come back zero-filled from the TMA unit (that is
//    the conv padding).  L2->SM operand traffic drops from taps x (A + B) to A + taps x B per K block.
//  * MH = 2 (L = 32): two 128-row accumulators share every weight tile.
//  * CL = 2: the two CTAs of a cluster work on neighbouring super-tiles of the same N-tile; each loads half of
//    every weight tile and multicasts it to both (weight traffic per CTA halves again).
//  * Epilogue: GroupNorm statistics reduced with lane shuffles + one shared-memory exchange per tile; the
//    normalise / Mish / time-bias / residual arithmetic runs on packed fp32x2 instructions; the bf16 tile goes
//    through a swizzled shared-memory staging buffer and leaves with a TMA store (the residual arrives the
//    same way), so global accesses are full lines instead of 16-byte pieces per row.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "f32x2.cuh"

namespace dad {

constexpr int T3_BN = 128;
constexpr int T3_BK = 64;
#ifndef DAD_T3_NB
#define DAD_T3_NB 6
#endif
constexpr int T3_NB = DAD_T3_NB;    // weight-tile ring (<= 8)
// Epilogue shape per GroupNorm width, measured per layer.  The narrow Conv1dBlocks (GroupNorm width 16
// or 32: C_out = 128 / 256) are bound by the epilogue's critical path -- an item of an L=32 layer is 4 units of
// work -- and gain 1-3 us per layer from a 4th epilogue warpgroup, which fits the register file only with 8-column
// TMEM chunks.  The wide, MMA-bound layers (GW 64 / 128) and the plain 1x1 convs (GW 0; up to K = 4096 for the
// HalfCheetah skip projections, where the narrow chunks cost 30 us) keep 3 warpgroups x 16-column chunks.
// -DDAD_T3_NWG / -DDAD_T3_CW force one shape for every instantiation (experiments).
__host__ __device__ constexpr bool t3_narrow(int gw) { return gw == 16 || gw == 32; }
__host__ __device__ constexpr int t3_nwg(int gw) {
#ifdef DAD_T3_NWG
  return DAD_T3_NWG;
#else
  return t3_narrow(gw) ? 4 : 3;
#endif
}
__host__ __device__ constexpr int t3_cw(int gw) {
#ifdef DAD_T3_CW
  return DAD_T3_CW;
#else
  return t3_narrow(gw) ? 8 : 16;
#endif
}
__host__ __device__ constexpr int t3_threads(int gw) { return 64 + 128 * t3_nwg(gw); }   // producer, MMA issuer, epilogue warpgroups
// epilogue unit width in columns: 64 where the GroupNorm width allows it (finer units = shorter accumulator
// residency and tails, half the staging memory), 128 for GroupNorm width 128
__host__ __device__ constexpr int t3_unit_cols(int gw) { return gw == 128 ? 128 : 64; }
__host__ __device__ constexpr int t3_stage_out_bytes(int gw) { return 128 * t3_unit_cols(gw) * 2; }   // bf16 staging per warpgroup
// MODE: how the CTAs of a launch cooperate
constexpr int T3_SINGLE = 0;        // one CTA per tile, tcgen05 cta_group::1
constexpr int T3_MCAST = 1;         // 2-CTA cluster, neighbouring M tiles, weight tiles multicast (cta_group::1)
constexpr int T3_PAIR = 2;          // 2-CTA cluster, ONE tcgen05 cta_group::2 MMA of M = 256: each CTA holds 128
                                    // rows of A and half of the weight tile, the leader issues for both

struct ConvT3Params {
  const float *bias, *gamma, *beta, *ttab;
  const LoopState *ls;
  unsigned long long *prof;
  int B;                     // samples in this launch
  int L;                     // positions per sample
  int S_t;                   // samples per super-tile (multiple of 8)
  int Cout;
  int n_mst, n_tiles_n;      // M super-tiles, N tiles
  int kch1, kch2;            // 64-channel blocks of source 1 / 2
  int taps;
  int tap_row[kMaxTaps];     // row offset of the tap's view into the haloed tile = (tap_off + halo_lo) * S_t
  int halo_lo;
  int a_tx_bytes;            // (L + halo) * S_t * 128: bytes one activation box delivers
  int a_stage_bytes;         // the same rounded up to 1024
  int n_a_stages;            // 2..4
  int b_stage_bytes;         // bytes of weight tile each CTA keeps per K block
  int has_res;
  int split_tail;            // 256-wide items: split the partial last round into 128-wide half entries
  int debug;
};

struct T3Smem {
  // byte offsets from the 1024-aligned base
  int a_ring, b_ring, stage_out, bars, params, scratch, total;
};

__host__ __device__ inline T3Smem t3_smem_layout(int a_stage_bytes, int n_a, int b_stage_bytes, int cout_pad, int S_t, int gw) {
  const int ng = gw > 0 ? t3_unit_cols(gw) / gw : 1;
  T3Smem s;
  s.a_ring = 0;
  s.b_ring = s.a_ring + n_a * a_stage_bytes;
  s.stage_out = s.b_ring + T3_NB * b_stage_bytes;
  s.bars = s.stage_out + t3_nwg(gw) * t3_stage_out_bytes(gw);
  s.params = s.bars + 512;
  s.scratch = s.params + 20 * cout_pad;                           // 16 B per channel (pairs) + 4 B bias
  s.total = s.scratch + t3_nwg(gw) * 4 * S_t * ng * 8 + 1024 /*alignment slack*/;
  return s;
}

template <int GW, int MH, int MODE, int NS>
__global__ void __launch_bounds__(t3_threads(GW), 1)
conv_t3_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW2,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO, const ConvT3Params p) {
  static_assert(NS == 1 || (MODE == T3_PAIR && MH == 1), "256-wide items need the CTA pair and one accumulator half");
  constexpr int CL = (MODE == T3_SINGLE) ? 1 : 2;
  constexpr int BN_ITEM = NS * T3_BN;                     // output channels per work item
  constexpr int ACC = 512 / BN_ITEM;                      // TMEM accumulator stages
  constexpr int CW = t3_cw(GW);                           // columns per TMEM load / epilogue chunk (8 or 16)
  constexpr int T3_NWG = t3_nwg(GW);                      // epilogue warpgroups; each takes every T3_NWG-th unit
  constexpr int T3_THREADS = t3_threads(GW);
  constexpr int UC = t3_unit_cols(GW);                    // columns per epilogue unit
  constexpr int UPI = BN_ITEM / UC;                       // units per whole item
  constexpr int UPH = (UPI > 1) ? UPI / 2 : 1;            // units per half entry (256-wide items only)
  constexpr int STAGE_OUT = t3_stage_out_bytes(GW);
  constexpr int NCHUNK = UC / CW;
  constexpr int NG = (GW > 0) ? UC / GW : 1;              // GroupNorm groups per unit
  constexpr int GPC = (GW > 0 && GW < CW) ? CW / GW : 1;  // groups per column chunk
  constexpr int CPG = (GW >= CW) ? GW / CW : 1;           // chunks per group
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1u);

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_tiles_n = p.n_tiles_n;                      // items along N (BN_ITEM wide)
  const int cout_pad = n_tiles_n * BN_ITEM;
  const T3Smem lay = t3_smem_layout(p.a_stage_bytes, p.n_a_stages, p.b_stage_bytes, cout_pad, p.S_t, GW);
  const uint32_t s_base = ptx::smem_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bars);
  uint64_t *full_a = bars;                 // [4]
  uint64_t *empty_a = bars + 4;            // [4]
  uint64_t *full_b = bars + 8;             // [T3_NB] (<= 8)
  uint64_t *empty_b = bars + 16;           // [T3_NB]
  uint64_t *tempty = bars + 24;            // [ACC]
  uint64_t *res_bar = bars + 28;           // [T3_NWG] (<= 4)
  // "accumulator ready" is signalled PER CONSUMER: unit u (the k-th unit of warpgroup w = u % NWG, k = u / NWG)
  // completes ufull[w][k & 1].  Every barrier is then waited on in strictly consecutive phases by one warpgroup;
  // a per-accumulator-stage barrier would be revisited by a warpgroup only every few phases and its parity
  // test would alias.
  uint64_t *ufull = bars + 32;             // [T3_NWG][2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 40);
  const uint32_t s_pair = s_base + lay.params;                      // [cout_pad/2] x {g0,g1,b0,b1 | bias0,bias1,t0,t1}
  const uint32_t s_bias = s_pair + 16u * (uint32_t)cout_pad;        // [cout_pad] floats
  const uint32_t s_scr = s_base + lay.scratch;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if constexpr (CL > 1) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    cta_rank = __shfl_sync(0xffffffffu, cta_rank, 0);      // tells the compiler the value is warp-uniform
  }
  const int kch = p.kch1 + p.kch2;
  // persistent schedule: work item -> (group of CL consecutive M super-tiles, N item); the CTAs of a cluster
  // walk the same items in lockstep and take one super-tile of the group each
  const int n_groups_m = (p.n_mst + CL - 1) / CL;
  const int total_items = n_groups_m * n_tiles_n;
  const int first_item = blockIdx.x / CL, item_stride = gridDim.x / CL;
  // Entry j of this CTA (cluster): rounds j < full_rounds take whole items; with 256-wide items the remaining
  // items (a partial last round) are split into 128-wide HALF entries so that every cluster stays busy.
  const int full_rounds = (NS == 2 && p.split_tail) ? total_items / item_stride : (1 << 28);
  const int items_full = (NS == 2 && p.split_tail) ? full_rounds * item_stride : total_items;
  const int n_half = 2 * (total_items - items_full);
  auto entry = [&](int j, int &item, bool &half, int &ns_only) -> bool {
    half = false;
    ns_only = 0;
    if (j < full_rounds) {
      item = first_item + j * item_stride;
      return item < total_items;
    }
    const int g = (j - full_rounds) * item_stride + first_item;
    if (g >= n_half) return false;
    item = items_full + (g >> 1);
    half = true;
    ns_only = g & 1;
    return true;
  };

  ptx::griddep_launch();                          // the next kernel may begin its own set-up
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA1);
    ptx::prefetch_tmap(&tmA2);
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmW2);
    ptx::prefetch_tmap(&tmR);
    ptx::prefetch_tmap(&tmO);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&full_a[s], 1);
      ptx::mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < T3_NB; ++s) {
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty_b[s], MODE == T3_MCAST ? 2 : 1);   // multicast: both CTAs must have consumed the slot
    }
    for (int s = 0; s < 2 * T3_NWG; ++s) ptx::mbar_init(&ufull[s], 1);
    for (int s = 0; s < ACC; ++s) {
      // every 128-column unit of the item is drained by 4 warps; pair: the epilogue warps of both CTAs
      ptx::mbar_init(&tempty[s], (MODE == T3_PAIR ? 8 : 4) * UPI);
    }
    for (int s = 0; s < T3_NWG; ++s) ptx::mbar_init(&res_bar[s], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (MODE == T3_PAIR) { ptx::tmem_alloc_2sm(tmem_slot, 512); ptx::tmem_relinquish_2sm(); }
    else { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  }
  // everything above touched no global data; from here on the previous kernels' results are needed
  ptx::griddep_wait();
  if (warp >= 2) {
    const int step = p.ls->step;
    const float *tt = p.ttab ? p.ttab + (size_t)step * p.Cout : nullptr;     // uniform timestep only (host guarantees)
    for (int n = threadIdx.x - 64; n < cout_pad; n += T3_THREADS - 64) {
      const bool in = n < p.Cout;
      float g = 0.f, e = 0.f;
      if constexpr (GW > 0) {
        if (in) { g = p.gamma[n]; e = p.beta[n]; }
      }
      const float bi = in ? p.bias[n] : 0.f;
      const float tv = (tt && in) ? tt[n] : 0.f;
      const uint32_t pb = s_pair + (uint32_t)(n >> 1) * 32u + (uint32_t)(n & 1) * 4u;
      ptx::sts32(pb + 0, g);
      ptx::sts32(pb + 8, e);
      ptx::sts32(pb + 16, bi);
      ptx::sts32(pb + 24, tv);
      ptx::sts32(s_bias + (uint32_t)n * 4u, bi);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) ptx::cluster_sync();      // the peer's barriers exist before anything is signalled at them
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // The producer and MMA warps stay CONVERGED: all 32 lanes walk the loops with warp-uniform values and only
  // the asynchronous instructions are issued by one elected lane.  (Inside `if (lane == 0)` the compiler has to
  // assume per-lane operands and wraps every UTMALDG / UTCHMMA in an elect-broadcast loop of ~20 instructions.)
  if (warp == 0) {
    // ================================ TMA producer ================================
    if (!(DAD_DEBUG_BITS(p) & 2)) {
      const bool leader_lane = ptx::elect_one();
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      // pair mode: both CTAs' loads complete on the LEADER's full barriers (the MMA issuer lives there)
      const uint32_t lead_full_a = (MODE == T3_PAIR) ? ptx::mapa(ptx::smem_u32(&full_a[0]), 0) : 0;
      const uint32_t lead_full_b = (MODE == T3_PAIR) ? ptx::mapa(ptx::smem_u32(&full_b[0]), 0) : 0;
      for (int j = 0;; ++j) {
        int item, ns_only;
        bool half;
        if (!entry(j, item, half, ns_only)) break;
        const int gm = item / n_tiles_n, tn = item - gm * n_tiles_n;
        const int tm = gm * CL + (int)cta_rank;
        const int b0 = tm * p.S_t, n0 = tn * BN_ITEM + ns_only * T3_BN;
        for (int ch = 0; ch < kch; ++ch) {
          // one haloed activation box per 64-channel block, shared by all taps
          ptx::mbar_wait(&empty_a[sa], pha ^ 1);
          uint8_t *dst = smem + lay.a_ring + sa * p.a_stage_bytes;
          const CUtensorMap *am = (ch < p.kch1) ? &tmA1 : &tmA2;
          const int c0 = (ch < p.kch1 ? ch : ch - p.kch1) * T3_BK;
          if (leader_lane) {
            if (DAD_DEBUG_BITS(p) & 64) {
              if (MODE != T3_PAIR || cta_rank == 0) ptx::mbar_arrive(&full_a[sa]);   // profiling: MMAs without TMA
            } else if constexpr (MODE == T3_PAIR) {
              if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_a[sa], 2u * (uint32_t)p.a_tx_bytes);
              ptx::tma_load_3d_2sm(dst, am, lead_full_a + 8u * sa, c0, b0, -p.halo_lo);
            } else {
              ptx::mbar_arrive_expect_tx(&full_a[sa], (uint32_t)p.a_tx_bytes);
              ptx::tma_load_3d(dst, am, &full_a[sa], c0, b0, -p.halo_lo);
            }
          }
          if (++sa == p.n_a_stages) { sa = 0; pha ^= 1; }
          for (int t = 0; t < p.taps; ++t) {
            ptx::mbar_wait(&empty_b[sb], phb ^ 1);
            uint8_t *wdst = smem + lay.b_ring + sb * p.b_stage_bytes;
            const int k0 = (t * kch + ch) * T3_BK;
            if (!leader_lane) {
            } else if (DAD_DEBUG_BITS(p) & 64) {
              if (MODE != T3_PAIR || cta_rank == 0) ptx::mbar_arrive(&full_b[sb]);
            } else if constexpr (MODE == T3_SINGLE) {
              ptx::mbar_arrive_expect_tx(&full_b[sb], (uint32_t)p.b_stage_bytes);
              ptx::tma_load_2d(wdst, &tmW, &full_b[sb], k0, n0);
            } else if constexpr (MODE == T3_MCAST) {
              // this CTA fetches rows [rank*64, rank*64+64) of the weight tile for the whole cluster
              ptx::mbar_arrive_expect_tx(&full_b[sb], (uint32_t)p.b_stage_bytes);
              ptx::tma_load_2d_mc(wdst + cta_rank * (p.b_stage_bytes / 2), &tmW, &full_b[sb], k0,
                                  n0 + (int)cta_rank * (BN_ITEM / 2), MC_MASK);
            } else if (half) {
              // half entry (128 output channels): 64 weight rows per CTA, fetched with the 64-row box map
              if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_b[sb], (uint32_t)p.b_stage_bytes);
              ptx::tma_load_2d_2sm(wdst, &tmW2, lead_full_b + 8u * sb, k0, n0 + (int)cta_rank * (T3_BN / 2));
            } else {
              // pair: this CTA keeps its half of the item's output channels; the MMA reads both halves
              if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_b[sb], 2u * (uint32_t)p.b_stage_bytes);
              ptx::tma_load_2d_2sm(wdst, &tmW, lead_full_b + 8u * sb, k0, n0 + (int)cta_rank * (BN_ITEM / 2));
            }
            if (++sb == T3_NB) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    // One thread feeds the tensor core (the pair leader's, in pair mode); descriptors advance by adding constants.
    if (MODE != T3_PAIR || cta_rank == 0) {
      const bool leader_lane = ptx::elect_one();
      constexpr uint32_t idesc = ptx::make_idesc_bf16(MODE == T3_PAIR ? 256 : 128, BN_ITEM);
      const uint64_t dconst = ptx::make_smem_desc_sw128(0);
      const uint32_t a_lo0 = ((s_base + lay.a_ring) >> 4), a_lo_step = (uint32_t)p.a_stage_bytes >> 4;
      const uint32_t b_lo0 = ((s_base + lay.b_ring) >> 4), b_lo_step = (uint32_t)p.b_stage_bytes >> 4;
      const uint32_t tap_step = p.taps > 1 ? (uint32_t)(p.tap_row[1] - p.tap_row[0]) * 8u : 0u;   // rows * 128 B >> 4
      const uint32_t tap_first = (uint32_t)p.tap_row[0] * 8u;
      const int n_taps = p.taps, n_a = p.n_a_stages;
      const bool run = !(DAD_DEBUG_BITS(p) & 2);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int it = 0;
      constexpr uint32_t idesc_half = ptx::make_idesc_bf16(MODE == T3_PAIR ? 256 : 128, T3_BN);
      for (;; ++it) {
        int item, ns_only;
        bool half;
        if (!entry(it, item, half, ns_only)) break;
        const uint32_t idesc_e = half ? idesc_half : idesc;
        uint32_t d_tmem[MH];
#pragma unroll
        for (int h = 0; h < MH; ++h) {
          const int u = it * MH + h;
          const int as = u % ACC;
          ptx::mbar_wait(&tempty[as], ((u / ACC) & 1) ^ 1);     // epilogue has drained this accumulator
          d_tmem[h] = tmem_base + as * BN_ITEM;
        }
        ptx::tc_fence_after();
        if (run) {
          for (int ch = 0; ch < kch; ++ch) {
            ptx::mbar_wait(&full_a[sa], pha);
            uint64_t da = dconst | (uint64_t)(a_lo0 + sa * a_lo_step + tap_first);
            for (int t = 0; t < n_taps; ++t) {
              ptx::mbar_wait(&full_b[sb], phb);
              ptx::tc_fence_after();
              const uint64_t db = dconst | (uint64_t)(b_lo0 + sb * b_lo_step);
              const uint32_t acc_kb = (ch | t) != 0;     // the first K block of an item overwrites the accumulator
              if (leader_lane) {
#pragma unroll
              for (int h = 0; h < MH; ++h) {
                // tap t = the same tile viewed tap_row rows further down (a multiple of the 1024 B swizzle atom)
#pragma unroll
                for (int k = 0; k < T3_BK / 16; ++k) {
                  const uint32_t acc = (k != 0) ? 1u : acc_kb;
                  if constexpr (MODE == T3_PAIR)
                    ptx::umma_bf16_2sm(d_tmem[h], da + (uint64_t)(h * 1024 + 2 * k), db + (uint64_t)(2 * k), idesc_e, acc);
                  else
                    ptx::umma_bf16(d_tmem[h], da + (uint64_t)(h * 1024 + 2 * k), db + (uint64_t)(2 * k), idesc_e, acc);
                }
              }
              if constexpr (MODE == T3_SINGLE) ptx::umma_commit(&empty_b[sb]);
              else if constexpr (MODE == T3_MCAST) ptx::umma_commit_mc(&empty_b[sb], MC_MASK);
              else ptx::umma_commit_2sm_mc(&empty_b[sb], MC_MASK);
              }
              __syncwarp();
              if (++sb == T3_NB) { sb = 0; phb ^= 1; }
              da += tap_step;
            }
            if (leader_lane) {
              if constexpr (MODE == T3_PAIR) ptx::umma_commit_2sm_mc(&empty_a[sa], MC_MASK);
              else ptx::umma_commit(&empty_a[sa]);
            }
            __syncwarp();
            if (++sa == n_a) { sa = 0; pha ^= 1; }
          }
        }
        if (leader_lane) {
          // units of this entry: NS for a whole item, one for a half entry (numbered after all whole items)
          const int u0 = half ? UPI * full_rounds + (it - full_rounds) * UPH : it * UPI;
#pragma unroll
          for (int ns = 0; ns < UPI; ++ns) {
            if (half && ns >= UPH) break;
            const int u = u0 + ns, k = u / T3_NWG;
            uint64_t *bar = &ufull[(u - k * T3_NWG) * 2 + (k & 1)];
            if constexpr (MODE == T3_PAIR) ptx::umma_commit_2sm_mc(bar, MC_MASK);
            else ptx::umma_commit(bar);
          }
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue ====================================
    const int wg = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                  // row within a 128-row half-tile
    const int s_smp = r % p.S_t;                  // sample within the super-tile (position-major rows)
    const int pos_per_half = 128 / p.S_t;
    const bool elected = (threadIdx.x - 64) % 128 == 0;
    const float inv_n = 1.0f / (float)(p.L * (GW > 0 ? GW : 1));
    uint8_t *stg_ptr = smem + lay.stage_out + wg * STAGE_OUT;
    const uint32_t stg = s_base + lay.stage_out + wg * STAGE_OUT;            // [UC/64 boxes][128 rows][128 B], swizzled
    const uint32_t scr = s_scr + (uint32_t)wg * (4u * p.S_t * NG * 8u);      // [4 warps][S_t][NG] x (sum, sumsq)
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t swz = (uint32_t)(r & 7);
    const uint32_t tempty0 = (MODE == T3_PAIR) ? ptx::mapa(ptx::smem_u32(&tempty[0]), 0) : ptx::smem_u32(&tempty[0]);
    uint32_t res_phase = 0;
    long long pc_wait = 0, pc_p1 = 0, pc_p2 = 0, pc_n = 0, pc_t0 = 0, pc_t1 = 0;

    // Residual tiles travel through the staging buffer: the box for unit (item, ns, h) is requested as soon as
    // the previous unit's store has finished reading the buffer.
    // unit u -> (entry index, item, 128-column sub-tile, half entry?)
    // `ns` = column block (UC wide) of the item the unit covers, `acol` = its column offset inside the accumulator
    auto unit = [&](int u, int &j, int &item, int &ns, int &acol, bool &half) -> bool {
      int ns_only;
      if (u < UPI * full_rounds) {
        j = u / UPI;
        ns = u - j * UPI;
        acol = ns * UC;
        return entry(j, item, half, ns_only);
      }
      const int v = u - UPI * full_rounds;
      j = full_rounds + v / UPH;
      const int s2 = v - (v / UPH) * UPH;
      const bool ok = entry(j, item, half, ns_only);
      ns = ns_only * UPH + s2;
      acol = s2 * UC;                       // a half entry accumulates in the first 128 columns of its stage
      return ok;
    };
    auto request_residual = [&](int item, int ns, int h) {
      const int gm = item / n_tiles_n, tn = item - gm * n_tiles_n;
      const int tm = gm * CL + (int)cta_rank;
      const int ch0 = tn * BN_ITEM + ns * UC;
      ptx::mbar_arrive_expect_tx(&res_bar[wg], STAGE_OUT);
      ptx::tma_load_3d(stg_ptr, &tmR, &res_bar[wg], ch0, tm * p.S_t, h * pos_per_half);
      if constexpr (UC == 128) ptx::tma_load_3d(stg_ptr + 16384, &tmR, &res_bar[wg], ch0 + 64, tm * p.S_t, h * pos_per_half);
    };
    auto release_acc = [&](int as, bool twice) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // a half entry has half the units the barrier expects: arrive for the missing ones too
        for (int rep = 0; rep < (twice ? 2 : 1); ++rep) {
          if constexpr (MODE == T3_PAIR) ptx::mbar_arrive_cluster(tempty0 + 8u * as);
          else ptx::mbar_arrive(&tempty[as]);
        }
      }
    };

    // units: u = it * NS + ns (one 128-column sub-tile of work item `it`, all MH halves); warpgroup wg takes u = wg (mod NWG)
    if (p.has_res && elected) {
      int j0, item0, ns0, ac0;
      bool half0;
      if (unit(wg, j0, item0, ns0, ac0, half0)) request_residual(item0, ns0, 0);
    }

    for (int u = wg;; u += T3_NWG) {
      int it, item, ns, acol;
      bool half;
      if (!unit(u, it, item, ns, acol, half)) break;
      if (DAD_PROF_PTR(p)) pc_t0 = clock64();
      const int gm = item / n_tiles_n, tn = item - gm * n_tiles_n;
      const int tm = gm * CL + (int)cta_rank;
      const int b0 = tm * p.S_t;
      const bool tile_ok = tm < p.n_mst;
      int nit, nitem = total_items, nns = 0, nacol;                      // this warpgroup's next unit
      bool nhalf;
      if (!unit(u + T3_NWG, nit, nitem, nns, nacol, nhalf)) nitem = total_items;
      const bool twice = half && NS == 2;
      uint32_t t_addr[MH];
      {
        const int k = u / T3_NWG;
        ptx::mbar_wait(&ufull[wg * 2 + (k & 1)], (k >> 1) & 1);
      }
#pragma unroll
      for (int h = 0; h < MH; ++h) t_addr[h] = tmem_base + ((it * MH + h) % ACC) * BN_ITEM + ((uint32_t)(q * 32) << 16);
      ptx::tc_fence_after();
      if (DAD_PROF_PTR(p)) { pc_t1 = clock64(); pc_wait += pc_t1 - pc_t0; pc_t0 = pc_t1; }
      if ((DAD_DEBUG_BITS(p) & 1) || !tile_ok) {
        // nothing to write for this super-tile: consume the residual that was prefetched for it and move on
        if (p.has_res) { ptx::mbar_wait(&res_bar[wg], res_phase); res_phase ^= 1; }
#pragma unroll
        for (int h = 0; h < MH; ++h) release_acc((it * MH + h) % ACC, twice);
        ptx::named_bar_sync(1 + wg, 128);
        if (p.has_res && elected && nitem < total_items) request_residual(nitem, nns, 0);
        continue;
      }

      {
        const int n0 = tn * BN_ITEM + ns * UC;
        const uint32_t col0 = (uint32_t)acol;
        if constexpr (GW > 0) {
          // ---- pass 1: GroupNorm statistics of (conv + bias) over the L positions x GW channels of each sample.
          // This thread's rows (one per half) belong to ONE sample; lanes with equal (lane % S_t) share it.
          f32x2 run1 = pk2(0.f, 0.f), run2 = pk2(0.f, 0.f);
#pragma unroll 1
          for (int c = 0; c < NCHUNK; ++c) {
            f32x2 s1[GPC], s2[GPC];
#pragma unroll
            for (int g = 0; g < GPC; ++g) { s1[g] = pk2(0.f, 0.f); s2[g] = pk2(0.f, 0.f); }
            const uint32_t sb = s_bias + (uint32_t)(n0 + c * CW) * 4u;
            f32x2 bb[CW / 2];
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              const float4 b4 = ptx::lds128(sb + j * 16);
              bb[2 * j] = pk2(b4.x, b4.y);
              bb[2 * j + 1] = pk2(b4.z, b4.w);
            }
#pragma unroll
            for (int h = 0; h < MH; ++h) {
              uint32_t v[32];
              if constexpr (CW == 16) ptx::tmem_ld16(t_addr[h] + col0 + c * CW, v); else ptx::tmem_ld8(t_addr[h] + col0 + c * CW, v);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < CW / 2; ++j) {
                const f32x2 x = fadd2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), bb[j]);
                const int g = (GW < CW) ? (2 * j) / GW : 0;       // compile-time (GW is even)
                s1[g] = fadd2(s1[g], x);
                s2[g] = ffma2(x, x, s2[g]);
              }
            }
            if constexpr (GW >= CW) {
              run1 = fadd2(run1, s1[0]);
              run2 = fadd2(run2, s2[0]);
              if ((c + 1) % CPG != 0) continue;
              s1[0] = run1; s2[0] = run2;
              run1 = pk2(0.f, 0.f); run2 = pk2(0.f, 0.f);
            }
            const int g0 = (GW >= CW) ? c / CPG : c * GPC;
#pragma unroll
            for (int g = 0; g < GPC; ++g) {
              float a0, a1, q0, q1;
              upk2(s1[g], a0, a1);
              upk2(s2[g], q0, q1);
              float t1 = a0 + a1, t2 = q0 + q1;
              for (int o = p.S_t; o < 32; o <<= 1) {
                t1 += __shfl_xor_sync(0xffffffffu, t1, o);
                t2 += __shfl_xor_sync(0xffffffffu, t2, o);
              }
              if (lane < p.S_t) ptx::sts64(scr + (uint32_t)((q * p.S_t + lane) * NG + g0 + g) * 8u, t1, t2);
            }
          }
        }
        // the staging buffer is free again once the previous store has read it; the elected thread checked that
        // before it prefetched this unit's residual (or checks it here when there is none)
        if (!p.has_res && elected) ptx::bulk_wait_read0();
        ptx::named_bar_sync(1 + wg, 128);          // statistics exchanged, staging buffer reusable
        if (DAD_PROF_PTR(p)) { pc_t1 = clock64(); pc_p1 += pc_t1 - pc_t0; pc_t0 = pc_t1; }

        // ---- pass 2: normalise, Mish, (+ time bias | + residual), convert, stage, TMA store
#pragma unroll 1
        for (int h = 0; h < MH; ++h) {
          if (p.has_res) { ptx::mbar_wait(&res_bar[wg], res_phase); res_phase ^= 1; }
          f32x2 rg2[GPC], nm2[GPC];      // (rstd, -mean) of the current group(s), carried across the chunks of a wide group
#pragma unroll
          for (int g = 0; g < GPC; ++g) { rg2[g] = pk2(0.f, 0.f); nm2[g] = pk2(0.f, 0.f); }
#pragma unroll 1
          for (int c = 0; c < NCHUNK; ++c) {
            uint32_t v[32];
            if constexpr (CW == 16) ptx::tmem_ld16(t_addr[h] + col0 + c * CW, v); else ptx::tmem_ld8(t_addr[h] + col0 + c * CW, v);
            const int nc = n0 + c * CW;
            const uint32_t sp = s_pair + (uint32_t)(nc >> 1) * 32u;
            f32x2 y[CW / 2];
            if constexpr (GW > 0) {
              const int g0 = (GW >= CW) ? c / CPG : c * GPC;
              if (GW < CW || c % CPG == 0) {
#pragma unroll
              for (int g = 0; g < GPC; ++g) {
                float t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  const float2 pr = ptx::lds64(scr + (uint32_t)((w * p.S_t + s_smp) * NG + g0 + g) * 8u);
                  t1 += pr.x;
                  t2 += pr.y;
                }
                const float m = t1 * inv_n;
                const float var = fmaxf(t2 * inv_n - m * m, 0.f);
                const float rs = rsqrtf(var + kGnEps);
                rg2[g] = pk2(rs, rs);
                nm2[g] = pk2(-m, -m);
              }
              }
              f32x2 a2[CW / 2], bsh[CW / 2], tt2[CW / 2];
#pragma unroll
              for (int j = 0; j < CW / 2; ++j) {
                const float4 p1 = ptx::lds128(sp + j * 32);          // {gamma0, gamma1, beta0, beta1}
                const float4 p2 = ptx::lds128(sp + j * 32 + 16);     // {bias0, bias1, tt0, tt1}
                const int g = (GW < CW) ? (2 * j) / GW : 0;
                a2[j] = fmul2(pk2(p1.x, p1.y), rg2[g]);
                bsh[j] = ffma2(fadd2(pk2(p2.x, p2.y), nm2[g]), a2[j], pk2(p1.z, p1.w));
                tt2[j] = pk2(p2.z, p2.w);
              }
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < CW / 2; ++j) {
                const f32x2 xn = ffma2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), a2[j], bsh[j]);
                y[j] = fadd2((DAD_DEBUG_BITS(p) & 32) ? xn : mish2(xn), tt2[j]);
              }
            } else {
              f32x2 bb[CW / 2];
#pragma unroll
              for (int j = 0; j < CW / 2; ++j) {
                const float4 p2 = ptx::lds128(sp + j * 32 + 16);
                bb[j] = pk2(p2.x, p2.y);
              }
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < CW / 2; ++j)
                y[j] = fadd2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), bb[j]);
            }
            // this thread's CW columns = CW/8 16-byte pieces of its row in the swizzled staging box
            const uint32_t box = stg + (uint32_t)((c * CW) >> 6) * 16384u + row_off;
            const uint32_t pc0 = (uint32_t)(((c * CW) & 63) >> 3);
#pragma unroll
            for (int pc = 0; pc < CW / 8; ++pc) {
              const uint32_t ad = box + (((pc0 + pc) ^ swz) << 4);
              if (p.has_res) {
                const uint4 r0 = ptx::lds128u(ad);
                const uint32_t rw[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  y[4 * pc + j] = fadd2(y[4 * pc + j], pk2(__uint_as_float(rw[j] << 16), __uint_as_float(rw[j] & 0xffff0000u)));
              }
              uint32_t ow[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float lo, hi;
                upk2(y[4 * pc + j], lo, hi);
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(ow[j]) : "f"(hi), "f"(lo));
              }
              ptx::sts128u(ad, make_uint4(ow[0], ow[1], ow[2], ow[3]));
            }
          }
          // this unit no longer needs accumulator h (the barrier counts all units of the item)
          release_acc((it * MH + h) % ACC, twice);
          // staged tile -> global with a TMA store (rows of samples >= B land in workspace padding)
          ptx::fence_proxy_async();
          ptx::named_bar_sync(1 + wg, 128);
          // the unit that uses the staging buffer next
          int nx_item = item, nx_ns = ns, nx_h = h + 1;
          if (nx_h == MH) { nx_h = 0; nx_ns = nns; nx_item = nitem; }
          const bool same_item = (h + 1 < MH);
          if (elected) {
            if (!(DAD_DEBUG_BITS(p) & 4)) {
              ptx::tma_store_3d(&tmO, stg, n0, b0, h * pos_per_half);
              if constexpr (UC == 128) ptx::tma_store_3d(&tmO, stg + 16384, n0 + 64, b0, h * pos_per_half);
            }
            ptx::bulk_commit();
            if (p.has_res && nx_item < total_items) {
              ptx::bulk_wait_read0();
              request_residual(nx_item, nx_ns, nx_h);
            }
          }
          // without a residual: the next pass 2 of this item must not overwrite the buffer while it is being read
          // (across items the wait happens right before the statistics barrier)
          if (!p.has_res && same_item) {
            if (elected) ptx::bulk_wait_read0();
            ptx::named_bar_sync(1 + wg, 128);
          }
        }
      }
      if (DAD_PROF_PTR(p)) { pc_p2 += clock64() - pc_t0; pc_n += 1; }
    }
    if (elected) ptx::bulk_wait0();               // all stores of this warpgroup have landed before the CTA exits
    if (DAD_PROF_PTR(p) && lane == 0) {
      atomicAdd(DAD_PROF_PTR(p) + 0, (unsigned long long)pc_wait);
      atomicAdd(DAD_PROF_PTR(p) + 1, (unsigned long long)pc_p1);
      atomicAdd(DAD_PROF_PTR(p) + 2, (unsigned long long)pc_p2);
      atomicAdd(DAD_PROF_PTR(p) + 3, (unsigned long long)pc_n);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) ptx::cluster_sync();      // no CTA leaves while its peer may still signal or read it
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (MODE == T3_PAIR) ptx::tmem_dealloc_2sm(tmem_base, 512);
    else ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dad
