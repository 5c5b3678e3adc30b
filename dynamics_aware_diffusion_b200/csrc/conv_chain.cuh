// K1/K2 (v4): a CHAIN of stride-1 Conv1d-over-horizon layers in ONE persistent launch.
//
//   replaces  ResidualTemporalBlock.forward:  Conv1dBlock -> + time_mlp(t) -> Conv1dBlock -> + residual_conv(x)
//             (temporal_unet.py:106-122) and runs of consecutive blocks of one U-Net level
//             (temporal_unet.py:214-237), which share the sequence length, the width and the GroupNorm shape.
//
// The per-layer kernel of round 1 (conv_t3.cuh) paid, per launch, a prologue (barrier init, TMEM allocation,
// parameter tables), a pipeline fill and an un-overlapped last epilogue: ~7 us x 29 stride-1 launches per U-Net
// pass, and a hard floor for small batches.  Here the CTA pairs of one launch walk a flat list of work entries
// that spans ALL convolutions of the chain:
//
//   entry G = first + k * n_clusters   ->   conv ci = G / items_per_conv,  item = G % items_per_conv
//
// so a cluster whose share of conv c is finished starts on conv c+1 while the others still drain conv c (no
// partial last round per layer either).  Samples never interact inside the U-Net, so an item of conv c+1 (a group
// of sample tiles x an N tile) depends only on the items of conv c that cover the SAME sample tile.  Every
// (conv, sample tile) has a counter in global memory:
//
//   producer side  epilogue warpgroup: TMA store of an output unit -> cp.async.bulk.wait_group (completion, not
//                  just .read) -> fence.proxy.async.global -> red.release.gpu.add counter[conv][tile]
//   consumer side  TMA-producer warp (activation operand) / epilogue (residual operand):
//                  ld.acquire.gpu counter >= units_per_tile -> fence.proxy.async.global -> TMA load
//
// Waits only ever point at a LOWER conv index, and the grid is sized to be co-resident (occupancy query on the
// host), so the schedule cannot deadlock.  The counters are zeroed by stage_x_kernel, the first kernel of every
// U-Net pass.  The 1x1 residual convolution of a channel-changing block is one more conv of the chain with a
// plain (bias-only) epilogue.
//
// Inside an item nothing changed from conv_t3.cuh: position-major haloed activation tiles fetched once per
// 64-channel block (tap t = the same tile viewed t * S_t rows further down), CTA pairs issuing
// tcgen05.mma.cta_group::2 (M = 256, N = 128 or 256), accumulators in TMEM, converged producer / MMA warps, and the
// GroupNorm + Mish + time-bias + residual epilogue on packed fp32x2 with TMA-store staging.  New in the epilogue:
//  * per-unit parameters: the (gamma, beta, bias, time bias) of a unit's 64 / 128 columns are fetched while the
//    previous unit is processed, into a double-buffered 2.5 KB table per warpgroup (a whole-layer table would be
//    40 KB for C_out = 2048 and would have to be swapped at every conv of the chain);
//  * GroupNorm width 256 (C_out = 2048: HalfCheetah / Door bottlenecks): a group spans the two 128-column units of
//    an item, which two warpgroups process side by side; they exchange their partial statistics through shared
//    memory and one 256-thread named barrier.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "f32x2.cuh"

namespace dad {

constexpr int CH_BN = 128;
constexpr int CH_BK = 64;
#ifndef DAD_CH_MAX_CONVS
#define DAD_CH_MAX_CONVS 12
#endif
constexpr int CH_MAX_CONVS = DAD_CH_MAX_CONVS;
constexpr int CH_MAX_NB = 8;          // weight-tile ring depth (runtime, <= 8)
#ifndef DAD_CH_EARLY_PUBLISH
#define DAD_CH_EARLY_PUBLISH 1
#endif

__host__ __device__ constexpr bool ch_narrow(int gw) { return gw == 16 || gw == 32; }
// Epilogue shape per GroupNorm width: 16-column TMEM chunks everywhere (8 independent fp32x2 chains per thread and loop
// iteration: the epilogue is bound by dependent-instruction latency, not by a pipe; 8-column chunks were 2-8 % slower on
// the narrow layers, 32-column chunks spill).  Warpgroups: 4 for the narrow Conv1dBlocks (widths 16 / 32: an L=32 item
// is 4 units of work and TMEM holds only 2 such items), 3 for width 64 (MMA-bound), 2 for widths 128 / 256 (C_out >=
// 1024, K >= 5120: an item's MMAs take ~20 us, its two 128-column units ~8 us each; two warpgroups also free 32 KB of
// staging memory for the L = 4 bottleneck of the four-level U-Nets, and width 256 pairs them on the halves of a group).
#ifndef DAD_CH_NARROW_NWG
#define DAD_CH_NARROW_NWG 4
#endif
#ifndef DAD_CH_NARROW_CW
#define DAD_CH_NARROW_CW 16
#endif
__host__ __device__ constexpr int ch_nwg(int gw) { return gw >= 128 ? 2 : (ch_narrow(gw) ? DAD_CH_NARROW_NWG : 3); }
__host__ __device__ constexpr int ch_cw(int gw) { return ch_narrow(gw) ? DAD_CH_NARROW_CW : 16; }
__host__ __device__ constexpr int ch_threads(int gw) { return 64 + 128 * ch_nwg(gw); }
__host__ __device__ constexpr int ch_unit_cols(int gw) { return gw >= 128 ? 128 : 64; }
__host__ __device__ constexpr int ch_stage_out_bytes(int gw) { return 128 * ch_unit_cols(gw) * 2; }
__host__ __device__ constexpr int ch_groups_per_unit(int gw) { return ch_unit_cols(gw) / gw > 0 ? ch_unit_cols(gw) / gw : 1; }

// Scalar description of one convolution of the chain.
struct ChainConvMeta {
  const float *bias, *gamma, *beta, *ttab;     // ttab: [n_timesteps][Cout] time-bias table or nullptr
  unsigned *flag_out;                          // [n_mst] completion counters of THIS conv's output tiles (nullptr: no in-chain consumer)
  const unsigned *flag_a;                      // counters of the in-chain conv that produces source 1, or nullptr
  const unsigned *flag_r;                      // counters of the in-chain conv that produces the residual, or nullptr
  int need_a, need_r;                          // counter values that mean "tile complete" (per U-Net pass)
  int Cout;
  int kch1, kch2;                              // 64-channel blocks of source 1 / 2 (channel concat)
  int taps;
  int tap_first, tap_step;                     // descriptor offsets (16-byte units) of tap 0 and between taps
  int halo_lo;
  int a_tx_bytes;                              // bytes one haloed activation box delivers
  int has_res;
  int plain;                                   // 1: bias-only epilogue (1x1 residual conv), 0: GroupNorm + Mish + time bias
  int pad_[2];
};

// One convolution as the device sees it: its tensor maps and scalars.  The whole chain travels in the kernel's
// parameter space (__grid_constant__, ~0.75 KB per conv; CUDA >= 12.1 allows 32 KB of parameters), where TMA
// descriptors may live without any tensormap-proxy fencing.
struct alignas(128) ChainConv {
  CUtensorMap tmA1, tmA2, tmW, tmR, tmO;
  CUtensorMap tmW1;            // weights boxed for 128-wide items (64 rows per CTA): the NS = 1 launch of a chain planned with NS = 2
  ChainConvMeta m;
};

struct ChainParams {
  const LoopState *ls;
  unsigned *err;               // set before a trap when a dependency wait times out (diagnostic)
  int n_convs;
  int flag_epoch;              // counters are compared with need * flag_epoch (1 in a U-Net pass; timing loops count up)
  int B, L, S_t;
  int n_mst;                   // sample tiles
  int n_tiles_n;               // items along N (128 * NS columns each); the same for every conv of the chain
  int a_stage_bytes, n_a_stages, b_stage_bytes, nb_stages;
  int w_alt;                   // 1: this launch runs the NS = 1 instantiation of an NS = 2 chain (weights through tmW1)
  int debug;
};

struct ChainArgs {
  ChainConv convs[CH_MAX_CONVS];
  ChainParams p;
};

struct ChSmem {
  int a_ring, b_ring, stage_out, bars, wparams, scratch, total;
};

__host__ __device__ inline ChSmem ch_smem_layout(int a_stage_bytes, int n_a, int b_stage_bytes, int nb, int S_t, int gw) {
  const int uc = ch_unit_cols(gw), nwg = ch_nwg(gw);
  ChSmem s;
  s.a_ring = 0;
  s.b_ring = s.a_ring + n_a * a_stage_bytes;
  s.stage_out = s.b_ring + nb * b_stage_bytes;
  s.bars = s.stage_out + nwg * ch_stage_out_bytes(gw);
  s.wparams = s.bars + 512;
  s.scratch = s.wparams + nwg * 2 * 20 * uc;                               // [warpgroup][2 buffers] x 20 B per column
  s.total = s.scratch + nwg * 2 * 4 * S_t * ch_groups_per_unit(gw) * 8 + 1024 /*alignment slack*/;
  return s;
}

template <int CW>
__device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, uint32_t (&v)[32]) {
  if constexpr (CW == 32) ptx::tmem_ld32(taddr, v);
  else if constexpr (CW == 16) ptx::tmem_ld16(taddr, v);
  else ptx::tmem_ld8(taddr, v);
}

template <int GW, int MH, int NS>
__global__ void __launch_bounds__(ch_threads(GW), 1)
conv_chain_kernel(const __grid_constant__ ChainArgs args) {
  const ChainParams &p = args.p;
  const ChainConv *convs = args.convs;
  static_assert(GW == 16 || GW == 32 || GW == 64 || GW == 128 || GW == 256, "GroupNorm width");
  static_assert(NS == 1 || MH == 1, "256-wide items use one accumulator half per CTA");
  static_assert(GW != 256 || NS == 2, "a 256-column GroupNorm group needs 256-wide items");
  constexpr int BN_ITEM = NS * CH_BN;                     // output channels per work item
  constexpr int ACC = 512 / BN_ITEM;                      // TMEM accumulator stages
  constexpr int CW = ch_cw(GW);                           // columns per TMEM load / epilogue chunk
  constexpr int NWG = ch_nwg(GW);                         // epilogue warpgroups; unit u belongs to warpgroup u % NWG
  constexpr int UC = ch_unit_cols(GW);                    // columns per epilogue unit
  constexpr int UPI = BN_ITEM / UC;                       // units per item
  constexpr int STAGE_OUT = ch_stage_out_bytes(GW);
  constexpr int NCHUNK = UC / CW;
  constexpr bool XWG = GW > UC;                           // a group spans the units of two warpgroups (GW = 256)
  constexpr int NG = ch_groups_per_unit(GW);              // GroupNorm groups per unit (1 partial group when XWG)
  constexpr int GPC = (GW < CW) ? CW / GW : 1;            // groups per column chunk
  constexpr int CPG = (GW >= CW) ? (XWG ? NCHUNK : GW / CW) : 1;   // chunks of a unit that make up one (partial) group
  constexpr int PBUF = 20 * UC;                           // bytes of one per-unit parameter buffer
  constexpr bool SPLIT8 = ch_narrow(GW) && CW == 16;      // pass 1 keeps conv_t3's 8-column summation order
  static_assert(!XWG || (NWG == 2 && UPI == 2), "cross-warpgroup statistics: one warpgroup per half of the item");
  constexpr uint16_t MC_MASK = 3;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const ChSmem lay = ch_smem_layout(p.a_stage_bytes, p.n_a_stages, p.b_stage_bytes, p.nb_stages, p.S_t, GW);
  const uint32_t s_base = ptx::smem_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bars);
  uint64_t *full_a = bars;                 // [4]
  uint64_t *empty_a = bars + 4;            // [4]
  uint64_t *full_b = bars + 8;             // [CH_MAX_NB]
  uint64_t *empty_b = bars + 16;           // [CH_MAX_NB]
  uint64_t *tempty = bars + 24;            // [ACC]
  uint64_t *res_bar = bars + 28;           // [NWG]
  // "accumulator ready" per consumer: unit u (the k-th unit of warpgroup w = u % NWG) completes ufull[w][k & 3], so
  // every barrier is waited on in strictly consecutive phases by one warpgroup (conv_t3.cuh, parity aliasing).
  // FOUR slots per warpgroup: a thread may sit in a dependency wait (the elected thread polling a tile counter) while
  // the MMA warp keeps committing; at most ACC * UPI <= 8 units are in flight, a warpgroup owns every NWG-th of them,
  // and units k and k + 4 (4 * NWG + 1 > 8 apart) can never be in flight together -- with two slots k and k + 2 can
  // (NWG = 3), and a thread that misses two completions of one barrier waits for ever.
  uint64_t *ufull = bars + 32;             // [NWG][4]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 48);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  cta_rank = __shfl_sync(0xffffffffu, cta_rank, 0);      // tells the compiler the value is warp-uniform
  const int n_tiles_n = p.n_tiles_n;
  const int n_groups_m = (p.n_mst + 1) / 2;
  const int items_per_conv = n_groups_m * n_tiles_n;
  const int total_entries = items_per_conv * p.n_convs;
  const int first_entry = blockIdx.x >> 1, entry_stride = gridDim.x >> 1;
  // k-th entry of this cluster -> (conv, item)
  auto entry = [&](int k, int &ci, int &item) -> bool {
    const int g = first_entry + k * entry_stride;
    if (g >= total_entries) return false;
    ci = g / items_per_conv;
    item = g - ci * items_per_conv;
    return true;
  };

  ptx::griddep_launch();                          // the next kernel may begin its own set-up
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&full_a[s], 1);
      ptx::mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < CH_MAX_NB; ++s) {
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty_b[s], 1);
    }
    for (int s = 0; s < 4 * NWG; ++s) ptx::mbar_init(&ufull[s], 1);
    // every unit of the item is drained by 4 warps in each CTA of the pair
    for (int s = 0; s < ACC; ++s) ptx::mbar_init(&tempty[s], 8 * UPI);
    for (int s = 0; s < NWG; ++s) ptx::mbar_init(&res_bar[s], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  if (warp == 0 && lane < p.n_convs) {
    ptx::prefetch_tmap(&convs[lane].tmA1);
    ptx::prefetch_tmap(&convs[lane].tmA2);
    ptx::prefetch_tmap(&convs[lane].tmW);
    ptx::prefetch_tmap(&convs[lane].tmR);
    ptx::prefetch_tmap(&convs[lane].tmO);
  }
  // everything above touched no data of the previous kernels; from here on their results are needed
  ptx::griddep_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                            // the peer's barriers exist before anything is signalled at them
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t flag_mul = (uint32_t)p.flag_epoch;

  // Dependency wait on a tile counter: acquire, then order the TMA (async proxy) read after it.
  auto wait_tile = [&](const unsigned *flag, uint32_t need) {
    if (ptx::ld_acquire_gpu(flag) < need) {
      const unsigned long long t0 = ptx::globaltimer_ns();
      while (ptx::ld_acquire_gpu(flag) < need) {
        __nanosleep(64);
        if (ptx::globaltimer_ns() - t0 > 4000000000ull) {      // 4 s: a broken schedule must not hang the device
          if (p.err) atomicExch(p.err, 0xC0DE0000u | (uint32_t)(blockIdx.x & 0xffff));
          __trap();
        }
      }
#ifdef DAD_TUNING
      // dependency-stall accounting (tools/chain_stalls.py): [1] = waits that had to spin, [2] = ns spent spinning
      if (p.err) {
        atomicAdd(p.err + 1, 1u);
        atomicAdd(p.err + 2, (unsigned)(ptx::globaltimer_ns() - t0));
      }
#endif
    }
#ifdef DAD_TUNING
    if (p.err) atomicAdd(p.err + 3, 1u);                        // [3] = all dependency waits
#endif
    ptx::fence_proxy_async_global();
  };

  // The producer and MMA warps stay CONVERGED: all 32 lanes walk the loops with warp-uniform values and only
  // the asynchronous instructions are issued by one elected lane (conv_t3.cuh).
  if (warp == 0) {
    // ================================ TMA producer ================================
    const bool leader_lane = ptx::elect_one();
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    // both CTAs' loads complete on the LEADER's full barriers (the MMA issuer lives there)
    const uint32_t lead_full_a = ptx::mapa(ptx::smem_u32(&full_a[0]), 0);
    const uint32_t lead_full_b = ptx::mapa(ptx::smem_u32(&full_b[0]), 0);
    for (int k = 0;; ++k) {
      int ci, item;
      if (!entry(k, ci, item)) break;
      const ChainConvMeta &cm = convs[ci].m;
      const ChainConv *cv = convs + ci;
      const int gm = item / n_tiles_n, tn = item - gm * n_tiles_n;
      const int tm = gm * 2 + (int)cta_rank;
      const int b0 = tm * p.S_t, n0 = tn * BN_ITEM;
      const int kch = cm.kch1 + cm.kch2;
      const CUtensorMap *wmap = (NS == 1 && p.w_alt) ? &cv->tmW1 : &cv->tmW;
      // the activation tile of this entry was written by an earlier conv of this launch: wait for all its units
      if (cm.flag_a != nullptr && tm < p.n_mst) {
        if (leader_lane) wait_tile(cm.flag_a + tm, (uint32_t)cm.need_a * flag_mul);
        __syncwarp();
      }
      for (int ch = 0; ch < kch; ++ch) {
        // one haloed activation box per 64-channel block, shared by all taps
        ptx::mbar_wait(&empty_a[sa], pha ^ 1);
        uint8_t *dst = smem + lay.a_ring + sa * p.a_stage_bytes;
        const CUtensorMap *am = (ch < cm.kch1) ? &cv->tmA1 : &cv->tmA2;
        const int c0 = (ch < cm.kch1 ? ch : ch - cm.kch1) * CH_BK;
        if (leader_lane) {
          if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_a[sa], 2u * (uint32_t)cm.a_tx_bytes);
          ptx::tma_load_3d_2sm(dst, am, lead_full_a + 8u * sa, c0, b0, -cm.halo_lo);
        }
        if (++sa == p.n_a_stages) { sa = 0; pha ^= 1; }
        for (int t = 0; t < cm.taps; ++t) {
          ptx::mbar_wait(&empty_b[sb], phb ^ 1);
          uint8_t *wdst = smem + lay.b_ring + sb * p.b_stage_bytes;
          const int k0 = (t * kch + ch) * CH_BK;
          if (leader_lane) {
            // this CTA keeps its half of the item's output channels; the pair MMA reads both halves
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_b[sb], 2u * (uint32_t)p.b_stage_bytes);
            ptx::tma_load_2d_2sm(wdst, wmap, lead_full_b + 8u * sb, k0, n0 + (int)cta_rank * (BN_ITEM / 2));
          }
          if (++sb == p.nb_stages) { sb = 0; phb ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    // One thread of the pair LEADER feeds both tensor cores; descriptors advance by adding constants.
    if (cta_rank == 0) {
      const bool leader_lane = ptx::elect_one();
      constexpr uint32_t idesc = ptx::make_idesc_bf16(256, BN_ITEM);
      const uint64_t dconst = ptx::make_smem_desc_sw128(0);
      const uint32_t a_lo0 = ((s_base + lay.a_ring) >> 4), a_lo_step = (uint32_t)p.a_stage_bytes >> 4;
      const uint32_t b_lo0 = ((s_base + lay.b_ring) >> 4), b_lo_step = (uint32_t)p.b_stage_bytes >> 4;
      const int n_a = p.n_a_stages, nb = p.nb_stages;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      for (int it = 0;; ++it) {
        int ci, item;
        if (!entry(it, ci, item)) break;
        const ChainConvMeta &cm = convs[ci].m;
        const int kch = cm.kch1 + cm.kch2, n_taps = cm.taps;
        const uint32_t tap_first = (uint32_t)cm.tap_first, tap_step = (uint32_t)cm.tap_step;
        uint32_t d_tmem[MH];
#pragma unroll
        for (int h = 0; h < MH; ++h) {
          const int u = it * MH + h;
          const int as = u % ACC;
          ptx::mbar_wait(&tempty[as], ((u / ACC) & 1) ^ 1);     // the epilogues have drained this accumulator
          d_tmem[h] = tmem_base + as * BN_ITEM;
        }
        ptx::tc_fence_after();
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
                // tap t = the same tile viewed tap_step further down (a multiple of the 1024 B swizzle atom)
#pragma unroll
                for (int k = 0; k < CH_BK / 16; ++k) {
                  const uint32_t acc = (k != 0) ? 1u : acc_kb;
                  ptx::umma_bf16_2sm(d_tmem[h], da + (uint64_t)(h * 1024 + 2 * k), db + (uint64_t)(2 * k), idesc, acc);
                }
              }
              ptx::umma_commit_2sm_mc(&empty_b[sb], MC_MASK);
            }
            __syncwarp();
            if (++sb == nb) { sb = 0; phb ^= 1; }
            da += tap_step;
          }
          if (leader_lane) ptx::umma_commit_2sm_mc(&empty_a[sa], MC_MASK);
          __syncwarp();
          if (++sa == n_a) { sa = 0; pha ^= 1; }
        }
        if (leader_lane) {
#pragma unroll
          for (int ns = 0; ns < UPI; ++ns) {
            const int u = it * UPI + ns, k = u / NWG;
            ptx::umma_commit_2sm_mc(&ufull[(u - k * NWG) * 4 + (k & 3)], MC_MASK);
          }
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue ====================================
    const int wg = (warp - 2) >> 2;
    const int wt = (int)threadIdx.x - 64 - wg * 128;        // thread within the warpgroup
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                  // row within a 128-row half-tile
    const int s_smp = r % p.S_t;                  // sample within the tile (position-major rows)
    const int pos_per_half = 128 / p.S_t;
    const bool elected = wt == 0;
    const float inv_n = 1.0f / (float)(p.L * GW);
    const int step = p.ls->step;                  // one timestep for the whole batch (the host routes per-row timesteps elsewhere)
    uint8_t *stg_ptr = smem + lay.stage_out + wg * STAGE_OUT;
    const uint32_t stg = s_base + lay.stage_out + wg * STAGE_OUT;            // [UC/64 boxes][128 rows][128 B], swizzled
    const uint32_t pbuf0 = s_base + lay.wparams + (uint32_t)wg * (2u * PBUF);
    const uint32_t scr_bytes = 4u * p.S_t * NG * 8u;                          // one statistics buffer: [4 warps][S_t][NG] x (sum, sumsq)
    const uint32_t scr0 = s_base + lay.scratch;                               // [warpgroup][2 buffers]
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t swz = (uint32_t)(r & 7);
    const uint32_t tempty0 = ptx::mapa(ptx::smem_u32(&tempty[0]), 0);
    uint32_t res_phase = 0;
    unsigned *pend_flag = nullptr;                // counter of this warpgroup's last store, not yet published (elected thread)

    // k-th unit of this warpgroup -> (entry ordinal, conv, item, column block of the item)
    auto unit = [&](int k, int &it, int &ci, int &item, int &ns) -> bool {
      const int u = wg + k * NWG;
      it = u / UPI;
      ns = u - it * UPI;
      return entry(it, ci, item);
    };
    auto publish_pending = [&]() {                // elected thread: all stores of this warpgroup have completed
      if (pend_flag) {
        ptx::bulk_wait0();
        ptx::fence_proxy_async_global();
        ptx::red_release_gpu_add(pend_flag, 1u);
        pend_flag = nullptr;
      }
    };
    // Residual tiles travel through the staging buffer: the box for (unit, half h) is requested as soon as the
    // previous store has finished reading the buffer.
    auto request_residual = [&](int ci, int item, int ns, int h) {
      const ChainConvMeta &cm = convs[ci].m;
      const int gm = item / n_tiles_n, tn = item - gm * n_tiles_n;
      const int tm = gm * 2 + (int)cta_rank;
      const int ch0 = tn * BN_ITEM + ns * UC;
      if (cm.flag_r != nullptr && tm < p.n_mst) {
        const uint32_t need = (uint32_t)cm.need_r * flag_mul;
        // our own unpublished store may be what the producer chain is waiting for: never spin on top of it
        if (ptx::ld_acquire_gpu(cm.flag_r + tm) < need) publish_pending();
        wait_tile(cm.flag_r + tm, need);
      }
      const CUtensorMap *rm = &convs[ci].tmR;
      ptx::mbar_arrive_expect_tx(&res_bar[wg], STAGE_OUT);
      ptx::tma_load_3d(stg_ptr, rm, &res_bar[wg], ch0, tm * p.S_t, h * pos_per_half);
      if constexpr (UC == 128) ptx::tma_load_3d(stg_ptr + 16384, rm, &res_bar[wg], ch0 + 64, tm * p.S_t, h * pos_per_half);
    };
    auto release_acc = [&](int as) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_remote(tempty0 + 8u * as);
    };
    // (gamma, beta, bias, time bias) of one column of a unit, straight from global memory (L2-resident tables)
    auto load_col = [&](int ci, int item, int ns, float &g, float &e, float &bi, float &tv) {
      const ChainConvMeta &cm = convs[ci].m;
      const int tn = item % n_tiles_n;
      const int n = tn * BN_ITEM + ns * UC + wt;
      g = 0.f; e = 0.f; tv = 0.f;
      bi = __ldg(cm.bias + n);
      if (!cm.plain) { g = __ldg(cm.gamma + n); e = __ldg(cm.beta + n); }
      if (cm.ttab) tv = __ldg(cm.ttab + (size_t)step * cm.Cout + n);
    };
    // table layout of one buffer: [UC/2] x {g0, g1, b0, b1 | bias0, bias1, t0, t1} (32 B per column pair), then [UC] biases
    auto store_col = [&](uint32_t pb, float g, float e, float bi, float tv) {
      const uint32_t a = pb + (uint32_t)(wt >> 1) * 32u + (uint32_t)(wt & 1) * 4u;
      ptx::sts32(a + 0, g);
      ptx::sts32(a + 8, e);
      ptx::sts32(a + 16, bi);
      ptx::sts32(a + 24, tv);
      ptx::sts32(pb + 16u * UC + (uint32_t)wt * 4u, bi);
    };

    {
      int it0, ci0, item0, ns0;
      if (unit(0, it0, ci0, item0, ns0)) {
        if (wt < UC) {
          float g, e, bi, tv;
          load_col(ci0, item0, ns0, g, e, bi, tv);
          store_col(pbuf0, g, e, bi, tv);
        }
        if (convs[ci0].m.has_res && elected) request_residual(ci0, item0, ns0, 0);
      }
      ptx::named_bar_sync(1 + wg, 128);            // first parameter buffer visible to the warpgroup
    }

    for (int k = 0;; ++k) {
      int it, ci, item, ns;
      if (!unit(k, it, ci, item, ns)) break;
      const ChainConvMeta &cm = convs[ci].m;
      const bool has_res = cm.has_res != 0, plain = cm.plain != 0;
      const int gm = item / n_tiles_n, tn = item - gm * n_tiles_n;
      const int tm = gm * 2 + (int)cta_rank;
      const int b0 = tm * p.S_t;
      const bool tile_ok = tm < p.n_mst;
      const int n0 = tn * BN_ITEM + ns * UC;       // first output channel of the unit
      const uint32_t col0 = (uint32_t)(ns * UC);   // its first accumulator column
      const uint32_t pb = pbuf0 + (uint32_t)(k & 1) * PBUF;
      const uint32_t pb_bias = pb + 16u * UC;
      int nit, nci = 0, nitem = 0, nns = 0;                          // this warpgroup's next unit
      const bool have_next = unit(k + 1, nit, nci, nitem, nns);
      const bool next_res = have_next && convs[nci].m.has_res != 0;
      // its parameters are fetched now and parked in registers until the statistics barrier
      float ng = 0.f, ne = 0.f, nbi = 0.f, ntv = 0.f;
      if (have_next && wt < UC) load_col(nci, nitem, nns, ng, ne, nbi, ntv);
      {
        uint64_t *bar = &ufull[wg * 4 + (k & 3)];
        const uint32_t par = (uint32_t)(k >> 2) & 1u;
        if (!ptx::mbar_try_wait(bar, par)) {
          // idle anyway: publish the completed store now (a consumer elsewhere may be waiting for exactly this tile)
          if (elected) publish_pending();
          ptx::mbar_wait(bar, par);
        }
        __syncwarp();
      }
      // TMEM address of this warp's lane quarter in accumulator half h (computed, not indexed: a [MH] array indexed
      // by the rolled h loop of pass 2 lands in local memory)
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
      auto t_addr_of = [&](int h) -> uint32_t { return t_lane + (uint32_t)(((it * MH + h) % ACC) * BN_ITEM); };
      ptx::tc_fence_after();
      if (!tile_ok) {
        // nothing to write for this tile (odd tile count): hand the accumulators back and move on
        if (has_res) { ptx::mbar_wait(&res_bar[wg], res_phase); res_phase ^= 1; }
#pragma unroll
        for (int h = 0; h < MH; ++h) release_acc((it * MH + h) % ACC);
        if (have_next && wt < UC) store_col(pbuf0 + (uint32_t)((k + 1) & 1) * PBUF, ng, ne, nbi, ntv);
        ptx::named_bar_sync(1 + wg, 128);
        if (next_res && elected) {
          ptx::bulk_wait_read0();
          request_residual(nci, nitem, nns, 0);
        }
        continue;
      }

      // statistics buffers: double-buffered by item parity when two warpgroups share a group
      const uint32_t scr_mine = scr0 + (uint32_t)(wg * 2 + (XWG ? (it & 1) : 0)) * scr_bytes;
      const uint32_t scr_peer = scr0 + (uint32_t)((wg ^ 1) * 2 + (it & 1)) * scr_bytes;      // XWG only
      if (!plain) {
        // ---- pass 1: GroupNorm statistics of (conv + bias) over the L positions x GW channels of each sample.
        // This thread's rows (one per half) belong to ONE sample; lanes with equal (lane % S_t) share it.
        f32x2 run1 = pk2(0.f, 0.f), run2 = pk2(0.f, 0.f);
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          f32x2 s1[GPC], s2[GPC];
#pragma unroll
          for (int g = 0; g < GPC; ++g) { s1[g] = pk2(0.f, 0.f); s2[g] = pk2(0.f, 0.f); }
          // narrow layers: the sums of columns 8..15 of the chunk are kept apart and added afterwards, which is the
          // association order of the 8-column chunks of conv_t3.cuh -- the two kernel families stay bit-identical
          // (tests/test_gpu_chain.py::test_fusion_levels_bit_identical)
          f32x2 u1 = pk2(0.f, 0.f), u2 = pk2(0.f, 0.f);
          const uint32_t sb = pb_bias + (uint32_t)(c * CW) * 4u;
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
            tmem_ld_cw<CW>(t_addr_of(h) + col0 + c * CW, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CW / 2; ++j) {
              const f32x2 x = fadd2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), bb[j]);
              const int g = (GW < CW) ? (2 * j) / GW : 0;       // compile-time (GW is even)
              if (SPLIT8 && j >= 4) {
                u1 = fadd2(u1, x);
                u2 = ffma2(x, x, u2);
              } else {
                s1[g] = fadd2(s1[g], x);
                s2[g] = ffma2(x, x, s2[g]);
              }
            }
          }
          if constexpr (GW >= CW) {
            run1 = fadd2(run1, s1[0]);
            run2 = fadd2(run2, s2[0]);
            if constexpr (SPLIT8) {
              run1 = fadd2(run1, u1);
              run2 = fadd2(run2, u2);
            }
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
            if (lane < p.S_t) ptx::sts64(scr_mine + (uint32_t)((q * p.S_t + lane) * NG + g0 + g) * 8u, t1, t2);
          }
        }
      }
      // the next unit's parameters go to the other buffer (its previous user, unit k-1, is long finished)
      if (have_next && wt < UC) store_col(pbuf0 + (uint32_t)((k + 1) & 1) * PBUF, ng, ne, nbi, ntv);
      // the staging buffer is free again once the previous store has read it; the elected thread checked that
      // before it prefetched this unit's residual (or checks it here when there is none)
      if (elected) {
        if (!has_res) ptx::bulk_wait_read0();
#if DAD_CH_EARLY_PUBLISH
        // the previous unit's store was issued a whole pass 1 ago: it has landed, publish it NOW rather than after this
        // unit's pass 2 -- consumers of that tile (the next conv of the chain, on another cluster) are 3-4 work-list
        // rounds behind at most, and a late counter makes their producer warps spin (tools/chain_stalls.py)
        publish_pending();
#endif
      }
      if (XWG && !plain) ptx::named_bar_sync(7, 256);      // both halves of the group have their partial sums out
      else ptx::named_bar_sync(1 + wg, 128);               // statistics exchanged, staging buffer reusable

      // ---- pass 2: normalise, Mish, (+ time bias | + residual), convert, stage, TMA store
#pragma unroll 1
      for (int h = 0; h < MH; ++h) {
        const uint32_t t_h = t_addr_of(h) + col0;
        if (has_res) { ptx::mbar_wait(&res_bar[wg], res_phase); res_phase ^= 1; }
        f32x2 rg2[GPC], nm2[GPC];      // (rstd, -mean) of the current group(s), carried across the chunks of a wide group
#pragma unroll
        for (int g = 0; g < GPC; ++g) { rg2[g] = pk2(0.f, 0.f); nm2[g] = pk2(0.f, 0.f); }
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          uint32_t v[32];
          tmem_ld_cw<CW>(t_h + c * CW, v);
          const uint32_t sp = pb + (uint32_t)((c * CW) >> 1) * 32u;
          f32x2 y[CW / 2];
          if (!plain) {
            const int g0 = (GW >= CW) ? c / CPG : c * GPC;
            if (GW < CW || c % CPG == 0) {
#pragma unroll
              for (int g = 0; g < GPC; ++g) {
                float t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  const float2 pr = ptx::lds64(scr_mine + (uint32_t)((w * p.S_t + s_smp) * NG + g0 + g) * 8u);
                  t1 += pr.x;
                  t2 += pr.y;
                }
                if constexpr (XWG) {
#pragma unroll
                  for (int w = 0; w < 4; ++w) {
                    const float2 pr = ptx::lds64(scr_peer + (uint32_t)((w * p.S_t + s_smp) * NG + g0 + g) * 8u);
                    t1 += pr.x;
                    t2 += pr.y;
                  }
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
              y[j] = fadd2(mish2(xn), tt2[j]);
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
            if (has_res) {
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
        release_acc((it * MH + h) % ACC);
        // staged tile -> global with a TMA store (rows of samples >= B land in workspace padding)
        ptx::fence_proxy_async();
        ptx::named_bar_sync(1 + wg, 128);
        const bool same_unit = (h + 1 < MH);
        if (elected) {
          ptx::tma_store_3d(&convs[ci].tmO, stg, n0, b0, h * pos_per_half);
          if constexpr (UC == 128) ptx::tma_store_3d(&convs[ci].tmO, stg + 16384, n0 + 64, b0, h * pos_per_half);
          ptx::bulk_commit();
          // the PREVIOUS store of this warpgroup was issued a whole unit ago: it has completed by now, publish it
          if (pend_flag) {
            ptx::bulk_wait1();
            ptx::fence_proxy_async_global();
            ptx::red_release_gpu_add(pend_flag, 1u);
          }
          pend_flag = cm.flag_out ? cm.flag_out + tm : nullptr;
          // the residual box of whatever uses the staging buffer next
          if (same_unit ? has_res : next_res) {
            ptx::bulk_wait_read0();
            if (same_unit) request_residual(ci, item, ns, h + 1);
            else request_residual(nci, nitem, nns, 0);
          }
        }
        // without a residual: the next pass 2 of this unit must not overwrite the buffer while it is being read
        // (across units the wait happens right before the statistics barrier)
        if (!has_res && same_unit) {
          if (elected) ptx::bulk_wait_read0();
          ptx::named_bar_sync(1 + wg, 128);
        }
      }
    }
    if (elected) {
      publish_pending();
      ptx::bulk_wait0();                          // all stores of this warpgroup have landed before the CTA exits
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                            // no CTA leaves while its peer may still signal or read it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace dad
