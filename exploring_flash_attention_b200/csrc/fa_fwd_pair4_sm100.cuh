// fa_fwd_pair4_sm100.cuh — K1R: K1Q (fa_fwd_pair2_sm100.cuh) with FOUR softmax warpgroups: the two KV tiles in flight
// (S / P double-buffered, one 128-row Q tile per CTA of a pair) are each split by key halves between two warpgroups, so
// every scheduler holds four softmax warps instead of two.
//
// UNVERIFIED: written after round 1's GPU budget was spent; it compiles for sm_100a but has not run on a GPU.  It is
// reachable only with FA_B200_FWD_PAIR=3, is not part of any test, and first light is the first item of round 2.
//
// Why: K1 and K1Q both need ~1650 cycles per 128 x 128 score tile per SM against 1024 MMA cycles although K1Q keeps the
// tensor pipe fed; ncu shows the 8 softmax warps (2 per scheduler) issuing 33 % of the time, stalled on their own
// dependent TMEM-load -> max -> exp -> pack -> TMEM-store streams.  Same reference functions and recurrence
// as K1: flash_attention_v1/CUDA/flash_attention_v1.h:161-248, numpy_gpu_like_opt2.py:161-195.
//   warpgroup wg   buffer b = wg >> 1 (KV tiles G with G & 1 == b), key half h = wg & 1 (keys [64h, 64h+64) of the tile);
//                  thread <-> row (TMEM lane = (warp & 3) * 32 + lane), 64 score values per thread
//   per tile       the two halves exchange their half-row maxima through shared memory (one 256-thread named barrier);
//                  the running max crosses from the owners of tile j to the owners of tile j+1 as in K1Q (read BEFORE
//                  the exchange barrier, so the publisher of tile j+2 cannot overwrite it early); both halves take the same
//                  rescale decision; each rescales its 64 columns of O; four partial row sums are added in the epilogue
//   epilogue       warpgroup wg normalises O columns [32 wg, 32 wg + 32) into staging block wg >> 1; one TMA store per block
//   registers      640 threads x 96, no setmaxnreg: 64 score values per softmax thread fit, and the MMA / TMA warps need ~72
// TMEM, MMA order, ring and barriers are K1Q's (p_full counts 2 x 256 arrivals, o_free 2 x 512).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fa_fwd_pair2_sm100.cuh"

namespace fa {

template <int DT>
struct FwdPair4Traits {
  static_assert(DT == DT_BF16 || DT == DT_F16, "pair kernel serves 16-bit storage");
  static constexpr int D = 128;
  static constexpr uint32_t FMT = (DT == DT_BF16) ? FMT_BF16 : FMT_F16;
  static constexpr int BM = 128, BN = 128;
  static constexpr int BLK_BYTES = 128 * 128;        // [128 rows x 128 B]
  static constexpr int HALF_BLK = 64 * 128;          // [64 rows x 128 B]
  static constexpr int Q_BYTES = 2 * BLK_BYTES;      // 128 rows x 256 B
  static constexpr int STAGE_BYTES = BLK_BYTES;      // per CTA: 2 x [64 keys x 64 d] of K, or [128 keys x 64 d] of V
  static constexpr int NS = 9;
  static constexpr int STAGING_BYTES = 2 * BLK_BYTES;
  static constexpr int TM_S = 0, TM_O = 256, TM_P = 384;
  static constexpr int NUM_BARS = 1 + 1 + 2 * NS + 2 + 2 + 2 + 1 + 1 + 2;
  static constexpr int XCHG_BYTES = (2 + 4 + 8) * 128 * 4;   // hand_m[2][128], l_buf[4][128], mx_buf[2][2][2][128]
  static constexpr int SMEM_BYTES = 1024 + Q_BYTES + NS * STAGE_BYTES + STAGING_BYTES + NUM_BARS * 8 + 16 + XCHG_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static constexpr int THREADS = 640;
};

template <int DT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1)
fa_fwd_pair4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FwdParams p) {
  using T = FwdPair4Traits<DT>;
  constexpr int D = T::D, BN = T::BN, BLK_BYTES = T::BLK_BYTES, HALF_BLK = T::HALF_BLK;
  constexpr int STAGE_BYTES = T::STAGE_BYTES, NS = T::NS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + T::Q_BYTES;
  uint8_t* sOut = sKV + NS * STAGE_BYTES;   // [2] one [128 rows x 128 B] staging block per softmax warpgroup
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + T::STAGING_BYTES);
  uint64_t* q_full = bars;              // [1]  leader: Q of both CTAs landed
  uint64_t* q_empty = q_full + 1;       // [1]  both:   every QK of this item retired (multicast commit)
  uint64_t* kv_full = q_empty + 1;      // [NS] leader: both halves of the stage landed
  uint64_t* kv_empty = kv_full + NS;    // [NS] both:   MMAs that read the stage retired (multicast commit)
  uint64_t* s_full = kv_empty + NS;     // [2]  both:   S[b] holds a new tile, every earlier MMA retired (multicast commit)
  uint64_t* p_full = s_full + 2;        // [2]  leader: P[b] written (and O rescaled) in both CTAs (2 x 256 arrivals)
  uint64_t* pv_done = p_full + 2;       // [2]  both:   the PV that read P[b] retired (multicast commit)
  uint64_t* o_done = pv_done + 2;       // [1]  both:   last PV of this item retired (multicast commit)
  uint64_t* o_free = o_done + 1;        // [1]  leader: O read out in both CTAs (2 x 512 arrivals)
  uint64_t* hand = o_free + 1;          // [2]  local:  half 0 of buffer b published the running max after its tile (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hand + 2);
  float* hand_m = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + T::NUM_BARS * 8 + 16);  // [2][128]
  float* l_buf = hand_m + 256;                                                                          // [4][128]
  float* mx_buf = l_buf + 512;   // [2 buffers][2 use parities][2 halves][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_pair_items = p.BH * ((p.L + 255) / 256);
  const int first_item = blockIdx.x >> 1, item_stride = gridDim.x >> 1;

  if (warp == 17 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_done, 1);
    mbar_init(o_free, 1024);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 512);
      mbar_init(&pv_done[b], 1);
      mbar_init(&hand[b], 128);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 16) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmO);
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 16) {
    if (warp == 16) {
      // ===================================== TMA producer (both CTAs) ======================================
      if (elect_one_sync()) {
        int tt = 0;   // K/V stages issued so far (ring position), across items
        int nq = 0;   // Q loads issued so far
        for (int item = first_item; item < n_pair_items; item += item_stride) {
          const Pair2Item c = decode_pair2_item(item, p, rank);
          auto load_kv = [&](bool is_v, int j) {   // K_j: this CTA's 64 keys; V_j: this CTA's 64 columns
            const int stage = tt % NS;
            if (tt >= NS) mbar_wait(&kv_empty[stage], ((tt / NS) - 1) & 1);
            if (rank == 0) mbar_arrive_expect_tx(&kv_full[stage], 2 * STAGE_BYTES);
            const uint32_t bar = mapa_shared(smem_u32(&kv_full[stage]), 0);
            uint8_t* dst = sKV + stage * STAGE_BYTES;
            if (is_v) {
              tma_load_3d_pair(dst, &tmV, bar, int(rank) * 64, j * BN, c.bh);
            } else {
              tma_load_3d_pair(dst, &tmK, bar, 0, j * BN + int(rank) * 64, c.bh);
              tma_load_3d_pair(dst + HALF_BLK, &tmK, bar, 64, j * BN + int(rank) * 64, c.bh);
            }
            ++tt;
          };
          if (nq > 0) mbar_wait(q_empty, (nq - 1) & 1);
          if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * T::Q_BYTES);
          {
            const uint32_t bar = mapa_shared(smem_u32(q_full), 0);
            tma_load_3d_pair(sQ, &tmQ, bar, 0, c.q_row0, c.bh);
            tma_load_3d_pair(sQ + BLK_BYTES, &tmQ, bar, 64, c.q_row0, c.bh);
          }
          ++nq;
          // consumption order: K0 K1 | V0 K2 | V1 K3 | ... | V(n-2) | V(n-1)
          load_kv(false, 0);
          if (c.n_tiles > 1) load_kv(false, 1);
          for (int j = 0; j < c.n_tiles; ++j) {
            load_kv(true, j);
            if (j + 2 < c.n_tiles) load_kv(false, j + 2);
          }
        }
      }
    } else if (warp == 17) {
      // ===================================== MMA issuer (leader CTA only) ==================================
      if (rank == 0 && elect_one_sync()) {
        constexpr uint32_t idesc_qk = make_idesc(T::FMT, 256, BN, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc(T::FMT, 256, D, 0, 1);
        constexpr uint64_t hiK = make_smem_desc_hi(16, 1024, SWZ_128B);
        constexpr uint64_t hiV = make_smem_desc_hi(BLK_BYTES, 1024, SWZ_128B);
        const uint32_t sQ_addr = smem_u32(sQ), sKV_addr = smem_u32(sKV);
        int tt = 0;   // K/V stages consumed so far
        int g = 0;    // KV tiles started so far, across items: tile G uses S[G & 1], P[G & 1], phase (G >> 1) & 1
        int ni = 0;   // items processed (phase of q_full / o_done / o_free)
        auto qk = [&](int G) {   // S[G & 1] = Q K^T for both CTAs, from the next ring stage
          const int stage = tt % NS;
          mbar_wait(&kv_full[stage], (tt / NS) & 1);
          tc_fence_after();
          const uint32_t b_base = sKV_addr + stage * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ss_pair(tmem_base + T::TM_S + (G & 1) * BN, make_smem_desc(sQ_addr + (k >> 2) * BLK_BYTES + (k & 3) * 32, hiK),
                         make_smem_desc(b_base + (k >> 2) * HALF_BLK + (k & 3) * 32, hiK), idesc_qk, k > 0 ? 1u : 0u);
          tc_commit_pair(&s_full[G & 1], 3);
          tc_commit_pair(&kv_empty[stage], 3);
          ++tt;
        };
        auto pv = [&](int G, uint32_t acc) {   // O (+)= P[G & 1] V, from the next ring stage
          const int stage = tt % NS;
          mbar_wait(&kv_full[stage], (tt / NS) & 1);
          tc_fence_after();
          const uint32_t b_base = sKV_addr + stage * STAGE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk)
            umma_ts_pair(tmem_base + T::TM_O, tmem_base + T::TM_P + (G & 1) * 64 + kk * 8,
                         make_smem_desc(b_base + kk * 16 * 128, hiV), idesc_pv, (acc | (kk > 0)) ? 1u : 0u);
          tc_commit_pair(&pv_done[G & 1], 3);
          tc_commit_pair(&kv_empty[stage], 3);
          ++tt;
        };
        for (int item = first_item; item < n_pair_items; item += item_stride) {
          const Pair2Item c = decode_pair2_item(item, p, rank);
          const int n = c.n_tiles;
          mbar_wait(q_full, ni & 1);
          tc_fence_after();
          // S[G & 1] last held tile G - 2, which both CTAs' softmax had read when p_full(G - 2) completed (waited below)
          qk(g);
          if (n > 1) qk(g + 1);
          if (n <= 2) tc_commit_pair(q_empty, 3);
          for (int j = 0; j < n; ++j) {
            const int G = g + j;
            if (j == 0 && ni > 0) mbar_wait(o_free, (ni - 1) & 1);   // the previous item's O has been read out
            mbar_wait(&p_full[G & 1], (G >> 1) & 1);
            tc_fence_after();
            pv(G, j > 0 ? 1u : 0u);
            if (j + 1 == n) tc_commit_pair(o_done, 3);
            if (j + 2 < n) {
              qk(G + 2);
              if (j + 3 == n) tc_commit_pair(q_empty, 3);   // that was the last QK of this item
            }
          }
          g += n;
          ++ni;
        }
      }
    }
  } else {
    // ===================================== softmax warpgroups (both CTAs) =================================
    const int wg = warp >> 2;  // 0..3
    const int b = wg >> 1;     // S / P buffer: this warpgroup's KV tiles are those with G & 1 == b
    const int h = wg & 1;      // key half of the tile, also which 64 columns of O it rescales
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t tS = t_lane + T::TM_S + b * BN + h * 64;
    const uint32_t tP = t_lane + T::TM_P + b * 64 + h * 32;
    const uint32_t tO = t_lane + T::TM_O;
    uint8_t* sO = sOut + b * BLK_BYTES;          // staging block of output columns [64 b, 64 b + 64)
    const uint32_t sO_addr = smem_u32(sO);
    const bool storer = (h == 0) && ((warp & 3) == 0) && (lane == 0);
    const uint32_t p_full_ld = mapa_shared(smem_u32(&p_full[b]), 0);
    const uint32_t o_free_ld = mapa_shared(smem_u32(o_free), 0);
    int g = 0;        // KV tiles started so far, across items (same count as the MMA issuer's)
    int ni = 0;       // items processed
    int n_mine = 0;   // tiles this warpgroup has processed (parity of its mx_buf slot)
    int n_got = 0;    // hand-overs of the other buffer's owners this warpgroup has consumed (phase of hand[1 - b])

    for (int item = first_item; item < n_pair_items; item += item_stride) {
      const Pair2Item c = decode_pair2_item(item, p, rank);
      const int n = c.n_tiles;
      float m_seen = -CUDART_INF_F;   // the running max this warpgroup's partial row sum is scaled to
      float l = 0.f;

      for (int j = 0; j < n; ++j) {
        const int G = g + j;
        if ((G & 1) != b) continue;   // the other buffer's tile
        mbar_wait(&s_full[b], (G >> 1) & 1);   // also proves PV(G - 2) retired: P[b] may be overwritten
        tc_fence_after();
        uint32_t s[2][32];
        tmem_ld32(tS, s[0]);
        tmem_ld32(tS + 32, s[1]);
        tc_wait_ld();

        const int valid = p.Lk - (j * BN + h * 64);   // keys of this half that exist
        if (valid < 64) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int x = 0; x < 32; ++x)
              if (cc * 32 + x >= valid) s[cc][x] = __float_as_uint(-CUDART_INF_F);
        }
        float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          mx0 = fmaxf(mx0, __uint_as_float(s[0][x]));
          mx1 = fmaxf(mx1, __uint_as_float(s[1][x]));
        }
        // Running max after tile j-1, published by half 0 of the other buffer's owners.  Read it BEFORE the exchange
        // barrier below: the next publish into that slot (tile j+1) needs this tile's publish, which needs both halves
        // past the barrier.
        float m_prev = -CUDART_INF_F;
        if (j > 0) {
          mbar_wait(&hand[1 - b], n_got & 1);
          ++n_got;
          m_prev = hand_m[(1 - b) * 128 + row];
        }
        // half-row maxima of the two warpgroups of this tile (key 0 of a tile always exists: the joint max is finite)
        float* mxs = mx_buf + ((b * 2 + (n_mine & 1)) * 2) * 128;
        ++n_mine;
        mxs[h * 128 + row] = fmaxf(mx0, mx1);
        named_bar_sync(1 + b, 256);
        const float mx = fmaxf(fmaxf(mx0, mx1), mxs[(1 - h) * 128 + row]);

        // Lazy rescale: keep the stale max unless the new one is > 2^8 larger (in exp2 units); same decision in both halves.
        const bool need = (j == 0) || ((mx - m_prev) * p.scale_log2 > kRescaleThreshold);
        const float m_new = need ? mx : m_prev;
        if (h == 0) {
          hand_m[b * 128 + row] = m_new;
          mbar_arrive(&hand[b]);   // the owners of tile j+1 (or the epilogue) may go on
        }
        if (m_new != m_seen) {     // bring this warpgroup's partial row sum to the new scale (0 on first use)
          l *= ex2_approx((m_seen - m_new) * p.scale_log2);
          m_seen = m_new;
        }
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          // O must be quiescent: PV(G - 1) used the other buffer and was issued after QK(G); wait for its own commit.
          mbar_wait(&pv_done[1 - b], ((G - 1) >> 1) & 1);
          tc_fence_after();
          const float alpha = need ? ex2_approx((m_prev - m_new) * p.scale_log2) : 1.0f;
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {   // this half's 64 columns of O, 16 at a time (registers are tight here)
            uint32_t o[16];
            tmem_ld16(tO + h * 64 + cc * 16, o);
            tc_wait_ld();
#pragma unroll
            for (int x = 0; x < 16; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * alpha);
            tmem_st16(tO + h * 64 + cc * 16, o);
          }
        }

        const float neg_m = -m_new * p.scale_log2;
        float2 lsum[2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            float2 v = make_float2(__uint_as_float(s[cc][x]), __uint_as_float(s[cc][x + 1]));
            v = __ffma2_rn(v, make_float2(p.scale_log2, p.scale_log2), make_float2(neg_m, neg_m));
            if (FA_POLY_MOD > 0 && ((x >> 1) % (FA_POLY_MOD > 0 ? FA_POLY_MOD : 1)) == FA_POLY_MOD - 1) {
              v = exp2_poly2(v);
            } else {
              v.x = ex2_approx(v.x);
              v.y = ex2_approx(v.y);
            }
            lsum[cc] = __fadd2_rn(lsum[cc], v);
            s[cc][x] = __float_as_uint(v.x);
            s[cc][x + 1] = __float_as_uint(v.y);
          }
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {   // this half's 64 keys -> P[b] columns [32 h, 32 h + 32) as packed 16-bit pairs
          uint32_t pk[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const float a0 = __uint_as_float(s[cc][2 * x]), a1 = __uint_as_float(s[cc][2 * x + 1]);
            pk[x] = (DT == DT_BF16) ? pack_bf16x2(a0, a1) : pack_f16x2(a0, a1);
          }
          tmem_st16(tP + cc * 16, pk);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive_cluster(p_full_ld);
        l += (lsum[0].x + lsum[0].y) + (lsum[1].x + lsum[1].y);
      }

      // ------------------------------- epilogue: (O / l), 32 columns per warpgroup ----------------------------
      // The warpgroups that did not own the last tile still have to take its hand-over (final running max).
      if (((g + n - 1) & 1) != b) {
        mbar_wait(&hand[1 - b], n_got & 1);
        ++n_got;
        const float m_fin = hand_m[(1 - b) * 128 + row];
        if (m_fin != m_seen) {
          l *= ex2_approx((m_seen - m_fin) * p.scale_log2);
          m_seen = m_fin;
        }
      }
      l_buf[wg * 128 + row] = l;
      named_bar_sync(3, 512);
      const float l_tot = (l_buf[row] + l_buf[128 + row]) + (l_buf[256 + row] + l_buf[384 + row]);
      mbar_wait(o_done, ni & 1);
      ++ni;
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(tO + wg * 32, o);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive_cluster(o_free_ld);   // this quarter of O is in registers
      const float inv_l = 1.0f / l_tot;
      if (wg == 0 && p.lse_out != nullptr && c.q_row0 + row < p.L)
        p.lse_out[size_t(c.bh) * p.L + c.q_row0 + row] = m_seen * p.scale + __logf(l_tot);
      if (storer) tma_store_wait_read_all();   // the previous item's store has finished reading the staging block
      named_bar_sync(4 + b, 256);
#pragma unroll
      for (int u = 0; u < 4; ++u) {   // this warpgroup's four 16-byte chunks of the 128-byte block row
        auto pk2 = [&](int ee) {
          const float a0 = __uint_as_float(o[ee]) * inv_l, a1 = __uint_as_float(o[ee + 1]) * inv_l;
          return (DT == DT_BF16) ? pack_bf16x2(a0, a1) : pack_f16x2(a0, a1);
        };
        uint4 v;
        v.x = pk2(8 * u + 0);
        v.y = pk2(8 * u + 2);
        v.z = pk2(8 * u + 4);
        v.w = pk2(8 * u + 6);
        const int q = h * 4 + u;      // chunk within the block row: columns 64 b + 32 h + 8 u
        st_shared_v4(sO_addr + row * 128 + ((q ^ (row & 7)) << 4), v);
      }
      fence_proxy_async_smem();
      named_bar_sync(4 + b, 256);
      if (storer && c.q_row0 < p.L) {
        tma_store_3d(&tmO, sO, b * 64, c.q_row0, c.bh);
        tma_store_commit();
      }
      g += n;
    }
    if (storer) tma_store_wait_read_all();
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 16) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace fa
