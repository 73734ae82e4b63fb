// fa_fwd_pair_sm100.cuh — K1P: the fused-tile forward kernel (K1, fa_fwd_sm100.cuh) on CTA PAIRS for the dense d = 128
// 16-bit case (BASELINE.json configs[1] and [3]).  Same reference functions, recurrence and TMEM layout as K1:
//   flash_attention_v1/CUDA/flash_attention_v1.h:161-248, flash_attention_v1_opt1.h:264-351,
//   flash_attention_v1_tiled_d/CUDA/flash_attention_v1.h:230-309 (d <= 128), numpy_gpu_like_opt2.py:161-195.
//
// Why pairs: K1 at full tensor rate needs 125 of the SM's 128 B/clk of shared-memory bandwidth — per KV tile and CTA it
// TMA-writes K and V (64 KB), reads Q twice (64 KB) and reads K and V once per Q tile (128 KB) in 2048 MMA cycles.  A
// 2-CTA MMA of M = 256 lets the two SMs of a pair share every K / V tile: each SM loads and reads only HALF of it (keys
// [64r, 64r+64) of K as the N half of QK^T, columns [64r, 64r+64) of V as the N half of P.V) and the tensor cores
// exchange the halves.  Per SM that is 32 + 64 + 64 = 160 KB per KV tile, 78 B/clk.
//   pair item   one head x 512 query rows: CTA r of the pair owns rows [256 (2q + r), +256) as two 128-row Q tiles
//   MMAs        issued by the leader (cluster rank 0) for both SMs: S_i = Q_i K^T (M256 N128 K16, A = each CTA's Q_i,
//               B = K halves), O_i += P_i V (A = each CTA's P_i in its own TMEM, B = V column halves)
//   barriers    q_full / kv_full / p_full / o_free live in the leader (peer TMA and softmax threads signal remotely);
//               s_full / kv_empty / q_empty / o_done are signalled in both CTAs by multicast tcgen05.commit
//   everything else (persistent item loop, ping-pong of the two Q tiles, P in TMEM over S, two 64-key P halves,
//   lazy rescale, polynomial exp2 share, staged TMA-store epilogue) is K1's.
// Served: d = 128, bf16 / fp16, dense (no causal mask, no key-padding lengths, Lq == Lk, not SPLIT); fa_api.cu routes
// everything else to K1.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fa_fwd_sm100.cuh"

namespace fa {

// D[tmem, both CTAs] (+)= A[tmem, each CTA's own 128 lanes] * B[smem, N/2 rows per CTA]
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int DT>
struct FwdPairTraits {
  static_assert(DT == DT_BF16 || DT == DT_F16, "pair kernel serves 16-bit storage");
  static constexpr int D = 128;
  static constexpr uint32_t FMT = (DT == DT_BF16) ? FMT_BF16 : FMT_F16;
  static constexpr int BM = 128, BN = 128;
  static constexpr int BLK_BYTES = 128 * 128;        // [128 rows x 128 B]
  static constexpr int HALF_BLK = 64 * 128;          // [64 rows x 128 B]
  static constexpr int TILE_BYTES = 2 * BLK_BYTES;   // a Q tile: 128 rows x 256 B
  static constexpr int STAGE_BYTES = BLK_BYTES;      // per CTA: half a K tile (2 x [64 keys x 64 d]) or half a V tile
                                                     // ([128 keys x 64 d])
  static constexpr int NS = 8;                       // K/V ring depth (stages = whole K or V tiles of the pair)
  static constexpr int STAGING_BYTES = 2 * BLK_BYTES;
  static constexpr int TM_S = 0, TM_O = 256;
  static constexpr int NUM_BARS = 2 + 2 + 2 * NS + 2 + 4 + 2 + 2;
  static constexpr int SMEM_BYTES = 1024 + 2 * TILE_BYTES + NS * STAGE_BYTES + STAGING_BYTES + NUM_BARS * 8 + 16;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static constexpr int THREADS = 384;
};

struct PairItem {
  int q_row0, bh, n_tiles;
};

// pair item -> (head, 512-row query block); q-block fastest so neighbouring pairs stream the same head's K/V out of L2
__device__ __forceinline__ PairItem decode_pair_item(int item, const FwdParams& p, uint32_t rank) {
  const int n_qblocks = (p.n_qpairs + 1) / 2;
  PairItem c;
  c.bh = item / n_qblocks;
  c.q_row0 = ((item % n_qblocks) * 2 + int(rank)) * 256;
  c.n_tiles = (p.Lk + 127) / 128;
  return c;
}

template <int DT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
fa_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FwdParams p) {
  using T = FwdPairTraits<DT>;
  constexpr int D = T::D, BM = T::BM, BN = T::BN, BLK_BYTES = T::BLK_BYTES, HALF_BLK = T::HALF_BLK;
  constexpr int TILE_BYTES = T::TILE_BYTES, STAGE_BYTES = T::STAGE_BYTES, NS = T::NS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + 2 * TILE_BYTES;
  uint8_t* sOut = sKV + NS * STAGE_BYTES;   // [2] one 128-row x 128-B staging block per softmax warpgroup
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + T::STAGING_BYTES);
  uint64_t* q_full = bars;              // [2]  leader: Q_i of both CTAs landed
  uint64_t* q_empty = q_full + 2;       // [2]  both:   every QK_i of this item retired (multicast commit)
  uint64_t* kv_full = q_empty + 2;      // [NS] leader: both halves of the stage landed
  uint64_t* kv_empty = kv_full + NS;    // [NS] both:   MMAs that read the stage retired (multicast commit)
  uint64_t* s_full = kv_empty + NS;     // [2]  both:   S_i(j) ready, and every earlier MMA retired (multicast commit)
  uint64_t* p_full = s_full + 2;        // [2][2] leader: key-half h of P_i(j) in TMEM of both CTAs (2 x 128 arrivals)
  uint64_t* o_done = p_full + 4;        // [2]  both:   last PV_i of this item retired (multicast commit)
  uint64_t* o_free = o_done + 2;        // [2]  leader: O_i read out in both CTAs (2 x 128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_pair_items = p.BH * ((p.n_qpairs + 1) / 2);
  const int first_item = blockIdx.x >> 1, item_stride = gridDim.x >> 1;

  if (warp == 9 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 256);
      mbar_init(&p_full[2 * i + 1], 256);
      mbar_init(&o_done[i], 1);
      mbar_init(&o_free[i], 256);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
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

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 8) {
      // ===================================== TMA producer (both CTAs) ======================================
      if (elect_one_sync()) {
        int tt = 0;            // K/V stages issued so far (ring position), across items
        int nq0 = 0, nq1 = 0;  // Q_i loads issued so far
        for (int item = first_item; item < n_pair_items; item += item_stride) {
          const PairItem c = decode_pair_item(item, p, rank);
          auto load_q = [&](int i, int& n) {
            if (n > 0) mbar_wait(&q_empty[i], (n - 1) & 1);
            if (rank == 0) mbar_arrive_expect_tx(&q_full[i], 2 * TILE_BYTES);
            const uint32_t bar = mapa_shared(smem_u32(&q_full[i]), 0);
#pragma unroll
            for (int b = 0; b < 2; ++b)
              tma_load_3d_pair(sQ + i * TILE_BYTES + b * BLK_BYTES, &tmQ, bar, b * 64, c.q_row0 + i * BM, c.bh);
            ++n;
          };
          auto load_kv = [&](int t) {  // t = 2j -> K_j (this CTA's 64 keys), t = 2j+1 -> V_j (this CTA's 64 columns)
            const int stage = tt % NS;
            if (tt >= NS) mbar_wait(&kv_empty[stage], ((tt / NS) - 1) & 1);
            if (rank == 0) mbar_arrive_expect_tx(&kv_full[stage], 2 * STAGE_BYTES);
            const uint32_t bar = mapa_shared(smem_u32(&kv_full[stage]), 0);
            uint8_t* dst = sKV + stage * STAGE_BYTES;
            const int row = (t >> 1) * BN;
            if (t & 1) {
              tma_load_3d_pair(dst, &tmV, bar, int(rank) * 64, row, c.bh);
            } else {
              tma_load_3d_pair(dst, &tmK, bar, 0, row + int(rank) * 64, c.bh);
              tma_load_3d_pair(dst + HALF_BLK, &tmK, bar, 64, row + int(rank) * 64, c.bh);
            }
            ++tt;
          };
          load_q(0, nq0);
          load_kv(0);
          load_q(1, nq1);
          for (int t = 1; t < 2 * c.n_tiles; ++t) load_kv(t);
        }
      }
    } else if (warp == 9) {
      // ===================================== MMA issuer (leader CTA only) ==================================
      if (rank == 0 && elect_one_sync()) {
        constexpr uint32_t idesc_qk = make_idesc(T::FMT, 256, BN, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc(T::FMT, 256, D, 0, 1);
        constexpr uint64_t hiK = make_smem_desc_hi(16, 1024, SWZ_128B);         // K-major, 8-row swizzle atoms
        constexpr uint64_t hiV = make_smem_desc_hi(BLK_BYTES, 1024, SWZ_128B);  // MN-major: 64 columns x 8-key atoms
        const uint32_t sQ_addr = smem_u32(sQ), sKV_addr = smem_u32(sKV);

        auto qk = [&](int i, int stage) {  // S_i = Q_i K^T for both CTAs
          const uint32_t a_base = sQ_addr + i * TILE_BYTES, b_base = sKV_addr + stage * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ss_pair(tmem_base + T::TM_S + i * BN, make_smem_desc(a_base + (k >> 2) * BLK_BYTES + (k & 3) * 32, hiK),
                         make_smem_desc(b_base + (k >> 2) * HALF_BLK + (k & 3) * 32, hiK), idesc_qk, k > 0 ? 1u : 0u);
        };
        auto pv = [&](int i, int stage, uint32_t acc, int kk0, int kk1) {  // O_i (+)= P_i V  (K-steps kk0..kk1-1)
          const uint32_t b_base = sKV_addr + stage * STAGE_BYTES;
#pragma unroll
          for (int kk = kk0; kk < kk1; ++kk)
            umma_ts_pair(tmem_base + T::TM_O + i * D, tmem_base + T::TM_S + i * BN + kk * 8,
                         make_smem_desc(b_base + kk * 16 * 128, hiV), idesc_pv, (acc | (kk > 0)) ? 1u : 0u);
        };

        constexpr int KT = BN / 16;
        int tt = 0;            // K/V stages consumed so far (ring position), across items
        int nt[2] = {0, 0};    // KV tiles processed for Q tile i (phase of s_full / p_full), across items
        int ni = 0;            // items processed (phase of q_full / o_done / o_free)
        for (int item = first_item; item < n_pair_items; item += item_stride) {
          const PairItem c = decode_pair_item(item, p, rank);
          const int t0 = tt;   // ring index of K_0 of this item
          mbar_wait(&kv_full[t0 % NS], (t0 / NS) & 1);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            mbar_wait(&q_full[i], ni & 1);
            tc_fence_after();
            qk(i, t0 % NS);
            tc_commit_pair(&s_full[i], 3);
            if (c.n_tiles == 1) tc_commit_pair(&q_empty[i], 3);
          }
          tc_commit_pair(&kv_empty[t0 % NS], 3);  // K_0: both QK(0) are issued by now
          for (int j = 0; j < c.n_tiles; ++j) {
            const int tv = t0 + 2 * j + 1, tk = t0 + 2 * j + 2;
            bool k_waited = false;
            mbar_wait(&kv_full[tv % NS], (tv / NS) & 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (j == 0 && ni > 0) {
                // PV_i(0) overwrites O_i: the previous item's epilogue must have read it out of TMEM in both CTAs
                mbar_wait_cluster(&o_free[i], (ni - 1) & 1);
              }
              mbar_wait_cluster(&p_full[2 * i], nt[i] & 1);
              tc_fence_after();
              pv(i, tv % NS, j > 0 ? 1u : 0u, 0, KT / 2);
              mbar_wait_cluster(&p_full[2 * i + 1], nt[i] & 1);
              tc_fence_after();
              pv(i, tv % NS, 1u, KT / 2, KT);
              ++nt[i];
              if (j + 1 < c.n_tiles) {
                if (!k_waited) {
                  mbar_wait(&kv_full[tk % NS], (tk / NS) & 1);
                  tc_fence_after();
                  k_waited = true;
                }
                qk(i, tk % NS);
                tc_commit_pair(&s_full[i], 3);
                if (j + 2 == c.n_tiles) tc_commit_pair(&q_empty[i], 3);  // that was the last QK_i of this item
              } else {
                tc_commit_pair(&o_done[i], 3);
              }
            }
            tc_commit_pair(&kv_empty[tv % NS], 3);
            if (j + 1 < c.n_tiles) tc_commit_pair(&kv_empty[tk % NS], 3);
          }
          tt = t0 + 2 * c.n_tiles;
          ++ni;
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ===================================== softmax warpgroups (both CTAs) =================================
    const int i = warp >> 2;  // which Q tile
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t tS = t_lane + T::TM_S + i * BN;
    const uint32_t tO = t_lane + T::TM_O + i * D;
    uint8_t* sO = sOut + i * BLK_BYTES;
    const uint32_t sO_addr = smem_u32(sO);
    const bool storer = ((warp & 3) == 0) && (lane == 0);
    const uint32_t p_full_ld = mapa_shared(smem_u32(&p_full[2 * i]), 0);   // leader's p_full[2i]; [2i+1] is 8 bytes on
    const uint32_t o_free_ld = mapa_shared(smem_u32(&o_free[i]), 0);
    int nt = 0;  // KV tiles processed (phase of s_full / p_full), across items
    int ni = 0;  // items processed (phase of o_done)

    for (int item = first_item; item < n_pair_items; item += item_stride) {
      const PairItem c = decode_pair_item(item, p, rank);
      float m_used = -CUDART_INF_F;
      float l = 0.f;

      for (int j = 0; j < c.n_tiles; ++j, ++nt) {
        mbar_wait(&s_full[i], nt & 1);
        tc_fence_after();
        uint32_t s[4][32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld32(tS + cc * 32, s[cc]);
        tc_wait_ld();

        // Row max; only the last tile of the key range can be ragged (its masking lives in its own branch).
        const int valid = p.Lk - j * BN;
        float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
        if (valid >= BN) {
#pragma unroll
          for (int x = 0; x < 32; ++x) {
            mx0 = fmaxf(mx0, __uint_as_float(s[0][x]));
            mx1 = fmaxf(mx1, __uint_as_float(s[1][x]));
            mx2 = fmaxf(mx2, __uint_as_float(s[2][x]));
            mx3 = fmaxf(mx3, __uint_as_float(s[3][x]));
          }
        } else {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int x = 0; x < 32; ++x)
              if (cc * 32 + x >= valid) s[cc][x] = __float_as_uint(-CUDART_INF_F);
#pragma unroll
          for (int x = 0; x < 32; ++x) {
            mx0 = fmaxf(mx0, __uint_as_float(s[0][x]));
            mx1 = fmaxf(mx1, __uint_as_float(s[1][x]));
            mx2 = fmaxf(mx2, __uint_as_float(s[2][x]));
            mx3 = fmaxf(mx3, __uint_as_float(s[3][x]));
          }
          asm volatile("" ::: "memory");  // keep this a real branch
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

        if (j == 0) {
          m_used = mx;
        } else {
          // Lazy rescale: keep the stale max unless the new one is > 2^8 larger (in exp2 units).  PV_i(j-1) has
          // retired (s_full[i](j) implies it: same issuing thread, in-order pipe), so O_i is quiescent.
          const bool need = (mx - m_used) * p.scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? ex2_approx((m_used - mx) * p.scale_log2) : 1.0f;
            if (need) m_used = mx;
            l *= alpha;
#pragma unroll
            for (int cc = 0; cc < D / 32; ++cc) {
              uint32_t o[32];
              tmem_ld32(tO + cc * 32, o);
              tc_wait_ld();
#pragma unroll
              for (int x = 0; x < 32; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * alpha);
              tmem_st32(tO + cc * 32, o);
            }
          }
        }

        const float neg_m = -m_used * p.scale_log2;
        float2 lsum[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
        auto exp_blocks = [&](int c0, int c1) {
#pragma unroll
          for (int x = 0; x < 32; x += 2) {
#pragma unroll
            for (int cc = c0; cc < c1; ++cc) {
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
        };
        auto store_p = [&](int c0) {  // P columns 32*c0 .. 32*c0+63 -> TMEM as packed 16-bit pairs, in place over S
          uint32_t pk[32];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int x = 0; x < 16; ++x) {
              const float a = __uint_as_float(s[c0 + h][2 * x]), b = __uint_as_float(s[c0 + h][2 * x + 1]);
              pk[h * 16 + x] = (DT == DT_BF16) ? pack_bf16x2(a, b) : pack_f16x2(a, b);
            }
          tmem_st32(tS + (c0 / 2) * 32, pk);
        };
        exp_blocks(0, 2);
        store_p(0);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive_cluster(p_full_ld);
        exp_blocks(2, 4);
        store_p(2);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive_cluster(p_full_ld + 8);
        l += ((lsum[0].x + lsum[0].y) + (lsum[1].x + lsum[1].y)) + ((lsum[2].x + lsum[2].y) + (lsum[3].x + lsum[3].y));
      }

      // ------------------------------- epilogue: O_i / l -------------------------------------
      mbar_wait(&o_done[i], ni & 1);
      ++ni;
      tc_fence_after();
      uint32_t o[D / 32][32];
#pragma unroll
      for (int cc = 0; cc < D / 32; ++cc) tmem_ld32(tO + cc * 32, o[cc]);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive_cluster(o_free_ld);  // O_i is in registers: the leader may start the next item's PV_i
      const float inv_l = 1.0f / l;
      const int row0 = c.q_row0 + i * BM;
      if (p.lse_out != nullptr && row0 + row < p.L) p.lse_out[size_t(c.bh) * p.L + row0 + row] = m_used * p.scale + __logf(l);
      // One 128-byte column block at a time through this warpgroup's staging block (128B swizzle), TMA store.
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (storer) tma_store_wait_read_all();  // the previous store has finished reading the staging block
        named_bar_sync(1 + i, 128);
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // 16-byte chunks of one block row
          const int e = b * 64 + u * 8;
          auto pk2 = [&](int ee) {
            const float a0 = __uint_as_float(o[ee / 32][ee % 32]) * inv_l;
            const float a1 = __uint_as_float(o[ee / 32][ee % 32 + 1]) * inv_l;
            return (DT == DT_BF16) ? pack_bf16x2(a0, a1) : pack_f16x2(a0, a1);
          };
          uint4 v;
          v.x = pk2(e + 0);
          v.y = pk2(e + 2);
          v.z = pk2(e + 4);
          v.w = pk2(e + 6);
          st_shared_v4(sO_addr + row * 128 + ((u ^ (row & 7)) << 4), v);
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + i, 128);
        if (storer && row0 < p.L) {   // a tile past the last query row (odd tile count) stores nothing
          tma_store_3d(&tmO, sO, b * 64, row0, c.bh);
          tma_store_commit();
        }
      }
    }
    if (storer) tma_store_wait_read_all();  // smem must outlive the last store's read
  }

  // Neither CTA may exit (or free TMEM) while the other can still signal its barriers or read its shared memory.
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 8) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace fa
