// fa_tiled_d_pair_sm100.cuh — K2P: tiled-d flash-attention forward for 16-bit head dims 256 and 512 on a CTA PAIR
// (cluster of 2, tcgen05 cta_group::2).  Same contract as K2 (fa_tiled_d_sm100.cuh) and the same reference functions:
//   flash_attention_v1_tiled_d/CUDA/flash_attention_v1.h:230-309      flash_attention_kernel
//   flash_attention_v1_tiled_d/CUDA/flash_attention_v1_opt.h:366-445  flash_attention_kernel_opt (WMMA)
// with the QK^T d-chunk loop (:154-178) and the S.V d-slab loop (:209-226).
//
// Why a pair: at d = 512 one SM's TMEM (512 columns) holds either a 128-row O or S, not both, so K2 gives each CTA a
// 256-wide slab of O and both slabs recompute S (1.5x the algorithmic MMA work).  Here the two SMs of a pair share one
// 128-row Q tile BY ROWS: with a 2-CTA MMA of M = 128 each SM computes 64 rows at full tensor rate, and its 64 x N
// block of D is spread over all 128 TMEM lanes with N/2 columns (lane r: columns [0,N/2), lane 64+r: [N/2,N)).
//   TMEM per CTA   S[b] [64b, 64b+64) for b < NB,  O [64 NB, 64 NB + D/2)     (d = 512, NB = 4: all 512 columns)
//   MMAs (leader)  S = Q K^T : M128 N128 K16, A = Q rows of each CTA, B = K tile, keys [0,64) from the leader's smem and
//                  [64,128) from the peer's;   O[:, 128g .. 128g+127] += P V : A = P (smem, written by each CTA's own
//                  softmax warps), B = V columns 128g + 64*rank + [0,64) from each CTA.  No S is computed twice and
//                  every K / V byte is loaded by exactly one of the two SMs.
//   smem per CTA   Q 64 rows x D (resident), P NB x [64 rows x 128 keys], ring of 16 KB stages: a K stage = two d-chunks
//                  [64 keys x 64 d], a V stage = [128 keys x 64 d].
//   warps          0-3 softmax (thread t <-> TMEM lane t: row t & 63, key half t >> 6; the two threads of a row
//                  exchange their half-row max through smem), 4 TMA producer (both CTAs), 5 MMA issuer (leader only).
//   barriers       full[] / q_full / p_full[] live in the leader (the peer's TMA and softmax threads signal them
//                  remotely); empty[] / s_full[] / pv_done[] are signalled in both CTAs by multicast tcgen05.commit.
//   schedule       S and P are NB-deep (TMEM and smem have the room), and the leader issues QK(j+NB-1) BEFORE it waits
//                  for P(j): the tensor pipe always has a score tile queued while the softmax of the oldest tile and
//                  the cross-SM barrier hops complete.  Issue order  QK(0..NB-2) | QK(j+NB-1) PV(j) | ...
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fa_fwd_sm100.cuh"

#ifndef FA_PAIR_NB
#define FA_PAIR_NB 4            // depth of the S (TMEM) and P (smem) buffers: 2, 3 or 4 (B200, B16 H8 L4096 d512: 1005 /
                                // 1010 / 1036 TFLOP/s)
#endif
#ifndef FA_PAIR_WARP_ARRIVE
#define FA_PAIR_WARP_ARRIVE 0   // 1: one p_full arrival per softmax warp instead of one per thread
#endif
#ifndef FA_PAIR_PROBE
#define FA_PAIR_PROBE 0         // timing probes that BREAK the numerics: 1 no half-row max exchange, 2 no exponentials,
#endif                          // 4 no P stores (never ship non-zero)
#ifndef FA_PAIR_FULL_FENCE
#define FA_PAIR_FULL_FENCE 0    // 1: fence.proxy.async over every state space instead of shared::cta
#endif

namespace fa {

template <int D, int DT>
struct TiledDPairTraits {
  static_assert(DT != DT_F32 && (D == 256 || D == 512), "pair kernel serves 16-bit d = 256, 512");
  static constexpr uint32_t FMT = (DT == DT_BF16) ? FMT_BF16 : FMT_F16;
  static constexpr int BM = 128;                    // query rows per pair
  static constexpr int BMC = 64;                    // query rows per CTA
  static constexpr int BN = 128;                    // keys per KV tile (64 loaded by each CTA)
  static constexpr int CH = 64;                     // elements of one 128-byte swizzle row
  static constexpr int NKC = D / CH;                // d chunks of Q / K
  static constexpr int NG = D / 128;                // PV column groups (N = 128 per MMA, 64 columns from each CTA)
  static constexpr int HALF_BLK = 64 * 128;         // [64 rows x 128 B]
  static constexpr int STAGE_BYTES = 2 * HALF_BLK;  // 16 KB
  static constexpr int KST = NKC / 2;               // K stages per KV tile
  static constexpr int VST = NG;                    // V stages per KV tile
  static constexpr int Q_BYTES = NKC * HALF_BLK;
  static constexpr int P_BYTES = 2 * HALF_BLK;      // [64 rows x 128 keys] 16-bit
  static constexpr int NB = FA_PAIR_NB;             // S / P buffers
  static_assert(NB >= 2 && NB <= 4, "FA_PAIR_NB must be 2, 3 or 4");
  static constexpr int MISC_BYTES = 2048;           // barriers, TMEM slot, row-max / row-sum exchange
  static constexpr int NS = (227 * 1024 - 1024 - MISC_BYTES - Q_BYTES - NB * P_BYTES) / STAGE_BYTES;  // d=512, NB=4: 6
  static constexpr int NUM_BARS = 1 + 2 * NS + 3 * NB;
  static_assert(NUM_BARS * 8 + 16 + 3 * 128 * 4 <= MISC_BYTES, "misc area too small");
  static constexpr int SMEM_BYTES = 1024 + Q_BYTES + NB * P_BYTES + NS * STAGE_BYTES + MISC_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static constexpr int THREADS = 192;
  static constexpr int TM_S = 0, TM_O = NB * 64;
  static_assert(TM_O + NG * 64 <= 512, "S buffers and O must fit TMEM");
  static constexpr int TMEM_COLS = (TM_O + NG * 64 <= 256) ? 256 : 512;
  static constexpr int P_ARRIVALS = FA_PAIR_WARP_ARRIVE ? 8 : 256;   // both CTAs' softmax warps / threads
};

template <int D, int DT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
fa_tiled_d_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                       const FwdParams p) {
  using T = TiledDPairTraits<D, DT>;
  constexpr int BN = T::BN, CH = T::CH, HALF_BLK = T::HALF_BLK, STAGE_BYTES = T::STAGE_BYTES, NKC = T::NKC, NG = T::NG,
                NS = T::NS, KST = T::KST, NB = T::NB;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // NKC blocks [64 rows x 128 B]; reused as the O staging area
  uint8_t* sP = sQ + T::Q_BYTES;                        // NB buffers x 2 blocks [64 rows x 64 keys]
  uint8_t* sRing = sP + NB * T::P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + NS * STAGE_BYTES);
  uint64_t* q_full = bars;            // [1]  leader: both CTAs' Q blocks landed
  uint64_t* full = q_full + 1;        // [NS] leader: both CTAs' halves of a stage landed
  uint64_t* empty = full + NS;        // [NS] both:   MMAs that read the stage retired (multicast commit)
  uint64_t* s_full = empty + NS;      // [NB] both:   S[b] holds tile j, b = j % NB (multicast commit)
  uint64_t* p_full = s_full + NB;     // [NB] leader: both CTAs' softmax warps wrote P[b] (and rescaled O)
  uint64_t* pv_done = p_full + NB;    // [NB] both:   PV(j) retired (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + NB);
  float* mx_buf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + T::NUM_BARS * 8 + 16);  // [2][128]
  float* l_buf = mx_buf + 256;                                                                          // [128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // pairs ordered q-tile fastest: the q-tiles of one head run on neighbouring SM pairs and share K / V through L2
  const int pair = blockIdx.x >> 1;
  const int n_qtiles = (p.L + T::BM - 1) / T::BM;
  const int q_row0 = (pair % n_qtiles) * T::BM + int(rank) * T::BMC;   // first query row of THIS CTA
  const int bh = pair / n_qtiles;
  // causal: nothing right of this q-tile's diagonal block is loaded (both CTAs of the pair walk the same KV tiles)
  const int kv_end = p.causal ? min(p.L, (pair % n_qtiles) * T::BM + T::BM) : p.L;
  const int n_tiles = (kv_end + BN - 1) / BN;
  // (Tried and dropped: starting each q-tile's KV loop at a different tile so the q-tiles of a head do not pull the same
  //  K/V lines out of L2 at the same time — no measurable change, 1031 vs 1039 TFLOP/s at B16 H8 L4096 d512.)

  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < NB; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], T::P_ARRIVALS);
      mbar_init(&pv_done[b], 1);
    }
    fence_mbar_init();
  }
  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmO);
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, T::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised and its TMEM allocated before anything remote happens
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================================== TMA producer (both CTAs) =====================================
    if (elect_one_sync()) {
      if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * T::Q_BYTES);
      const uint32_t q_full_ld = mapa_shared(smem_u32(q_full), 0);
#pragma unroll
      for (int c = 0; c < NKC; ++c) tma_load_3d_pair(sQ + c * HALF_BLK, &tmQ, q_full_ld, c * CH, q_row0, bh);
      int it = 0;
      auto acquire = [&]() -> int {   // returns the stage; the leader arms its barrier for both CTAs' bytes
        const int stage = it % NS;
        if (it >= NS) mbar_wait_sleep(&empty[stage], ((it / NS) - 1) & 1);
        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
        ++it;
        return stage;
      };
      auto load_k = [&](int j) {      // this CTA's 64 keys of tile j, two d-chunks per stage
        for (int s = 0; s < KST; ++s) {
          const int stage = acquire();
          const uint32_t bar = mapa_shared(smem_u32(&full[stage]), 0);
          uint8_t* dst = sRing + stage * STAGE_BYTES;
          tma_load_3d_pair(dst, &tmK, bar, (2 * s) * CH, j * BN + int(rank) * 64, bh);
          tma_load_3d_pair(dst + HALF_BLK, &tmK, bar, (2 * s + 1) * CH, j * BN + int(rank) * 64, bh);
        }
      };
      auto load_v = [&](int j) {      // all 128 keys of tile j, this CTA's 64 columns of each 128-column group
        for (int g = 0; g < NG; ++g) {
          const int stage = acquire();
          const uint32_t bar = mapa_shared(smem_u32(&full[stage]), 0);
          tma_load_3d_pair(sRing + stage * STAGE_BYTES, &tmV, bar, g * 128 + int(rank) * 64, j * BN, bh);
        }
      };
      // consumption order: K(0) .. K(NB-2) | K(NB-1) V(0) | K(NB) V(1) | ... | V(n-1)
      for (int j = 0; j < NB - 1 && j < n_tiles; ++j) load_k(j);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + NB - 1 < n_tiles) load_k(j + NB - 1);
        load_v(j);
      }
    }
  } else if (warp == 5) {
    // ===================================== MMA issuer (leader CTA only) =================================
    if (rank == 0 && elect_one_sync()) {
      constexpr uint32_t idesc_qk = make_idesc(T::FMT, 128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(T::FMT, 128, 128, 0, 1);
      constexpr uint64_t hiK = make_smem_desc_hi(16, 1024, SWZ_128B);           // K-major rows of 128 B, 8-row groups
      constexpr uint64_t hiV = make_smem_desc_hi(STAGE_BYTES, 1024, SWZ_128B);  // MN-major: 64 d columns x 8-key groups
      const uint32_t sQ_addr = smem_u32(sQ), sP_addr = smem_u32(sP), ring_addr = smem_u32(sRing);
      int it = 0;
      auto qk = [&](int b) {  // S[b] = Q K^T, accumulated over the d chunks as they land
        for (int s = 0; s < KST; ++s) {
          const int stage = it % NS;
          mbar_wait_sleep(&full[stage], (it / NS) & 1);
          tc_fence_after();
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss_pair(tmem_base + T::TM_S + b * 64,
                           make_smem_desc(sQ_addr + (2 * s + e) * HALF_BLK + k * 32, hiK),
                           make_smem_desc(ring_addr + stage * STAGE_BYTES + e * HALF_BLK + k * 32, hiK), idesc_qk,
                           (s | e | k) ? 1u : 0u);
          tc_commit_pair(&empty[stage], 3);
          ++it;
        }
        tc_commit_pair(&s_full[b], 3);
      };
      auto pv = [&](int b, uint32_t acc) {  // O[:, 128g .. 128g+127] (+)= P[b] V(:, group g)
        for (int g = 0; g < NG; ++g) {
          const int stage = it % NS;
          mbar_wait_sleep(&full[stage], (it / NS) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk)
            umma_ss_pair(tmem_base + T::TM_O + g * 64,
                         make_smem_desc(sP_addr + b * T::P_BYTES + (kk >> 2) * HALF_BLK + (kk & 3) * 32, hiK),
                         make_smem_desc(ring_addr + stage * STAGE_BYTES + kk * 16 * 128, hiV), idesc_pv,
                         (acc | (kk > 0)) ? 1u : 0u);
          tc_commit_pair(&empty[stage], 3);
          ++it;
        }
        tc_commit_pair(&pv_done[b], 3);
      };
      mbar_wait_sleep(q_full, 0);
      tc_fence_after();
      for (int j = 0; j < NB - 1 && j < n_tiles; ++j) qk(j);
      int b = 0, b_ahead = NB - 1;   // j % NB, (j + NB - 1) % NB
      uint32_t par = 0;              // (j / NB) & 1
      for (int j = 0; j < n_tiles; ++j) {
        // QK(j+NB-1) goes in BEFORE the wait for P(j).  Its S buffer held tile j-1, which both CTAs' softmax had read
        // when they signalled p_full(j-1) (waited for one iteration ago).
        if (j + NB - 1 < n_tiles) qk(b_ahead);
        mbar_wait_cluster(&p_full[b], par);
        tc_fence_after();
        pv(b, j > 0 ? 1u : 0u);
        if (++b == NB) { b = 0; par ^= 1u; }
        if (++b_ahead == NB) b_ahead = 0;
      }
    }
  } else {
    // ===================================== softmax warpgroup (both CTAs) ================================
    const int t = threadIdx.x;          // TMEM lane
    const int row = t & 63;             // query row within this CTA
    const int half = t >> 6;            // which 64 keys of the tile / which 64 columns of each O group
    const uint32_t t_lane = tmem_base + (uint32_t(warp * 32) << 16);
    const uint32_t tO = t_lane + T::TM_O;
    const uint32_t sP_addr = smem_u32(sP);
    const uint32_t p_full_ld = mapa_shared(smem_u32(p_full), 0);   // the leader's p_full[0]; [1] is 8 bytes on
    float m_used = -CUDART_INF_F;
    float l = 0.f;

    int b = 0, b_prev = NB - 1;        // j % NB, (j - 1) % NB
    uint32_t par = 0, par_prev = 1;    // (j / NB) & 1, ((j - 1) / NB) & 1
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t tS = t_lane + T::TM_S + b * 64;
      // S[b] ready; in-order completion also proves PV(j-NB) retired, i.e. P[b] may be overwritten.
      mbar_wait_sleep(&s_full[b], par);
      tc_fence_after();
      uint32_t s[2][32];
      tmem_ld32(tS, s[0]);
      tmem_ld32(tS + 32, s[1]);
      tc_wait_ld();

      // keys of this thread's half that its row may attend to: the ragged end of the keys and, when causal, the diagonal
      // (a row's right half can be masked out entirely; its left half never is: key j*BN <= the row index on the diagonal)
      int valid = kv_end - (j * BN + half * 64);
      if (p.causal) valid = min(valid, q_row0 + row - (j * BN + half * 64) + 1);
      if (valid < 64) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int x = 0; x < 32; ++x)
            if (c * 32 + x >= valid) s[c][x] = __float_as_uint(-CUDART_INF_F);
      }
      float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        mx0 = fmaxf(mx0, __uint_as_float(s[0][x]));
        mx1 = fmaxf(mx1, __uint_as_float(s[1][x]));
      }
      // the other half of this row lives in thread t ^ 64: exchange the half-row maxima (key 0 of a tile always exists,
      // so the joint maximum is finite)
#if FA_PAIR_PROBE & 1
      const float mx = fmaxf(mx0, mx1);
#else
      mx_buf[(j & 1) * 128 + t] = fmaxf(mx0, mx1);
      named_bar_sync(1, 128);
      const float mx = fmaxf(fmaxf(mx0, mx1), mx_buf[(j & 1) * 128 + (t ^ 64)]);
#endif

      if (j == 0) {
        m_used = mx;
      } else {
        const bool need = (mx - m_used) * p.scale_log2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // O may only be touched once PV(j-1) has retired.  Its barrier last completed for tile j-1-NB (proved by
          // S(j) being ready) and cannot complete again before this thread signals P(j): the parity is unambiguous.
          mbar_wait_sleep(&pv_done[b_prev], par_prev);
          tc_fence_after();
          const float alpha = need ? ex2_approx((m_used - mx) * p.scale_log2) : 1.0f;
          if (need) m_used = mx;
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < NG * 2; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int x = 0; x < 32; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * alpha);
            tmem_st32(tO + c * 32, o);
          }
          tc_wait_st();
        }
      }

      const float neg_m = -m_used * p.scale_log2;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int x = 0; x < 32; ++x) {
#if FA_PAIR_PROBE & 2
        const float p0 = fmaf(__uint_as_float(s[0][x]), p.scale_log2, neg_m);
        const float p1 = fmaf(__uint_as_float(s[1][x]), p.scale_log2, neg_m);
#else
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[0][x]), p.scale_log2, neg_m));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[1][x]), p.scale_log2, neg_m));
#endif
        l0 += p0;
        l1 += p1;
        s[0][x] = __float_as_uint(p0);
        s[1][x] = __float_as_uint(p1);
      }
      l += l0 + l1;
      // P[b], block `half` ([64 rows x 64 keys], 128-byte rows, 128B swizzle): 8 chunks of 8 keys for this row
      const uint32_t p_row = sP_addr + b * T::P_BYTES + half * HALF_BLK + row * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = q >> 2, e = (q & 3) * 8;
        auto pk2 = [&](int i) {
          const float a0 = __uint_as_float(s[c][e + i]), a1 = __uint_as_float(s[c][e + i + 1]);
          return (DT == DT_BF16) ? pack_bf16x2(a0, a1) : pack_f16x2(a0, a1);
        };
        uint4 v;
        v.x = pk2(0);
        v.y = pk2(2);
        v.z = pk2(4);
        v.w = pk2(6);
#if FA_PAIR_PROBE & 4
        if (v.x == 0x12345678u) st_shared_v4(p_row + ((q ^ (row & 7)) << 4), v);
#else
        st_shared_v4(p_row + ((q ^ (row & 7)) << 4), v);
#endif
      }
      // st.shared -> visible to the tensor core (P is the A operand: only this SM reads it)
#if FA_PAIR_FULL_FENCE
      fence_proxy_async_all();
#else
      fence_proxy_async_smem();
#endif
      tc_fence_before();
#if FA_PAIR_WARP_ARRIVE
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(p_full_ld + b * 8);
#else
      mbar_arrive_cluster(p_full_ld + b * 8);
#endif
      b_prev = b;
      par_prev = par;
      if (++b == NB) { b = 0; par ^= 1u; }
    }

    // ------------------------------- epilogue: O / l -> 16-bit -> smem (Q's dead blocks) -> TMA store ------------
    l_buf[t] = l;
    named_bar_sync(1, 128);
    const float l_row = l + l_buf[t ^ 64];
    const float inv_l = 1.0f / l_row;
    if (p.lse_out != nullptr && half == 0 && q_row0 + row < p.L)
      p.lse_out[size_t(bh) * p.L + q_row0 + row] = m_used * p.scale + __logf(l_row);
    mbar_wait_sleep(&pv_done[b_prev], par_prev);   // the last PV (same parity argument as in the rescale branch)
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < NG * 2; ++c) {   // 32 columns at a time; group g = c >> 1 holds d = 128g + 64*half + [0,64)
      uint32_t o[32];
      tmem_ld32(tO + c * 32, o);
      tc_wait_ld();
      uint8_t* blk = sQ + (2 * (c >> 1) + half) * HALF_BLK + row * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        auto pk2 = [&](int e) {
          const float a0 = __uint_as_float(o[e]) * inv_l, a1 = __uint_as_float(o[e + 1]) * inv_l;
          return (DT == DT_BF16) ? pack_bf16x2(a0, a1) : pack_f16x2(a0, a1);
        };
        uint4 v;
        v.x = pk2(8 * u + 0);
        v.y = pk2(8 * u + 2);
        v.z = pk2(8 * u + 4);
        v.w = pk2(8 * u + 6);
        const int q = (c & 1) * 4 + u;   // 16-byte chunk within the 128-byte block row
        *reinterpret_cast<uint4*>(blk + ((q ^ (row & 7)) << 4)) = v;
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (t == 0 && q_row0 < p.L) {
#pragma unroll
      for (int c = 0; c < NKC; ++c) tma_store_3d(&tmO, sQ + c * HALF_BLK, c * CH, q_row0, bh);
      tma_store_commit();
      tma_store_wait_all();
    }
  }

  // Neither CTA may exit (or free TMEM) while the other can still signal its barriers or read its shared memory.
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 4) tmem_dealloc_pair(tmem_base, T::TMEM_COLS);
}

inline bool tiled_d_pair_supported(int d, int dtype) { return dtype != DT_F32 && (d == 256 || d == 512); }

}  // namespace fa
