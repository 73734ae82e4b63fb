// fa_fwd_sm100.cuh — K1 / K3a: flash-attention forward for head dims whose S and O accumulators
// fit TMEM together (256 + 2*D <= 512 columns), hand-written for sm_100a.
//
// Replaces (same (Q,K,V)->O semantics, [B,H,L,d] contiguous, dense, scale 1/sqrt(d)):
//   flash_attention_v1/CUDA/flash_attention_v1.h:161-248        flash_attention_kernel (scalar)
//   flash_attention_v1/CUDA/flash_attention_v1_opt1.h:264-351   flash_attention_kernel_opt1 (WMMA)
//   flash_attention_v1_tiled_d/CUDA/flash_attention_v1.h:230-309 (d <= 128 case)
//   flash_attention_v2/CUDA/flash_attention_v2.h:243-341        partial_attention_kernel (SPLIT = true)
// Recurrence (flash_attention_v1/numpy_gpu_like_opt2.py:161-195): per KV tile
//   S = Q K^T / sqrt(d); m' = max(m, rowmax S); P = exp(S - m'); l = l*alpha + rowsum P; O = O*alpha + P V.
//
// B200 mapping
//   CTA = one pair of 128-row Q tiles of one (b,h) head [x one KV split], 10 warps:
//     warps 0-3  softmax warpgroup for Q tile 0   (thread <-> S/O row: no shuffles for row max / row sum)
//     warps 4-7  softmax warpgroup for Q tile 1
//     warp  8    TMA producer (one lane): Q tiles once, then the K_j / V_j ring
//     warp  9    tcgen05.mma issuer (one lane)      (warps 10-11 idle: setmaxnreg works per warpgroup)
//   TMEM (512 cols): S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D); P overwrites S in place
//     (packed 16-bit pairs for bf16/fp16, fp32 words for tf32) and feeds the PV MMA as the TMEM A operand.
//   smem: Q 2 tiles + NS-stage K/V ring, all 128B-swizzled [128 rows x 128 B] blocks written by TMA.
//   The two Q tiles ping-pong on the tensor pipe: while warpgroup i runs softmax on S_i the MMA warp
//   issues PV/QK for tile 1-i.  O is rescaled lazily (only when the running max moves by > 2^8), by the
//   softmax warpgroup itself, so there is no separate correction stage on the critical path.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "sm100_ptx.cuh"

namespace fa {

enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };

struct FwdParams {
  int L;             // sequence length (queries == keys)
  int BH;            // B*H
  int kv_per_split;  // keys handled by one split (== L when SPLIT is false)
  int n_splits;
  float scale_log2;  // log2(e)/sqrt(d)
  float scale;       // 1/sqrt(d)
  float* o_accum;    // [n_splits][BH][L][D] fp32, each split normalised by its own l   (SPLIT only)
  float* lse_accum;  // [n_splits][BH][L]    fp32, m/sqrt(d) + ln(l)                     (SPLIT only)
};

template <int D, int DT>
struct FwdTraits {
  static constexpr int ES = (DT == DT_F32) ? 4 : 2;                       // element bytes
  static constexpr uint32_t KIND = (DT == DT_F32) ? KIND_TF32 : KIND_F16;
  static constexpr uint32_t FMT = (DT == DT_F32) ? FMT_TF32 : (DT == DT_BF16 ? FMT_BF16 : FMT_F16);
  static constexpr int BM = 128;                   // query rows per tile (= TMEM lanes)
  static constexpr int BN = 128;                   // keys per KV tile
  static constexpr int ROW_BYTES = D * ES;
  static_assert(ROW_BYTES % 128 == 0, "head-dim row must be a whole number of 128-byte swizzle rows");
  static constexpr int NBLK = ROW_BYTES / 128;     // 128-byte column blocks per tile
  static constexpr int BLK_ELEMS = 128 / ES;       // elements per block row
  static constexpr int BLK_BYTES = 128 * 128;      // 128 rows x 128 B
  static constexpr int TILE_BYTES = NBLK * BLK_BYTES;
  static constexpr int UK = 32 / ES;               // MMA K: 16 (16-bit) / 8 (tf32)
  static constexpr int NS = (TILE_BYTES >= 32768) ? 5 : 8;  // K/V ring depth
  static constexpr int TMEM_COLS = 512;
  static constexpr int TM_S = 0, TM_O = 256;
  static_assert(256 + 2 * D <= 512, "S and O accumulators must fit TMEM");
  static constexpr int NUM_BARS = 2 + 2 * NS + 2 + 4 + 2;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + (2 + NS) * TILE_BYTES + NUM_BARS * 8 + 16;
  static constexpr int THREADS = 384;  // 3 warpgroups: softmax0, softmax1, {TMA, MMA, 2 idle}
};

// Tuning knobs (A/B-tested on B200 through alternative builds; the defaults are what ships).
#ifndef FA_P_HALVES
#define FA_P_HALVES 1   // 1: publish P in two 64-key halves so the first half of PV overlaps the second half of exp
#endif
#ifndef FA_PACKED
#define FA_PACKED 1     // 1: packed fp32x2 FFMA2 / FADD2 for the scale-subtract and the row sum (halves their issue slots)
#endif

// Softmax rescale threshold in log2 units (P values stay <= 2^8; exact after the final O / l).
constexpr float kRescaleThreshold = 8.0f;

template <int D, int DT, bool SPLIT>
__global__ void __launch_bounds__(384, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FwdParams p) {
  using T = FwdTraits<D, DT>;
  constexpr int ES = T::ES, BM = T::BM, BN = T::BN, NBLK = T::NBLK, BLK_ELEMS = T::BLK_ELEMS;
  constexpr int BLK_BYTES = T::BLK_BYTES, TILE_BYTES = T::TILE_BYTES, UK = T::UK, NS = T::NS;
  constexpr uint32_t KIND = T::KIND;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + 2 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NS * TILE_BYTES);
  uint64_t* q_full = bars;              // [2]  TMA -> MMA
  uint64_t* kv_full = q_full + 2;       // [NS] TMA -> MMA
  uint64_t* kv_empty = kv_full + NS;    // [NS] MMA (tcgen05.commit) -> TMA
  uint64_t* s_full = kv_empty + NS;     // [2]  MMA -> softmax: S_i(j) ready (and every earlier MMA retired)
  uint64_t* p_full = s_full + 2;        // [2][2] softmax (128 arrivals) -> MMA: P_i(j) (key half h) in TMEM, O_i rescaled
  uint64_t* o_done = p_full + 4;        // [2]  MMA -> softmax: last PV_i retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_row0 = blockIdx.x * (2 * BM);
  const int bh = blockIdx.y;
  const int split = SPLIT ? blockIdx.z : 0;
  const int kv_begin = SPLIT ? split * p.kv_per_split : 0;
  const int kv_end = SPLIT ? min(p.L, kv_begin + p.kv_per_split) : p.L;
  const int n_tiles = (kv_end - kv_begin + BN - 1) / BN;
  const int n_q = (p.L - q_row0 > BM) ? 2 : 1;

  if (warp == 9 && lane == 0) {
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 128);
      mbar_init(&p_full[2 * i + 1], 128);
      mbar_init(&o_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      if (!SPLIT) tma_prefetch_desc(&tmO);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, T::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register re-split (launch gives every thread 168): the data-movement warpgroup keeps 40, each softmax
  // thread gets 232 so a full 128-column S row plus its packed P stays in registers.  setmaxnreg sits at the top
  // of each role branch (no control-flow merge after it) so ptxas allocates each role under its own budget.
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 8) {
    // ===================================== TMA producer =====================================
    if (elect_one_sync()) {
      auto load_tile = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int row) {
        mbar_arrive_expect_tx(bar, TILE_BYTES);
#pragma unroll
        for (int b = 0; b < NBLK; ++b) tma_load_3d(dst + b * BLK_BYTES, map, bar, b * BLK_ELEMS, row, bh);
      };
      auto load_kv = [&](int t) {  // t = 2j -> K_j, t = 2j+1 -> V_j
        const int stage = t % NS;
        if (t >= NS) mbar_wait(&kv_empty[stage], ((t / NS) - 1) & 1);
        load_tile(sKV + stage * TILE_BYTES, (t & 1) ? &tmV : &tmK, &kv_full[stage], kv_begin + (t >> 1) * BN);
      };
      load_tile(sQ, &tmQ, &q_full[0], q_row0);
      load_kv(0);
      if (n_q > 1) load_tile(sQ + TILE_BYTES, &tmQ, &q_full[1], q_row0 + BM);
      for (int t = 1; t < 2 * n_tiles; ++t) load_kv(t);
    }
  } else if (warp == 9) {
    // ===================================== MMA issuer ========================================
    if (elect_one_sync()) {
      constexpr uint32_t idesc_qk = make_idesc(T::FMT, BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(T::FMT, BM, D, 0, 1);
      constexpr uint64_t hiK = make_smem_desc_hi(16, 1024, SWZ_128B);         // K-major, 8-row atoms 1024 B apart
      // MN-major V: LBO = next 128-B column block; 8-key atoms 1024 B apart.  32-bit (tf32) MN-major operands only
      // exist in the 128B-swizzle / 32B-atom layout (4-key atoms, 512 B apart) — V's tensor map matches (fa_api.cu).
      constexpr uint64_t hiV = (DT == DT_F32) ? make_smem_desc_hi(BLK_BYTES, 512, SWZ_128B_BASE32B)
                                              : make_smem_desc_hi(BLK_BYTES, 1024, SWZ_128B);
      const uint32_t sQ_addr = smem_u32(sQ), sKV_addr = smem_u32(sKV);

      auto qk = [&](int i, int stage) {  // S_i = Q_i K^T
        const uint32_t a_base = sQ_addr + i * TILE_BYTES, b_base = sKV_addr + stage * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < D / UK; ++k) {
          const uint32_t off = (k / 4) * BLK_BYTES + (k % 4) * 32;
          umma_ss<KIND>(tmem_base + T::TM_S + i * BN, make_smem_desc(a_base + off, hiK),
                        make_smem_desc(b_base + off, hiK), idesc_qk, k > 0 ? 1u : 0u);
        }
      };
      auto pv = [&](int i, int stage, uint32_t acc, int kk0, int kk1) {  // O_i (+)= P_i V  (K-steps kk0..kk1-1)
        const uint32_t b_base = sKV_addr + stage * TILE_BYTES;
#pragma unroll
        for (int kk = kk0; kk < kk1; ++kk) {
          umma_ts<KIND>(tmem_base + T::TM_O + i * D, tmem_base + T::TM_S + i * BN + kk * 8,
                        make_smem_desc(b_base + kk * UK * 128, hiV), idesc_pv, (acc | (kk > 0)) ? 1u : 0u);
        }
      };

      mbar_wait(&kv_full[0], 0);
      for (int i = 0; i < n_q; ++i) {
        mbar_wait(&q_full[i], 0);
        tc_fence_after();
        qk(i, 0);
        tc_commit(&s_full[i]);
      }
      tc_commit(&kv_empty[0]);
      for (int j = 0; j < n_tiles; ++j) {
        const int tv = 2 * j + 1, tk = 2 * j + 2;
        const bool has_next = (j + 1 < n_tiles);
        mbar_wait(&kv_full[tv % NS], (tv / NS) & 1);
        for (int i = 0; i < n_q; ++i) {
          constexpr int KT = BN / UK;
          mbar_wait(&p_full[2 * i], j & 1);
          tc_fence_after();
          if (FA_P_HALVES) {
            pv(i, tv % NS, j > 0 ? 1u : 0u, 0, KT / 2);
            mbar_wait(&p_full[2 * i + 1], j & 1);
            tc_fence_after();
            pv(i, tv % NS, 1u, KT / 2, KT);
          } else {
            pv(i, tv % NS, j > 0 ? 1u : 0u, 0, KT);
          }
          if (!has_next) tc_commit(&o_done[i]);
          if (has_next) {
            if (i == 0) {
              mbar_wait(&kv_full[tk % NS], (tk / NS) & 1);
              tc_fence_after();
            }
            qk(i, tk % NS);
            tc_commit(&s_full[i]);
          }
        }
        tc_commit(&kv_empty[tv % NS]);
        if (has_next) tc_commit(&kv_empty[tk % NS]);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ===================================== softmax warpgroups =================================
    const int i = warp >> 2;  // which Q tile
    if (i < n_q) {
      const int row = (warp & 3) * 32 + lane;
      const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
      const uint32_t tS = t_lane + T::TM_S + i * BN;
      const uint32_t tO = t_lane + T::TM_O + i * D;
      float m_used = -CUDART_INF_F;
      float l = 0.f;

      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&s_full[i], j & 1);
        tc_fence_after();
        uint32_t s[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, s[c]);
        tc_wait_ld();

        // Row max.  Only the last tile of a key range can be ragged; its masking (128 compare+select pairs) lives in
        // its own branch together with a copy of the max tree so the compiler cannot if-convert it into every tile.
        const int valid = kv_end - (kv_begin + j * BN);
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
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int x = 0; x < 32; ++x)
              if (c * 32 + x >= valid) s[c][x] = __float_as_uint(-CUDART_INF_F);
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
          // Lazy rescale: keep the stale max unless the new one is > 2^8 larger (in exp2 units).
          const bool need = (mx - m_used) * p.scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            // s_full[i](j) retiring implies PV_i(j-1) retired (same issuing thread, in-order pipe), so O_i is quiescent.
            const float alpha = need ? ex2_approx((m_used - mx) * p.scale_log2) : 1.0f;
            if (need) m_used = mx;
            l *= alpha;
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + c * 32, o);
              tc_wait_ld();
#pragma unroll
              for (int x = 0; x < 32; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * alpha);
              tmem_st32(tO + c * 32, o);
            }
          }
        }

        const float neg_m = -m_used * p.scale_log2;
        // p = 2^(s*scale_log2 - m*scale_log2): column blocks c0..c1-1 of s, four independent streams per step
        float2 lsum[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
        auto exp_blocks = [&](int c0, int c1) {
#pragma unroll
          for (int x = 0; x < 32; x += 2) {
#pragma unroll
            for (int c = c0; c < c1; ++c) {
              float2 v = make_float2(__uint_as_float(s[c][x]), __uint_as_float(s[c][x + 1]));
#if FA_PACKED
              v = __ffma2_rn(v, make_float2(p.scale_log2, p.scale_log2), make_float2(neg_m, neg_m));
#else
              v.x = fmaf(v.x, p.scale_log2, neg_m);
              v.y = fmaf(v.y, p.scale_log2, neg_m);
#endif
              v.x = ex2_approx(v.x);
              v.y = ex2_approx(v.y);
#if FA_PACKED
              lsum[c] = __fadd2_rn(lsum[c], v);
#else
              lsum[c].x += v.x;
              lsum[c].y += v.y;
#endif
              s[c][x] = __float_as_uint(v.x);
              s[c][x + 1] = __float_as_uint(v.y);
            }
          }
        };
        auto store_p = [&](int c0, int c1) {  // P columns 32*c0 .. 32*c1-1 -> TMEM (in place over S)
          if constexpr (DT == DT_F32) {
#pragma unroll
            for (int c = c0; c < c1; ++c) tmem_st32(tS + c * 32, s[c]);
          } else {
#pragma unroll
            for (int c = c0; c < c1; c += 2) {
              uint32_t pk[32];
#pragma unroll
              for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int x = 0; x < 16; ++x) {
                  const float a = __uint_as_float(s[c + cc][2 * x]), b = __uint_as_float(s[c + cc][2 * x + 1]);
                  pk[cc * 16 + x] = (DT == DT_BF16) ? pack_bf16x2(a, b) : pack_f16x2(a, b);
                }
              tmem_st32(tS + (c / 2) * 32, pk);
            }
          }
        };
        if (FA_P_HALVES) {
          exp_blocks(0, 2);
          store_p(0, 2);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&p_full[2 * i]);
          exp_blocks(2, 4);
          store_p(2, 4);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&p_full[2 * i + 1]);
        } else {
          exp_blocks(0, 4);
          store_p(0, 4);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&p_full[2 * i]);
        }
        l += ((lsum[0].x + lsum[0].y) + (lsum[1].x + lsum[1].y)) + ((lsum[2].x + lsum[2].y) + (lsum[3].x + lsum[3].y));
      }

      // ------------------------------- epilogue: O_i / l -------------------------------------
      mbar_wait(&o_done[i], 0);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const int row_g = q_row0 + i * BM + row;
      if constexpr (SPLIT) {
        if (row_g < p.L) {
          const size_t ridx = (size_t(split) * p.BH + bh) * p.L + row_g;
          p.lse_accum[ridx] = m_used * p.scale + __logf(l);
        }
        float* dst = p.o_accum + ((size_t(split) * p.BH + bh) * p.L + row_g) * D;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + c * 32, o);
          tc_wait_ld();
          if (row_g < p.L) {
#pragma unroll
            for (int x = 0; x < 32; x += 4) {
              float4 v = make_float4(__uint_as_float(o[x]) * inv_l, __uint_as_float(o[x + 1]) * inv_l,
                                     __uint_as_float(o[x + 2]) * inv_l, __uint_as_float(o[x + 3]) * inv_l);
              *reinterpret_cast<float4*>(dst + c * 32 + x) = v;
            }
          }
        }
      } else {
        // Stage the tile in Q_i's (now dead) smem buffer in the same 128B-swizzled block layout, then TMA-store it.
        uint8_t* sO = sQ + i * TILE_BYTES;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + c * 32, o);
          tc_wait_ld();
          constexpr int CH = 32 * ES / 16;  // 16-byte chunks per 32 columns
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            uint4 v;
            if constexpr (DT == DT_F32) {
              v.x = __float_as_uint(__uint_as_float(o[4 * u + 0]) * inv_l);
              v.y = __float_as_uint(__uint_as_float(o[4 * u + 1]) * inv_l);
              v.z = __float_as_uint(__uint_as_float(o[4 * u + 2]) * inv_l);
              v.w = __float_as_uint(__uint_as_float(o[4 * u + 3]) * inv_l);
            } else {
              auto pk2 = [&](int e) {
                const float a = __uint_as_float(o[e]) * inv_l, b = __uint_as_float(o[e + 1]) * inv_l;
                return (DT == DT_BF16) ? pack_bf16x2(a, b) : pack_f16x2(a, b);
              };
              v.x = pk2(8 * u + 0);
              v.y = pk2(8 * u + 2);
              v.z = pk2(8 * u + 4);
              v.w = pk2(8 * u + 6);
            }
            const int q = c * CH + u;  // 16-byte chunk index within the row
            uint8_t* dst = sO + (q >> 3) * BLK_BYTES + row * 128 + (((q & 7) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = v;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + i, 128);
        if ((warp & 3) == 0 && lane == 0) {
#pragma unroll
          for (int b = 0; b < NBLK; ++b) tma_store_3d(&tmO, sO + b * BLK_BYTES, b * BLK_ELEMS, q_row0 + i * BM, bh);
          tma_store_commit();
          tma_store_wait_all();
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, T::TMEM_COLS);
}

}  // namespace fa
