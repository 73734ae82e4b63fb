// fa_splitkv_sm100.cuh — K3a': the V2 split-KV partial kernel for SHORT splits (kv_per_split <= 128 keys, i.e. every split
// is a single KV tile — the reference's own configuration: BK 16 x KV_TILES_PER_BLOCK 4 = 64 keys per split at C3).
//
// Replaces  partial_attention_kernel  flash_attention_v2/CUDA/flash_attention_v2.h:243-341 (same workspace contract as
// the fused-tile kernel's SPLIT mode: Oaccum [n_splits][BH][L][D] fp32 normalised by the split's own row sum, LSEaccum
// [n_splits][BH][L] = m/sqrt(d) + ln l).  Longer splits stay on fa_fwd_kernel<.., SPLIT = true>.
//
// Why a separate kernel.  With one tile per split there is no online-softmax recurrence at all (no running max, no O
// rescale), the work per (q-tile, split) unit is ~100 tensor-core cycles, and the kernel has to write 4-byte partials:
// 68 MB out for 25 MB in at C3.  The fused-tile kernel spends a fixed ~2.3 us per work item on that (prologue/epilogue of
// its 2-Q-tile pipeline, Q re-loaded for every split, a 64-key split computed as a half-masked 128-key tile, row-strided
// 16-byte stores of the fp32 rows).  Here instead:
//   unit = (head, 128-row q-tile, split).  The units of the whole problem are numbered q-tile-major / split-fastest and
//     cut into gridDim.x contiguous ranges (persistent CTAs, <= 1 % imbalance), so Q is loaded once per q-tile and stays in
//     shared memory (double-buffered) while that tile's splits stream through.
//   BN = 64 or 128 keys per tile (template): a 64-key split is a 128x64 score tile, not a half-masked 128x128 one.
//   TMEM  S[NSB] (BN columns each, P aliases S; NSB = 2..4 as TMEM allows) + O[2] (D columns each).
//   A unit is ~420 softmax + ~100 epilogue warp instructions per SM sub-partition, issued at the ~2.5 cycles per
//   instruction a sub-partition reaches with two or three resident warps (tests/gpu_probe/probe_pipes.cu): ~1300 cycles
//   per 64-key unit is the issue-bound rate measured with the stores removed (0.7 us/unit; 1.06 us/unit with them at C3,
//   1.3 us/unit when the partials stream to HBM).  So the roles are spread over warps and units are overlapped:
//     warps 0-3 / 4-7   softmax of even / odd units: S -> P.  The row sum is known before P is stored (single tile), so P
//                       is stored NORMALISED, O = P V is final, and this warpgroup also writes the row's LSE.
//     warps 8-11        epilogue of every unit: O -> registers -> swizzled fp32 staging in shared memory -> TMA store
//                       (full 128-byte lines instead of row-strided 16-byte stores from registers).
//     warp 12 / 14      TMA producers: Q and the K ring / the V ring (separate rings, so V tiles waiting for their PV do
//                       not pin the slots the K prefetch needs: the prefetch distance sets the unit rate, see NR).
//     warp 13           tcgen05.mma issuer: QK runs NSB-1 units ahead of PV.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fa_fwd_sm100.cuh"

#ifndef FA_SKV_WAIT
#define FA_SKV_WAIT mbar_wait   // polling or mbar_wait_sleep (suspend-time hint), A/B-tested per kernel
#endif

namespace fa {

template <int D, int DT, int BN_>
struct SplitTileTraits {
  using F = FwdTraits<D, DT>;
  static constexpr int BM = 128, BN = BN_;
  static_assert(BN == 64 || BN == 128, "one KV tile of 64 or 128 keys per split");
  static constexpr int Q_BLK_BYTES = 128 * F::SWB;           // one column block of Q: 128 rows x SWB bytes
  static constexpr int KV_BLK_BYTES = BN * F::SWB;           // one column block of a K / V tile: BN rows x SWB bytes
  static constexpr int Q_BYTES = F::NBLK * Q_BLK_BYTES;
  static constexpr int KV_BYTES = F::NBLK * KV_BLK_BYTES;
  static constexpr int STG_BLK_BYTES = 128 * 128;            // fp32 staging block: 128 rows x 32 columns (one 128-byte row each)
  static constexpr int STG_COLS = (D == 64 && Q_BYTES <= 16384) ? 64 : 32;   // columns staged per hand-over to the TMA store
  static constexpr int STG_BUF_BYTES = (STG_COLS / 32) * STG_BLK_BYTES;
  static constexpr int STG_BYTES = 2 /*buffers*/ * STG_BUF_BYTES;
  // K and V tiles travel through SEPARATE rings (one producer warp each).  QK runs NSB-1 units ahead of PV, so in a shared
  // ring the V tiles still waiting for their PV would pin the slots the K prefetch needs; with 8 KB tiles and ~1 us of DRAM
  // latency the prefetch distance, not any pipe, sets the unit rate (measured: 2 units ahead = 2000 cycles per unit).
  static constexpr int NR_RAW = (216 * 1024 - 2 * Q_BYTES - STG_BYTES) / (2 * KV_BYTES);
  static constexpr int NR = NR_RAW > 8 ? 8 : NR_RAW;         // depth of each ring (tiles)
  static_assert(NR >= 1, "no room for the K/V rings");
  static constexpr int NSB_TMEM = (512 - 2 * D) / BN;        // S buffers: TMEM holds NSB*BN + 2*D columns
  static constexpr int NSB = NSB_TMEM > 4 ? 4 : NSB_TMEM;
  static_assert(NSB >= 2, "need at least two S buffers");
  static constexpr int NUM_BARS = 2 + 2 + 4 * NR + 2 * NSB + 2 + 2;
  static constexpr int SMEM_BYTES = 1024 + 2 * Q_BYTES + 2 * NR * KV_BYTES + STG_BYTES + NUM_BARS * 8 + 16;
  static constexpr int TM_S = 0, TM_O = NSB * BN;
  static constexpr int TMEM_NEED = NSB * BN + 2 * D;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static_assert(TMEM_NEED <= 512, "S and O buffers must fit TMEM");
  static constexpr int THREADS = 512;
};

struct SplitTileParams {
  int L;             // query rows per head (== keys per head)
  int BH;
  int kv_per_split;  // <= BN
  int n_splits;
  int n_qtiles;      // ceil(L / 128)
  long long n_units; // BH * n_qtiles * n_splits
  float scale_log2, scale;
  float* lse_accum;  // [n_splits][BH][L]
};

// Walks units in order (split fastest, then q-tile, then head) without a 64-bit division per step.
struct UnitCursor {
  int split, qt, bh;
  __device__ __forceinline__ UnitCursor(long long u, const SplitTileParams& p) {
    split = int(u % p.n_splits);
    const long long qt_all = u / p.n_splits;
    qt = int(qt_all % p.n_qtiles);
    bh = int(qt_all / p.n_qtiles);
  }
  __device__ __forceinline__ void next(const SplitTileParams& p) {
    if (++split == p.n_splits) {
      split = 0;
      if (++qt == p.n_qtiles) {
        qt = 0;
        ++bh;
      }
    }
  }
};

template <int D, int DT, int BN>
__global__ void __launch_bounds__(512, 1)
fa_splitkv_tile_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmOacc,
                       const SplitTileParams p) {
  using T = SplitTileTraits<D, DT, BN>;
  using F = typename T::F;
  constexpr int NR = T::NR, NSB = T::NSB, NBLK = F::NBLK, BLK_ELEMS = F::BLK_ELEMS, UK = F::UK;
  constexpr uint32_t KIND = F::KIND;
  constexpr int NB = BN / 32;   // 32-column blocks of a score tile

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2] Q tiles
  uint8_t* sK = sQ + 2 * T::Q_BYTES;                   // [NR] K tiles
  uint8_t* sV = sK + NR * T::KV_BYTES;                 // [NR] V tiles
  uint8_t* sStg = sV + NR * T::KV_BYTES;               // [2] fp32 staging buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + T::STG_BYTES);
  uint64_t* q_full = bars;              // [2]   TMA -> MMA
  uint64_t* q_empty = q_full + 2;       // [2]   MMA (commit) -> TMA: every QK of that q-tile retired
  uint64_t* k_full = q_empty + 2;       // [NR]  TMA -> MMA
  uint64_t* k_empty = k_full + NR;      // [NR]  MMA (commit) -> TMA
  uint64_t* v_full = k_empty + NR;      // [NR]  TMA -> MMA
  uint64_t* v_empty = v_full + NR;      // [NR]  MMA (commit) -> TMA
  uint64_t* s_full = v_empty + NR;      // [NSB] MMA -> softmax: S[sb] of this unit ready
  uint64_t* p_full = s_full + NSB;      // [NSB] softmax (128 arrivals) -> MMA: P[sb] in TMEM
  uint64_t* o_done = p_full + NSB;      // [2]   MMA -> epilogue: O[ob] = P V retired
  uint64_t* o_free = o_done + 2;        // [2]   epilogue (128 arrivals) -> MMA: O[ob] read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 13 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&o_done[i], 1);
      mbar_init(&o_free[i], 128);
    }
    for (int i = 0; i < NSB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    for (int s = 0; s < NR; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 12) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmOacc);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, T::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the combine kernel may be scheduled as CTAs of this grid retire (it waits for our completion)

  // this CTA's contiguous range of units; unit u -> (q-tile index over all heads, split), split fastest
  const long long u_begin = p.n_units * blockIdx.x / gridDim.x;
  const long long u_end = p.n_units * (blockIdx.x + 1) / gridDim.x;
  const int n_local = int(u_end - u_begin);

  // Register re-split (launch gives every thread 128): 168 per softmax thread, 96 per epilogue thread, 80 for the
  // data-movement warpgroup  (168 + 168 + 96 + 80 = 4 * 128).
  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (warp == 12) {
      // ===================================== TMA producer: Q and K ==============================
      if (elect_one_sync()) {
        int nq = 0;   // Q tiles loaded so far
        UnitCursor uc(u_begin, p);
        for (int n = 0; n < n_local; ++n, uc.next(p)) {
          const int split = uc.split, bh = uc.bh, q_row0 = uc.qt * T::BM;
          if (n == 0 || split == 0) {
            const int qi = nq & 1;
            if (nq >= 2) FA_SKV_WAIT(&q_empty[qi], ((nq >> 1) - 1) & 1);
            mbar_arrive_expect_tx(&q_full[qi], T::Q_BYTES);
#pragma unroll
            for (int b = 0; b < NBLK; ++b)
              tma_load_3d(sQ + qi * T::Q_BYTES + b * T::Q_BLK_BYTES, &tmQ, &q_full[qi], b * BLK_ELEMS, q_row0, bh);
            ++nq;
          }
          const int stage = n % NR;
          if (n >= NR) FA_SKV_WAIT(&k_empty[stage], ((n / NR) - 1) & 1);
          mbar_arrive_expect_tx(&k_full[stage], T::KV_BYTES);
#pragma unroll
          for (int b = 0; b < NBLK; ++b)
            tma_load_3d(sK + stage * T::KV_BYTES + b * T::KV_BLK_BYTES, &tmK, &k_full[stage], b * BLK_ELEMS,
                        split * p.kv_per_split, bh);
        }
      }
    } else if (warp == 14) {
      // ===================================== TMA producer: V ====================================
      if (elect_one_sync()) {
        UnitCursor uc(u_begin, p);
        for (int n = 0; n < n_local; ++n, uc.next(p)) {
          const int stage = n % NR;
          if (n >= NR) FA_SKV_WAIT(&v_empty[stage], ((n / NR) - 1) & 1);
          mbar_arrive_expect_tx(&v_full[stage], T::KV_BYTES);
#pragma unroll
          for (int b = 0; b < NBLK; ++b)
            tma_load_3d(sV + stage * T::KV_BYTES + b * T::KV_BLK_BYTES, &tmV, &v_full[stage], b * BLK_ELEMS,
                        uc.split * p.kv_per_split, uc.bh);
        }
      }
    } else if (warp == 13) {
      // ===================================== MMA issuer ========================================
      if (elect_one_sync()) {
        constexpr uint32_t idesc_qk = make_idesc(F::FMT, T::BM, BN, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc(F::FMT, T::BM, D, 0, 1);
        constexpr uint64_t hiK = make_smem_desc_hi(16, 8 * F::SWB, F::SWZ);
        constexpr uint64_t hiV = (DT == DT_F32) ? make_smem_desc_hi(T::KV_BLK_BYTES, 512, SWZ_128B_BASE32B)
                                                : make_smem_desc_hi(T::KV_BLK_BYTES, 8 * F::SWB, F::SWZ);
        const uint32_t sQ_addr = smem_u32(sQ), sK_addr = smem_u32(sK), sV_addr = smem_u32(sV);
        int nq = 0;   // Q tiles consumed so far (the one in use is nq - 1)
        int qk_split = int(u_begin % p.n_splits);   // split index of the next unit qk() is called for (called in order)
        auto qk = [&](int n) {   // S[n % NSB] = Q K^T of local unit n
          const int split = qk_split;
          qk_split = (qk_split + 1 == p.n_splits) ? 0 : qk_split + 1;
          if (n == 0 || split == 0) {
            FA_SKV_WAIT(&q_full[nq & 1], (nq >> 1) & 1);
            ++nq;
          }
          const int qi = (nq - 1) & 1;
          const int sb = n % NSB;
          const int stage = n % NR;
          FA_SKV_WAIT(&k_full[stage], (n / NR) & 1);
          tc_fence_after();
          const uint32_t a_base = sQ_addr + qi * T::Q_BYTES, b_base = sK_addr + stage * T::KV_BYTES;
#pragma unroll
          for (int k = 0; k < D / UK; ++k) {
            const uint32_t a_off = (k / F::KPR) * T::Q_BLK_BYTES + (k % F::KPR) * 32;
            const uint32_t b_off = (k / F::KPR) * T::KV_BLK_BYTES + (k % F::KPR) * 32;
            umma_ss<KIND>(tmem_base + T::TM_S + sb * BN, make_smem_desc(a_base + a_off, hiK),
                          make_smem_desc(b_base + b_off, hiK), idesc_qk, k > 0 ? 1u : 0u);
          }
          tc_commit(&s_full[sb]);
          tc_commit(&k_empty[stage]);
          // last QK that reads this Q tile: the next unit starts a new q-tile, or the range ends
          if (n + 1 == n_local || qk_split == 0) tc_commit(&q_empty[qi]);
        };
        // QK runs NSB-1 units ahead: S[sb] of unit n+NSB-1 was last read (as P) by PV(n-1), issued before it, in order.
        for (int n = 0; n < NSB - 1 && n < n_local; ++n) qk(n);
        for (int n = 0; n < n_local; ++n) {
          if (n + NSB - 1 < n_local) qk(n + NSB - 1);
          const int sb = n % NSB, ks = n / NSB;
          const int ob = n & 1, ko = n >> 1;
          const int stage = n % NR;
          FA_SKV_WAIT(&v_full[stage], (n / NR) & 1);
          if (ko > 0) FA_SKV_WAIT(&o_free[ob], (ko - 1) & 1);   // O[ob] of unit n-2 has been read out
          FA_SKV_WAIT(&p_full[sb], ks & 1);
          tc_fence_after();
          const uint32_t b_base = sV_addr + stage * T::KV_BYTES;
#pragma unroll
          for (int kk = 0; kk < BN / UK; ++kk)
            umma_ts<KIND>(tmem_base + T::TM_O + ob * D, tmem_base + T::TM_S + sb * BN + kk * (UK * F::ES / 4),
                          make_smem_desc(b_base + kk * UK * F::SWB, hiV), idesc_pv, kk > 0 ? 1u : 0u);
          tc_commit(&o_done[ob]);
          tc_commit(&v_empty[stage]);
        }
      }
    }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    // ===================================== softmax warpgroups =================================
    // One tile per split: no running max, no rescale, and the row sum is known before P is stored — so P is stored already
    // normalised (P / l) and O = P V needs no epilogue arithmetic at all.
    const int w = warp >> 2;   // this warpgroup takes local units n with (n & 1) == w
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    UnitCursor uc(u_begin + w, p);
    for (int n = w; n < n_local; n += 2, uc.next(p), uc.next(p)) {
      const int sb = n % NSB, ks = n / NSB;
      const int split = uc.split, bh = uc.bh, q_row0 = uc.qt * T::BM;
      const int kv_begin = split * p.kv_per_split;
      const int valid = min(p.L, kv_begin + p.kv_per_split) - kv_begin;   // 1 .. BN keys of this tile count
      const uint32_t tS = t_lane + T::TM_S + sb * BN;

      FA_SKV_WAIT(&s_full[sb], ks & 1);
      tc_fence_after();
      uint32_t s[NB][32];
#pragma unroll
      for (int c = 0; c < NB; ++c) tmem_ld32(tS + c * 32, s[c]);
      tc_wait_ld();
      if (valid < BN) {
#pragma unroll
        for (int c = 0; c < NB; ++c)
#pragma unroll
          for (int x = 0; x < 32; ++x)
            if (c * 32 + x >= valid) s[c][x] = __float_as_uint(-CUDART_INF_F);
      }
      float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};   // four independent chains
#pragma unroll
      for (int c = 0; c < NB; ++c)
#pragma unroll
        for (int x = 0; x < 32; x += 8)
#pragma unroll
          for (int y = 0; y < 4; ++y)
            mx[y] = fmaxf(mx[y], fmaxf(__uint_as_float(s[c][x + 2 * y]), __uint_as_float(s[c][x + 2 * y + 1])));
      const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      const float neg_m = -m * p.scale_log2;
      float2 lsum[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
      for (int c = 0; c < NB; ++c)
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          float2 v = __ffma2_rn(make_float2(__uint_as_float(s[c][x]), __uint_as_float(s[c][x + 1])),
                                make_float2(p.scale_log2, p.scale_log2), make_float2(neg_m, neg_m));
          v.x = ex2_approx(v.x);
          v.y = ex2_approx(v.y);
          lsum[(x >> 1) & 3] = __fadd2_rn(lsum[(x >> 1) & 3], v);
          s[c][x] = __float_as_uint(v.x);
          s[c][x + 1] = __float_as_uint(v.y);
        }
      const float2 l2 = __fadd2_rn(__fadd2_rn(lsum[0], lsum[1]), __fadd2_rn(lsum[2], lsum[3]));
      const float l = l2.x + l2.y;
      const float inv_l = 1.0f / l;
      const float2 inv2 = make_float2(inv_l, inv_l);
      if constexpr (DT == DT_F32) {
#pragma unroll
        for (int c = 0; c < NB; ++c) {
#pragma unroll
          for (int x = 0; x < 32; x += 2) {
            const float2 v = __fmul2_rn(make_float2(__uint_as_float(s[c][x]), __uint_as_float(s[c][x + 1])), inv2);
            s[c][x] = __float_as_uint(v.x);
            s[c][x + 1] = __float_as_uint(v.y);
          }
          tmem_st32(tS + c * 32, s[c]);   // fp32 P in place over S (tf32 A operand)
        }
      } else {
#pragma unroll
        for (int c = 0; c < NB; c += 2) {   // two 32-key blocks -> 32 packed columns
          uint32_t pk[32];
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const float2 v0 = __fmul2_rn(make_float2(__uint_as_float(s[c][2 * x]), __uint_as_float(s[c][2 * x + 1])), inv2);
            const float2 v1 =
                __fmul2_rn(make_float2(__uint_as_float(s[c + 1][2 * x]), __uint_as_float(s[c + 1][2 * x + 1])), inv2);
            pk[x] = (DT == DT_BF16) ? pack_bf16x2(v0.x, v0.y) : pack_f16x2(v0.x, v0.y);
            pk[16 + x] = (DT == DT_BF16) ? pack_bf16x2(v1.x, v1.y) : pack_f16x2(v1.x, v1.y);
          }
          tmem_st32(tS + c * 16, pk);
        }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[sb]);
      const int row_g = q_row0 + row;
      if (row_g < p.L) p.lse_accum[(size_t(split) * p.BH + bh) * p.L + row_g] = m * p.scale + __logf(l);
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    // ===================================== epilogue warpgroup =================================
    // O[ob] (already normalised) -> registers -> swizzled fp32 staging -> TMA store, SC columns (1 or 2 128-byte blocks)
    // at a time through two staging buffers.
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t stg_addr = smem_u32(sStg);
    const bool storer = (warp == 8) && (lane == 0);
    constexpr int SC = T::STG_COLS;
    int chunk_no = 0;   // staging buffers written so far (buffer = chunk_no & 1)
    UnitCursor uc(u_begin, p);
    for (int n = 0; n < n_local; ++n, uc.next(p)) {
      const int ob = n & 1, ko = n >> 1;
      const int split = uc.split, bh = uc.bh, q_row0 = uc.qt * T::BM;
      const uint32_t tO = t_lane + T::TM_O + ob * D;
      FA_SKV_WAIT(&o_done[ob], ko & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < D / 32; c0 += SC / 32, ++chunk_no) {
        const int buf = chunk_no & 1;
        uint32_t o[SC / 32][32];
#pragma unroll
        for (int cc = 0; cc < SC / 32; ++cc) tmem_ld32(tO + (c0 + cc) * 32, o[cc]);
        tc_wait_ld();
        if (c0 + SC / 32 >= D / 32) {   // last columns of O[ob] are in registers: PV of unit n+2 may overwrite it
          tc_fence_before();
          mbar_arrive(&o_free[ob]);
        }
        if (storer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store 2 buffers ago has read `buf`
        named_bar_sync(1, 128);
#pragma unroll
        for (int cc = 0; cc < SC / 32; ++cc)
#pragma unroll
          for (int v4 = 0; v4 < 8; ++v4)
            st_shared_v4(stg_addr + buf * T::STG_BUF_BYTES + cc * T::STG_BLK_BYTES + row * 128 + ((v4 ^ (row & 7)) << 4),
                         make_uint4(o[cc][4 * v4], o[cc][4 * v4 + 1], o[cc][4 * v4 + 2], o[cc][4 * v4 + 3]));
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (storer) {
#pragma unroll
          for (int cc = 0; cc < SC / 32; ++cc)
            tma_store_3d(&tmOacc, sStg + buf * T::STG_BUF_BYTES + cc * T::STG_BLK_BYTES, (c0 + cc) * 32, q_row0,
                         split * p.BH + bh);
          tma_store_commit();
        }
      }
    }
    if (storer) tma_store_wait_all();   // every partial row is in global memory before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, T::TMEM_COLS);
}

}  // namespace fa
