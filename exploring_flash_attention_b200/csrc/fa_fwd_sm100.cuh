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
//   PERSISTENT grid (one CTA per SM).  A work item = one pair of 128-row Q tiles of one (b,h) head [x one KV split];
//   CTA c walks items c, c+gridDim.x, ... (q-pair index fastest, so neighbouring SMs stream the same head's K/V out
//   of L2).  Barrier phases, the K/V ring and TMEM live across items, so the next item's Q/K/V loads and first QK^T
//   overlap the previous item's epilogue instead of paying a launch-style prologue/epilogue per tile pair.
//   12 warps:
//     warps 0-3  softmax warpgroup for Q tile 0   (thread <-> S/O row: no shuffles for row max / row sum)
//     warps 4-7  softmax warpgroup for Q tile 1
//     warp  8    TMA producer (one elected lane): Q tiles, then the K_j / V_j ring
//     warp  9    tcgen05.mma issuer (one elected lane)      (warps 10-11 idle: setmaxnreg works per warpgroup)
//   TMEM (512 cols): S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D); P overwrites S in place
//     (packed 16-bit pairs for bf16/fp16, fp32 words for tf32) and feeds the PV MMA as the TMEM A operand.
//     16-bit d <= 64 (SEP_P): P0/P1 get their own 64 columns after O, which lets QK_i(j+1) run under softmax_i(j).
//   smem: Q 2 tiles + NS-stage K/V ring + one output staging block per warpgroup, all swizzled [128 rows x 128 B]
//     blocks moved by TMA (64-byte rows / 64B swizzle when the head-dim row is only 64 bytes: 16-bit d = 32).
//   The two Q tiles ping-pong on the tensor pipe: while warpgroup i runs softmax on S_i the MMA warp issues PV/QK for
//   tile 1-i.  P is published in two 64-key halves so PV starts while the second half is still being exponentiated.
//   O is rescaled lazily (only when the running max moves by > 2^8) by the softmax warpgroup itself.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "sm100_ptx.cuh"

namespace fa {

enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };

struct FwdParams {
  int L;             // query rows per head
  int Lk;            // keys per head (== L except for partial / context-parallel calls)
  int BH;            // B*H
  int H;             // heads per batch entry (indexes kv_lens)
  const int* kv_lens;  // optional [B] int32: batch entry b attends to its first kv_lens[b] keys only (non-SPLIT, may be null)
  int kv_per_split;  // keys handled by one split (== L when SPLIT is false)
  int n_splits;
  int n_qpairs;      // ceil(L / 256)
  int n_items;       // BH * n_splits * n_qpairs
  float scale_log2;  // log2(e)/sqrt(d)
  float scale;       // 1/sqrt(d)
  float* o_accum;    // [n_splits][BH][L][D] fp32, each split normalised by its own l   (SPLIT only)
  float* lse_accum;  // [n_splits][BH][L]    fp32, m/sqrt(d) + ln(l)                     (SPLIT only)
  float* lse_out;    // optional [BH][L] fp32: log-sum-exp of the scaled scores of each row      (non-SPLIT, may be null)
  int causal;        // != 0: query row r attends to keys 0..r only   (SPLIT: only with n_splits == 1 and L == Lk)
  int out_head_rows; // SPLIT: rows between consecutive heads of o_accum / lse_accum (== L unless the caller writes into a
                     // row window of a taller partial buffer)
  long long* trace;  // FA_TRACE builds only: [3 roles][256 tiles][8 slots] SM-clock stamps of CTA 0 (tests/gpu_probe/trace_k1.py)
};

#ifdef FA_TRACE
// SM cycle counter read that cannot be hoisted above the computation of `dep`.
__device__ __forceinline__ long long trace_clock(uint32_t dep = 0) {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep) : "memory");
  return t;
}
#define FA_TR(role, tile, slot, dep)                                                              \
  do {                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && (tile) < 256)                                    \
      p.trace[((role) * 256 + (tile)) * 8 + (slot)] = trace_clock(dep);                           \
  } while (0)
#else
#define FA_TR(role, tile, slot, dep) do { } while (0)
#endif

// (Tried and dropped: O rows straight from registers to global to buy a 5th K/V ring stage — the row-strided 16-byte
//  stores cost 11 % at L=1024 for +1 % at L=16384.)

template <int D, int DT>
struct FwdTraits {
  static constexpr int ES = (DT == DT_F32) ? 4 : 2;                       // element bytes
  static constexpr uint32_t KIND = (DT == DT_F32) ? KIND_TF32 : KIND_F16;
  static constexpr uint32_t FMT = (DT == DT_F32) ? FMT_TF32 : (DT == DT_BF16 ? FMT_BF16 : FMT_F16);
  static constexpr int BM = 128;                   // query rows per tile (= TMEM lanes)
  static constexpr int BN = 128;                   // keys per KV tile
  static constexpr int ROW_BYTES = D * ES;
  static constexpr int SWB = ROW_BYTES >= 128 ? 128 : 64;   // swizzle span = smem row pitch of a block (bytes)
  static_assert(ROW_BYTES % SWB == 0 && ROW_BYTES >= 64, "head-dim row must be a whole number of swizzle rows");
  static constexpr uint32_t SWZ = (SWB == 128) ? SWZ_128B : SWZ_64B;
  static constexpr int NBLK = ROW_BYTES / SWB;     // column blocks per tile
  static constexpr int BLK_ELEMS = SWB / ES;       // elements per block row
  static constexpr int BLK_BYTES = 128 * SWB;      // 128 rows x SWB bytes
  static constexpr int KPR = SWB / 32;             // MMA K-steps per block row
  static constexpr int TILE_BYTES = NBLK * BLK_BYTES;
  static constexpr int UK = 32 / ES;               // MMA K: 16 (16-bit) / 8 (tf32)
  static constexpr int NS = (TILE_BYTES >= 32768) ? 4 : 8;  // K/V ring depth
  static constexpr int STAGING_BYTES = 2 * BLK_BYTES;      // one output staging block per softmax warpgroup
  static constexpr int TMEM_COLS = 512;
  static constexpr int TM_S = 0, TM_O = 256;
  static_assert(256 + 2 * D <= 512, "S and O accumulators must fit TMEM");
  // 16-bit P_i is 64 TMEM columns.  When they fit next to S and O (d <= 64) P gets its own columns instead of aliasing
  // S_i, so QK_i(j+1) can be issued as soon as S_i(j) is in registers and runs under softmax_i(j) instead of after it.
  static constexpr bool SEP_P = (DT != DT_F32) && (256 + 2 * D + 128 <= 512);
  static constexpr int TM_P = SEP_P ? 256 + 2 * D : TM_S;
  static constexpr int P_STRIDE = SEP_P ? 64 : BN;   // columns between P_0 and P_1
  static constexpr int NUM_BARS = 2 + 2 + 2 * NS + 2 + 4 + 2 + 2 + 4;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + (2 + NS) * TILE_BYTES + STAGING_BYTES + NUM_BARS * 8 + 16;
  static constexpr int THREADS = 384;  // 3 warpgroups: softmax0, softmax1, {TMA, MMA, 2 idle}
};

// Tuning knobs (A/B-tested on B200 through alternative builds; the defaults are what ships).
#ifndef FA_P_HALVES
#define FA_P_HALVES 1   // 1: publish P in two 64-key halves so the first half of PV overlaps the second half of exp
#endif
#ifndef FA_P_SPLIT
#define FA_P_SPLIT 2    // 32-key blocks of P in the first published part (2: halves; 3: three quarters then one quarter,
                        // measured +-0 on B200)
#endif
#ifndef FA_PACKED
#define FA_PACKED 1     // 1: packed fp32x2 FFMA2 / FADD2 for the scale-subtract and the row sum (halves their issue slots)
#endif

// (Tried and dropped, round 2 re-measured without register spills: issuing the next item's first QK_i right behind this
//  item's last PV_i — C4 slice +1 %, C2 +-0, causal C2 -3.5 %.  One tcgen05.mma issuing warp per Q tile: the tiles fall into
//  lock-step (both exponentiate, then both queue MMAs), period 3200 -> 3860 cycles; pacing the softmax warpgroups against
//  each other restores the offset but not the period (3550).  Speculative exponentials against the running max with the
//  row-max FMNMX stream folded beside them (redo from TMEM when a row's max really moves): -8 %, the two code paths cost
//  registers and the compiler does not interleave the streams.  The per-Q-tile chain QK -> softmax -> PV bounds K1.)

#ifndef FA_POLY_MOD
#define FA_POLY_MOD 4   // N > 0: one element pair in N takes exp2 on the FMA pipes (Cody-Waite + degree-3 polynomial);
                        // measured on B200: N=4 gives +5 % at d=32, +2 % at d=128 L=1024, N=2/3 no better
#endif

#ifndef FA_K1_TMA_WAIT
#define FA_K1_TMA_WAIT mbar_wait_sleep   // the TMA producer's ring-slot waits are not on the per-tile chain: it sleeps on them
                                         // instead of polling beside the softmax warps of its sub-partition (+0.3..0.6 %)
#endif

// Softmax rescale threshold in log2 units (P values stay <= 2^8; exact after the final O / l).
constexpr float kRescaleThreshold = 8.0f;

// 2^x for a pair of fp32 on the FMA/ALU pipes: n = round(x) via the 1.5*2^23 trick, 2^f on [-0.5, 0.5] by a degree-3
// minimax polynomial (max relative error 7.5e-5, below the 2^-9 rounding of a bf16/fp16 P and the 2^-11 of tf32),
// then n is added into the exponent field.  Inputs are <= kRescaleThreshold; anything below -126 flushes to 2^-126.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.0f, 12582912.0f));          // low mantissa bits = round(x)
  const float2 r = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));        // float(round(x))
  const float2 f = __ffma2_rn(r, make_float2(-1.0f, -1.0f), x);                   // x - round(x)
  float2 q = __ffma2_rn(f, make_float2(0.05517166f, 0.05517166f), make_float2(0.24261113f, 0.24261113f));
  q = __ffma2_rn(q, f, make_float2(0.69326097f, 0.69326097f));
  q = __ffma2_rn(q, f, make_float2(0.99992806f, 0.99992806f));
  q.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  q.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return q;
}

struct ItemCoord {
  int q_row0, bh, split, kv_begin, kv_end, n_tiles, n_q;
  int nt0, nt1;  // KV tiles Q tile 0 / 1 actually needs (causal: up to its diagonal tile); n_tiles = the larger of the two
  // (not an array: a runtime index put the struct in local memory and left LDL/STL in every role's loop)
  __device__ __forceinline__ int nt(int i) const { return i ? nt1 : nt0; }
};

template <bool SPLIT>
__device__ __forceinline__ ItemCoord decode_item(int item, const FwdParams& p) {
  ItemCoord c;
  int qp = item % p.n_qpairs;
  int rest = item / p.n_qpairs;
  if (p.causal) {
    // Causal items grow with the q-pair index.  Order ALL items longest-first (q-pair descending, head fastest) so the
    // static round-robin over CTAs is an LPT schedule; head-major order would hand every CTA the same q-pair index
    // whenever gridDim.x is a multiple of n_qpairs.
    qp = p.n_qpairs - 1 - item / p.BH;
    rest = item % p.BH;
  }
  c.split = SPLIT ? rest % p.n_splits : 0;
  c.bh = SPLIT ? rest / p.n_splits : rest;
  c.q_row0 = qp * 256;
  int kv_len = p.Lk;
  if (!SPLIT && p.kv_lens != nullptr) kv_len = max(1, min(p.Lk, __ldg(p.kv_lens + c.bh / p.H)));  // key-padding mask
  c.kv_begin = SPLIT ? c.split * p.kv_per_split : 0;
  c.kv_end = SPLIT ? min(kv_len, c.kv_begin + p.kv_per_split) : kv_len;
  c.n_tiles = (c.kv_end - c.kv_begin + 127) / 128;
  c.n_q = (p.L - c.q_row0 > 128) ? 2 : 1;
  c.nt0 = c.nt1 = c.n_tiles;
  if (p.causal) {
    c.nt0 = min(c.n_tiles, c.q_row0 / 128 + 1);
    c.nt1 = min(c.n_tiles, c.q_row0 / 128 + 2);
    c.n_tiles = c.n_q > 1 ? c.nt1 : c.nt0;
  }
  return c;
}

template <int D, int DT, bool SPLIT>
__global__ void __launch_bounds__(384, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FwdParams p) {
  using T = FwdTraits<D, DT>;
  constexpr int BM = T::BM, BN = T::BN, NBLK = T::NBLK, BLK_ELEMS = T::BLK_ELEMS;
  constexpr int BLK_BYTES = T::BLK_BYTES, TILE_BYTES = T::TILE_BYTES, UK = T::UK, NS = T::NS;
  constexpr uint32_t KIND = T::KIND;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + 2 * TILE_BYTES;
  uint8_t* sOut = sKV + NS * TILE_BYTES;  // [2] one 128-row x 128-B staging block per softmax warpgroup
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + T::STAGING_BYTES);
  uint64_t* q_full = bars;              // [2]  TMA -> MMA: Q_i of this item landed
  uint64_t* q_empty = q_full + 2;       // [2]  MMA (commit) -> TMA: every QK_i of this item retired, Q_i may be overwritten
  uint64_t* kv_full = q_empty + 2;      // [NS] TMA -> MMA
  uint64_t* kv_empty = kv_full + NS;    // [NS] MMA (commit) -> TMA
  uint64_t* s_full = kv_empty + NS;     // [2]  MMA -> softmax: S_i(j) ready (and every earlier MMA retired)
  uint64_t* p_full = s_full + 2;        // [2][2] softmax (128 arrivals) -> MMA: key-half h of P_i(j) in TMEM, O_i rescaled
  uint64_t* o_done = p_full + 4;        // [2]  MMA -> softmax: last PV_i of this item retired
  uint64_t* o_free = o_done + 2;        // [2]  softmax (128 arrivals) -> MMA: O_i read out, next item may overwrite it
  uint64_t* s_free = o_free + 2;        // [2]  SEP_P: softmax (128 arrivals) -> MMA: S_i(j) is in registers
  uint64_t* pv_done = s_free + 2;       // [2]  SEP_P: MMA (commit) -> softmax: PV_i(j) retired (P_i and O_i quiescent)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 9 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 128);
      mbar_init(&p_full[2 * i + 1], 128);
      mbar_init(&o_done[i], 1);
      mbar_init(&o_free[i], 128);
      mbar_init(&s_free[i], 128);
      mbar_init(&pv_done[i], 1);
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
      if (!SPLIT) tma_prefetch_desc(&tmO);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, T::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if constexpr (SPLIT) pdl_launch_dependents();   // the combine kernel behind a split-KV launch (see launch_combine)

  // Register re-split (launch gives every thread 168): the data-movement warpgroup keeps 72, each softmax thread gets
  // 216 (2*128*216 + 128*72 = 384*168 exactly: an inc that cannot be met from the launch allocation blocks forever).  setmaxnreg sits at the top of each role branch (no control-flow merge after it) so ptxas allocates each
  // role under its own budget.
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 8) {
      // ===================================== TMA producer =====================================
      if (elect_one_sync()) {
        int tt = 0;            // K/V tiles issued so far (ring position), across items
        int nq0 = 0, nq1 = 0;  // Q_i loads issued so far
        auto load_tile = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int row, int bh) {
          mbar_arrive_expect_tx(bar, TILE_BYTES);
#pragma unroll
          for (int b = 0; b < NBLK; ++b) tma_load_3d(dst + b * BLK_BYTES, map, bar, b * BLK_ELEMS, row, bh);
        };
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
          const ItemCoord c = decode_item<SPLIT>(item, p);
          auto load_q = [&](int i, int& n) {
            if (n > 0) FA_K1_TMA_WAIT(&q_empty[i], (n - 1) & 1);
            load_tile(sQ + i * TILE_BYTES, &tmQ, &q_full[i], c.q_row0 + i * BM, c.bh);
            ++n;
          };
          auto load_kv = [&](int t) {  // t = 2j -> K_j, t = 2j+1 -> V_j
            const int stage = tt % NS;
            if (tt >= NS) FA_K1_TMA_WAIT(&kv_empty[stage], ((tt / NS) - 1) & 1);
            load_tile(sKV + stage * TILE_BYTES, (t & 1) ? &tmV : &tmK, &kv_full[stage], c.kv_begin + (t >> 1) * BN, c.bh);
            ++tt;
          };
          load_q(0, nq0);
          load_kv(0);
          if (c.n_q > 1) load_q(1, nq1);
          for (int t = 1; t < 2 * c.n_tiles; ++t) load_kv(t);
        }
      }
    } else if (warp == 9) {
      // ===================================== MMA issuer ========================================
      if (elect_one_sync()) {
        constexpr uint32_t idesc_qk = make_idesc(T::FMT, BM, BN, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc(T::FMT, BM, D, 0, 1);
        constexpr uint64_t hiK = make_smem_desc_hi(16, 8 * T::SWB, T::SWZ);  // K-major, 8-row swizzle atoms
        // MN-major V: LBO = next 128-B column block; 8-key atoms 1024 B apart.  32-bit (tf32) MN-major operands only
        // exist in the 128B-swizzle / 32B-atom layout (4-key atoms, 512 B apart) — V's tensor map matches (fa_api.cu).
        constexpr uint64_t hiV = (DT == DT_F32) ? make_smem_desc_hi(BLK_BYTES, 512, SWZ_128B_BASE32B)
                                                : make_smem_desc_hi(BLK_BYTES, 8 * T::SWB, T::SWZ);
        constexpr int KT = BN / UK;                 // MMA K-steps per KV tile
        constexpr int KS = FA_P_SPLIT * 32 / UK;     // ... of which the first published part of P covers KS
        const uint32_t sQ_addr = smem_u32(sQ), sKV_addr = smem_u32(sKV), tmem_base_ = tmem_base;

        // With P aliased over S the issuer sits on the per-tile critical chain.  Left to itself the compiler hoists every
        // per-K-step descriptor / TMEM address out of the item loop, runs out of this warpgroup's 72 registers and
        // reloads them from local memory (LDL) between "P is ready" and the first MMA; recomputing them from `opaque`
        // bases is a few integer adds (C2 +1 %, C3 +5 %, causal +3 %).  The SEP_P schedule was faster with the hoisting.
        auto opaque = [](uint32_t v) {
          if constexpr (!T::SEP_P) asm volatile("" : "+r"(v));
          return v;
        };
        auto qk = [&](int i, int stage) {  // S_i = Q_i K^T
          const uint32_t a_base = opaque(sQ_addr) + i * TILE_BYTES, b_base = opaque(sKV_addr) + stage * TILE_BYTES;
          const uint32_t tmem_base = opaque(tmem_base_);
#pragma unroll
          for (int k = 0; k < D / UK; ++k) {
            const uint32_t off = (k / T::KPR) * BLK_BYTES + (k % T::KPR) * 32;
            umma_ss<KIND>(tmem_base + T::TM_S + i * BN, make_smem_desc(a_base + off, hiK),
                          make_smem_desc(b_base + off, hiK), idesc_qk, k > 0 ? 1u : 0u);
          }
        };
        auto pv = [&](int i, int stage, uint32_t acc, int kk0, int kk1) {  // O_i (+)= P_i V  (K-steps kk0..kk1-1)
          const uint32_t b_base = opaque(sKV_addr) + stage * TILE_BYTES;
          const uint32_t tmem_base = opaque(tmem_base_);
#pragma unroll
          for (int kk = kk0; kk < kk1; ++kk) {
            umma_ts<KIND>(tmem_base + T::TM_O + i * D, tmem_base + T::TM_P + i * T::P_STRIDE + kk * (UK * T::ES / 4),
                          make_smem_desc(b_base + kk * UK * T::SWB, hiV), idesc_pv, (acc | (kk > 0)) ? 1u : 0u);
          }
        };

        int tt = 0;            // K/V tiles consumed so far (ring position), across items
        int nt[2] = {0, 0};    // KV tiles processed for Q tile i (phase of s_full / p_full), across items
        int ni[2] = {0, 0};    // items processed for Q tile i (phase of q_full / o_done / o_free)
        if constexpr (T::SEP_P) {
          // P_i has its own TMEM columns: per KV step issue both QK(j+1) first (each as soon as its S_i(j) has been
          // read out), then both PV(j) as their P halves arrive.  S_i(j+1) is then ready before softmax_i(j) ends.
          int nqk[2] = {0, 0};  // QK_i issued so far (phase of s_free), across items
          for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const ItemCoord c = decode_item<SPLIT>(item, p);
            const int t0 = tt;
            mbar_wait(&kv_full[t0 % NS], (t0 / NS) & 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (i >= c.n_q) continue;
              mbar_wait(&q_full[i], ni[i] & 1);
              if (nqk[i] > 0) mbar_wait(&s_free[i], (nqk[i] - 1) & 1);
              tc_fence_after();
              qk(i, t0 % NS);
              tc_commit(&s_full[i]);
              ++nqk[i];
              if (c.nt(i) == 1) tc_commit(&q_empty[i]);
            }
            tc_commit(&kv_empty[t0 % NS]);
            for (int j = 0; j < c.n_tiles; ++j) {
              const int tv = t0 + 2 * j + 1, tk = t0 + 2 * j + 2;
              if (j + 1 < c.n_tiles) {
                mbar_wait(&kv_full[tk % NS], (tk / NS) & 1);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  if (i >= c.n_q || j + 1 >= c.nt(i)) continue;
                  mbar_wait(&s_free[i], (nqk[i] - 1) & 1);
                  tc_fence_after();
                  qk(i, tk % NS);
                  tc_commit(&s_full[i]);
                  ++nqk[i];
                  if (j + 2 == c.nt(i)) tc_commit(&q_empty[i]);
                }
                tc_commit(&kv_empty[tk % NS]);
              }
              mbar_wait(&kv_full[tv % NS], (tv / NS) & 1);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                if (i >= c.n_q || j >= c.nt(i)) continue;
                if (j == 0 && ni[i] > 0) mbar_wait(&o_free[i], (ni[i] - 1) & 1);
                mbar_wait(&p_full[2 * i], nt[i] & 1);
                tc_fence_after();
                if (FA_P_HALVES) {
                  pv(i, tv % NS, j > 0 ? 1u : 0u, 0, KS);
                  mbar_wait(&p_full[2 * i + 1], nt[i] & 1);
                  tc_fence_after();
                  pv(i, tv % NS, 1u, KS, KT);
                } else {
                  pv(i, tv % NS, j > 0 ? 1u : 0u, 0, KT);
                }
                ++nt[i];
                tc_commit(&pv_done[i]);
                if (j + 1 == c.nt(i)) tc_commit(&o_done[i]);
              }
              tc_commit(&kv_empty[tv % NS]);
            }
            tt = t0 + 2 * c.n_tiles;
            ++ni[0];
            if (c.n_q > 1) ++ni[1];
          }
        } else
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
          const ItemCoord c = decode_item<SPLIT>(item, p);
          const int t0 = tt;   // ring index of K_0 of this item
          mbar_wait(&kv_full[t0 % NS], (t0 / NS) & 1);
#pragma unroll
          for (int i = 0; i < 2; ++i) {  // compile-time i: the per-tile counters stay in registers
            if (i >= c.n_q) continue;
            mbar_wait(&q_full[i], ni[i] & 1);
            tc_fence_after();
            qk(i, t0 % NS);
            tc_commit(&s_full[i]);
            if (c.nt(i) == 1) tc_commit(&q_empty[i]);
          }
          tc_commit(&kv_empty[t0 % NS]);  // K_0: both QK(0) are issued by now
          for (int j = 0; j < c.n_tiles; ++j) {
            const int tv = t0 + 2 * j + 1, tk = t0 + 2 * j + 2;
            bool k_waited = false;
            mbar_wait(&kv_full[tv % NS], (tv / NS) & 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (i >= c.n_q || j >= c.nt(i)) continue;  // causal: tile 0 stops one KV tile before tile 1
              if (j == 0 && ni[i] > 0) {
                // PV_i(0) overwrites O_i: the previous item's epilogue must have read it out of TMEM
                mbar_wait(&o_free[i], (ni[i] - 1) & 1);
              }
              mbar_wait(&p_full[2 * i], nt[i] & 1);
              FA_TR(2, nt[i], 4 * i + 0, 0);
              tc_fence_after();
              if (FA_P_HALVES) {
                pv(i, tv % NS, j > 0 ? 1u : 0u, 0, KS);
                FA_TR(2, nt[i], 4 * i + 1, 0);
                mbar_wait(&p_full[2 * i + 1], nt[i] & 1);
                FA_TR(2, nt[i], 4 * i + 2, 0);
                tc_fence_after();
                pv(i, tv % NS, 1u, KS, KT);
              } else {
                pv(i, tv % NS, j > 0 ? 1u : 0u, 0, KT);
              }
              ++nt[i];
              if (j + 1 < c.nt(i)) {
                if (!k_waited) {
                  mbar_wait(&kv_full[tk % NS], (tk / NS) & 1);
                  tc_fence_after();
                  k_waited = true;
                }
                qk(i, tk % NS);
                tc_commit(&s_full[i]);
                FA_TR(2, nt[i] - 1, 4 * i + 3, 0);
                if (j + 2 == c.nt(i)) tc_commit(&q_empty[i]);  // that was the last QK_i of this item
              } else {
                tc_commit(&o_done[i]);
              }
            }
            tc_commit(&kv_empty[tv % NS]);
            if (j + 1 < c.n_tiles) tc_commit(&kv_empty[tk % NS]);
          }
          tt = t0 + 2 * c.n_tiles;
          ++ni[0];
          if (c.n_q > 1) ++ni[1];
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ===================================== softmax warpgroups =================================
    const int i = warp >> 2;  // which Q tile
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t tS = t_lane + T::TM_S + i * BN;
    const uint32_t tO = t_lane + T::TM_O + i * D;
    const uint32_t tP = t_lane + T::TM_P + i * T::P_STRIDE;  // 16-bit P_i (aliases S_i unless SEP_P)
    uint8_t* sO = sOut + i * BLK_BYTES;
    const uint32_t sO_addr = smem_u32(sO);
    const bool storer = ((warp & 3) == 0) && (lane == 0);
    int nt = 0;  // KV tiles processed (phase of s_full / p_full), across items
    int ni = 0;  // items processed (phase of o_done)

    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const ItemCoord c = decode_item<SPLIT>(item, p);
      if (i >= c.n_q) continue;  // this item has a single Q tile
      float m_used = -CUDART_INF_F;
      float l = 0.f;

      const int my_tiles = c.nt(i);
      const int row_q = c.q_row0 + i * BM + row;   // my query row inside the head
      for (int j = 0; j < my_tiles; ++j, ++nt) {
        // (Tried: __nanosleep(150..300) here when P aliases S — S_i(j) cannot be ready before PV_i(j-1) + QK_i(j), >= 768
        //  tensor cycles after this warp published P — so the warp would not poll beside the other tile's warp: +-0.)
        mbar_wait(&s_full[i], nt & 1);
        if (row == 0) FA_TR(i, nt, 0, 0);
        tc_fence_after();
        uint32_t s[4][32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld32(tS + cc * 32, s[cc]);
        tc_wait_ld();
        if (row == 0) FA_TR(i, nt, 1, s[3][31]);
        if constexpr (T::SEP_P) {
          tc_fence_before();
          mbar_arrive(&s_free[i]);  // S_i(j) is in registers: QK_i(j+1) may overwrite it while this tile's exp runs
        }

        // valid = number of leading columns of this tile my row may attend to: the ragged end of the key range and,
        // when causal, the diagonal (key index <= query row); only the last tile of a row can be cut.
        int valid = c.kv_end - (c.kv_begin + j * BN);
        if (p.causal) valid = min(valid, row_q - j * BN + 1);
        const bool full_tile = __all_sync(0xffffffffu, valid >= BN);
        auto row_max = [&]() {   // four independent FMNMX3 streams
          float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
          for (int x = 0; x < 32; ++x) {
            mx0 = fmaxf(mx0, __uint_as_float(s[0][x]));
            mx1 = fmaxf(mx1, __uint_as_float(s[1][x]));
            mx2 = fmaxf(mx2, __uint_as_float(s[2][x]));
            mx3 = fmaxf(mx3, __uint_as_float(s[3][x]));
          }
          return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        };
        // Moves the running max to mx on the rows that `need` it and rescales l and O_i (in TMEM) on those rows.
        // PV_i(j-1) has retired, so O_i is quiescent: without SEP_P s_full[i](j) implies it (same issuing thread, in-order
        // pipe); with SEP_P the pv_done wait below does.
        auto rescale = [&](float mx, bool need) {
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
        };
        // p = 2^(s*scale_log2 - m*scale_log2): column blocks c0..c1-1 of s, four independent streams per step
        float2 lsum[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
        auto exp_blocks = [&](int c0, int c1, float neg_m) {
#pragma unroll
          for (int x = 0; x < 32; x += 2) {
#pragma unroll
            for (int cc = c0; cc < c1; ++cc) {
              float2 v = make_float2(__uint_as_float(s[cc][x]), __uint_as_float(s[cc][x + 1]));
#if FA_PACKED
              v = __ffma2_rn(v, make_float2(p.scale_log2, p.scale_log2), make_float2(neg_m, neg_m));
#else
              v.x = fmaf(v.x, p.scale_log2, neg_m);
              v.y = fmaf(v.y, p.scale_log2, neg_m);
#endif
              // MUFU does 16 ex2/clk/SM, as many cycles per KV tile as the tensor pipe needs at d=128 and 2-4x more
              // at d<=64, so every FA_POLY_MOD-th element pair is evaluated on the FMA pipes instead.
              // (16-bit d = 64 — P in its own TMEM columns, QK ahead of the softmax — does best with one pair in eight:
              //  +6.5 % at L = 8192, profiles/r2_poly_ab.txt; every other instantiation with one in FA_POLY_MOD)
              constexpr int PM = (FA_POLY_MOD == 4 && D == 64 && DT != DT_F32) ? 8 : FA_POLY_MOD;
              if (PM > 0 && ((x >> 1) % (PM > 0 ? PM : 1)) == PM - 1) {
                v = exp2_poly2(v);
              } else {
                v.x = ex2_approx(v.x);
                v.y = ex2_approx(v.y);
              }
#if FA_PACKED
              lsum[cc] = __fadd2_rn(lsum[cc], v);
#else
              lsum[cc].x += v.x;
              lsum[cc].y += v.y;
#endif
              s[cc][x] = __float_as_uint(v.x);
              s[cc][x + 1] = __float_as_uint(v.y);
            }
          }
        };
        auto pack_block = [&](int cc, uint32_t* pk) {   // 32 fp32 P values of block cc -> 16 packed 16-bit pairs
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const float a = __uint_as_float(s[cc][2 * x]), b = __uint_as_float(s[cc][2 * x + 1]);
            pk[x] = (DT == DT_BF16) ? pack_bf16x2(a, b) : pack_f16x2(a, b);
          }
        };
        auto store_p = [&](int c0, int c1) {  // P key blocks c0 .. c1-1 -> TMEM (in place over S unless SEP_P)
          if constexpr (DT == DT_F32) {
#pragma unroll
            for (int cc = c0; cc < c1; ++cc) tmem_st32(tS + cc * 32, s[cc]);
          } else {   // 16 packed columns per key block
#pragma unroll
            for (int cc = c0; cc < c1; ++cc) {
              if ((cc & 1) == 0 && cc + 1 < c1) {
                uint32_t pk[32];
                pack_block(cc, pk);
                pack_block(cc + 1, pk + 16);
                tmem_st32(tP + cc * 16, pk);
              } else if ((cc & 1) == 1 && cc > c0) {
                // second block of a pair stored above
              } else {
                uint32_t pk[16];
                pack_block(cc, pk);
                tmem_st16(tP + cc * 16, pk);
              }
            }
          }
        };
        constexpr int P1 = FA_P_HALVES ? FA_P_SPLIT : 4;   // key blocks in the first published part of P

        {
          // Row max.  Only the last tile of a key range can be ragged; its masking (128 compare+select pairs) lives in
          // its own branch together with a copy of the max tree so the compiler cannot if-convert it into every tile.
          float mx;
          if (full_tile) {
            mx = row_max();
          } else {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
              for (int x = 0; x < 32; ++x)
                if (cc * 32 + x >= valid) s[cc][x] = __float_as_uint(-CUDART_INF_F);
            mx = row_max();
            asm volatile("" ::: "memory");  // keep this a real branch
          }
          if (row == 0) FA_TR(i, nt, 2, __float_as_uint(mx));
          if constexpr (T::SEP_P) {
            // QK_i(j) was issued ahead of PV_i(j-1) here, so s_full no longer implies that PV retired: wait for it before
            // O_i is rescaled or P_i overwritten (it was issued a whole tile period ago; this does not spin in steady state).
            if (nt > 0) {
              mbar_wait(&pv_done[i], (nt - 1) & 1);
              tc_fence_after();
            }
          }
          if (j == 0) {
            m_used = mx;
          } else {
            // Lazy rescale: keep the stale max unless the new one is > 2^8 larger (in exp2 units).
            const bool need = (mx - m_used) * p.scale_log2 > kRescaleThreshold;
            if (__any_sync(0xffffffffu, need)) rescale(mx, need);
          }
          exp_blocks(0, P1, -m_used * p.scale_log2);
        }
        const float neg_m = -m_used * p.scale_log2;
        if (row == 0) FA_TR(i, nt, 3, s[P1 - 1][31]);
        store_p(0, P1);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[2 * i]);
        if (row == 0) FA_TR(i, nt, 4, 0);
        if (P1 < 4) {
          exp_blocks(P1, 4, neg_m);
          if (row == 0) FA_TR(i, nt, 5, s[3][31]);
          store_p(P1, 4);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&p_full[2 * i + 1]);
          if (row == 0) FA_TR(i, nt, 6, 0);
        }
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
      mbar_arrive(&o_free[i]);  // O_i is in registers: the MMA warp may start the next item's PV_i
      const float inv_l = 1.0f / l;
      const int row_g = c.q_row0 + i * BM + row;
      if (!SPLIT && p.lse_out != nullptr && row_g < p.L)
        p.lse_out[size_t(c.bh) * p.L + row_g] = m_used * p.scale + __logf(l);
      if constexpr (SPLIT) {
        if (row_g < p.L) {
          const size_t ridx = (size_t(c.split) * p.BH + c.bh) * p.out_head_rows + row_g;
          p.lse_accum[ridx] = m_used * p.scale + __logf(l);
          float* dst = p.o_accum + ridx * D;
#pragma unroll
          for (int cc = 0; cc < D / 32; ++cc)
#pragma unroll
            for (int x = 0; x < 32; x += 4) {
              float4 v = make_float4(__uint_as_float(o[cc][x]) * inv_l, __uint_as_float(o[cc][x + 1]) * inv_l,
                                     __uint_as_float(o[cc][x + 2]) * inv_l, __uint_as_float(o[cc][x + 3]) * inv_l);
              *reinterpret_cast<float4*>(dst + cc * 32 + x) = v;
            }
        }
      } else {
        // One 128-byte column block at a time through this warpgroup's staging block (same 128B swizzle), TMA store.
#pragma unroll
        for (int b = 0; b < NBLK; ++b) {
          if (storer) tma_store_wait_read_all();  // the previous store has finished reading the staging block
          named_bar_sync(1 + i, 128);
#pragma unroll
          for (int u = 0; u < T::SWB / 16; ++u) {  // 16-byte chunks of one block row
            uint4 v;
            if constexpr (DT == DT_F32) {
              const int e = b * BLK_ELEMS + u * 4;  // output column of the first element of this chunk
              v.x = __float_as_uint(__uint_as_float(o[e / 32][e % 32 + 0]) * inv_l);
              v.y = __float_as_uint(__uint_as_float(o[e / 32][e % 32 + 1]) * inv_l);
              v.z = __float_as_uint(__uint_as_float(o[e / 32][e % 32 + 2]) * inv_l);
              v.w = __float_as_uint(__uint_as_float(o[e / 32][e % 32 + 3]) * inv_l);
            } else {
              const int e = b * BLK_ELEMS + u * 8;
              auto pk2 = [&](int ee) {
                const float a0 = __uint_as_float(o[ee / 32][ee % 32]) * inv_l;
                const float a1 = __uint_as_float(o[ee / 32][ee % 32 + 1]) * inv_l;
                return (DT == DT_BF16) ? pack_bf16x2(a0, a1) : pack_f16x2(a0, a1);
              };
              v.x = pk2(e + 0);
              v.y = pk2(e + 2);
              v.z = pk2(e + 4);
              v.w = pk2(e + 6);
            }
            // 128B swizzle: chunk ^= row % 8; 64B swizzle: chunk ^= (row / 2) % 4   (address bits [4,7) ^ [7,10))
            const int sw = (T::SWB == 128) ? (row & 7) : ((row >> 1) & 3);
            st_shared_v4(sO_addr + row * T::SWB + ((u ^ sw) << 4), v);
          }
          fence_proxy_async_smem();
          named_bar_sync(1 + i, 128);
          if (storer) {
            tma_store_3d(&tmO, sO, b * BLK_ELEMS, c.q_row0 + i * BM, c.bh);
            tma_store_commit();
          }
        }
      }
    }
    if (!SPLIT && storer) tma_store_wait_read_all();  // smem must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, T::TMEM_COLS);
}

}  // namespace fa
