/* fa_b200.h — C ABI of libfa_b200.so, the B200 (sm_100a) flash-attention forward path.
 *
 * Every entry point is the drop-in for one host launcher of tyler-utah/exploring_flash_attention
 * (paths below are relative to that repository).  Pointers are plain device (or, for the *_host
 * variants, host) pointers to contiguous row-major [B,H,L,d] tensors, exactly the reference layout
 * (base offset (b*H+h)*L*d, flash_attention_v1/CUDA/flash_attention_v1.h:182).  No torch types.
 *
 * Differences from the reference launchers, all deliberate:
 *   - return an int status instead of assert()/void  (reference: flash_attention_v1.h:263-264);
 *   - asynchronous on the caller's stream, no per-call cudaGetDeviceProperties / cudaDeviceSynchronize
 *     (reference: flash_attention_v1.h:280-292);
 *   - head dim and dtype are runtime arguments (reference: compile-time -DD, -DUSE_FP64);
 *   - V2 workspace is caller-owned (reference cudaMalloc/cudaFree per call, flash_attention_v2.h:461-463,506-508);
 *   - 64-bit element offsets.
 * There is no CPU fallback anywhere in this library.
 */
#ifndef FA_B200_H
#define FA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types of Q, K, V, O */
#define FA_DTYPE_F32 0  /* fp32 storage, tf32 tensor-core products, fp32 accumulation */
#define FA_DTYPE_BF16 1 /* bf16 storage, fp32 accumulation */
#define FA_DTYPE_F16 2  /* fp16 storage, fp32 accumulation (the reference's DATA_TYPE=__half) */

#define FA_DTYPE_F64 3  /* fa_naive_attention only: fp64 storage and math (the reference's float64 oracle runs) */

/* status codes */
#define FA_OK 0
#define FA_ERR_SHAPE (-1)         /* non-positive B/H/L/d, bad tile hints */
#define FA_ERR_DTYPE (-2)         /* unknown dtype */
#define FA_ERR_ALIGN (-3)         /* base pointer not 16-byte aligned / NULL */
#define FA_ERR_UNSUPPORTED_D (-4) /* head dim not served for this dtype */
#define FA_ERR_CUDA (-5)          /* CUDA runtime / driver error, see fa_last_error() */
#define FA_ERR_WORKSPACE (-6)     /* workspace too small */

/* Thread-local description of the last non-zero status returned on this thread. */
const char* fa_last_error(void);

/* Library / device facts: returns sm count of the current device (<=0 on error). */
int fa_device_sm_count(void);

/* ---- V1 -------------------------------------------------------------------------------------
 * Replaces  void flash_attention_v1(const DATA_TYPE* Q, K, V, DATA_TYPE* O, int B, int H, int L, int d_runtime)
 *           flash_attention_v1/CUDA/flash_attention_v1.h:251-293  and  flash_attention_v1_opt1(...)
 *           flash_attention_v1/CUDA/flash_attention_v1_opt1.h:354-396.
 * d in {32,64,128} for 16-bit dtypes, {32,64} for FA_DTYPE_F32; d in {256,512} (16-bit) and {128,256} (fp32) are
 * routed to the tiled-d kernel. */
int fa_v1_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d, int dtype,
                  void* stream /* cudaStream_t */);

/* ---- V1, extended (the reference lists "dynamic sequence lengths and causal masking" as future work,
 * flash_attention_v1/README_v1.md:169).  Same kernel as fa_v1_forward plus:
 *   LSE   optional [B*H*L] fp32 output: log(sum_j exp(q_i.k_j / sqrt(d))) per query row (NULL to skip);
 *   flags FA_FLAG_CAUSAL: query row i attends to keys 0..i only (KV tiles above the diagonal are never loaded).
 * Every (d, dtype) fa_v1_forward serves; rows of 512-1024 bytes run on the tiled-d kernels (16-bit d = 512 on the CTA-
 * pair kernel, 16-bit d = 256 and fp32 d = 128/256 on the slab kernel). */
#define FA_FLAG_CAUSAL 1u
int fa_v1_forward_ex(const void* Q, const void* K, const void* V, void* O, float* LSE, int B, int H, int L, int d,
                     int dtype, unsigned flags, void* stream);

/* Key-padding mask and rectangular attention ("dynamic sequence lengths"):
 *   Q, O [B,H,Lq,d];  K, V [B,H,Lk,d];
 *   kv_lens optional device int32[B]: every head of batch entry b attends to its first kv_lens[b] keys only
 *           (values are clamped to [1, Lk]; NULL = all Lk keys).  KV tiles past the length are never loaded.
 *   FA_FLAG_CAUSAL needs Lq == Lk.  Every (d, dtype) fa_v1_forward serves. */
int fa_v1_forward_varlen(const void* Q, const void* K, const void* V, void* O, float* LSE, const int* kv_lens, int B,
                         int H, int Lq, int Lk, int d, int dtype, unsigned flags, void* stream);

/* One un-merged partial of attention (the multi-GPU generalisation of V2's split-KV,
 * flash_attention_v2/README.md:5-21): queries [B,H,Lq,d] against ONE shard of keys/values [B,H,Lk,d].
 *   Opartial   [B*H*Lq*d] fp32, normalised by this shard's own row sums;
 *   LSEpartial [B*H*Lq]   fp32 log-sum-exp of the scaled scores over this shard.
 * N such partials stored back to back ([N][B*H][Lq][d] / [N][B*H][Lq]) are exactly fa_v2_combine's input.
 *   q_head_rows / kv_head_rows / out_head_rows: rows between consecutive (b,h) heads of Q / K,V / the partial buffers,
 *           0 = dense.  With the row offset folded into the pointers this addresses a row window of taller tensors
 *           (e.g. the late half of the local queries against the early half of a neighbour's keys in a zig-zag ring).
 *   flags   FA_FLAG_CAUSAL (needs Lq == Lk): row i attends to keys 0..i of this shard (the diagonal block). */
int fa_partial_forward(const void* Q, const void* K, const void* V, float* Opartial, float* LSEpartial, int B, int H,
                       int Lq, int Lk, int d, int dtype, long long q_head_rows, long long kv_head_rows,
                       long long out_head_rows, unsigned flags, void* stream);

/* ---- V1 tiled-d -----------------------------------------------------------------------------
 * Replaces  void flash_attention_v1[_opt](..., int d_runtime, int d_tile_qk_runtime, int d_tile_v_runtime)
 *           flash_attention_v1_tiled_d/CUDA/flash_attention_v1.h:312-354, flash_attention_v1_opt.h:448-490.
 * d_tile_qk / d_tile_v are validated like the reference (positive, divide d) and are streaming-chunk
 * hints: they change scheduling only, never results. */
int fa_v1_tiled_d_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d,
                          int d_tile_qk, int d_tile_v, int dtype, void* stream);

/* The same contract served by the CTA-pair kernel (two SMs share one 128-row query tile through 2-CTA tensor-core
 * MMAs, so no score tile is computed twice at d = 512): 16-bit dtypes, d in {256, 512}.  fa_v1_tiled_d_forward
 * routes 16-bit d = 512 to it BY DEFAULT (environment variable FA_B200_TILED_D_PAIR: 1 = the default, 2 = d = 256 as
 * well, 0 = never: the single-CTA slab kernel; any other value is ignored with a warning). */
int fa_v1_tiled_d_pair_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d,
                               int dtype, void* stream);

/* ---- V2 split-KV ----------------------------------------------------------------------------
 * Replaces  void flash_attention_v2(Q,K,V,O,B,H,L,d,d_tile_qk,d_tile_v,kv_tiles_per_block)
 *           flash_attention_v2/CUDA/flash_attention_v2.h:438-509 (partial_attention_kernel :243-341,
 *           reduction_kernel :356-435).
 * A split covers kv_per_split consecutive keys (= BK_ref * kv_tiles_per_block in reference terms);
 * n_splits = ceil(L / kv_per_split).  Workspace layout (split-major, fp32):
 *   Oaccum   [n_splits][B*H][L][d]  each split normalised by its own l
 *   LSEaccum [n_splits][B*H][L]     m/sqrt(d) + ln(l)
 * which carries the same information as the reference's (O_unnormalised, m, l) triple
 * (flash_attention_v2.h:321-340).  Every (d, dtype) fa_v1_forward serves: rows of at most 256 bytes on the fused-tile
 * kernel, 16-bit d = 256/512 and fp32 d = 128/256 (the reference V2's default D = 128 in USE_FP64 mode) on the slab
 * tiled-d kernel. */
int fa_v2_num_splits(int L, int kv_per_split);
size_t fa_v2_workspace_bytes(int B, int H, int L, int d, int kv_per_split);
int fa_v2_splitkv_forward(const void* Q, const void* K, const void* V, float* Oaccum, float* LSEaccum, int B, int H,
                          int L, int d, int kv_per_split, int dtype, void* stream);
int fa_v2_combine(const float* Oaccum, const float* LSEaccum, void* O, int B, int H, int L, int d, int n_splits,
                  int dtype, void* stream);
/* split-KV + combine with a caller-owned workspace of at least fa_v2_workspace_bytes() bytes */
int fa_v2_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d, int kv_per_split,
                  int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* ---- backward (the reference has no backward pass: README.md:80-84 lists only "Flash Attention V3"
 * as future work).  Gradients of O = softmax(Q K^T / sqrt(d)) V with respect to Q, K, V given dO:
 *   Q, K, V, O, dO  [B,H,L,d] (O and LSE as produced by fa_v1_forward_ex with the same flags);  LSE [B*H*L] fp32;
 *   dQ, dK, dV      [B,H,L,d] outputs in the same dtype;  flags: FA_FLAG_CAUSAL.
 * The probabilities are recomputed tile by tile from Q, K and LSE (never stored).  workspace: caller-owned,
 * fa_v1_backward_workspace_bytes() bytes, 256-byte aligned (row statistics).  bf16 / fp16, d in {64, 128}. */
size_t fa_v1_backward_workspace_bytes(int B, int H, int L);
int fa_v1_backward(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* LSE, void* dQ,
                   void* dK, void* dV, int B, int H, int L, int d, int dtype, unsigned flags, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- independent evaluation ---------------------------------------------------------------
 * Replaces  naive_attention(Q, K, V)  common/reference.py:7-21 — the reference's own ORACLE, which its scripts compare
 * every kernel with.  It therefore shares nothing with the kernels above: the [Lq x Lk] score matrix is materialised
 * in `workspace` (scores = Q K^T / sqrt(d); row softmax; O = probs V, exactly reference.py:16-21), on the CUDA cores
 * in full fp32 (FA_DTYPE_F32) or fp64 (FA_DTYPE_F64) — no tensor cores, no tf32, no online softmax.  Any d, any Lq, Lk.
 * Q [n_heads][Lq][d]; K, V [n_heads][Lk][d]; O like Q.  Heads are processed in groups of as many score matrices as the
 * workspace holds (at least one: fa_naive_attention_workspace_bytes(1, Lq, Lk, dtype)). */
size_t fa_naive_attention_workspace_bytes(int n_heads, int Lq, int Lk, int dtype);
int fa_naive_attention(const void* Q, const void* K, const void* V, void* O, int n_heads, int Lq, int Lk, int d,
                       int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* ---- host-buffer convenience (what the reference drivers do around their launchers:
 * cudaMalloc + cudaMemcpy H2D x3 + launch + cudaMemcpy D2H, flash_attention_v1/CUDA/driver.cu:184-247).
 * Q,K,V,O are HOST pointers (pinned for full PCIe speed); device staging buffers are cached inside the
 * library between calls.  variant: 0 = V1, 1 = tiled-d, 2 = V2 (kv_per_split used).  Synchronous. */
int fa_forward_host(int variant, const void* Qh, const void* Kh, const void* Vh, void* Oh, int B, int H, int L, int d,
                    int kv_per_split, int dtype);
void fa_release_host_staging(void);   /* frees the CURRENT device's staging buffers (one cached set per device) */

/* Strided block copy on the copy engines (cudaMemcpy2DAsync, cudaMemcpyDefault): `height` rows of `width` bytes.  Used
 * by the sequence-parallel all-to-all path (sharding.alltoall_attention) to move [heads][rows][d] blocks between peer-
 * mapped buffers without taking SMs from the attention kernel. */
int fa_copy_2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* stream);
/* n copies of the same block shape and pitches in one call (the exchange issues 8-32 per step). */
int fa_copy_2d_multi_async(int n, void* const* dst, size_t dpitch, const void* const* src, size_t spitch, size_t width,
                           size_t height, void* stream);

/* n contiguous copies of `bytes` each in one call (cudaMemcpyAsync): what sharding.alltoall_attention pulls peer blocks
 * with — 1-D copies run on the copy engines beside the persistent attention kernel, strided ones did not. */
int fa_copy_multi_async(int n, void* const* dst, const void* const* src, size_t bytes, void* stream);

/* Diagnostics: how many TMA tensor maps were served from the per-thread cache / had to be encoded by the driver since
 * the library was loaded (launching on the same buffers step after step must not re-encode; the reference re-derives
 * its launch state on every call, flash_attention_v1.h:280-292).  Either pointer may be NULL. */
void fa_debug_map_cache_stats(unsigned long long* hits, unsigned long long* misses);

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H */
