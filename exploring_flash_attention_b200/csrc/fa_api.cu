// fa_api.cu — the C ABI declared in include/fa_b200.h: argument validation, TMA tensor-map
// construction, kernel dispatch.  Host-side counterpart of the reference launchers
// (flash_attention_v1/CUDA/flash_attention_v1.h:251-293, flash_attention_v1_tiled_d/CUDA/flash_attention_v1.h:312-354,
// flash_attention_v2/CUDA/flash_attention_v2.h:438-509) without their per-call device queries, allocations and syncs.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>

#include "../../include/fa_b200.h"
#include "fa_combine_sm100.cuh"
#include "fa_bwd_sm100.cuh"
#include "fa_fwd_sm100.cuh"
#include "fa_naive_sm100.cuh"
#include "fa_splitkv_sm100.cuh"
#include "fa_tiled_d_pair_sm100.cuh"
#include "fa_tiled_d_sm100.cuh"

#ifndef FA_TILED_D_PAIR_DEFAULT
#define FA_TILED_D_PAIR_DEFAULT 1   // 1: 16-bit d = 512 runs on CTA pairs (B200: 1036 vs 830 TFLOP/s for the slab kernel at
                                    // B16 H8 L4096); d = 256 stays on the slab kernel (1340 vs 680)
#endif

#ifdef FA_TRACE
// Tuning builds only (never the shipped library): device buffer the fused-tile kernel's CTA 0 writes SM-clock stamps to.
static long long* g_trace = nullptr;
extern "C" void fa_debug_set_trace(void* dev_buf) { g_trace = static_cast<long long*>(dev_buf); }
#endif

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define FA_CUDA_TRY(expr)                                                                           \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return fail(FA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

int elem_size(int dtype) { return dtype == FA_DTYPE_F32 ? 4 : 2; }

// SM count of the current device, queried once per device (not per call like the reference's
// cudaGetDeviceProperties, flash_attention_v1.h:280-281).  The per-device slots are atomics: two host threads racing on
// the first call both store the same value.
constexpr int kMaxDevices = 64;

int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

int sm_count() {
  static std::atomic<int> cached[kMaxDevices];
  const int dev = current_device();
  if (dev < 0) return -1;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return -1;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device) instead of on every launch:
// each launch_* template instantiation owns one of these.
struct SmemAttrOnce {
  std::atomic<unsigned long long> done{0};
  template <typename Kern>
  int ensure(Kern kern, int bytes) {
    const int dev = current_device();
    if (dev < 0) return fail(FA_ERR_CUDA, "cannot query the current device");
    if ((done.load(std::memory_order_acquire) >> dev) & 1ull) return FA_OK;
    FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.fetch_or(1ull << dev, std::memory_order_release);
    return FA_OK;
  }
};

// [BH][L][D] row-major tensor, box = one 128-byte-wide, `box_rows`-row block of one head, 128B swizzle.
// Rows past L are zero-filled on load and clipped on store, so tiles never leak into the next head.
int encode_map(CUtensorMap* m, const void* ptr, int dtype, int D, int L, int BH, int box_rows, bool mn_major_operand,
               long long head_stride_rows) {
  EncodeFn enc = get_encode_fn();
  if (!enc) return fail(FA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int es = elem_size(dtype);
  CUtensorMapDataType dt = dtype == FA_DTYPE_F32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : dtype == FA_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                    : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint64_t dims[3] = {cuuint64_t(D), cuuint64_t(L), cuuint64_t(BH)};
  cuuint64_t strides[2] = {cuuint64_t(D) * es, cuuint64_t(head_stride_rows > 0 ? head_stride_rows : L) * D * es};
  const int swb = (D * es >= 128) ? 128 : 64;  // swizzle span = bytes of one block row (64 only for 16-bit d = 32)
  cuuint32_t box[3] = {cuuint32_t(swb / es), cuuint32_t(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  // 32-bit MN-major UMMA operands (fp32 V) must use the 32B-atom flavour of the 128B swizzle.
  const CUtensorMapSwizzle swz = (swb == 64)                      ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : (mn_major_operand && es == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                                 : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
  return FA_OK;
}

// A tensor map is a pure function of (pointer, dtype, shape, box, layout): callers that launch on the same buffers step
// after step (every training / serving loop) get it from a small per-thread direct-mapped cache instead of a driver call
// per operand per launch (the reference re-derives all launch state per call, flash_attention_v1.h:280-292).
// 32 sets x 4 ways, round-robin replacement inside a set.
struct MapKey {
  const void* ptr;
  long long head_stride_rows;
  int dtype, D, L, BH, box_rows, mn_major;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && head_stride_rows == o.head_stride_rows && dtype == o.dtype && D == o.D && L == o.L &&
           BH == o.BH && box_rows == o.box_rows && mn_major == o.mn_major;
  }
};
struct MapSlot {
  MapKey key{};
  bool valid = false;
  alignas(64) CUtensorMap map;
};
constexpr int kMapCacheSets = 32, kMapCacheWays = 4;   // 4-way sets: the 4-7 operands of one launch never evict each other
std::atomic<unsigned long long> g_map_hits{0}, g_map_misses{0};

int make_map(CUtensorMap* m, const void* ptr, int dtype, int D, int L, int BH, int box_rows, bool mn_major_operand = false,
             long long head_stride_rows = 0 /* rows between consecutive heads; 0 = L (dense) */) {
  thread_local MapSlot cache[kMapCacheSets][kMapCacheWays];
  thread_local unsigned char next_way[kMapCacheSets] = {};
  const MapKey key{ptr, head_stride_rows, dtype, D, L, BH, box_rows, mn_major_operand ? 1 : 0};
  unsigned long long h = reinterpret_cast<uintptr_t>(ptr) >> 8;
  h = (h ^ (h >> 17)) * 0x9E3779B97F4A7C15ull + (unsigned long long)(box_rows * 31 + (mn_major_operand ? 7 : 0) + L * 131 + BH);
  const int set = int((h >> 20) % kMapCacheSets);
  for (int w = 0; w < kMapCacheWays; ++w) {
    MapSlot& slot = cache[set][w];
    if (slot.valid && slot.key == key) {
      *m = slot.map;
      g_map_hits.fetch_add(1, std::memory_order_relaxed);
      return FA_OK;
    }
  }
  const int rc = encode_map(m, ptr, dtype, D, L, BH, box_rows, mn_major_operand, head_stride_rows);
  if (rc != FA_OK) return rc;
  MapSlot& slot = cache[set][next_way[set]];
  next_way[set] = (unsigned char)((next_way[set] + 1) % kMapCacheWays);
  slot.key = key;
  slot.map = *m;
  slot.valid = true;
  g_map_misses.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

int check_common(const void* Q, const void* K, const void* V, const void* O, int B, int H, int L, int d, int dtype) {
  if (B <= 0 || H <= 0 || L <= 0 || d <= 0) return fail(FA_ERR_SHAPE, "B, H, L, d must be positive");
  if (dtype != FA_DTYPE_F32 && dtype != FA_DTYPE_BF16 && dtype != FA_DTYPE_F16)
    return fail(FA_ERR_DTYPE, "dtype must be FA_DTYPE_F32, FA_DTYPE_BF16 or FA_DTYPE_F16");
  const void* ptrs[4] = {Q, K, V, O};
  for (const void* p : ptrs)
    if (p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) != 0)
      return fail(FA_ERR_ALIGN, "Q, K, V, O must be non-null and 16-byte aligned");
  return FA_OK;
}

// Optional arguments of the fused-tile kernel beyond the reference's (Q,K,V,O,B,H,L,d).
struct FwdExtra {
  int Lk = 0;                    // keys per head when it differs from the query count (0: same as L)
  int H = 0;                     // heads per batch entry (needed with kv_lens)
  const int* kv_lens = nullptr;  // device [B] int32 key-padding lengths
  // Row windows of taller tensors (rows between consecutive heads; 0 = dense).  The row offset is in the pointer.
  long long q_head_rows = 0, kv_head_rows = 0, out_head_rows = 0;
};

template <int D, int DT, bool SPLIT>
int launch_fwd(const void* Q, const void* K, const void* V, void* O, int BH, int L, int kv_per_split, int n_splits,
               float* o_accum, float* lse_accum, cudaStream_t stream, float* lse_out = nullptr, int causal = 0,
               const FwdExtra& ex = FwdExtra()) {
  using T = fa::FwdTraits<D, DT>;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  const int Lk = ex.Lk > 0 ? ex.Lk : L;
  if ((rc = make_map(&tmQ, Q, DT, D, L, BH, 128, false, ex.q_head_rows)) != FA_OK) return rc;
  if ((rc = make_map(&tmK, K, DT, D, Lk, BH, 128, false, ex.kv_head_rows)) != FA_OK) return rc;
  if ((rc = make_map(&tmV, V, DT, D, Lk, BH, 128, /*mn_major_operand=*/true, ex.kv_head_rows)) != FA_OK) return rc;
  if (SPLIT) {
    tmO = tmQ;  // unused by the split epilogue
  } else if ((rc = make_map(&tmO, O, DT, D, L, BH, 128)) != FA_OK) {
    return rc;
  }
  fa::FwdParams p{};
  p.L = L;
  p.Lk = Lk;
  p.BH = BH;
  p.H = ex.H > 0 ? ex.H : 1;
  p.kv_lens = SPLIT ? nullptr : ex.kv_lens;
  p.kv_per_split = SPLIT ? kv_per_split : Lk;
  p.n_splits = SPLIT ? n_splits : 1;
  p.n_qpairs = (L + 255) / 256;
  const long long items = (long long)BH * p.n_splits * p.n_qpairs;
  if (items > 0x7fffffffLL) return fail(FA_ERR_SHAPE, "too many (head, split, q-tile) work items");
  p.n_items = int(items);
  p.scale = 1.0f / std::sqrt(float(D));
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.o_accum = o_accum;
  p.lse_accum = lse_accum;
  p.lse_out = SPLIT ? nullptr : lse_out;
  p.causal = (SPLIT && (n_splits != 1 || Lk != L)) ? 0 : causal;
  p.out_head_rows = int(ex.out_head_rows > 0 ? ex.out_head_rows : L);
#ifdef FA_TRACE
  p.trace = g_trace;
#endif
  auto kern = fa::fa_fwd_kernel<D, DT, SPLIT>;
  static SmemAttrOnce smem_attr;
  if ((rc = smem_attr.ensure(kern, T::SMEM_BYTES)) != FA_OK) return rc;
  // persistent: one CTA per SM (smem and TMEM admit exactly one), each walking items blockIdx.x, +gridDim.x, ...
  const int sms = sm_count();
  if (sms <= 0) return fail(FA_ERR_CUDA, "cannot query the SM count of the current device");
  const int grid = p.n_items < sms ? p.n_items : sms;
  kern<<<grid, T::THREADS, T::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

template <bool SPLIT>
int dispatch_fwd(const void* Q, const void* K, const void* V, void* O, int BH, int L, int d, int dtype,
                 int kv_per_split, int n_splits, float* o_accum, float* lse_accum, cudaStream_t s,
                 float* lse_out = nullptr, int causal = 0, const FwdExtra& ex = FwdExtra()) {
#define FA_CASE(DD, DTT)                                                                                         \
  if (d == DD && dtype == DTT)                                                                                   \
    return launch_fwd<DD, DTT, SPLIT>(Q, K, V, O, BH, L, kv_per_split, n_splits, o_accum, lse_accum, s, lse_out, \
                                      causal, ex);
  FA_CASE(128, fa::DT_BF16)
  FA_CASE(64, fa::DT_BF16)
  FA_CASE(128, fa::DT_F16)
  FA_CASE(64, fa::DT_F16)
  FA_CASE(32, fa::DT_BF16)
  FA_CASE(32, fa::DT_F16)
  FA_CASE(32, fa::DT_F32)
  FA_CASE(64, fa::DT_F32)
#undef FA_CASE
  return fail(FA_ERR_UNSUPPORTED_D,
              "fused-tile kernel serves d in {32,64,128} for bf16/fp16 and d in {32,64} for fp32; got d=" +
                  std::to_string(d) + " dtype=" + std::to_string(dtype));
}

// K3a': split-KV partials when every split is a single KV tile (kv_per_split <= 128), fa_splitkv_sm100.cuh.
template <int D, int DT, int BN>
int launch_splitkv_tile(const void* Q, const void* K, const void* V, int BH, int L, int kv_per_split, int n_splits,
                        float* o_accum, float* lse_accum, cudaStream_t stream) {
  using T = fa::SplitTileTraits<D, DT, BN>;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((long long)n_splits * BH > 0x7fffffffLL) return fail(FA_ERR_SHAPE, "too many (split, head) partial slabs");
  if ((rc = make_map(&tmQ, Q, DT, D, L, BH, 128)) != FA_OK) return rc;
  if ((rc = make_map(&tmK, K, DT, D, L, BH, BN)) != FA_OK) return rc;
  if ((rc = make_map(&tmV, V, DT, D, L, BH, BN, /*mn_major_operand=*/true)) != FA_OK) return rc;
  // the fp32 partials as a [n_splits*BH][L][D] tensor, stored 32 columns (one 128-byte row) x 128 rows at a time
  if ((rc = make_map(&tmO, o_accum, FA_DTYPE_F32, D, L, n_splits * BH, 128)) != FA_OK) return rc;
  fa::SplitTileParams p{};
  p.L = L;
  p.BH = BH;
  p.kv_per_split = kv_per_split;
  p.n_splits = n_splits;
  p.n_qtiles = (L + 127) / 128;
  p.n_units = (long long)BH * p.n_qtiles * n_splits;
  p.scale = 1.0f / std::sqrt(float(D));
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.lse_accum = lse_accum;
  auto kern = fa::fa_splitkv_tile_kernel<D, DT, BN>;
  static SmemAttrOnce smem_attr;
  if ((rc = smem_attr.ensure(kern, T::SMEM_BYTES)) != FA_OK) return rc;
  const int sms = sm_count();
  if (sms <= 0) return fail(FA_ERR_CUDA, "cannot query the SM count of the current device");
  const int grid = p.n_units < sms ? int(p.n_units) : sms;
  kern<<<grid, T::THREADS, T::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

int dispatch_splitkv_tile(const void* Q, const void* K, const void* V, int BH, int L, int d, int dtype, int kv_per_split,
                          int n_splits, float* o_accum, float* lse_accum, cudaStream_t s) {
#define FA_CASE(DD, DTT)                                                                                              \
  if (d == DD && dtype == DTT) {                                                                                      \
    if (kv_per_split <= 64)                                                                                           \
      return launch_splitkv_tile<DD, DTT, 64>(Q, K, V, BH, L, kv_per_split, n_splits, o_accum, lse_accum, s);         \
    return launch_splitkv_tile<DD, DTT, 128>(Q, K, V, BH, L, kv_per_split, n_splits, o_accum, lse_accum, s);          \
  }
  FA_CASE(128, fa::DT_BF16)
  FA_CASE(64, fa::DT_BF16)
  FA_CASE(32, fa::DT_BF16)
  FA_CASE(128, fa::DT_F16)
  FA_CASE(64, fa::DT_F16)
  FA_CASE(32, fa::DT_F16)
  FA_CASE(32, fa::DT_F32)
  FA_CASE(64, fa::DT_F32)
#undef FA_CASE
  return fail(FA_ERR_UNSUPPORTED_D, "split-KV tile kernel: unsupported d / dtype");
}

// FA_B200_SPLITKV_TILE=0 keeps short splits on the fused-tile kernel's SPLIT mode (A/B measurements).  Read once.
bool splitkv_tile_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("FA_B200_SPLITKV_TILE");
    return !(e && e[0] == '0' && e[1] == '\0');
  }();
  return on;
}

// K4: backward (fa_bwd_sm100.cuh): prep (Delta, LSE in log2 units) + dK/dV pass + dQ pass.
template <int D, int DT>
int launch_backward(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* LSE, void* dQ,
                    void* dK, void* dV, int BH, int L, int causal, float* ws, cudaStream_t stream) {
  using T = fa::BwdTraits<D, DT>;
  const int n_tiles = (L + 127) / 128;
  const int Lp = n_tiles * 128;
  if ((long long)BH * n_tiles > 0x7fffffffLL) return fail(FA_ERR_SHAPE, "too many (head, tile) blocks");
  float* lse2 = ws;
  float* delta = ws + size_t(BH) * Lp;
  const long long prep_rows = (long long)BH * Lp;
  fa::fa_bwd_prep_kernel<D, DT><<<unsigned((prep_rows + 7) / 8), 256, 0, stream>>>(O, dO, LSE, lse2, delta, L, Lp, BH);
  FA_CUDA_TRY(cudaGetLastError());
  CUtensorMap tQ, tK, tV, tdO, tdQ, tdK, tdV;
  int rc;
  if ((rc = make_map(&tQ, Q, DT, D, L, BH, 128)) != FA_OK) return rc;
  if ((rc = make_map(&tK, K, DT, D, L, BH, 128)) != FA_OK) return rc;
  if ((rc = make_map(&tV, V, DT, D, L, BH, 128)) != FA_OK) return rc;
  if ((rc = make_map(&tdO, dO, DT, D, L, BH, 128)) != FA_OK) return rc;
  if ((rc = make_map(&tdQ, dQ, DT, D, L, BH, 128)) != FA_OK) return rc;
  if ((rc = make_map(&tdK, dK, DT, D, L, BH, 128)) != FA_OK) return rc;
  if ((rc = make_map(&tdV, dV, DT, D, L, BH, 128)) != FA_OK) return rc;
  fa::BwdParams p{};
  p.L = L;
  p.Lp = Lp;
  p.BH = BH;
  p.causal = causal;
  p.scale = 1.0f / std::sqrt(float(D));
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.lse2 = lse2;
  p.delta = delta;
  auto k_dkv = fa::fa_bwd_kernel<D, DT, true>;
  auto k_dq = fa::fa_bwd_kernel<D, DT, false>;
  static SmemAttrOnce attr_dkv, attr_dq;
  if ((rc = attr_dkv.ensure(k_dkv, T::SMEM_BYTES)) != FA_OK) return rc;
  if ((rc = attr_dq.ensure(k_dq, T::SMEM_BYTES)) != FA_OK) return rc;
  const unsigned grid = unsigned(BH * n_tiles);
  k_dkv<<<grid, T::THREADS, T::SMEM_BYTES, stream>>>(tK, tV, tQ, tdO, tdK, tdV, p);
  FA_CUDA_TRY(cudaGetLastError());
  k_dq<<<grid, T::THREADS, T::SMEM_BYTES, stream>>>(tQ, tdO, tK, tV, tdQ, tdQ, p);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

// Optional arguments of the slab tiled-d kernel beyond the reference's dense (Q,K,V,O): key ranges (V2 splits / partials),
// causal masking, the log-sum-exp output.
struct TiledDExtra {
  int kv_per_split = 0, n_splits = 1;   // kv_per_split 0: one range covering every key
  float* o_accum = nullptr;             // non-null: fp32 partial rows + lse_accum instead of O
  float* lse_accum = nullptr;
  float* lse_out = nullptr;
  int causal = 0;
  FwdExtra fx;
};

template <int D, int DT>
int launch_tiled_d(const void* Q, const void* K, const void* V, void* O, int BH, int L, cudaStream_t stream,
                   const TiledDExtra& ex = TiledDExtra()) {
  using T = fa::TiledDTraits<D, DT>;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  const int Lk = ex.fx.Lk > 0 ? ex.fx.Lk : L;
  if ((rc = make_map(&tmQ, Q, DT, D, L, BH, 128, false, ex.fx.q_head_rows)) != FA_OK) return rc;
  if ((rc = make_map(&tmK, K, DT, D, Lk, BH, 128, false, ex.fx.kv_head_rows)) != FA_OK) return rc;
  if ((rc = make_map(&tmV, V, DT, D, Lk, BH, 128, /*mn_major_operand=*/true, ex.fx.kv_head_rows)) != FA_OK) return rc;
  if (ex.o_accum != nullptr) {
    tmO = tmQ;  // unused by the partial epilogue
  } else if ((rc = make_map(&tmO, O, DT, D, L, BH, 128)) != FA_OK) {
    return rc;
  }
  fa::FwdParams p{};
  p.L = L;
  p.Lk = Lk;
  p.BH = BH;
  p.H = ex.fx.H > 0 ? ex.fx.H : 1;
  p.kv_lens = ex.o_accum ? nullptr : ex.fx.kv_lens;
  p.kv_per_split = ex.kv_per_split > 0 ? ex.kv_per_split : Lk;
  p.n_splits = ex.n_splits;
  p.n_qpairs = 0;
  p.n_items = 0;
  p.scale = 1.0f / std::sqrt(float(D));
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.o_accum = ex.o_accum;
  p.lse_accum = ex.lse_accum;
  p.lse_out = ex.o_accum ? nullptr : ex.lse_out;
  p.causal = (ex.n_splits == 1 && Lk == L) ? ex.causal : 0;
  p.out_head_rows = int(ex.fx.out_head_rows > 0 ? ex.fx.out_head_rows : L);
  auto kern = fa::fa_tiled_d_kernel<D, DT>;
  static SmemAttrOnce smem_attr;
  if ((rc = smem_attr.ensure(kern, T::SMEM_BYTES)) != FA_OK) return rc;
  const long long blocks = (long long)((L + 127) / 128) * BH * ex.n_splits * T::NSLAB;
  if (blocks > 0x7fffffffLL) return fail(FA_ERR_SHAPE, "too many (head, split, q-tile, slab) blocks");
  dim3 grid((unsigned)blocks);
  kern<<<grid, T::THREADS, T::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

// K2P: one CTA pair per 128-row q-tile (fa_tiled_d_pair_sm100.cuh).
template <int D, int DT>
int launch_tiled_d_pair(const void* Q, const void* K, const void* V, void* O, int BH, int L, cudaStream_t stream,
                        float* lse_out = nullptr, int causal = 0) {
  using T = fa::TiledDPairTraits<D, DT>;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((rc = make_map(&tmQ, Q, DT, D, L, BH, 64)) != FA_OK) return rc;    // each CTA: its 64 query rows
  if ((rc = make_map(&tmK, K, DT, D, L, BH, 64)) != FA_OK) return rc;    // each CTA: its 64 keys of a tile
  if ((rc = make_map(&tmV, V, DT, D, L, BH, 128, /*mn_major_operand=*/true)) != FA_OK) return rc;
  if ((rc = make_map(&tmO, O, DT, D, L, BH, 64)) != FA_OK) return rc;
  fa::FwdParams p{};
  p.L = L;
  p.Lk = L;
  p.BH = BH;
  p.kv_per_split = L;
  p.n_splits = 1;
  p.scale = 1.0f / std::sqrt(float(D));
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.lse_out = lse_out;
  p.causal = causal;
  auto kern = fa::fa_tiled_d_pair_kernel<D, DT>;
  static SmemAttrOnce smem_attr;
  if ((rc = smem_attr.ensure(kern, T::SMEM_BYTES)) != FA_OK) return rc;
  const long long blocks = 2LL * ((L + 127) / 128) * BH;   // the kernel carries __cluster_dims__(2, 1, 1)
  if (blocks > 0x7fffffffLL) return fail(FA_ERR_SHAPE, "too many (head, q-tile) CTA pairs");
  kern<<<dim3((unsigned)blocks), T::THREADS, T::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

// Which kernel serves 16-bit d = 512 / 256: FA_B200_TILED_D_PAIR=1 (the default) selects the CTA-pair kernel for d = 512,
// =2 for d = 256 as well, =0 the single-CTA slab kernel.  Read once per process; any other value is ignored.
int tiled_d_pair_mode() {
  static const int mode = [] {
    const char* e = std::getenv("FA_B200_TILED_D_PAIR");
    if (e == nullptr || e[0] == '\0') return FA_TILED_D_PAIR_DEFAULT;
    if ((e[0] == '0' || e[0] == '1' || e[0] == '2') && e[1] == '\0') return e[0] - '0';
    std::fprintf(stderr, "libfa_b200: FA_B200_TILED_D_PAIR=%s ignored (accepted: 0, 1, 2)\n", e);
    return FA_TILED_D_PAIR_DEFAULT;
  }();
  return mode;
}

// The slab kernel, with its optional arguments (split ranges, causal, LSE, key padding, Lq != Lk).
int dispatch_tiled_d_slab(const void* Q, const void* K, const void* V, void* O, int BH, int L, int d, int dtype,
                          cudaStream_t s, const TiledDExtra& ex = TiledDExtra()) {
  if (d == 256 && dtype == fa::DT_BF16) return launch_tiled_d<256, fa::DT_BF16>(Q, K, V, O, BH, L, s, ex);
  if (d == 512 && dtype == fa::DT_BF16) return launch_tiled_d<512, fa::DT_BF16>(Q, K, V, O, BH, L, s, ex);
  if (d == 256 && dtype == fa::DT_F16) return launch_tiled_d<256, fa::DT_F16>(Q, K, V, O, BH, L, s, ex);
  if (d == 512 && dtype == fa::DT_F16) return launch_tiled_d<512, fa::DT_F16>(Q, K, V, O, BH, L, s, ex);
  if (d == 128 && dtype == fa::DT_F32) return launch_tiled_d<128, fa::DT_F32>(Q, K, V, O, BH, L, s, ex);
  if (d == 256 && dtype == fa::DT_F32) return launch_tiled_d<256, fa::DT_F32>(Q, K, V, O, BH, L, s, ex);
  return fail(FA_ERR_UNSUPPORTED_D, "tiled-d kernel serves d in {256,512} for bf16/fp16 and d in {128,256} for fp32; got d=" + std::to_string(d) +
                                        " dtype=" + std::to_string(dtype));
}

// Dense (Q,K,V)->O: the CTA-pair kernel where it wins (16-bit d = 512 by default), else the slab kernel.
int dispatch_tiled_d(const void* Q, const void* K, const void* V, void* O, int BH, int L, int d, int dtype,
                     cudaStream_t s) {
  const int pair_mode = tiled_d_pair_mode();
  if (pair_mode >= 1 && d == 512 && dtype == fa::DT_BF16) return launch_tiled_d_pair<512, fa::DT_BF16>(Q, K, V, O, BH, L, s);
  if (pair_mode >= 1 && d == 512 && dtype == fa::DT_F16) return launch_tiled_d_pair<512, fa::DT_F16>(Q, K, V, O, BH, L, s);
  if (pair_mode >= 2 && d == 256 && dtype == fa::DT_BF16) return launch_tiled_d_pair<256, fa::DT_BF16>(Q, K, V, O, BH, L, s);
  if (pair_mode >= 2 && d == 256 && dtype == fa::DT_F16) return launch_tiled_d_pair<256, fa::DT_F16>(Q, K, V, O, BH, L, s);
  return dispatch_tiled_d_slab(Q, K, V, O, BH, L, d, dtype, s);
}

// Does the fused-tile kernel (K1) serve this (d, dtype)?  Otherwise the row is 512-1024 bytes and the slab kernel does.
bool fused_tile_serves(int d, int dtype) { return d <= 128 && !(dtype == FA_DTYPE_F32 && d > 64); }

// The combine is launched with programmatic stream serialization: it may be scheduled while the kernel before it in the
// stream (normally the split-KV kernel, which calls griddepcontrol.launch_dependents) is draining, and waits in
// griddepcontrol.wait until that kernel's partials are complete — the ~3 us launch gap between the two short kernels of
// a V2 call disappears.  After any other kernel the attribute only means "start after it has finished", as usual.
template <typename Kern>
int launch_pdl(Kern kern, unsigned blocks, unsigned threads, cudaStream_t s, const float* o_accum, const float* lse_accum,
               void* O, long long rows, int n_splits) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, o_accum, lse_accum, O, rows, n_splits));
  return FA_OK;
}

template <int D, int DT>
int launch_combine(const float* o_accum, const float* lse_accum, void* O, long long rows, int n_splits,
                   cudaStream_t s) {
  constexpr int G = (D / 4 < 32) ? D / 4 : 32;
  constexpr int NV = D / (4 * G);
  if (n_splits <= fa::kCombineMaxSplitsFast) {
    constexpr int RPB = (fa::kCombineThreads / G) * ((NV >= 2) ? 1 : 2);
    const long long blocks = (rows + RPB - 1) / RPB;
    return launch_pdl(fa::fa_combine_kernel<D, DT>, (unsigned)blocks, fa::kCombineThreads, s, o_accum, lse_accum, O, rows, n_splits);
  }
  constexpr int RPB = fa::kCombineThreads / G;
  const long long blocks = (rows + RPB - 1) / RPB;
  return launch_pdl(fa::fa_combine_generic_kernel<D, DT>, (unsigned)blocks, fa::kCombineThreads, s, o_accum, lse_accum, O, rows, n_splits);
}

// Cached device staging for the host-buffer entry point: ONE set per device (buffers, streams and events belong to the
// device they were created on), each behind its own mutex so host threads driving different GPUs do not serialise.
struct HostStaging {
  void* buf[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t bytes = 0;
  void* ws = nullptr;
  size_t ws_bytes = 0;
#ifndef FA_HOST_CHUNKS
#define FA_HOST_CHUNKS 8
#endif
#ifndef FA_HOST_GEOMETRIC
#define FA_HOST_GEOMETRIC 1
#endif
  static constexpr int kChunks = FA_HOST_CHUNKS;  // head chunks in flight through the H2D / compute / D2H pipeline
  cudaStream_t stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_h2d[kChunks] = {}, ev_comp[kChunks] = {};
  bool streams_ready = false;
  std::mutex mu;

  void release_buffers() {
    for (auto& b : buf) {
      if (b) cudaFree(b);
      b = nullptr;
    }
    if (ws) cudaFree(ws);
    ws = nullptr;
    bytes = ws_bytes = 0;
  }
  // Nothing of a failed call may still be in flight towards the caller's host buffers when the error is returned.
  void drain() {
    if (!streams_ready) return;
    for (auto& st : stream) cudaStreamSynchronize(st);
  }
} g_stage[kMaxDevices];

template <typename T>
int naive_attention_impl(const T* Q, const T* K, const T* V, T* O, int n_heads, int Lq, int Lk, int d, T* scores,
                         size_t ws_bytes, cudaStream_t s) {
  const size_t per_head = size_t(Lq) * Lk * sizeof(T);
  const int chunk = int(std::min<size_t>(size_t(n_heads), std::min<size_t>(ws_bytes / per_head, 65535)));
  if (chunk < 1) return fail(FA_ERR_WORKSPACE, "workspace must hold the [Lq x Lk] scores of at least one head: " + std::to_string(per_head) + " bytes");
  const T alpha = T(1) / std::sqrt(T(d));
  for (int h0 = 0; h0 < n_heads; h0 += chunk) {
    const int nh = std::min(chunk, n_heads - h0);
    const T *q = Q + size_t(h0) * Lq * d, *k = K + size_t(h0) * Lk * d, *v = V + size_t(h0) * Lk * d;
    T* o = O + size_t(h0) * Lq * d;
    fa::naive_gemm_kernel<T, true><<<dim3((Lk + 63) / 64, (Lq + 63) / 64, nh), 256, 0, s>>>(
        q, k, scores, Lq, Lk, d, alpha, (long long)Lq * d, (long long)Lk * d, (long long)Lq * Lk);
    fa::naive_softmax_rows_kernel<T><<<dim3(unsigned(size_t(nh) * Lq)), 256, 0, s>>>(scores, Lk);
    fa::naive_gemm_kernel<T, false><<<dim3((d + 63) / 64, (Lq + 63) / 64, nh), 256, 0, s>>>(
        scores, v, o, Lq, d, Lk, T(1), (long long)Lq * Lk, (long long)Lk * d, (long long)Lq * d);
    FA_CUDA_TRY(cudaGetLastError());
  }
  return FA_OK;
}

}  // namespace

extern "C" {

const char* fa_last_error(void) { return g_err.c_str(); }

int fa_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

int fa_v1_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d, int dtype,
                  void* stream) {
  int rc = check_common(Q, K, V, O, B, H, L, d, dtype);
  if (rc != FA_OK) return rc;
  if (d > 128 || (dtype == FA_DTYPE_F32 && d > 64))
    return fa_v1_tiled_d_forward(Q, K, V, O, B, H, L, d, d >= 64 ? 64 : d, d >= 64 ? 64 : d, dtype, stream);
  return dispatch_fwd<false>(Q, K, V, O, B * H, L, d, dtype, L, 1, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int fa_v1_forward_ex(const void* Q, const void* K, const void* V, void* O, float* LSE, int B, int H, int L, int d,
                     int dtype, unsigned flags, void* stream) {
  int rc = check_common(Q, K, V, O, B, H, L, d, dtype);
  if (rc != FA_OK) return rc;
  if (flags & ~unsigned(FA_FLAG_CAUSAL)) return fail(FA_ERR_SHAPE, "unknown flag bits");
  if (!fused_tile_serves(d, dtype)) {
    TiledDExtra tx;
    tx.lse_out = LSE;
    tx.causal = (flags & FA_FLAG_CAUSAL) ? 1 : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (tx.lse_out == nullptr && !tx.causal) return dispatch_tiled_d(Q, K, V, O, B * H, L, d, dtype, s);
    // 16-bit d = 512: the CTA-pair kernel serves causal masking and the LSE output itself
    if (tiled_d_pair_mode() >= 1 && d == 512 && dtype == fa::DT_BF16)
      return launch_tiled_d_pair<512, fa::DT_BF16>(Q, K, V, O, B * H, L, s, tx.lse_out, tx.causal);
    if (tiled_d_pair_mode() >= 1 && d == 512 && dtype == fa::DT_F16)
      return launch_tiled_d_pair<512, fa::DT_F16>(Q, K, V, O, B * H, L, s, tx.lse_out, tx.causal);
    return dispatch_tiled_d_slab(Q, K, V, O, B * H, L, d, dtype, s, tx);
  }
  return dispatch_fwd<false>(Q, K, V, O, B * H, L, d, dtype, L, 1, nullptr, nullptr, static_cast<cudaStream_t>(stream),
                             LSE, (flags & FA_FLAG_CAUSAL) ? 1 : 0);
}

int fa_v1_forward_varlen(const void* Q, const void* K, const void* V, void* O, float* LSE, const int* kv_lens, int B,
                         int H, int Lq, int Lk, int d, int dtype, unsigned flags, void* stream) {
  int rc = check_common(Q, K, V, O, B, H, Lq, d, dtype);
  if (rc != FA_OK) return rc;
  if (Lk <= 0) return fail(FA_ERR_SHAPE, "Lk must be positive");
  if (flags & ~unsigned(FA_FLAG_CAUSAL)) return fail(FA_ERR_SHAPE, "unknown flag bits");
  if ((flags & FA_FLAG_CAUSAL) && Lq != Lk) return fail(FA_ERR_SHAPE, "causal masking needs Lq == Lk");
  FwdExtra ex;
  ex.Lk = Lk;
  ex.H = H;
  ex.kv_lens = kv_lens;
  if (!fused_tile_serves(d, dtype)) {
    TiledDExtra tx;
    tx.lse_out = LSE;
    tx.causal = (flags & FA_FLAG_CAUSAL) ? 1 : 0;
    tx.fx = ex;
    return dispatch_tiled_d_slab(Q, K, V, O, B * H, Lq, d, dtype, static_cast<cudaStream_t>(stream), tx);
  }
  return dispatch_fwd<false>(Q, K, V, O, B * H, Lq, d, dtype, Lk, 1, nullptr, nullptr, static_cast<cudaStream_t>(stream),
                             LSE, (flags & FA_FLAG_CAUSAL) ? 1 : 0, ex);
}

int fa_partial_forward(const void* Q, const void* K, const void* V, float* Opartial, float* LSEpartial, int B, int H,
                       int Lq, int Lk, int d, int dtype, long long q_head_rows, long long kv_head_rows,
                       long long out_head_rows, unsigned flags, void* stream) {
  int rc = check_common(Q, K, V, Opartial, B, H, Lq, d, dtype);
  if (rc != FA_OK) return rc;
  if (Lk <= 0) return fail(FA_ERR_SHAPE, "Lk must be positive");
  if (LSEpartial == nullptr) return fail(FA_ERR_ALIGN, "LSEpartial must be non-null");
  if (flags & ~unsigned(FA_FLAG_CAUSAL)) return fail(FA_ERR_SHAPE, "unknown flag bits");
  if ((flags & FA_FLAG_CAUSAL) && Lq != Lk) return fail(FA_ERR_SHAPE, "causal masking needs Lq == Lk");
  if ((q_head_rows != 0 && q_head_rows < Lq) || (kv_head_rows != 0 && kv_head_rows < Lk) ||
      (out_head_rows != 0 && out_head_rows < Lq))
    return fail(FA_ERR_SHAPE, "head strides (in rows) must be 0 (dense) or at least the row count");
  if (out_head_rows > 0x7fffffffLL) return fail(FA_ERR_SHAPE, "out_head_rows too large");
  FwdExtra ex;
  ex.Lk = Lk;
  ex.H = H;
  ex.q_head_rows = q_head_rows;
  ex.kv_head_rows = kv_head_rows;
  ex.out_head_rows = out_head_rows;
  if (!fused_tile_serves(d, dtype)) {
    TiledDExtra tx;
    tx.o_accum = Opartial;
    tx.lse_accum = LSEpartial;
    tx.causal = (flags & FA_FLAG_CAUSAL) ? 1 : 0;
    tx.fx = ex;
    return dispatch_tiled_d_slab(Q, K, V, nullptr, B * H, Lq, d, dtype, static_cast<cudaStream_t>(stream), tx);
  }
  return dispatch_fwd<true>(Q, K, V, nullptr, B * H, Lq, d, dtype, /*kv_per_split=*/Lk, /*n_splits=*/1, Opartial,
                            LSEpartial, static_cast<cudaStream_t>(stream), nullptr, (flags & FA_FLAG_CAUSAL) ? 1 : 0, ex);
}

int fa_v1_tiled_d_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d,
                          int d_tile_qk, int d_tile_v, int dtype, void* stream) {
  int rc = check_common(Q, K, V, O, B, H, L, d, dtype);
  if (rc != FA_OK) return rc;
  // Same argument contract as the reference launcher (flash_attention_v1_tiled_d/CUDA/flash_attention_v1.h:323-328).
  if (d_tile_qk <= 0 || d_tile_v <= 0 || d % d_tile_qk != 0 || d % d_tile_v != 0)
    return fail(FA_ERR_SHAPE, "d_tile_qk and d_tile_v must be positive divisors of d");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (fused_tile_serves(d, dtype)) return dispatch_fwd<false>(Q, K, V, O, B * H, L, d, dtype, L, 1, nullptr, nullptr, s);
  return dispatch_tiled_d(Q, K, V, O, B * H, L, d, dtype, s);
}

int fa_v1_tiled_d_pair_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d,
                               int dtype, void* stream) {
  int rc = check_common(Q, K, V, O, B, H, L, d, dtype);
  if (rc != FA_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d == 512 && dtype == FA_DTYPE_BF16) return launch_tiled_d_pair<512, fa::DT_BF16>(Q, K, V, O, B * H, L, s);
  if (d == 512 && dtype == FA_DTYPE_F16) return launch_tiled_d_pair<512, fa::DT_F16>(Q, K, V, O, B * H, L, s);
  if (d == 256 && dtype == FA_DTYPE_BF16) return launch_tiled_d_pair<256, fa::DT_BF16>(Q, K, V, O, B * H, L, s);
  if (d == 256 && dtype == FA_DTYPE_F16) return launch_tiled_d_pair<256, fa::DT_F16>(Q, K, V, O, B * H, L, s);
  return fail(FA_ERR_UNSUPPORTED_D, "the CTA-pair tiled-d kernel serves d in {256,512} for bf16/fp16; got d=" +
                                        std::to_string(d) + " dtype=" + std::to_string(dtype));
}

int fa_v2_num_splits(int L, int kv_per_split) {
  if (L <= 0 || kv_per_split <= 0) return 0;
  return (L + kv_per_split - 1) / kv_per_split;
}

size_t fa_v2_workspace_bytes(int B, int H, int L, int d, int kv_per_split) {
  const int ns = fa_v2_num_splits(L, kv_per_split);
  if (ns <= 0 || B <= 0 || H <= 0 || d <= 0) return 0;
  const size_t rows = size_t(B) * H * L;
  const size_t o_bytes = ((size_t(ns) * rows * d * sizeof(float)) + 255) & ~size_t(255);
  const size_t lse_bytes = ((size_t(ns) * rows * sizeof(float)) + 255) & ~size_t(255);
  return o_bytes + lse_bytes;
}

int fa_v2_splitkv_forward(const void* Q, const void* K, const void* V, float* Oaccum, float* LSEaccum, int B, int H,
                          int L, int d, int kv_per_split, int dtype, void* stream) {
  int rc = check_common(Q, K, V, Oaccum, B, H, L, d, dtype);
  if (rc != FA_OK) return rc;
  if (kv_per_split <= 0) return fail(FA_ERR_SHAPE, "kv_per_split must be positive");
  if (LSEaccum == nullptr) return fail(FA_ERR_ALIGN, "LSEaccum must be non-null");
  const int ns = fa_v2_num_splits(L, kv_per_split);
  if (!fused_tile_serves(d, dtype)) {
    // rows of 512-1024 bytes (16-bit d = 256/512, fp32 d = 128/256 — the reference V2's own default D = 128 in its
    // USE_FP64 mode): the slab tiled-d kernel with a key range per CTA
    TiledDExtra tx;
    tx.kv_per_split = kv_per_split;
    tx.n_splits = ns;
    tx.o_accum = Oaccum;
    tx.lse_accum = LSEaccum;
    return dispatch_tiled_d_slab(Q, K, V, nullptr, B * H, L, d, dtype, static_cast<cudaStream_t>(stream), tx);
  }
  // every split a single KV tile (the reference's own C3 geometry: 64 keys per split): the dedicated kernel
  if (kv_per_split <= 128 && splitkv_tile_enabled())
    return dispatch_splitkv_tile(Q, K, V, B * H, L, d, dtype, kv_per_split, ns, Oaccum, LSEaccum,
                                 static_cast<cudaStream_t>(stream));
  return dispatch_fwd<true>(Q, K, V, nullptr, B * H, L, d, dtype, kv_per_split, ns, Oaccum, LSEaccum,
                            static_cast<cudaStream_t>(stream));
}

int fa_v2_combine(const float* Oaccum, const float* LSEaccum, void* O, int B, int H, int L, int d, int n_splits,
                  int dtype, void* stream) {
  if (B <= 0 || H <= 0 || L <= 0 || d <= 0 || n_splits <= 0) return fail(FA_ERR_SHAPE, "B, H, L, d, n_splits must be positive");
  if (dtype != FA_DTYPE_F32 && dtype != FA_DTYPE_BF16 && dtype != FA_DTYPE_F16) return fail(FA_ERR_DTYPE, "unknown dtype");
  if (!Oaccum || !LSEaccum || !O || (reinterpret_cast<uintptr_t>(Oaccum) & 15) || (reinterpret_cast<uintptr_t>(O) & 15))
    return fail(FA_ERR_ALIGN, "Oaccum, LSEaccum, O must be non-null; Oaccum and O 16-byte aligned");
  const long long rows = (long long)B * H * L;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define FA_CASE(DD)                                                                                        \
  if (d == DD) {                                                                                           \
    if (dtype == FA_DTYPE_F32) return launch_combine<DD, fa::DT_F32>(Oaccum, LSEaccum, O, rows, n_splits, s);   \
    if (dtype == FA_DTYPE_BF16) return launch_combine<DD, fa::DT_BF16>(Oaccum, LSEaccum, O, rows, n_splits, s); \
    return launch_combine<DD, fa::DT_F16>(Oaccum, LSEaccum, O, rows, n_splits, s);                          \
  }
  FA_CASE(32)
  FA_CASE(64)
  FA_CASE(128)
  FA_CASE(256)
  FA_CASE(512)
#undef FA_CASE
  return fail(FA_ERR_UNSUPPORTED_D, "combine serves d in {32,64,128,256,512}");
}

int fa_v2_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d, int kv_per_split,
                  int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(Q, K, V, O, B, H, L, d, dtype);
  if (rc != FA_OK) return rc;
  if (kv_per_split <= 0) return fail(FA_ERR_SHAPE, "kv_per_split must be positive");
  const size_t need = fa_v2_workspace_bytes(B, H, L, d, kv_per_split);
  if (workspace == nullptr || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return fail(FA_ERR_WORKSPACE, "workspace must be 256-byte aligned and hold " + std::to_string(need) + " bytes");
  const int ns = fa_v2_num_splits(L, kv_per_split);
  const size_t rows = size_t(B) * H * L;
  float* o_accum = static_cast<float*>(workspace);
  float* lse_accum =
      reinterpret_cast<float*>(static_cast<char*>(workspace) + (((size_t(ns) * rows * d * sizeof(float)) + 255) & ~size_t(255)));
  rc = fa_v2_splitkv_forward(Q, K, V, o_accum, lse_accum, B, H, L, d, kv_per_split, dtype, stream);
  if (rc != FA_OK) return rc;
  return fa_v2_combine(o_accum, lse_accum, O, B, H, L, d, ns, dtype, stream);
}

size_t fa_v1_backward_workspace_bytes(int B, int H, int L) {
  if (B <= 0 || H <= 0 || L <= 0) return 0;
  return size_t(2) * size_t(B) * H * (size_t((L + 127) / 128) * 128) * sizeof(float);
}

int fa_v1_backward(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* LSE, void* dQ,
                   void* dK, void* dV, int B, int H, int L, int d, int dtype, unsigned flags, void* workspace,
                   size_t workspace_bytes, void* stream) {
  int rc = check_common(Q, K, V, dQ, B, H, L, d, dtype);
  if (rc != FA_OK) return rc;
  const void* more[4] = {O, dO, dK, dV};
  for (const void* ptr : more)
    if (ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
      return fail(FA_ERR_ALIGN, "O, dO, dK, dV must be non-null and 16-byte aligned");
  if (LSE == nullptr) return fail(FA_ERR_ALIGN, "LSE must be non-null (fa_v1_forward_ex produces it)");
  if (flags & ~unsigned(FA_FLAG_CAUSAL)) return fail(FA_ERR_SHAPE, "unknown flag bits");
  const size_t need = fa_v1_backward_workspace_bytes(B, H, L);
  if (workspace == nullptr || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return fail(FA_ERR_WORKSPACE, "workspace must be 256-byte aligned and hold " + std::to_string(need) + " bytes");
  const int causal = (flags & FA_FLAG_CAUSAL) ? 1 : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
#define FA_CASE(DD, DTT) \
  if (d == DD && dtype == DTT) return launch_backward<DD, DTT>(Q, K, V, O, dO, LSE, dQ, dK, dV, B * H, L, causal, ws, s);
  FA_CASE(128, fa::DT_BF16)
  FA_CASE(64, fa::DT_BF16)
  FA_CASE(128, fa::DT_F16)
  FA_CASE(64, fa::DT_F16)
#undef FA_CASE
  return fail(FA_ERR_UNSUPPORTED_D, "backward serves d in {64,128} for bf16/fp16; got d=" + std::to_string(d) +
                                        " dtype=" + std::to_string(dtype));
}

// ---- independent evaluation (drop-in for the reference's oracle naive_attention, common/reference.py:7-21) ----
size_t fa_naive_attention_workspace_bytes(int n_heads, int Lq, int Lk, int dtype) {
  if (n_heads <= 0 || Lq <= 0 || Lk <= 0 || (dtype != FA_DTYPE_F32 && dtype != FA_DTYPE_F64)) return 0;
  return size_t(n_heads) * Lq * Lk * (dtype == FA_DTYPE_F64 ? 8 : 4);
}

int fa_naive_attention(const void* Q, const void* K, const void* V, void* O, int n_heads, int Lq, int Lk, int d,
                       int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_heads <= 0 || Lq <= 0 || Lk <= 0 || d <= 0) return fail(FA_ERR_SHAPE, "n_heads, Lq, Lk, d must be positive");
  if (dtype != FA_DTYPE_F32 && dtype != FA_DTYPE_F64) return fail(FA_ERR_DTYPE, "naive attention computes in FA_DTYPE_F32 or FA_DTYPE_F64");
  if (!Q || !K || !V || !O || !workspace) return fail(FA_ERR_ALIGN, "Q, K, V, O, workspace must be non-null");
  if ((long long)n_heads * Lq > 0x7fffffffLL) return fail(FA_ERR_SHAPE, "too many rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == FA_DTYPE_F64)
    return naive_attention_impl<double>(static_cast<const double*>(Q), static_cast<const double*>(K), static_cast<const double*>(V),
                                        static_cast<double*>(O), n_heads, Lq, Lk, d, static_cast<double*>(workspace), workspace_bytes, s);
  return naive_attention_impl<float>(static_cast<const float*>(Q), static_cast<const float*>(K), static_cast<const float*>(V),
                                     static_cast<float*>(O), n_heads, Lq, Lk, d, static_cast<float*>(workspace), workspace_bytes, s);
}

// Strided block copy on the copy engines (no SMs): `height` rows of `width` bytes, row pitches dpitch / spitch, between any
// two device pointers the current device can address (its own memory or peer-mapped symmetric memory).  The all-to-all
// sequence-parallel path moves [heads][rows][d] blocks with it while the persistent attention kernel owns every SM.
int fa_copy_2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* stream) {
  if (!dst || !src || width == 0 || height == 0 || dpitch < width || spitch < width)
    return fail(FA_ERR_SHAPE, "fa_copy_2d_async: null pointer, empty block or pitch smaller than the row width");
  FA_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return FA_OK;
}

// n such copies with common block shape and pitches, one call: the head exchange issues 8-32 of them per step and the
// per-call cost of the Python -> ctypes path (10-20 us) would otherwise exceed the copies themselves.
int fa_copy_2d_multi_async(int n, void* const* dst, size_t dpitch, const void* const* src, size_t spitch, size_t width,
                           size_t height, void* stream) {
  if (n < 0 || !dst || !src) return fail(FA_ERR_SHAPE, "fa_copy_2d_multi_async: bad arguments");
  for (int i = 0; i < n; ++i) {
    const int rc = fa_copy_2d_async(dst[i], dpitch, src[i], spitch, width, height, stream);
    if (rc != FA_OK) return rc;
  }
  return FA_OK;
}

// n contiguous copies of `bytes` each (cudaMemcpyAsync, cudaMemcpyDefault): plain 1-D copies are the ones the copy
// engines run beside a persistent kernel; the strided form above did not overlap with it on B200.
int fa_copy_multi_async(int n, void* const* dst, const void* const* src, size_t bytes, void* stream) {
  if (n < 0 || !dst || !src || bytes == 0) return fail(FA_ERR_SHAPE, "fa_copy_multi_async: bad arguments");
  for (int i = 0; i < n; ++i)
    FA_CUDA_TRY(cudaMemcpyAsync(dst[i], src[i], bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return FA_OK;
}

void fa_release_host_staging(void) {
  const int dev = current_device();
  if (dev < 0) return;
  HostStaging& st = g_stage[dev];
  std::lock_guard<std::mutex> lk(st.mu);
  st.drain();
  st.release_buffers();
}

void fa_debug_map_cache_stats(unsigned long long* hits, unsigned long long* misses) {
  if (hits) *hits = g_map_hits.load();
  if (misses) *misses = g_map_misses.load();
}

int fa_forward_host(int variant, const void* Qh, const void* Kh, const void* Vh, void* Oh, int B, int H, int L, int d,
                    int kv_per_split, int dtype) {
  // every argument is validated before anything is allocated or enqueued
  if (!Qh || !Kh || !Vh || !Oh) return fail(FA_ERR_ALIGN, "host pointers must be non-null");
  if (B <= 0 || H <= 0 || L <= 0 || d <= 0) return fail(FA_ERR_SHAPE, "B, H, L, d must be positive");
  if (dtype != FA_DTYPE_F32 && dtype != FA_DTYPE_BF16 && dtype != FA_DTYPE_F16) return fail(FA_ERR_DTYPE, "unknown dtype");
  if (variant < 0 || variant > 2) return fail(FA_ERR_SHAPE, "variant must be 0 (V1), 1 (tiled-d) or 2 (V2)");
  if (variant == 2 && kv_per_split <= 0) return fail(FA_ERR_SHAPE, "kv_per_split must be positive");
  const int dev = current_device();
  if (dev < 0) return fail(FA_ERR_CUDA, "cannot query the current device");
  HostStaging& S = g_stage[dev];
  std::lock_guard<std::mutex> lk(S.mu);
  const size_t bytes = size_t(B) * H * L * d * elem_size(dtype);
  if (bytes > S.bytes) {
    S.drain();
    for (auto& b : S.buf) {
      if (b) cudaFree(b);
      b = nullptr;
    }
    S.bytes = 0;
    for (auto& b : S.buf) FA_CUDA_TRY(cudaMalloc(&b, bytes));
    S.bytes = bytes;
  }
  // Heads are independent, so the batch is pipelined in head chunks over three streams: while chunk c computes, chunk
  // c+1 is on the H2D copy engine and chunk c-1 on the D2H engine (PCIe is full duplex).  The reference drivers copy,
  // launch and copy back strictly in sequence (flash_attention_v1/CUDA/driver.cu:184-247).
  if (!S.streams_ready) {
    for (auto& st : S.stream) FA_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& e : S.ev_h2d) FA_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : S.ev_comp) FA_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    S.streams_ready = true;
  }
  // (Tried: Q, K and V on three H2D streams so the ~12 us of copy set-up per cudaMemcpyAsync overlap — 4.47 vs 4.16 ms
  //  per C2 step: the concurrent copies share the link and every chunk's LAST operand arrives later.)
  cudaStream_t s_h2d = S.stream[0], s_comp = S.stream[1], s_d2h = S.stream[2];
  const int BH = B * H;
  // Chunk sizes halve (1/2, 1/4, ... of the heads, the last two equal): every cudaMemcpyAsync costs ~12 us of copy-
  // engine set-up, so few large copies up front, and a small last chunk so little D2H is left when the H2D stream ends.
  int bounds[HostStaging::kChunks + 1];
  int n_chunks = 0;
  bounds[0] = 0;
#if FA_HOST_GEOMETRIC
  {
    const int floor_heads = BH / 16 > 0 ? BH / 16 : 1;
    int done = 0;
    while (done < BH) {
      int take = (BH - done + 1) / 2;
      if (take < floor_heads || n_chunks == HostStaging::kChunks - 1) take = BH - done;
      done += take;
      bounds[++n_chunks] = done;
    }
  }
#else
  n_chunks = BH < HostStaging::kChunks ? BH : HostStaging::kChunks;
  for (int c = 1; c <= n_chunks; ++c) bounds[c] = int((long long)BH * c / n_chunks);
#endif
  const size_t head_bytes = size_t(L) * d * elem_size(dtype);
  if (variant == 2) {
    int max_heads = 0;
    for (int c = 0; c < n_chunks; ++c) max_heads = std::max(max_heads, bounds[c + 1] - bounds[c]);
    const size_t ws_need = fa_v2_workspace_bytes(1, max_heads, L, d, kv_per_split);
    if (ws_need > S.ws_bytes) {
      S.drain();
      if (S.ws) cudaFree(S.ws);
      S.ws = nullptr;
      S.ws_bytes = 0;
      FA_CUDA_TRY(cudaMalloc(&S.ws, ws_need));
      S.ws_bytes = ws_need;
    }
  }
  // An error below leaves earlier chunks' copies in flight: drain the three streams before handing the error back.
  auto bail = [&](int rc) {
    S.drain();
    return rc;
  };
#define FA_HOST_TRY(expr)                                                                                       \
  do {                                                                                                          \
    cudaError_t _e = (expr);                                                                                    \
    if (_e != cudaSuccess) return bail(fail(FA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e))); \
  } while (0)
  for (int c = 0; c < n_chunks; ++c) {
    const int h0 = bounds[c], h1 = bounds[c + 1];
    const int nh = h1 - h0;
    if (nh == 0) continue;
    const size_t off = size_t(h0) * head_bytes, cb = size_t(nh) * head_bytes;
    char *dQ = static_cast<char*>(S.buf[0]) + off, *dK = static_cast<char*>(S.buf[1]) + off;
    char *dV = static_cast<char*>(S.buf[2]) + off, *dO = static_cast<char*>(S.buf[3]) + off;
    FA_HOST_TRY(cudaMemcpyAsync(dQ, static_cast<const char*>(Qh) + off, cb, cudaMemcpyHostToDevice, s_h2d));
    FA_HOST_TRY(cudaMemcpyAsync(dK, static_cast<const char*>(Kh) + off, cb, cudaMemcpyHostToDevice, s_h2d));
    FA_HOST_TRY(cudaMemcpyAsync(dV, static_cast<const char*>(Vh) + off, cb, cudaMemcpyHostToDevice, s_h2d));
    FA_HOST_TRY(cudaEventRecord(S.ev_h2d[c], s_h2d));
    FA_HOST_TRY(cudaStreamWaitEvent(s_comp, S.ev_h2d[c], 0));
    int rc;
    if (variant == 0) {
      rc = fa_v1_forward(dQ, dK, dV, dO, 1, nh, L, d, dtype, s_comp);
    } else if (variant == 1) {
      const int dt = d >= 64 ? 64 : d;
      rc = fa_v1_tiled_d_forward(dQ, dK, dV, dO, 1, nh, L, d, dt, dt, dtype, s_comp);
    } else {
      rc = fa_v2_forward(dQ, dK, dV, dO, 1, nh, L, d, kv_per_split, dtype, S.ws, S.ws_bytes, s_comp);
    }
    if (rc != FA_OK) return bail(rc);
    FA_HOST_TRY(cudaEventRecord(S.ev_comp[c], s_comp));
    FA_HOST_TRY(cudaStreamWaitEvent(s_d2h, S.ev_comp[c], 0));
    FA_HOST_TRY(cudaMemcpyAsync(static_cast<char*>(Oh) + off, dO, cb, cudaMemcpyDeviceToHost, s_d2h));
  }
  FA_HOST_TRY(cudaStreamSynchronize(s_d2h));
#undef FA_HOST_TRY
  return FA_OK;
}

}  // extern "C"
