// fa_combine_sm100.cuh — K3b: split-KV combine (bandwidth kernel).
//
// Replaces reduction_kernel, flash_attention_v2/CUDA/flash_attention_v2.h:356-435 (byte-identical copy
// in flash_attention_v2_opt.h:477-556) and the Python reduction_kernel
// (flash_attention_v2/numpy_gpu_like.py:229-288):
//     m_g = max_k m_k;  s_k = exp(m_k - m_g);  O = sum_k O_k s_k / sum_k l_k s_k.
// With split-normalised partials Õ_k = O_k / l_k and LSE_k = m_k/sqrt(d) + ln l_k this is
//     LSE = log sum_k exp(LSE_k);  O = sum_k exp(LSE_k - LSE) Õ_k.
//
// The reference walks workspace_O with stride BQ*D across splits using scalar half loads, one thread per
// row for max/sum and three __syncthreads (its own TODO list, flash_attention_v2.h:348-354).  Here:
//   * workspace is split-major fp32 so every split is one contiguous stream;
//   * a group of G = min(32, d/4) lanes owns one query row; each lane moves 16-byte vectors
//     (ld.global.nc.L1::no_allocate.v4) => fully coalesced 128-byte lines;
//   * the per-row max / sum over splits are warp-shuffle reductions inside the lane group —
//     no shared memory, no block barrier;
//   * all split loads of a thread are issued before the first use (memory-level parallelism).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fa_fwd_sm100.cuh"

namespace fa {

__device__ __forceinline__ float4 ld_stream_f4(const float* ptr) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(ptr));
  return r;
}

template <int DT>
__device__ __forceinline__ void store_out4(void* O, size_t elem_idx, float4 v) {
  if constexpr (DT == DT_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(O) + elem_idx) = v;
  } else {
    uint2 pk;
    if constexpr (DT == DT_BF16) {
      pk.x = pack_bf16x2(v.x, v.y);
      pk.y = pack_bf16x2(v.z, v.w);
    } else {
      pk.x = pack_f16x2(v.x, v.y);
      pk.y = pack_f16x2(v.z, v.w);
    }
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(O) + elem_idx) = pk;
  }
}

constexpr int kCombineThreads = 256;
constexpr int kCombineMaxUnroll = 8;  // splits kept in flight per thread

// D: head dim (multiple of 4).  One lane group of G lanes per row; NV float4 vectors per lane.
template <int D, int DT>
__global__ void __launch_bounds__(kCombineThreads)
fa_combine_kernel(const float* __restrict__ o_accum, const float* __restrict__ lse_accum, void* __restrict__ O,
                  long long rows, int n_splits) {
  constexpr int G = (D / 4 < 32) ? D / 4 : 32;
  constexpr int NV = D / (4 * G);
  constexpr int ROWS_PER_BLOCK = kCombineThreads / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);  // lane within the row group
  const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + threadIdx.x / G;
  const bool live = row < rows;
  const long long r = live ? row : rows - 1;  // keep whole warps converged for the shuffles
  const size_t split_stride = size_t(rows) * D;
  const float* src = o_accum + size_t(r) * D + gl * 4;

  // ---- LSE over splits: lanes of the group take splits gl, gl+G, ... ; shuffle-reduce max and sum.
  float my_max = -CUDART_INF_F;
  for (int k = gl; k < n_splits; k += G) my_max = fmaxf(my_max, __ldg(lse_accum + size_t(k) * rows + r));
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) my_max = fmaxf(my_max, __shfl_xor_sync(0xffffffffu, my_max, off));
  float my_sum = 0.f;
  for (int k = gl; k < n_splits; k += G) my_sum += __expf(__ldg(lse_accum + size_t(k) * rows + r) - my_max);
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) my_sum += __shfl_xor_sync(0xffffffffu, my_sum, off);
  const float inv_sum = 1.0f / my_sum;

  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int k0 = 0; k0 < n_splits; k0 += kCombineMaxUnroll) {
    float4 x[kCombineMaxUnroll][NV];
    float w[kCombineMaxUnroll];
#pragma unroll
    for (int u = 0; u < kCombineMaxUnroll; ++u) {
      const int k = k0 + u;
      if (k < n_splits) {
#pragma unroll
        for (int v = 0; v < NV; ++v) x[u][v] = ld_stream_f4(src + size_t(k) * split_stride + v * (G * 4));
        w[u] = __ldg(lse_accum + size_t(k) * rows + r);
      }
    }
#pragma unroll
    for (int u = 0; u < kCombineMaxUnroll; ++u) {
      const int k = k0 + u;
      if (k < n_splits) {
        const float wk = __expf(w[u] - my_max) * inv_sum;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          acc[v].x = fmaf(wk, x[u][v].x, acc[v].x);
          acc[v].y = fmaf(wk, x[u][v].y, acc[v].y);
          acc[v].z = fmaf(wk, x[u][v].z, acc[v].z);
          acc[v].w = fmaf(wk, x[u][v].w, acc[v].w);
        }
      }
    }
  }
  if (live) {
#pragma unroll
    for (int v = 0; v < NV; ++v) store_out4<DT>(O, size_t(row) * D + v * (G * 4) + gl * 4, acc[v]);
  }
}

}  // namespace fa
