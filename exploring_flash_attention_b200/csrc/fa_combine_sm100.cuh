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
constexpr int kCombineMaxUnroll = 8;  // generic kernel: splits kept in flight per thread
constexpr int kCombineChunk = 4;      // fast kernel: splits loaded per batch
constexpr int kCombineMaxSplitsFast = 32;

// Fast path (n_splits <= 32).  A group of G = min(32, D/4) lanes owns RPT query rows (interleaved so that adjacent
// groups touch adjacent rows); lane gl of a group owns split gl's LSE (+G, +2G.. for small D), so
//   LSE_max / sum exp  are xor-shuffle reductions inside the group, and weight w_k is fetched with one shuffle.
// Every global load of a batch — the LSE words and RPT*4*NV 16-byte Oaccum vectors — is issued before the first
// dependent instruction, so one memory round trip covers both the softmax-over-splits and the first 4 splits.
template <int D, int DT>
__global__ void __launch_bounds__(kCombineThreads)
fa_combine_kernel(const float* __restrict__ o_accum, const float* __restrict__ lse_accum, void* __restrict__ O,
                  long long rows, int n_splits) {
  constexpr int G = (D / 4 < 32) ? D / 4 : 32;
  constexpr int NV = D / (4 * G);
  constexpr int KK = kCombineMaxSplitsFast / G;          // LSE slots per lane
  constexpr int RPT = (NV >= 2) ? 1 : 2;                 // rows per lane group
  constexpr int GROUPS = kCombineThreads / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int gbase = lane & ~(G - 1);
  const long long row0 = (long long)blockIdx.x * (GROUPS * RPT) + threadIdx.x / G;
  const size_t split_stride = size_t(rows) * D;

  long long r[RPT];
  bool live[RPT];
#pragma unroll
  for (int t = 0; t < RPT; ++t) {
    const long long row = row0 + (long long)t * GROUPS;
    live[t] = row < rows;
    r[t] = live[t] ? row : rows - 1;  // dead groups shadow the last row so whole warps stay converged for the shuffles
  }

  pdl_wait();   // launched with programmatic stream serialization: the split-KV kernel's partials are complete from here on
  // ---- issue: LSE words first (they come back first), then the first batch of Oaccum vectors
  float e[RPT][KK];
#pragma unroll
  for (int t = 0; t < RPT; ++t)
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
      const int k = kk * G + gl;
      e[t][kk] = (k < n_splits) ? __ldg(lse_accum + size_t(k) * rows + r[t]) : -CUDART_INF_F;
    }
  float4 x[RPT][kCombineChunk][NV];
  auto issue = [&](int k0) {
#pragma unroll
    for (int t = 0; t < RPT; ++t)
#pragma unroll
      for (int u = 0; u < kCombineChunk; ++u)
        if (k0 + u < n_splits) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
            x[t][u][v] = ld_stream_f4(o_accum + size_t(k0 + u) * split_stride + size_t(r[t]) * D + v * (G * 4) + gl * 4);
        }
  };
  issue(0);

  // ---- softmax over splits (per row): w_k = exp(LSE_k - max) / sum, kept by the lane that owns split k
  float w[RPT][KK];
#pragma unroll
  for (int t = 0; t < RPT; ++t) {
    float mx = e[t][0];
#pragma unroll
    for (int kk = 1; kk < KK; ++kk) mx = fmaxf(mx, e[t][kk]);
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
      w[t][kk] = __expf(e[t][kk] - mx);  // exp(-inf) = 0 for the unused slots
      sum += w[t][kk];
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) w[t][kk] *= inv;
  }

  float4 acc[RPT][NV];
#pragma unroll
  for (int t = 0; t < RPT; ++t)
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[t][v] = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
    for (int j0 = 0; j0 < G; j0 += kCombineChunk) {
      const int k0 = kk * G + j0;
      if (k0 >= n_splits) break;
      if (k0 > 0) issue(k0);
#pragma unroll
      for (int u = 0; u < kCombineChunk; ++u) {
#pragma unroll
        for (int t = 0; t < RPT; ++t) {
          const float wk = __shfl_sync(0xffffffffu, w[t][kk], gbase + ((j0 + u) & (G - 1)));
          if (k0 + u < n_splits) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              acc[t][v].x = fmaf(wk, x[t][u][v].x, acc[t][v].x);
              acc[t][v].y = fmaf(wk, x[t][u][v].y, acc[t][v].y);
              acc[t][v].z = fmaf(wk, x[t][u][v].z, acc[t][v].z);
              acc[t][v].w = fmaf(wk, x[t][u][v].w, acc[t][v].w);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < RPT; ++t)
    if (live[t]) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
        store_out4<DT>(O, size_t(row0 + (long long)t * GROUPS) * D + v * (G * 4) + gl * 4, acc[t][v]);
    }
}

// Generic path (any n_splits): same data movement, LSE re-read per split instead of kept in lane slots.
template <int D, int DT>
__global__ void __launch_bounds__(kCombineThreads)
fa_combine_generic_kernel(const float* __restrict__ o_accum, const float* __restrict__ lse_accum, void* __restrict__ O,
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

  pdl_wait();   // see fa_combine_kernel
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
