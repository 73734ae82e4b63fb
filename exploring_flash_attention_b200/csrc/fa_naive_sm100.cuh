// fa_naive_sm100.cuh — the package's INDEPENDENT attention evaluation: the drop-in for the reference's own oracle
// naive_attention (common/reference.py:7-21), so that a reference script ported onto this package still compares the
// fused kernels with something that is not them.
//
// Deliberately nothing like the fused kernels: the [Lq x Lk] score matrix is materialised in global memory exactly as
// reference.py:17-21 does (scores = Q K^T * scale; scores -= rowmax; probs = exp(scores); probs /= rowsum; O = probs V),
// products run on the CUDA cores in the storage precision of the caller's NumPy buffers (fp32, or fp64 for the
// reference's float64 runs) with no tf32 truncation, no tensor cores, no online softmax, no exp2 polynomial.
// It is a validation helper (a few TFLOP/s), not a hot path.
#pragma once
#include <cuda_runtime.h>

namespace fa {

// C[m][n] = alpha * sum_k A[m][k] * (B_IS_NK ? B[n][k] : B[k][n]); row-major, one matrix per blockIdx.z.
// 64 x 64 output tile per 256-thread block, 4 x 4 per thread, 16-wide k chunks through shared memory.
template <typename T, bool B_IS_NK>
__global__ void __launch_bounds__(256) naive_gemm_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                          T* __restrict__ C, int M, int N, int K, T alpha,
                                                          long long strideA, long long strideB, long long strideC) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ T sA[TK][TM + 1];
  __shared__ T sB[TK][TN + 1];
  A += blockIdx.z * strideA;
  B += blockIdx.z * strideB;
  C += blockIdx.z * strideC;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  T acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      const int kk = e % TK, mm = e / TK;
      const int m = m0 + mm, k = k0 + kk;
      sA[kk][mm] = (m < M && k < K) ? A[(long long)m * K + k] : T(0);
    }
    for (int e = threadIdx.x; e < TN * TK; e += 256) {
      int kk, nn;
      if (B_IS_NK) {
        kk = e % TK;
        nn = e / TK;
      } else {
        nn = e % TN;
        kk = e / TN;
      }
      const int n = n0 + nn, k = k0 + kk;
      T v = T(0);
      if (n < N && k < K) v = B_IS_NK ? B[(long long)n * K + k] : B[(long long)k * N + n];
      sB[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[(long long)m * N + n] = alpha * acc[i][j];
    }
}

__device__ __forceinline__ float naive_exp(float x) { return expf(x); }
__device__ __forceinline__ double naive_exp(double x) { return exp(x); }

// In-place row softmax of [rows][n] (reference.py:18-20): one 256-thread block per row.
template <typename T>
__global__ void __launch_bounds__(256) naive_softmax_rows_kernel(T* __restrict__ S, int n) {
  __shared__ T red[8];
  T* row = S + (long long)blockIdx.x * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto block_reduce = [&](T v, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T other = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_max ? (other > v ? other : v) : v + other;
    }
    __syncthreads();   // red[] may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    T r = red[0];
    for (int w = 1; w < 8; ++w) r = is_max ? (red[w] > r ? red[w] : r) : r + red[w];
    return r;
  };
  T mx = row[0];
  for (int c = threadIdx.x; c < n; c += 256) mx = row[c] > mx ? row[c] : mx;
  mx = block_reduce(mx, true);
  T sum = T(0);
  for (int c = threadIdx.x; c < n; c += 256) {
    const T e = naive_exp(row[c] - mx);
    row[c] = e;
    sum += e;
  }
  sum = block_reduce(sum, false);
  for (int c = threadIdx.x; c < n; c += 256) row[c] = row[c] / sum;
}

}  // namespace fa
