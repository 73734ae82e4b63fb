// driver_v1.cu — a compiled C++ consumer of include/fa_b200.h: what the reference's V1 driver
// (flash_attention_v1/CUDA/driver.cu:135-294) becomes once INTEGRATION.md §1's change is applied — the kernel header and
// its launcher call (driver.cu:220-238) replaced by fa_v1_forward from libfa_b200.so, nothing else linked from this repo.
// Same flow: srand(42) U[-1,1] fp16 Q,K,V of B32 H8 L1024 d32 (driver.cu:71-75,137-168), device buffers, 10 warm-up + 50
// timed launches, copy back, max-abs against a CPU evaluation, "Test PASSED" under 1e-3 (driver.cu:275).  The CPU side
// here is a plain scalar softmax(QK^T/sqrt d)V on a sample of heads (argv[1], default 4): the reference's OpenMP
// standard_attention_cpu over all 256 heads takes ~10 s and is timed by bench.py, not here.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -I include tests/drivers/driver_v1.cu \
//        -L exploring_flash_attention_b200/csrc -lfa_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../exploring_flash_attention_b200/csrc' -o tests/drivers/driver_v1
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fa_b200.h"

#define CHECK_CUDA(x)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

#define CHECK_FA(x)                                                              \
  do {                                                                           \
    int rc_ = (x);                                                               \
    if (rc_ != FA_OK) {                                                          \
      std::fprintf(stderr, "libfa_b200 status %d: %s\n", rc_, fa_last_error()); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

static void fill_uniform(std::vector<__half>& v) {
  for (auto& x : v) x = __float2half(2.0f * (float(std::rand()) / float(RAND_MAX)) - 1.0f);
}

// One head, fp32 math on the fp16-rounded inputs, materialised score row.
static void cpu_head(const __half* Q, const __half* K, const __half* V, float* O, int L, int d) {
  std::vector<float> s(L);
  const float scale = 1.0f / std::sqrt(float(d));
  for (int i = 0; i < L; ++i) {
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) {
      float acc = 0.f;
      for (int c = 0; c < d; ++c) acc += __half2float(Q[i * d + c]) * __half2float(K[j * d + c]);
      s[j] = acc * scale;
      mx = std::fmax(mx, s[j]);
    }
    float sum = 0.f;
    for (int j = 0; j < L; ++j) {
      s[j] = std::exp(s[j] - mx);
      sum += s[j];
    }
    for (int c = 0; c < d; ++c) {
      float acc = 0.f;
      for (int j = 0; j < L; ++j) acc += s[j] * __half2float(V[j * d + c]);
      O[i * d + c] = acc / sum;
    }
  }
}

int main(int argc, char** argv) {
  const int B = 32, H = 8, L = 1024, d = 32;
  const int check_heads = argc > 1 ? std::atoi(argv[1]) : 4;
  std::srand(42);
  std::printf("Flash Attention CUDA Test\nKernel: libfa_b200 fa_v1_forward (sm_100a)\nPrecision: FP16 (half)\n");
  std::printf("B=%d, H=%d, L=%d, d=%d\nTotal attention heads: %d\nSMs: %d\n\n", B, H, L, d, B * H, fa_device_sm_count());

  const size_t n = size_t(B) * H * L * d, bytes = n * sizeof(__half);
  std::vector<__half> hQ(n), hK(n), hV(n), hO(n);
  fill_uniform(hQ);
  fill_uniform(hK);
  fill_uniform(hV);

  __half *dQ, *dK, *dV, *dO;
  CHECK_CUDA(cudaMalloc(&dQ, bytes));
  CHECK_CUDA(cudaMalloc(&dK, bytes));
  CHECK_CUDA(cudaMalloc(&dV, bytes));
  CHECK_CUDA(cudaMalloc(&dO, bytes));
  CHECK_CUDA(cudaMemcpy(dQ, hQ.data(), bytes, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dK, hK.data(), bytes, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dV, hV.data(), bytes, cudaMemcpyHostToDevice));

  std::printf("Warming up (10 runs)...\n");
  for (int i = 0; i < 10; ++i) CHECK_FA(fa_v1_forward(dQ, dK, dV, dO, B, H, L, d, FA_DTYPE_F16, nullptr));
  CHECK_CUDA(cudaDeviceSynchronize());   // the reference launcher synchronised internally (flash_attention_v1.h:292)
  std::printf("Running timed iterations (50 runs)...\n");
  const int runs = 50;
  const auto t0 = std::chrono::high_resolution_clock::now();
  for (int i = 0; i < runs; ++i) CHECK_FA(fa_v1_forward(dQ, dK, dV, dO, B, H, L, d, FA_DTYPE_F16, nullptr));
  CHECK_CUDA(cudaDeviceSynchronize());
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count() / runs;
  std::printf("Average GPU time (%d runs): %.4f ms  (%.1f TFLOP/s)\n\n", runs, ms, 4.0 * B * H * double(L) * L * d / (ms * 1e-3) / 1e12);
  CHECK_CUDA(cudaMemcpy(hO.data(), dO, bytes, cudaMemcpyDeviceToHost));

  // the argument contract: a bad head dim comes back as a status, not an abort (reference: assert, flash_attention_v1.h:263)
  if (fa_v1_forward(dQ, dK, dV, dO, B, H, L, 48, FA_DTYPE_F16, nullptr) != FA_ERR_UNSUPPORTED_D) {
    std::fprintf(stderr, "expected FA_ERR_UNSUPPORTED_D for d=48\n");
    return 1;
  }

  float max_abs = 0.f;
  std::vector<float> ref(size_t(L) * d);
  for (int c = 0; c < check_heads; ++c) {
    const int head = int((long long)c * (B * H - 1) / (check_heads > 1 ? check_heads - 1 : 1));
    const size_t off = size_t(head) * L * d;
    cpu_head(hQ.data() + off, hK.data() + off, hV.data() + off, ref.data(), L, d);
    for (size_t e = 0; e < size_t(L) * d; ++e) max_abs = std::fmax(max_abs, std::fabs(ref[e] - __half2float(hO[off + e])));
  }
  std::printf("Results Comparison (%d sampled heads):\nMax absolute difference: %g\n", check_heads, max_abs);
  const bool pass = max_abs < 1e-3f;
  std::printf(pass ? "\nTest PASSED - Results match!\n" : "\nTest FAILED - Results differ significantly!\n");
  cudaFree(dQ);
  cudaFree(dK);
  cudaFree(dV);
  cudaFree(dO);
  return pass ? 0 : 1;
}
