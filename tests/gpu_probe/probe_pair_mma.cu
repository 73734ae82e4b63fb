// probe_pair_mma.cu — how fast does the tensor pipe run for each tcgen05.mma shape / cta_group when every SM issues a long
// back-to-back stream of MMAs from shared memory (SS, bf16, K = 16)?  Answers whether 2-CTA MMAs with M = 128 (64 rows
// per SM) reach the same rate as M = 256: per MMA each SM fetches the peer's half of B, N*16 bytes per M*N/512 cycles.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_pair_mma probe_pair_mma.cu && ./probe_pair_mma
#include <cstdio>
#include <cuda_runtime.h>

#include "../../exploring_flash_attention_b200/csrc/sm100_ptx.cuh"

using namespace fa;

template <int CG, int M, int N>
__global__ void __launch_bounds__(128, 1) mma_stream_kernel(int n_mma, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // up to 128 rows x 128 B (K-major, 128B swizzle), 4 K-steps
  uint8_t* sB = smem + 16384;         // up to 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    if (CG == 2) tmem_alloc_pair(slot, 512); else tmem_alloc(slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const bool leader = (CG == 1) || cluster_ctarank() == 0;
  if (warp == 1 && leader && elect_one_sync()) {
    constexpr uint32_t idesc = make_idesc(FMT_BF16, M, N, 0, 0);
    constexpr uint64_t hi = make_smem_desc_hi(16, 1024, SWZ_128B);
    const uint32_t a = smem_u32(sA), b = smem_u32(sB);
    for (int i = 0; i < n_mma; ++i) {
      const int k = i & 3;
      if (CG == 2)
        umma_ss_pair(tmem + (i & 4 ? 256 : 0), make_smem_desc(a + k * 32, hi), make_smem_desc(b + k * 32, hi), idesc, i > 7);
      else
        umma_ss<KIND_F16>(tmem + (i & 4 ? 256 : 0), make_smem_desc(a + k * 32, hi), make_smem_desc(b + k * 32, hi), idesc,
                          i > 7);
    }
    if (CG == 2) tc_commit_pair(bar, 3); else tc_commit(bar);
  }
  __syncwarp();
  if (warp == 2) {
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t r[16];
    tmem_ld16(tmem + (uint32_t(64) << 16), r);   // warp 2 may only touch TMEM lanes 64..95
    tc_wait_ld();
    if (sink && r[0] == 0x12345678u) sink[0] = 1.f;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 2) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

template <int CG, int M, int N>
void run(const char* name, int sms) {
  auto kern = mma_stream_kernel<CG, M, N>;
  const int smem = 1024 + 16384 + 32768 + 64;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 40000;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms - (sms % CG));
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float* sink = nullptr;
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, n_mma, sink);
    cudaEventRecord(e1);
    cudaError_t err2 = cudaDeviceSynchronize();
    if (err != cudaSuccess || err2 != cudaSuccess) {
      printf("%-28s FAILED: %s / %s\n", name, cudaGetErrorString(err), cudaGetErrorString(err2));
      return;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double ns_per_mma = best * 1e6 / n_mma;
  const double flop = 2.0 * M * N * 16;                           // per MMA (whole CTA group)
  const double tflops = flop * n_mma * (cfg.gridDim.x / CG) / (best * 1e-3) / 1e12;
  printf("%-28s %8.2f ns/MMA  %8.1f TFLOP/s on %d SMs  (%.3f ms)\n", name, ns_per_mma, tflops, cfg.gridDim.x, best);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<1, 128, 128>("cta_group::1 M128 N128", sms);
  run<1, 128, 256>("cta_group::1 M128 N256", sms);
  run<1, 128, 64>("cta_group::1 M128 N64", sms);
  run<2, 128, 128>("cta_group::2 M128 N128", sms);
  run<2, 128, 256>("cta_group::2 M128 N256", sms);
  run<2, 256, 128>("cta_group::2 M256 N128", sms);
  run<2, 256, 256>("cta_group::2 M256 N256", sms);
  run<2, 256, 64>("cta_group::2 M256 N64", sms);
  return 0;
}
