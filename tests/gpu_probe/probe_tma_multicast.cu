// probe_tma_multicast.cu — how many bytes per second can every SM pull out of L2 into shared memory with TMA, and does
// cluster multicast (one L2 read delivered to the same smem offset of several CTAs) raise the per-SM delivered rate?
// The tiled-d kernels stream 64-128 KB of K/V per KV tile per SM out of L2; at d = 512 that, not the tensor pipe, bounds them.
//
// Every CTA streams a region of a 64 MB bf16 [65536 rows x 512] buffer (L2 resident) through a ring of 16 KB stages:
//   unicast     each CTA loads every box itself
//   multicast c clusters of c CTAs; CTA r issues the loads with it % c == r, multicast to all c CTAs
// and the clusters either all walk the SAME 8 MB region in lockstep (what the q-tiles of one head do), walk it from
// staggered starting points, or walk distinct 0.5 MB regions.  Prints delivered GB/s per SM and in aggregate.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_tma_multicast probe_tma_multicast.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../exploring_flash_attention_b200/csrc/sm100_ptx.cuh"

using namespace fa;

constexpr int NS = 8;
constexpr int ROWS = 65536, COLS = 512;   // 64 MB, L2 resident

__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "h"(mask)
      : "memory");
}

// BOX_ROWS x 64 elements (128 B) per TMA op; a stage is always 16 KB = 128 rows.
template <int CSZ, int BOX_ROWS>
__global__ void __launch_bounds__(96, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int n_iters, int region_rows,
                                                         int distinct, int stagger) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * 16384);
  uint64_t* empty = full + NS;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = (CSZ > 1) ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CSZ);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync_all();
  constexpr int OPS = 128 / BOX_ROWS;   // TMA ops per stage
  const int cluster_id = blockIdx.x / CSZ;
  const int region_base = distinct ? (cluster_id * region_rows) % ROWS : 0;
  const int it0 = stagger * cluster_id;
  if (warp == 0 && elect_one_sync()) {
    // producer: arm the local barrier for every stage; issue the loads this CTA is responsible for
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % NS;
      const bool mine = (it % CSZ) == int(rank);
      if (it >= NS && mine) mbar_wait_cluster(&empty[s], ((it / NS) - 1) & 1);   // every CTA of the cluster drained it
      // The local full barrier may only be re-armed once the local consumer has passed it: the consumer's arrival on
      // empty (at the issuer) implies that, but a non-issuing CTA has no such wait, so it tracks its own consumer.
      if (it >= NS && !mine) {
        volatile int* prog = reinterpret_cast<volatile int*>(empty + NS);
        while (*prog < it - NS + 1) { }
      }
      mbar_arrive_expect_tx(&full[s], 16384);
      if (mine) {
        const int e = it + it0;
        const int row = region_base + (e * 128) % region_rows, col = ((e * 128) / region_rows * 64) % COLS;
        for (int o = 0; o < OPS; ++o) {
          if (CSZ > 1)
            tma_load_3d_mc(smem + s * 16384 + o * BOX_ROWS * 128, &tm, &full[s], col, row + o * BOX_ROWS, 0,
                           uint16_t((1u << CSZ) - 1));
          else
            tma_load_3d(smem + s * 16384 + o * BOX_ROWS * 128, &tm, &full[s], col, row + o * BOX_ROWS, 0);
        }
      }
    }
  } else if (warp == 1 && elect_one_sync()) {
    // consumer: wait for the stage, release it at the CTA that will refill it
    volatile int* prog = reinterpret_cast<volatile int*>(empty + NS);
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % NS;
      mbar_wait(&full[s], (it / NS) & 1);
      *prog = it + 1;
      const int next_issuer = (it + NS) % CSZ;
      mbar_arrive_cluster(mapa_shared(smem_u32(&empty[s]), next_issuer));
    }
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync_all();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CSZ, int BOX_ROWS>
void run(const char* name, void* buf, EncodeFn enc, int sms, int region_rows = 8192, int distinct = 0, int stagger = 0) {
  CUtensorMap tm;
  cuuint64_t dims[3] = {COLS, ROWS, 1};
  cuuint64_t strides[2] = {COLS * 2, cuuint64_t(ROWS) * COLS * 2};
  cuuint32_t box[3] = {64, BOX_ROWS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("%-28s tensor map failed\n", name);
    return;
  }
  auto kern = stream_kernel<CSZ, BOX_ROWS>;
  const int smem = 1024 + NS * 16384 + 2 * NS * 8 + 64;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_iters = 16384;   // 256 MB per SM
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms - (sms % CSZ));
  cfg.blockDim = dim3(96);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CSZ;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, tm, n_iters, region_rows, distinct, stagger);
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
  const double per_sm = double(n_iters) * 16384 / (best * 1e-3) / 1e9;
  printf("%-28s %7.1f GB/s per SM delivered  %7.2f TB/s aggregate on %d SMs  (%.3f ms)\n", name, per_sm,
         per_sm * cfg.gridDim.x / 1e3, cfg.gridDim.x, best);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeFn enc = reinterpret_cast<EncodeFn>(p);
  void* buf = nullptr;
  cudaMalloc(&buf, size_t(ROWS) * COLS * 2);
  cudaMemset(buf, 0, size_t(ROWS) * COLS * 2);
  run<1, 128>("unicast lockstep", buf, enc, sms);
  run<1, 128>("unicast staggered x37", buf, enc, sms, 8192, 0, 37);
  run<1, 128>("unicast distinct 0.5MB", buf, enc, sms, 512, 1, 0);
  run<1, 64>("unicast distinct, 8KB boxes", buf, enc, sms, 512, 1, 0);
  run<2, 128>("multicast x2 lockstep", buf, enc, sms);
  run<2, 128>("multicast x2 staggered", buf, enc, sms, 8192, 0, 37);
  run<2, 128>("multicast x2 distinct", buf, enc, sms, 512, 1, 0);
  run<4, 128>("multicast x4 staggered", buf, enc, sms, 8192, 0, 37);
  run<4, 128>("multicast x4 distinct", buf, enc, sms, 512, 1, 0);
  return 0;
}
