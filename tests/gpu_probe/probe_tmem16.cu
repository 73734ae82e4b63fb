// probe_tmem16.cu — dev probe: register <-> (lane, column) mapping of the 16-lane tcgen05.ld / tcgen05.st shapes
// (16x256b, 16x128b) and whether a warp may address the upper 16 lanes of its 32-lane TMEM quadrant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../exploring_flash_attention_b200/csrc -o probe_tmem16 probe_tmem16.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"

using namespace fa;

// out[warp][half][lane_in_warp][32]: what 16x256b.x8 returns at lanes 32q+16h, columns 0..63
__global__ void k(uint32_t* out_ld, uint32_t* out_st) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const uint32_t t_lane = base + (uint32_t(warp * 32) << 16);
  // pattern: value(lane L, column c) = L * 256 + c, written one row per thread (32x32b)
  uint32_t v[32];
  for (int cb = 0; cb < 4; ++cb) {
    for (int x = 0; x < 32; ++x) v[x] = (warp * 32 + lane) * 256 + cb * 32 + x;
    tmem_st32(t_lane + cb * 32, v);
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t r[32];
    tmem_ld16x256b_x8(t_lane + (uint32_t(16 * h) << 16) + 64, r);   // columns 64..127
    tc_wait_ld();
    for (int x = 0; x < 32; ++x) out_ld[((warp * 2 + h) * 32 + lane) * 32 + x] = r[x];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // store side: 16x128b.x8 at columns 0..31 of the upper / lower half, value = tag | reg index, read back by rows
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
    for (int x = 0; x < 16; ++x) r[x] = 0x80000000u | (h << 24) | (lane << 8) | x;
    tmem_st16x128b_x8(t_lane + (uint32_t(16 * h) << 16), r);
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tmem_ld32(t_lane, v);
  tc_wait_ld();
  for (int x = 0; x < 32; ++x) out_st[(warp * 32 + lane) * 32 + x] = v[x];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 128);
}

int main() {
  uint32_t *d_ld, *d_st;
  cudaMalloc(&d_ld, 4 * 2 * 32 * 32 * 4);
  cudaMalloc(&d_st, 128 * 32 * 4);
  k<<<1, 128>>>(d_ld, d_st);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t ld[4 * 2 * 32 * 32], st[128 * 32];
  cudaMemcpy(ld, d_ld, sizeof(ld), cudaMemcpyDeviceToHost);
  cudaMemcpy(st, d_st, sizeof(st), cudaMemcpyDeviceToHost);
  // expected: reg 4x+e of thread t at (warp, h): lane = 32*warp + 16h + t/4 + (e>=2 ? 8 : 0), column = 64 + 8x + 2(t%4) + (e&1)
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int h = 0; h < 2; ++h) for (int t = 0; t < 32; ++t) for (int r = 0; r < 32; ++r) {
    const int x = r / 4, e = r % 4;
    const uint32_t want = uint32_t(32 * w + 16 * h + t / 4 + (e >= 2 ? 8 : 0)) * 256 + 64 + 8 * x + 2 * (t % 4) + (e & 1);
    const uint32_t got = ld[((w * 2 + h) * 32 + t) * 32 + r];
    if (got != want && bad++ < 12) printf("LD mismatch w%d h%d t%d r%d: got lane %u col %u, want lane %u col %u\n", w, h, t, r, got / 256, got % 256, want / 256, want % 256);
  }
  printf("16x256b.x8 load mapping: %s (%d mismatches)\n", bad ? "DIFFERENT" : "as expected", bad);
  // expected store: reg 2x+e of thread t at half h lands at lane 32w + 16h + t/4 + 8e, column 4x + t%4
  int bad2 = 0;
  for (int L = 0; L < 128; ++L) for (int c = 0; c < 32; ++c) {
    const int w = L / 32, h = (L % 32) / 16, rr = L % 16, e = rr / 8, t = (rr % 8) * 4 + c % 4, x = c / 4;
    const uint32_t want = 0x80000000u | (h << 24) | (t << 8) | (2 * x + e);
    const uint32_t got = st[L * 32 + c];
    if (got != want && bad2++ < 12) printf("ST mismatch lane %d col %d: got %08x want %08x\n", L, c, got, want);
    (void)w;
  }
  printf("16x128b.x8 store mapping: %s (%d mismatches)\n", bad2 ? "DIFFERENT" : "as expected", bad2);
  return 0;
}
