// probe_umma.cu — single-CTA check of the hand-built UMMA/TMA encodings used by fa_fwd_sm100.cuh:
//   (1) TMA 128B-swizzled loads + K-major smem descriptors + SS tcgen05.mma      ->  S = Q K^T
//   (2) tcgen05.st of P over S in place + MN-major V descriptor + TS tcgen05.mma ->  O = P V   (P = S/16)
// against a CPU computation, for each (head dim, dtype) the library instantiates.  Test tool only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_umma probe_umma.cu && ./probe_umma
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../exploring_flash_attention_b200/csrc/fa_fwd_sm100.cuh"

using namespace fa;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e)); \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

template <int D, int DT>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, float* outS, float* outO) {
  using T = FwdTraits<D, DT>;
  constexpr int NBLK = T::NBLK, BLK_ELEMS = T::BLK_ELEMS, BLK_BYTES = T::BLK_BYTES, TILE_BYTES = T::TILE_BYTES;
  constexpr int UK = T::UK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = smem + 2 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * TILE_BYTES);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 3 * TILE_BYTES);
    for (int b = 0; b < NBLK; ++b) {
      tma_load_3d(sQ + b * BLK_BYTES, &tmQ, &bars[0], b * BLK_ELEMS, 0, 0);
      tma_load_3d(sK + b * BLK_BYTES, &tmK, &bars[0], b * BLK_ELEMS, 0, 0);
      tma_load_3d(sV + b * BLK_BYTES, &tmV, &bars[0], b * BLK_ELEMS, 0, 0);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc_qk = make_idesc(T::FMT, 128, 128, 0, 0);
    constexpr uint64_t hiK = make_smem_desc_hi(16, 1024, SWZ_128B);
    for (int k = 0; k < D / UK; ++k) {
      const uint32_t off = (k / 4) * BLK_BYTES + (k % 4) * 32;
      umma_ss<T::KIND>(tb, make_smem_desc(smem_u32(sQ) + off, hiK), make_smem_desc(smem_u32(sK) + off, hiK), idesc_qk,
                       k > 0);
    }
    tc_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = threadIdx.x;
  const uint32_t t_lane = tb + (uint32_t(warp * 32) << 16);
  uint32_t s[4][32];
  for (int c = 0; c < 4; ++c) tmem_ld32(t_lane + c * 32, s[c]);
  tc_wait_ld();
  for (int c = 0; c < 4; ++c)
    for (int x = 0; x < 32; ++x) {
      outS[row * 128 + c * 32 + x] = __uint_as_float(s[c][x]);
      s[c][x] = __float_as_uint(__uint_as_float(s[c][x]) * 0.0625f);
    }
  if constexpr (DT == DT_F32) {
    for (int c = 0; c < 4; ++c) tmem_st32(t_lane + c * 32, s[c]);
  } else {
    uint32_t pk[2][32];
    for (int c = 0; c < 4; ++c)
      for (int x = 0; x < 16; ++x) {
        const float a = __uint_as_float(s[c][2 * x]), b = __uint_as_float(s[c][2 * x + 1]);
        pk[c >> 1][(c & 1) * 16 + x] = (DT == DT_BF16) ? pack_bf16x2(a, b) : pack_f16x2(a, b);
      }
    tmem_st32(t_lane, pk[0]);
    tmem_st32(t_lane + 32, pk[1]);
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    constexpr uint32_t idesc_pv = make_idesc(T::FMT, 128, D, 0, 1);
    constexpr uint64_t hiV = (DT == DT_F32) ? make_smem_desc_hi(BLK_BYTES, 512, SWZ_128B_BASE32B)
                                            : make_smem_desc_hi(BLK_BYTES, 1024, SWZ_128B);
    for (int kk = 0; kk < 128 / UK; ++kk)
      umma_ts<T::KIND>(tb + 256, tb + kk * 8, make_smem_desc(smem_u32(sV) + kk * UK * 128, hiV), idesc_pv, kk > 0);
    tc_commit(&bars[2]);
  }
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  for (int c = 0; c < D / 32; ++c) {
    uint32_t o[32];
    tmem_ld32(t_lane + 256 + c * 32, o);
    tc_wait_ld();
    for (int x = 0; x < 32; ++x) outO[row * D + c * 32 + x] = __uint_as_float(o[x]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float round_to(float v, int dt) {
  if (dt == DT_BF16) return __bfloat162float(__float2bfloat16(v));
  if (dt == DT_F16) return __half2float(__float2half(v));
  return v;
}
static float tf32_trunc(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}

template <int D, int DT>
int run(EncodeFn enc) {
  using T = FwdTraits<D, DT>;
  const int es = T::ES;
  const size_t n = 128 * D;
  std::vector<float> q(n), k(n), v(n);
  srand(7 + D + DT);
  for (size_t i = 0; i < n; ++i) {
    q[i] = round_to(rand() / float(RAND_MAX) * 2 - 1, DT);
    k[i] = round_to(rand() / float(RAND_MAX) * 2 - 1, DT);
    v[i] = round_to(rand() / float(RAND_MAX) * 2 - 1, DT);
  }
  std::vector<uint8_t> hq(n * es), hk(n * es), hv(n * es);
  auto pack = [&](const std::vector<float>& src, std::vector<uint8_t>& dst) {
    for (size_t i = 0; i < n; ++i) {
      if (DT == DT_F32) {
        memcpy(&dst[i * 4], &src[i], 4);
      } else if (DT == DT_BF16) {
        __nv_bfloat16 b = __float2bfloat16(src[i]);
        memcpy(&dst[i * 2], &b, 2);
      } else {
        __half b = __float2half(src[i]);
        memcpy(&dst[i * 2], &b, 2);
      }
    }
  };
  pack(q, hq);
  pack(k, hk);
  pack(v, hv);
  void *dq, *dk, *dv;
  float *dS, *dO;
  CK(cudaMalloc(&dq, n * es));
  CK(cudaMalloc(&dk, n * es));
  CK(cudaMalloc(&dv, n * es));
  CK(cudaMalloc(&dS, 128 * 128 * 4));
  CK(cudaMalloc(&dO, 128 * D * 4));
  CK(cudaMemcpy(dq, hq.data(), n * es, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, hk.data(), n * es, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, hv.data(), n * es, cudaMemcpyHostToDevice));
  CUtensorMap maps[3];
  void* ptrs[3] = {dq, dk, dv};
  for (int i = 0; i < 3; ++i) {
    CUtensorMapDataType dt = DT == DT_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                             : DT == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                             : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t dims[3] = {cuuint64_t(D), 128, 1};
    cuuint64_t strides[2] = {cuuint64_t(D) * es, cuuint64_t(128) * D * es};
    cuuint32_t box[3] = {cuuint32_t(128 / es), 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle swz = (i == 2 && DT == DT_F32) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(&maps[i], dt, 3, ptrs[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("encode failed %d\n", int(r));
      return 2;
    }
  }
  const int smem = 1024 + 3 * T::TILE_BYTES + 64;
  CK(cudaFuncSetAttribute(probe_kernel<D, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<D, DT><<<1, 128, smem>>>(maps[0], maps[1], maps[2], dS, dO);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> S(128 * 128), O(128 * D);
  CK(cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  // CPU
  double errS = 0, errO = 0, magS = 0, magO = 0;
  std::vector<float> Pref(128 * 128);
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      double acc = 0;
      for (int c = 0; c < D; ++c) {
        float a = q[i * D + c], b = k[j * D + c];
        if (DT == DT_F32) a = tf32_trunc(a), b = tf32_trunc(b);
        acc += double(a) * b;
      }
      errS = fmax(errS, fabs(acc - S[i * 128 + j]));
      magS = fmax(magS, fabs(acc));
      // P as the GPU rounds it (from the GPU's own S, so (2) is tested independently of (1)'s rounding)
      float pv = S[i * 128 + j] * 0.0625f;
      Pref[i * 128 + j] = DT == DT_F32 ? tf32_trunc(pv) : round_to(pv, DT);
    }
  for (int i = 0; i < 128; ++i)
    for (int c = 0; c < D; ++c) {
      double acc = 0;
      for (int j = 0; j < 128; ++j) {
        float b = v[j * D + c];
        if (DT == DT_F32) b = tf32_trunc(b);
        acc += double(Pref[i * 128 + j]) * b;
      }
      errO = fmax(errO, fabs(acc - O[i * D + c]));
      magO = fmax(magO, fabs(acc));
    }
  const bool okS = errS < 2e-2 * fmax(1.0, magS) * (DT == DT_F32 ? 1 : 0.05), okO = errO < 2e-2 * fmax(1.0, magO);
  printf("probe D=%d dtype=%d : S max|err|=%.3e (max|S|=%.2f) %s ; O max|err|=%.3e (max|O|=%.2f) %s\n", D, DT, errS,
         magS, okS ? "PASS" : "FAIL", errO, magO, okO ? "PASS" : "FAIL");
  if (!okS) {
    printf("  S[0][0..7] gpu:");
    for (int j = 0; j < 8; ++j) printf(" %.4f", S[j]);
    printf("\n");
  }
  if (!okO) {
    printf("  O[0][0..7] gpu:");
    for (int j = 0; j < 8; ++j) printf(" %.4f", O[j]);
    printf("\n");
  }
  cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dS); cudaFree(dO);
  return (okS && okO) ? 0 : 1;
}

int main(int argc, char** argv) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaFree(0));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
  EncodeFn enc = reinterpret_cast<EncodeFn>(p);
  const int which = argc > 1 ? atoi(argv[1]) : -1;
  int rc = 0;
  if (which < 0 || which == 0) rc |= run<128, DT_BF16>(enc);
  if (which < 0 || which == 1) rc |= run<64, DT_BF16>(enc);
  if (which < 0 || which == 2) rc |= run<128, DT_F16>(enc);
  if (which < 0 || which == 3) rc |= run<32, DT_F32>(enc);
  if (which < 0 || which == 4) rc |= run<64, DT_F32>(enc);
  printf(rc ? "PROBE: FAIL\n" : "PROBE: ALL PASS\n");
  return rc;
}
