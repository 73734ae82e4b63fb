// probe_pipes.cu — dev probe: issue cost (cycles per warp instruction per SM sub-partition) of the instructions the
// softmax inner loop is made of, with 1 and 2 warps per scheduler.  Build: nvcc -arch=sm_100a -O3 -o probe_pipes probe_pipes.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ILP 8
#define ITERS 512

template <int OP>
__device__ __forceinline__ void body(float2 (&a)[ILP], float c, float d) {
#pragma unroll
  for (int k = 0; k < ILP; ++k) {
    if constexpr (OP == 0) {  // FFMA2 reg,reg,reg
      a[k] = __ffma2_rn(a[k], make_float2(c, c), make_float2(d, d));
    } else if constexpr (OP == 1) {  // FADD2 reg,reg
      a[k] = __fadd2_rn(a[k], make_float2(d, d));
    } else if constexpr (OP == 2) {  // FFMA scalar reg,reg,reg
      a[k].x = fmaf(a[k].x, c, d);
    } else if constexpr (OP == 3) {  // FFMA scalar imm
      a[k].x = fmaf(a[k].x, 1.0009765625f, d);
    } else if constexpr (OP == 4) {  // FADD scalar
      a[k].x = a[k].x + d;
    } else if constexpr (OP == 5) {  // MUFU.EX2
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k].x));
    } else if constexpr (OP == 6) {  // FMNMX
      a[k].x = fmaxf(a[k].x, d);
    } else if constexpr (OP == 7) {  // FMNMX3
      asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[k].x) : "f"(c), "f"(d));
    } else if constexpr (OP == 8) {  // F2FP bf16x2 pack
      uint32_t r;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[k].x), "f"(a[k].y));
      a[k].x = __uint_as_float(r);
    } else if constexpr (OP == 9) {  // IMAD (x + t<<23 as mad)
      a[k].x = __int_as_float(__float_as_int(a[k].y) * 8388608 + __float_as_int(a[k].x));
    } else if constexpr (OP == 10) {  // FFMA2 with immediate multiplicand and addend
      a[k] = __ffma2_rn(a[k], make_float2(1.0009765625f, 1.0009765625f), make_float2(0.5f, 0.5f));
    } else if constexpr (OP == 11) {  // FMUL2
      a[k] = __fmul2_rn(a[k], make_float2(c, c));
    } else if constexpr (OP == 12) {  // ex2 f16x2 packed
      uint32_t r = __float_as_uint(a[k].x);
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
      a[k].x = __uint_as_float(r);
    } else if constexpr (OP == 13) {  // ex2 bf16x2 packed
      uint32_t r = __float_as_uint(a[k].x);
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r));
      a[k].x = __uint_as_float(r);
    } else if constexpr (OP == 14) {  // mix: FFMA2 + 2 MUFU + FADD2 per pair (the MUFU path of the softmax)
      a[k] = __ffma2_rn(a[k], make_float2(c, c), make_float2(d, d));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k].x));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k].y));
      a[(k + 1) % ILP] = __fadd2_rn(a[(k + 1) % ILP], a[k]);
    } else if constexpr (OP == 15) {  // scalar mix: 2 FFMA + 2 MUFU + 2 FADD per pair
      a[k].x = fmaf(a[k].x, c, d);
      a[k].y = fmaf(a[k].y, c, d);
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k].x));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k].y));
      a[(k + 1) % ILP].x += a[k].x;
      a[(k + 1) % ILP].y += a[k].y;
    } else if constexpr (OP == 16) {  // HFMA2 bf16
      uint32_t r = __float_as_uint(a[k].x), cc = __float_as_uint(c);
      asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(r) : "r"(cc));
      a[k].x = __uint_as_float(r);
    } else if constexpr (OP == 17) {  // LOP3/PRMT unpack-like: shl
      a[k].x = __uint_as_float(__float_as_uint(a[k].x) << 16);
    } else if constexpr (OP == 18) {  // FHADD: fp32 += fp16 (mixed-precision add)
      uint16_t hh = (uint16_t)__float_as_uint(a[k].y);
      asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(a[k].x) : "h"(hh));
    } else if constexpr (OP == 19) {  // HADD2.F16
      uint32_t r = __float_as_uint(a[k].x), cc = __float_as_uint(c);
      asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r) : "r"(cc));
      a[k].x = __uint_as_float(r);
    } else if constexpr (OP == 20) {  // candidate fp16 softmax pair: FFMA2 + F2FP.F16 + MUFU.f16x2 + HADD2
      float2 v = __ffma2_rn(a[k], make_float2(c, c), make_float2(d, d));
      uint32_t r;
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v.y), "f"(v.x));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
      uint32_t acc = __float_as_uint(a[(k + 1) % ILP].y);
      asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(acc) : "r"(r));
      a[(k + 1) % ILP].y = __uint_as_float(acc);
      a[k].x = __uint_as_float(r);
    }
  }
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, float c, float d) {
  float2 a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) body<OP>(a, c, d);
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_body) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 16}) {
    k<OP><<<148, warps * 32>>>(out, cyc, 0.999f, 1e-4f);
    k<OP><<<148, warps * 32>>>(out, cyc, 0.999f, 1e-4f);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = double(h) / (double(ITERS) * ILP * instr_per_body) / (warps / 4);
    printf("%-34s warps/SMSP %d: %.2f cyc per warp-instr per SMSP  (%lld cyc)\n", name, warps / 4, per, h);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("FFMA2 r,r,r", 1);
  run<10>("FFMA2 r,imm,imm", 1);
  run<1>("FADD2 r,r", 1);
  run<11>("FMUL2 r,r", 1);
  run<2>("FFMA r,r,r", 1);
  run<3>("FFMA r,imm,r", 1);
  run<4>("FADD r,r", 1);
  run<5>("MUFU.EX2", 1);
  run<12>("MUFU.EX2 f16x2", 1);
  run<13>("MUFU.EX2 bf16x2", 1);
  run<6>("FMNMX", 1);
  run<7>("FMNMX3", 1);
  run<8>("F2FP.BF16 pack", 1);
  run<9>("IMAD", 1);
  run<16>("HFMA2.BF16", 1);
  run<17>("SHL", 1);
  run<18>("FHADD f32+=f16", 1);
  run<19>("HADD2.F16", 1);
  run<20>("mix FFMA2+F2FP+MUFUf16x2+HADD2 (per pair)", 1);
  run<14>("mix FFMA2+2MUFU+FADD2 (per pair)", 1);
  run<15>("mix 2FFMA+2MUFU+2FADD (per pair)", 1);
  return 0;
}
