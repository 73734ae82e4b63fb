/* standard_attention.c — plain-C restatement of the reference's OpenMP CPU attention,
 * common/standard.h:28-102 (standard_attention_cpu).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py):
 * used as the parity checker and as the reported CPU baseline, never by the product path.
 *
 * Same structure as the reference: one OpenMP iteration per (b,h) head (standard.h:41-43), a materialised
 * float scores[L*L] per head (:52), fp32 arithmetic on storage-typed inputs (:58-62), row max / exp / sum /
 * normalise (:67-88), then scores @ V with an fp32 accumulator rounded to the storage type (:91-99).
 * Storage types: 0 = fp32, 1 = bf16, 2 = fp16 (the reference's DATA_TYPE=__half), 3 = fp64 (USE_FP64=1).
 * Differences: 64-bit offsets (the reference uses int, standard.h:46), a thread-count argument, an int status.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static float h2f(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal */
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7F800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f; memcpy(&f, &bits, 4); return f;
}

static uint16_t f2h(float f) { /* round to nearest even, like __float2half */
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u; x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (x > 0x7F800000u ? 0x200u : 0));
  if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);           /* overflow -> inf */
  if (x < 0x33000001u) return (uint16_t)sign;                         /* underflow -> 0 */
  int e = (int)(x >> 23) - 127; uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
  int shift; uint32_t base;
  if (e < -14) { shift = 13 + (-14 - e); base = 0; } else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7FFFFFu; }
  uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) ++q;
  return (uint16_t)(sign | (base + q));
}

static float bf2f(uint16_t b) { uint32_t bits = (uint32_t)b << 16; float f; memcpy(&f, &bits, 4); return f; }
static uint16_t f2bf(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  if ((x & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((x >> 16) | 0x40u);
  x += 0x7FFFu + ((x >> 16) & 1u);
  return (uint16_t)(x >> 16);
}

static inline float load_elem(const void* p, size_t i, int dt) {
  switch (dt) {
    case 0: return ((const float*)p)[i];
    case 1: return bf2f(((const uint16_t*)p)[i]);
    case 2: return h2f(((const uint16_t*)p)[i]);
    default: return (float)((const double*)p)[i];
  }
}
static inline void store_elem(void* p, size_t i, int dt, float v) {
  switch (dt) {
    case 0: ((float*)p)[i] = v; break;
    case 1: ((uint16_t*)p)[i] = f2bf(v); break;
    case 2: ((uint16_t*)p)[i] = f2h(v); break;
    default: ((double*)p)[i] = (double)v; break;
  }
}

/* Row-wise softmax in place: max, exp, sum, normalise (standard.h:67-88). */
static void softmax_rows(float* scores, int L) {
  for (int i = 0; i < L; ++i) {
    float* row = scores + (size_t)i * L;
    float mx = row[0];
    for (int j = 1; j < L; ++j) if (row[j] > mx) mx = row[j];
    float sum = 0.0f;
    for (int j = 0; j < L; ++j) { row[j] = expf(row[j] - mx); sum += row[j]; }
    for (int j = 0; j < L; ++j) row[j] /= sum;
  }
}

/* Heads [head_begin, head_end) of the flattened B*H axis are computed (a bounded sample for the CPU baseline);
 * pass 0, B*H for the whole tensor.  Returns 0, or -1 on bad arguments / allocation failure. */
int oracle_standard_attention_cpu(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d,
                                  int dtype, int n_threads, int head_begin, int head_end) {
  if (B <= 0 || H <= 0 || L <= 0 || d <= 0 || dtype < 0 || dtype > 3) return -1;
  if (head_begin < 0 || head_end > B * H || head_begin > head_end) return -1;
  const float scale = 1.0f / sqrtf((float)d);                                   /* standard.h:38 */
  int failed = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int bh = head_begin; bh < head_end; ++bh) {                              /* :41-43 */
    const size_t base = (size_t)bh * L * d;                                     /* :46 */
    float* scores = (float*)malloc((size_t)L * L * sizeof(float));              /* :52 */
    if (!scores) { failed = 1; continue; }
    if (dtype == 3) {
      /* USE_FP64=1: DATA_TO_FLOAT is the identity, so each product is formed in double and added to the float
       * accumulator in double before rounding back to float (standard.h:58-62, :93-97). */
      const double* q = (const double*)Q + base; const double* k = (const double*)K + base;
      const double* v = (const double*)V + base; double* o = (double*)O + base;
      for (int i = 0; i < L; ++i)
        for (int j = 0; j < L; ++j) {
          float sum = 0.0f;
          for (int c = 0; c < d; ++c) sum = (float)((double)sum + q[(size_t)i * d + c] * k[(size_t)j * d + c]);
          scores[(size_t)i * L + j] = sum * scale;
        }
      softmax_rows(scores, L);
      for (int i = 0; i < L; ++i)
        for (int c = 0; c < d; ++c) {
          float sum = 0.0f;
          for (int j = 0; j < L; ++j) sum = (float)((double)sum + (double)scores[(size_t)i * L + j] * v[(size_t)j * d + c]);
          o[(size_t)i * d + c] = (double)sum;
        }
      free(scores);
      continue;
    }
    float* q = (float*)malloc((size_t)L * d * sizeof(float) * 3);
    if (!q) { failed = 1; free(scores); continue; }
    float* k = q + (size_t)L * d; float* v = k + (size_t)L * d;
    for (size_t i = 0; i < (size_t)L * d; ++i) {                                /* DATA_TO_FLOAT hoisted out of the loops */
      q[i] = load_elem(Q, base + i, dtype); k[i] = load_elem(K, base + i, dtype); v[i] = load_elem(V, base + i, dtype);
    }
    for (int i = 0; i < L; ++i)                                                 /* :55-64 */
      for (int j = 0; j < L; ++j) {
        float sum = 0.0f;
        for (int c = 0; c < d; ++c) sum += q[(size_t)i * d + c] * k[(size_t)j * d + c];
        scores[(size_t)i * L + j] = sum * scale;
      }
    softmax_rows(scores, L);                                                    /* :67-88 */
    for (int i = 0; i < L; ++i)                                                 /* :91-99 */
      for (int c = 0; c < d; ++c) {
        float sum = 0.0f;
        for (int j = 0; j < L; ++j) sum += scores[(size_t)i * L + j] * v[(size_t)j * d + c];
        store_elem(O, base + (size_t)i * d + c, dtype, sum);
      }
    free(scores); free(q);
  }
  return failed ? -1 : 0;
}

/* The reference drivers' input synthesis: srand(seed) once, then ((float)rand() / RAND_MAX) * 2.0f - 1.0f per element,
 * Q then K then V (flash_attention_v1/CUDA/driver.cu:71-75, :137, :168-170).  glibc's rand() stream is reproduced by
 * calling glibc itself; `skip` values are drawn and discarded first (to reach K or V without materialising Q). */
void oracle_driver_uniform(unsigned seed, size_t skip, size_t n, float* out) {
  srand(seed);
  for (size_t i = 0; i < skip; ++i) (void)rand();
  for (size_t i = 0; i < n; ++i) out[i] = ((float)rand() / (float)RAND_MAX) * 2.0f - 1.0f;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
