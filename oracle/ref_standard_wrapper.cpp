// ref_standard_wrapper.cpp — C-ABI shim around the UNMODIFIED reference header common/standard.h
// (compiled from where it lies under /root/reference via -I; no reference source is copied here).
// Built by oracle/Makefile into oracle/_ref/ (git-ignored, shipped to the GPU box as a prebuilt .so).
// TEST INFRASTRUCTURE ONLY: validates oracle/standard_attention.c and serves as bench.py's
// `--impl reference` / cpu_baseline.kind="reference".
#include "common/standard.h"  // standard_attention_cpu(const DATA_TYPE*, ..., int B, int H, int L, int d)

#if USE_FP64
#define REF_FN ref_standard_attention_cpu_f64
#else
#define REF_FN ref_standard_attention_cpu_f16
#endif

extern "C" void REF_FN(const void* Q, const void* K, const void* V, void* O, int B, int H, int L, int d, int n_threads) {
  if (n_threads > 0) omp_set_num_threads(n_threads);
  standard_attention_cpu(static_cast<const DATA_TYPE*>(Q), static_cast<const DATA_TYPE*>(K),
                         static_cast<const DATA_TYPE*>(V), static_cast<DATA_TYPE*>(O), B, H, L, d);
}
extern "C" int ref_max_threads(void) { return omp_get_max_threads(); }
