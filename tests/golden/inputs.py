"""Seeded inputs shared by make_golden.py (which runs the reference on them) and the tests (which run the oracle
and the CUDA path on them).  Same convention as every reference script: default_rng(seed), Q then K then V, each
standard_normal((L, d)) cast to the run dtype (e.g. flash_attention_v1/numpy_gpu_like_opt2.py:245-252)."""
import numpy as np

# name -> (seed, L, d): ragged L (not a multiple of the 8-row tiles), d served by every CUDA instantiation
SMALL_CASES = {"v1": (1, 100, 64), "td": (2, 44, 64), "v2": (3, 52, 64)}
# the reference scripts' own __main__ configurations: name -> (L, d, dtype)
MAIN_CASES = {"v1_opt2_main": (1024, 32, np.float64), "v1_basic_main": (2048, 32, np.float16),
              "tiled_d_main": (2048, 128, np.float16), "v2_main": (256, 128, np.float16)}


def qkv(seed, L, d, dtype):
    rng = np.random.default_rng(seed)
    return tuple(rng.standard_normal((L, d)).astype(dtype) for _ in range(3))
