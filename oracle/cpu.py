"""ctypes access to the C checkers: oracle/liboracle.so (plain-C restatement of common/standard.h:28-102) and
oracle/_ref/libref_standard_*.so (the reference's own standard.h compiled from /root/reference).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_DT = {"float32": 0, "bfloat16": 1, "float16": 2, "float64": 3}


def build(quiet: bool = True) -> None:
    """make -C oracle (liboracle.so always; _ref only where /root/reference exists)."""
    res = subprocess.run(["make", "-C", str(HERE)], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    if not quiet:
        print(res.stdout)


def _load(path: Path) -> ctypes.CDLL:
    if not path.exists():
        raise FileNotFoundError(f"{path} not built (run `make -C oracle`)")
    return ctypes.CDLL(str(path))


def max_threads() -> int:
    return _load(HERE / "liboracle.so").oracle_max_threads()


def standard_attention_cpu(Q, K, V, dtype_name=None, n_threads=0, head_begin=0, head_end=None):
    """Restated standard_attention_cpu. Q,K,V: [B,H,L,d] numpy arrays; for bf16 pass uint16 bit patterns with
    dtype_name='bfloat16'. Returns O (same dtype/shape); only heads [head_begin, head_end) are written."""
    lib = _load(HERE / "liboracle.so")
    B, H, L, d = Q.shape
    name = dtype_name or str(Q.dtype)
    Q, K, V = (np.ascontiguousarray(x) for x in (Q, K, V))
    O = np.zeros_like(Q)
    head_end = B * H if head_end is None else head_end
    fn = lib.oracle_standard_attention_cpu
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 8
    rc = fn(Q.ctypes.data, K.ctypes.data, V.ctypes.data, O.ctypes.data, B, H, L, d, _DT[name], n_threads, head_begin,
            head_end)
    if rc != 0:
        raise RuntimeError("oracle_standard_attention_cpu failed")
    return O


def driver_inputs(B, H, L, d, dtype=np.float16, seed=42):
    """Q, K, V exactly as the reference drivers synthesise them (srand(42), U[-1,1], Q->K->V, rounded to `dtype`)."""
    lib = _load(HERE / "liboracle.so")
    fn = lib.oracle_driver_uniform
    fn.restype = None
    fn.argtypes = [ctypes.c_uint, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
    n = B * H * L * d
    buf = np.empty(3 * n, dtype=np.float32)
    fn(seed, 0, 3 * n, buf.ctypes.data)
    return tuple(buf[k * n:(k + 1) * n].reshape(B, H, L, d).astype(dtype) for k in range(3))


def have_ref() -> bool:
    return (HERE / "_ref" / "libref_standard_f16.so").exists()


def ref_standard_attention_cpu(Q, K, V, n_threads=0):
    """The reference's own standard_attention_cpu (common/standard.h) on float16 or float64 [B,H,L,d] arrays."""
    name = str(Q.dtype)
    if name not in ("float16", "float64"):
        raise ValueError("the reference CPU path is built for __half (USE_FP64=0) and double (USE_FP64=1) only")
    tag = "f16" if name == "float16" else "f64"
    lib = _load(HERE / "_ref" / f"libref_standard_{tag}.so")
    fn = getattr(lib, f"ref_standard_attention_cpu_{tag}")
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5
    B, H, L, d = Q.shape
    Q, K, V = (np.ascontiguousarray(x) for x in (Q, K, V))
    O = np.zeros_like(Q)
    fn(Q.ctypes.data, K.ctypes.data, V.ctypes.data, O.ctypes.data, B, H, L, d, n_threads)
    return O


def ref_max_threads() -> int:
    return _load(HERE / "_ref" / "libref_standard_f16.so").ref_max_threads()
