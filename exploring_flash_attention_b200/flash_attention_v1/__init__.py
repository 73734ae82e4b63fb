"""Drop-in for the reference's V1 entry points.

  flash_attention_tiled(Q, K, V, O, L, d, Bq=8, Bk=8)   flash_attention_v1/numpy_gpu_like_opt2.py:198 (also the
      numpy_gpu_like / _1D / _opt1 variants: same signature, same result)
  flash_attention_tiled_2d(Q, K, V, Bq=8, Bk=8)          flash_attention_v1/numpy_basic.py:69 (returns O)
  flash_attention_v1(Q, K, V, O, B, H, L, d)             flash_attention_v1/CUDA/flash_attention_v1.h:251 (device tensors)

Bq / Bk are accepted and validated like the reference (positive) but only describe the caller's tiling: the sm_100a
kernel always uses its own 2x128-row / 128-key tiles — tile sizes change scheduling, never the result.
"""
from __future__ import annotations

from .. import ops
from .._lib import FlashAttentionError
from .._numpy_bridge import store_head, to_device_head


def _check_tiles(**tiles):
    for name, v in tiles.items():
        if int(v) <= 0:
            raise FlashAttentionError(-1, f"{name} must be positive")


def flash_attention_tiled(Q, K, V, O, L, d, Bq=8, Bk=8):
    """One head, 1-D flattened [L*d] buffers (NumPy or torch), O written in place."""
    _check_tiles(Bq=Bq, Bk=Bk)
    q, k, v = (to_device_head(x, L, d) for x in (Q, K, V))
    store_head(O, ops.flash_attention_v1(q, k, v, sync=True), L, d)


def flash_attention_tiled_2d(Q, K, V, Bq=8, Bk=8):
    """numpy_basic.py form: [L,d] in, [L,d] out."""
    import numpy as np
    L, d = Q.shape
    O = np.zeros((L, d), dtype=np.asarray(Q).dtype)
    flash_attention_tiled(Q, K, V, O, L, d, Bq, Bk)
    return O


def flash_attention_v1(Q, K, V, O, B, H, L, d):
    """Launcher form on device tensors [B,H,L,d]; like the reference it returns after the kernel has finished."""
    if tuple(Q.shape) != (B, H, L, d):
        raise FlashAttentionError(-1, f"Q has shape {tuple(Q.shape)}, expected {(B, H, L, d)}")
    return ops.flash_attention_v1(Q, K, V, O, sync=True)
