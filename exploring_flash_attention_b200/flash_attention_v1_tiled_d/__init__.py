"""Drop-in for the reference's tiled-d entry points (head dim streamed in chunks; d up to 512).

  flash_attention_tiled(Q, K, V, O, L, d, Bq=8, Bk=8, d_tile_qk=16, d_tile_v=16)
        flash_attention_v1_tiled_d/numpy_gpu_like.py:224
  flash_attention_tiled_global(Q, K, V, Bq=8, Bk=8, d_tile_qk=16, d_tile_v=16)
        flash_attention_v1_tiled_d/numpy_basic.py:99
  flash_attention_v1(Q, K, V, O, B, H, L, d, d_tile_qk, d_tile_v)
        flash_attention_v1_tiled_d/CUDA/flash_attention_v1.h:312 (device tensors)

d_tile_qk / d_tile_v must be positive (the NumPy reference handles a ragged last chunk; the CUDA launcher form also
requires them to divide d, like the reference's asserts) and act as streaming-chunk hints only.
"""
from __future__ import annotations

import numpy as np

from .. import ops
from .._lib import FlashAttentionError
from .._numpy_bridge import store_head, to_device_head

BQ, BK, D_TILE_QK, D_TILE_V = 8, 8, 16, 16  # module constants of the reference scripts


def _hint(d, tile):
    """Largest divisor of d that is <= the requested chunk (the kernel wants chunks that divide d)."""
    tile = max(1, min(int(tile), d))
    while d % tile:
        tile -= 1
    return tile


def flash_attention_tiled(Q, K, V, O, L, d, Bq=8, Bk=8, d_tile_qk=16, d_tile_v=16):
    for name, v in dict(Bq=Bq, Bk=Bk, d_tile_qk=d_tile_qk, d_tile_v=d_tile_v).items():
        if int(v) <= 0:
            raise FlashAttentionError(-1, f"{name} must be positive")
    q, k, v = (to_device_head(x, L, d) for x in (Q, K, V))
    out = ops.flash_attention_v1_tiled_d(q, k, v, d_tile_qk=_hint(d, d_tile_qk), d_tile_v=_hint(d, d_tile_v), sync=True)
    store_head(O, out, L, d)


def flash_attention_tiled_global(Q, K, V, Bq=8, Bk=8, d_tile_qk=16, d_tile_v=16):
    L, d = Q.shape
    O = np.zeros((L, d), dtype=np.asarray(Q).dtype)
    flash_attention_tiled(Q, K, V, O, L, d, Bq, Bk, d_tile_qk, d_tile_v)
    return O


def flash_attention_v1(Q, K, V, O, B, H, L, d, d_tile_qk, d_tile_v):
    if tuple(Q.shape) != (B, H, L, d):
        raise FlashAttentionError(-1, f"Q has shape {tuple(Q.shape)}, expected {(B, H, L, d)}")
    return ops.flash_attention_v1_tiled_d(Q, K, V, O, d_tile_qk=d_tile_qk, d_tile_v=d_tile_v, sync=True)
