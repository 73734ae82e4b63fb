"""Drop-in for the reference's V2 split-KV entry points.

  flash_attention_tiled_v2(Q, K, V, O, workspace_O, workspace_m, workspace_l, L, d, Bq=8, Bk=8, d_tile_qk=16,
                           d_tile_v=16, kv_tiles_per_block=1)        flash_attention_v2/numpy_gpu_like.py:343
  partial_attention_kernel(...) / reduction_kernel(...)               :174 / :229  (whole-grid forms, see below)
  flash_attention_v2(Q, K, V, O, B, H, L, d, d_tile_qk, d_tile_v, kv_tiles_per_block)
                                                                      flash_attention_v2/CUDA/flash_attention_v2.h:438

A split ("kv block") covers Bk * kv_tiles_per_block consecutive keys, exactly the reference's partition
(numpy_gpu_like.py:378-385, flash_attention_v2.h:449-451).  The device workspace is split-major fp32
(Oaccum [S,BH,L,d] normalised partials, LSEaccum [S,BH,L]).  When the caller passes the reference's workspace dicts
they are filled per (q_tile_idx, kv_block_idx) with the equivalent triple (O = normalised partial, m = LSE, l = 1):
the reference's own merge  O = sum_k O_k e^{m_k-m_g} / sum_k l_k e^{m_k-m_g}  (numpy_gpu_like.py:269-288) applied
to it yields the same output.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import FlashAttentionError
from .._numpy_bridge import store_head, to_device_head

BQ, BK, D_TILE_QK, D_TILE_V, KV_TILES_PER_BLOCK = 8, 8, 16, 16, 4  # module constants of the reference script


def _splits(L, Bk, kv_tiles_per_block):
    for name, v in dict(Bk=Bk, kv_tiles_per_block=kv_tiles_per_block).items():
        if int(v) <= 0:
            raise FlashAttentionError(-1, f"{name} must be positive")
    return int(Bk) * int(kv_tiles_per_block)


def _fill_workspace_dicts(workspace_O, workspace_m, workspace_l, Oacc, LSE, L, d, Bq, dtype):
    S = Oacc.shape[0]
    Oh = Oacc.reshape(S, L, d).cpu().numpy()
    Lh = LSE.reshape(S, L).cpu().numpy()
    for qt in range((L + Bq - 1) // Bq):
        r0, r1 = qt * Bq, min(qt * Bq + Bq, L)
        for kb in range(S):
            o = np.zeros(Bq * d, dtype=dtype); o[: (r1 - r0) * d] = Oh[kb, r0:r1].reshape(-1)
            m = np.full(Bq, -np.inf, dtype=dtype); m[: r1 - r0] = Lh[kb, r0:r1]
            l = np.zeros(Bq, dtype=dtype); l[: r1 - r0] = 1
            workspace_O[(qt, kb)], workspace_m[(qt, kb)], workspace_l[(qt, kb)] = o, m, l


def flash_attention_tiled_v2(Q, K, V, O, workspace_O, workspace_m, workspace_l, L, d, Bq=8, Bk=8, d_tile_qk=16,
                             d_tile_v=16, kv_tiles_per_block=1):
    """One head; 1-D flattened buffers; O written in place; workspace dicts filled (may be None to skip)."""
    if int(Bq) <= 0:
        raise FlashAttentionError(-1, "Bq must be positive")
    kv_per_split = _splits(L, Bk, kv_tiles_per_block)
    q, k, v = (to_device_head(x, L, d) for x in (Q, K, V))
    Oacc, LSE = ops.flash_attention_v2_splitkv(q, k, v, kv_per_split)
    out = ops.flash_attention_v2_combine(Oacc, LSE, q.dtype, (1, 1, L, d))
    torch.cuda.current_stream().synchronize()
    store_head(O, out, L, d)
    if workspace_O is not None and workspace_m is not None and workspace_l is not None:
        dt = np.float32 if isinstance(Q, torch.Tensor) else np.asarray(Q).dtype
        _fill_workspace_dicts(workspace_O, workspace_m, workspace_l, Oacc, LSE, L, d, int(Bq), dt)


def partial_attention_kernel(Q, K, V, L, d, Bk=8, kv_tiles_per_block=1):
    """Whole-grid form of KERNEL 1: returns the device workspace (Oaccum [S,1,L,d], LSEaccum [S,1,L])."""
    q, k, v = (to_device_head(x, L, d) for x in (Q, K, V))
    return ops.flash_attention_v2_splitkv(q, k, v, _splits(L, Bk, kv_tiles_per_block))


def reduction_kernel(Oaccum, LSEaccum, O_final, L, d):
    """Whole-grid form of KERNEL 2: merges a device workspace into O_final (in place)."""
    dt = torch.float16 if (not isinstance(O_final, torch.Tensor) and O_final.dtype == np.float16) else torch.float32
    if isinstance(O_final, torch.Tensor):
        dt = O_final.dtype if O_final.dtype != torch.float64 else torch.float32
    out = ops.flash_attention_v2_combine(Oaccum, LSEaccum, dt, (1, 1, L, d))
    torch.cuda.current_stream().synchronize()
    store_head(O_final, out, L, d)


def flash_attention_v2(Q, K, V, O, B, H, L, d, d_tile_qk, d_tile_v, kv_tiles_per_block, Bk=16, workspace=None):
    """Launcher form on device tensors [B,H,L,d] (reference BK is the compile-time 16, flash_attention_v2.h:35)."""
    if tuple(Q.shape) != (B, H, L, d):
        raise FlashAttentionError(-1, f"Q has shape {tuple(Q.shape)}, expected {(B, H, L, d)}")
    if d_tile_qk <= 0 or d_tile_v <= 0 or d % d_tile_qk or d % d_tile_v:
        raise FlashAttentionError(-1, "d_tile_qk and d_tile_v must be positive divisors of d")
    return ops.flash_attention_v2(Q, K, V, _splits(L, Bk, kv_tiles_per_block), O=O, workspace=workspace, sync=True)
