"""Drop-in for the reference's V2 split-KV entry points.

  flash_attention_tiled_v2(Q, K, V, O, workspace_O, workspace_m, workspace_l, L, d, Bq=8, Bk=8, d_tile_qk=16,
                           d_tile_v=16, kv_tiles_per_block=1)        flash_attention_v2/numpy_gpu_like.py:343
  partial_attention_kernel(...) / reduction_kernel(...)               :174 / :229  (reference signatures, one block /
      one q tile per call) and partial_attention_grid / reduction_grid (whole-grid forms)
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


def partial_attention_grid(Q, K, V, L, d, Bk=8, kv_tiles_per_block=1):
    """Whole-grid form of KERNEL 1: returns the device workspace (Oaccum [S,1,L,d], LSEaccum [S,1,L])."""
    q, k, v = (to_device_head(x, L, d) for x in (Q, K, V))
    return ops.flash_attention_v2_splitkv(q, k, v, _splits(L, Bk, kv_tiles_per_block))


def reduction_grid(Oaccum, LSEaccum, O_final, L, d):
    """Whole-grid form of KERNEL 2: merges a device workspace into O_final (in place)."""
    dt = torch.float16 if (not isinstance(O_final, torch.Tensor) and O_final.dtype == np.float16) else torch.float32
    if isinstance(O_final, torch.Tensor):
        dt = O_final.dtype if O_final.dtype != torch.float64 else torch.float32
    out = ops.flash_attention_v2_combine(Oaccum, LSEaccum, dt, (1, 1, L, d))
    torch.cuda.current_stream().synchronize()
    store_head(O_final, out, L, d)


def partial_attention_kernel(Q, K, V, workspace_O, workspace_m, workspace_l, q_tile_idx, kv_block_idx, L, d, Bq, Bk,
                             d_tile_qk, d_tile_v, kv_block_start, kv_block_end):
    """Reference signature (flash_attention_v2/numpy_gpu_like.py:174-226): the partial of ONE (q tile, kv block).

    The GPU grid computes every block of the head at once, so this per-block form runs the split-KV kernel for the head
    and keeps the requested entry; a caller that loops over blocks like the reference's driver loop (:378-391) gets the
    same dict contents, just with redundant launches — use flash_attention_tiled_v2 / partial_attention_grid instead.
    The kv block must be one of the uniform blocks of the reference's partition (kv_block_idx * tiles .. + tiles)."""
    for name, val in dict(Bq=Bq, Bk=Bk, d_tile_qk=d_tile_qk, d_tile_v=d_tile_v).items():
        if int(val) <= 0:
            raise FlashAttentionError(-1, f"{name} must be positive")
    n_kv_tiles = (L + Bk - 1) // Bk
    tiles = kv_block_end - kv_block_start
    if kv_block_idx > 0:
        tiles = kv_block_start // kv_block_idx            # uniform block size implied by the block's position
    if tiles <= 0 or kv_block_start != kv_block_idx * tiles or kv_block_end != min(kv_block_start + tiles, n_kv_tiles):
        raise FlashAttentionError(-1, "kv block is not part of a uniform kv_tiles_per_block partition")
    Oacc, LSE = partial_attention_grid(Q, K, V, L, d, Bk, tiles)
    torch.cuda.current_stream().synchronize()
    dt = np.float32 if isinstance(Q, torch.Tensor) else np.asarray(Q).dtype
    r0, r1 = q_tile_idx * Bq, min(q_tile_idx * Bq + Bq, L)
    o = np.zeros(Bq * d, dtype=dt)
    o[: (r1 - r0) * d] = Oacc[kv_block_idx, 0, r0:r1].reshape(-1).cpu().numpy()
    m = np.full(Bq, -np.inf, dtype=dt)
    m[: r1 - r0] = LSE[kv_block_idx, 0, r0:r1].cpu().numpy()
    l = np.zeros(Bq, dtype=dt)
    l[: r1 - r0] = 1
    workspace_O[(q_tile_idx, kv_block_idx)] = o
    workspace_m[(q_tile_idx, kv_block_idx)] = m
    workspace_l[(q_tile_idx, kv_block_idx)] = l


def reduction_kernel(workspace_O, workspace_m, workspace_l, O_final, q_tile_idx, num_kv_blocks, L, d, Bq):
    """Reference signature (flash_attention_v2/numpy_gpu_like.py:229-288): merge the partials of ONE q tile on the GPU
    with the combine kernel.  Accepts any (O, m, l) triples that satisfy the reference's merge formula — the ones written
    by partial_attention_kernel above (O normalised, m = LSE, l = 1) or the reference's own (un-normalised O, m, l)."""
    r0, r1 = q_tile_idx * Bq, min(q_tile_idx * Bq + Bq, L)
    n = r1 - r0
    dev = torch.device("cuda", torch.cuda.current_device())
    O_parts = np.stack([np.asarray(workspace_O[(q_tile_idx, k)], dtype=np.float64)[: n * d].reshape(n, d)
                        for k in range(num_kv_blocks)])
    m = np.stack([np.asarray(workspace_m[(q_tile_idx, k)], dtype=np.float64)[:n] for k in range(num_kv_blocks)])
    l = np.stack([np.asarray(workspace_l[(q_tile_idx, k)], dtype=np.float64)[:n] for k in range(num_kv_blocks)])
    with np.errstate(divide="ignore", invalid="ignore"):
        lse = m + np.log(l)                                  # (O, m, l) -> split-normalised partial + LSE
        On = np.where(l[..., None] > 0, O_parts / l[..., None], 0.0)
    Oacc = torch.from_numpy(On.astype(np.float32)).to(dev).reshape(num_kv_blocks, 1, n, d).contiguous()
    LSE = torch.from_numpy(lse.astype(np.float32)).to(dev).reshape(num_kv_blocks, 1, n).contiguous()
    out = ops.flash_attention_v2_combine(Oacc, LSE, torch.float32, (1, 1, n, d))
    torch.cuda.current_stream().synchronize()
    res = out.reshape(n, d).cpu().numpy()
    if isinstance(O_final, torch.Tensor):
        O_final.reshape(L, d)[r0:r1] = torch.from_numpy(res).to(O_final.dtype)
    else:
        O_final.reshape(L, d)[r0:r1] = res.astype(O_final.dtype)


def flash_attention_v2(Q, K, V, O, B, H, L, d, d_tile_qk, d_tile_v, kv_tiles_per_block, Bk=16, workspace=None):
    """Launcher form on device tensors [B,H,L,d] (reference BK is the compile-time 16, flash_attention_v2.h:35)."""
    if tuple(Q.shape) != (B, H, L, d):
        raise FlashAttentionError(-1, f"Q has shape {tuple(Q.shape)}, expected {(B, H, L, d)}")
    if d_tile_qk <= 0 or d_tile_v <= 0 or d % d_tile_qk or d % d_tile_v:
        raise FlashAttentionError(-1, "d_tile_qk and d_tile_v must be positive divisors of d")
    return ops.flash_attention_v2(Q, K, V, _splits(L, Bk, kv_tiles_per_block), O=O, workspace=workspace, sync=True)
