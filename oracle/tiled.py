"""NumPy restatements of the reference's tiled algorithms (the semantics the CUDA kernels implement).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function keeps the reference's signature and buffer
conventions (1-D flattened row-major buffers, output written in place) but replaces the per-element Python loops by
per-tile NumPy expressions, so a whole head runs in milliseconds instead of ~11 s.  State buffers keep the input
dtype exactly like the reference (m, l, O_acc, S, alpha are allocated with Q.dtype).

  flash_attention_tiled        <- flash_attention_v1/numpy_gpu_like_opt2.py:198-241  (process_kv_tile :135-195)
  flash_attention_tiled_d      <- flash_attention_v1_tiled_d/numpy_gpu_like.py:224-270 (process_kv_tile :171-221,
                                  mat_mul_scaled_d_tiled :20-66, mat_scale_rows_mul_add_d_tiled :68-105)
  flash_attention_tiled_global <- flash_attention_v1_tiled_d/numpy_basic.py:99-151
  partial_attention_kernel / reduction_kernel / flash_attention_tiled_v2
                               <- flash_attention_v2/numpy_gpu_like.py:174-226 / :229-288 / :343-405
"""
from __future__ import annotations

import numpy as np


def _process_kv_tile(Q_t, K_t, V_t, m, l, O_acc, d, d_tile_qk=None, d_tile_v=None):
    """One KV-tile update of the streaming softmax (numpy_gpu_like_opt2.py:161-195).

    Q_t [bq,d], K_t/V_t [bk,d]; m,l [bq]; O_acc [bq,d]; all updated in place, all in the buffers' dtype.
    With d_tile_* set, the two contractions are evaluated chunk by chunk along d exactly as the tiled-d variant does
    (flash_attention_v1_tiled_d/numpy_gpu_like.py:40-66, :89-105): partial QK^T sums are accumulated in S's dtype,
    and O columns are produced one d_tile_v slab at a time.
    """
    dt = Q_t.dtype
    scale = dt.type(1.0 / np.sqrt(d))
    if d_tile_qk is None:
        S = ((Q_t @ K_t.T) * scale).astype(dt)                         # mat_mul_scaled :14-33
    else:
        S = np.zeros((Q_t.shape[0], K_t.shape[0]), dtype=dt)
        for d0 in range(0, d, d_tile_qk):                               # tiled-d :40-62
            S += (Q_t[:, d0:d0 + d_tile_qk] @ K_t[:, d0:d0 + d_tile_qk].T).astype(dt)
        S = (S * scale).astype(dt)                                      # :64-66
    new_max = np.maximum(m, S.max(axis=1))                              # opt2 :174-180
    alpha = np.exp(m - new_max).astype(dt)                              # :181  (exp(-inf)=0 on the first tile)
    m[:] = new_max                                                      # :183
    P = np.exp(S - m[:, None]).astype(dt)                               # mat_sub_vec_exp :65-80
    l[:] = (l * alpha + P.sum(axis=1, dtype=dt)).astype(dt)             # row_sum_mul_add_inplace :99-116
    if d_tile_v is None:
        O_acc[:] = (O_acc * alpha[:, None] + (P @ V_t).astype(dt)).astype(dt)   # mat_scale_rows_mul_add :35-63
    else:
        for d0 in range(0, d, d_tile_v):                                # tiled-d :89-105
            sl = slice(d0, d0 + d_tile_v)
            O_acc[:, sl] = (O_acc[:, sl] * alpha[:, None] + (P @ V_t[:, sl]).astype(dt)).astype(dt)


def _attention_rows(Q2, K2, V2, q_start, q_end, k_tiles, Bk, d, d_tile_qk, d_tile_v):
    """Streaming state (m, l, O_acc) for query rows [q_start, q_end) over the given KV tile indices."""
    dt = Q2.dtype
    L = K2.shape[0]
    bq = q_end - q_start
    m = np.full(bq, -np.inf, dtype=dt)
    l = np.zeros(bq, dtype=dt)
    O_acc = np.zeros((bq, d), dtype=dt)
    for t in k_tiles:
        k0, k1 = t * Bk, min(t * Bk + Bk, L)
        _process_kv_tile(Q2[q_start:q_end], K2[k0:k1], V2[k0:k1], m, l, O_acc, d, d_tile_qk, d_tile_v)
    return m, l, O_acc


def flash_attention_tiled(Q, K, V, O, L, d, Bq=8, Bk=8):
    """V1 tile loop; 1-D flattened [L*d] buffers, O written in place (numpy_gpu_like_opt2.py:198-241)."""
    Q2, K2, V2 = (np.asarray(x).reshape(L, d) for x in (Q, K, V))
    O2 = O.reshape(L, d)
    n_kt = (L + Bk - 1) // Bk
    with np.errstate(over="ignore", invalid="ignore"):
        for q0 in range(0, L, Bq):
            q1 = min(q0 + Bq, L)
            m, l, O_acc = _attention_rows(Q2, K2, V2, q0, q1, range(n_kt), Bk, d, None, None)
            O2[q0:q1] = O_acc / l[:, None]                              # mat_div_vec_store :82-97


def flash_attention_tiled_d(Q, K, V, O, L, d, Bq=8, Bk=8, d_tile_qk=16, d_tile_v=16):
    """Tiled-d V1 (reference name: flash_attention_tiled, flash_attention_v1_tiled_d/numpy_gpu_like.py:224-270)."""
    Q2, K2, V2 = (np.asarray(x).reshape(L, d) for x in (Q, K, V))
    O2 = O.reshape(L, d)
    n_kt = (L + Bk - 1) // Bk
    with np.errstate(over="ignore", invalid="ignore"):
        for q0 in range(0, L, Bq):
            q1 = min(q0 + Bq, L)
            m, l, O_acc = _attention_rows(Q2, K2, V2, q0, q1, range(n_kt), Bk, d, d_tile_qk, d_tile_v)
            O2[q0:q1] = O_acc / l[:, None]


def flash_attention_tiled_global(Q, K, V, Bq=8, Bk=8, d_tile_qk=16, d_tile_v=16):
    """2-D [L,d] in, [L,d] out (flash_attention_v1_tiled_d/numpy_basic.py:99-151)."""
    L, d = Q.shape
    O = np.zeros((L, d), dtype=Q.dtype)
    flash_attention_tiled_d(Q.reshape(-1), K.reshape(-1), V.reshape(-1), O.reshape(-1), L, d, Bq, Bk, d_tile_qk,
                            d_tile_v)
    return O


def partial_attention_kernel(Q, K, V, workspace_O, workspace_m, workspace_l, q_tile_idx, kv_block_idx, L, d, Bq, Bk,
                             d_tile_qk, d_tile_v, kv_block_start, kv_block_end):
    """One (q tile, kv block) partial: un-normalised O_acc, m, l into the workspace dicts
    (flash_attention_v2/numpy_gpu_like.py:174-226)."""
    Q2, K2, V2 = (np.asarray(x).reshape(L, d) for x in (Q, K, V))
    q0 = q_tile_idx * Bq
    q1 = min(q0 + Bq, L)
    with np.errstate(over="ignore", invalid="ignore"):
        m, l, O_acc = _attention_rows(Q2, K2, V2, q0, q1, range(kv_block_start, kv_block_end), Bk, d, d_tile_qk,
                                      d_tile_v)
    full = np.zeros(Bq * d, dtype=Q2.dtype)            # reference keeps [Bq*d] even for a short last tile (:204)
    full[: (q1 - q0) * d] = O_acc.reshape(-1)
    mm = np.full(Bq, -np.inf, dtype=Q2.dtype); mm[: q1 - q0] = m
    ll = np.zeros(Bq, dtype=Q2.dtype); ll[: q1 - q0] = l
    workspace_O[(q_tile_idx, kv_block_idx)] = full
    workspace_m[(q_tile_idx, kv_block_idx)] = mm
    workspace_l[(q_tile_idx, kv_block_idx)] = ll


def reduction_kernel(workspace_O, workspace_m, workspace_l, O_final, q_tile_idx, num_kv_blocks, L, d, Bq):
    """Merge the partials of one q tile (flash_attention_v2/numpy_gpu_like.py:229-288):
    m_g = max_k m_k; s_k = float32(exp(m_k - m_g)); O = sum_k O_k s_k / sum_k l_k s_k."""
    q0 = q_tile_idx * Bq
    q1 = min(q0 + Bq, L)
    n = q1 - q0
    pm = np.stack([workspace_m[(q_tile_idx, k)][:n] for k in range(num_kv_blocks)])            # [K, n]
    pl = np.stack([workspace_l[(q_tile_idx, k)][:n] for k in range(num_kv_blocks)])
    pO = np.stack([workspace_O[(q_tile_idx, k)][: n * d].reshape(n, d) for k in range(num_kv_blocks)])  # [K,n,d]
    m_global = pm.max(axis=0)                                                                   # :269-272
    scales = np.exp(pm - m_global[None, :]).astype(np.float32)                                  # :275-278
    l_global = (pl.astype(np.float64) * scales).sum(axis=0)                                     # :276-279 (python float acc)
    numer = (pO.astype(np.float64) * scales[:, :, None]).sum(axis=0)                            # :283-286
    O_final.reshape(L, d)[q0:q1] = (numer / l_global[:, None])                                  # :287


def flash_attention_tiled_v2(Q, K, V, O, workspace_O, workspace_m, workspace_l, L, d, Bq=8, Bk=8, d_tile_qk=16,
                             d_tile_v=16, kv_tiles_per_block=1):
    """Two-kernel split-KV simulation (flash_attention_v2/numpy_gpu_like.py:343-405)."""
    num_q_tiles = (L + Bq - 1) // Bq
    num_kv_tiles = (L + Bk - 1) // Bk
    num_kv_blocks = (num_kv_tiles + kv_tiles_per_block - 1) // kv_tiles_per_block
    for qt in range(num_q_tiles):
        for kb in range(num_kv_blocks):
            s = kb * kv_tiles_per_block
            e = min(s + kv_tiles_per_block, num_kv_tiles)
            partial_attention_kernel(Q, K, V, workspace_O, workspace_m, workspace_l, qt, kb, L, d, Bq, Bk, d_tile_qk,
                                     d_tile_v, s, e)
    for qt in range(num_q_tiles):
        reduction_kernel(workspace_O, workspace_m, workspace_l, O, qt, num_kv_blocks, L, d, Bq)
