"""NumPy restatement of the reference oracle, common/reference.py:7-21 (naive_attention).

TEST INFRASTRUCTURE (see oracle/__init__.py).  `naive_attention` follows the reference line by line, including
its dtype behaviour (Q @ K.T is evaluated in the input dtype, then promoted by the np.float64 scale — SURVEY.md
§3.5).  The parity tests use `naive_attention_f64`, which up-casts the already-rounded inputs first so the
oracle carries no rounding of its own.
"""
from __future__ import annotations

import numpy as np


def naive_attention(Q, K, V):
    """softmax(Q K^T / sqrt(d)) V for one [L, d] head — common/reference.py:7-21."""
    L, d = Q.shape
    scale = 1.0 / np.sqrt(d)                               # reference.py:16
    scores = (Q @ K.T) * scale                             # :17
    scores = scores - scores.max(axis=1, keepdims=True)    # :18
    probs = np.exp(scores)                                 # :19
    probs = probs / probs.sum(axis=1, keepdims=True)       # :20
    return probs @ V                                       # :21


def naive_attention_f64(Q, K, V):
    """Same math on float64 copies of the (already rounded) inputs. [L, d] -> [L, d] float64."""
    return naive_attention(np.asarray(Q, dtype=np.float64), np.asarray(K, dtype=np.float64),
                           np.asarray(V, dtype=np.float64))


def naive_attention_ex_f64(Q, K, V, causal=False, kv_len=None):
    """Extended oracle for SURVEY.md §8(f)-1 (not in the reference, which defers "dynamic sequence lengths and causal
    masking", flash_attention_v1/README_v1.md:169): same math as naive_attention with an optional causal mask (row i
    sees keys 0..i), an optional key-padding length (only the first kv_len keys are attended to), rectangular shapes
    (Q [Lq,d], K/V [Lk,d]) and the per-row log-sum-exp of the scaled scores.  float64 -> (O [Lq,d], LSE [Lq])."""
    Q, K, V = (np.asarray(x, dtype=np.float64) for x in (Q, K, V))
    Lq, d = Q.shape
    Lk = K.shape[0]
    s = (Q @ K.T) * (1.0 / np.sqrt(d))
    if causal:
        assert Lq == Lk
        s = np.where(np.tril(np.ones((Lq, Lk), dtype=bool)), s, -np.inf)
    if kv_len is not None:
        s = np.where(np.arange(Lk)[None, :] < kv_len, s, -np.inf)
    m = s.max(axis=1, keepdims=True)
    p = np.exp(s - m)
    l = p.sum(axis=1, keepdims=True)
    return (p / l) @ V, (m + np.log(l))[:, 0]


def merge_partials_f64(O_parts, LSE_parts):
    """Merge of normalised attention partials over disjoint key shards (the reduction of
    flash_attention_v2/numpy_gpu_like.py:269-288 restated on (O~, LSE)):
    O = sum_k exp(LSE_k - LSE) O~_k with LSE = log sum_k exp(LSE_k).  [N,...,L,d], [N,...,L] -> [...,L,d]."""
    O_parts = np.asarray(O_parts, dtype=np.float64)
    LSE_parts = np.asarray(LSE_parts, dtype=np.float64)
    m = LSE_parts.max(axis=0)
    w = np.exp(LSE_parts - m)
    return (w[..., None] * O_parts).sum(axis=0) / w.sum(axis=0)[..., None]


def naive_attention_batched_f64(Q, K, V, heads=None, rows=None):
    """[B,H,L,d] (or [BH,L,d]) inputs -> float64 outputs for the selected flat head indices and query rows.

    `rows` (slice or index array) limits the query rows so L=16384 stays tractable: only a [rows, L] score block is
    ever materialised (common/reference.py materialises [L, L]).
    """
    Q = np.asarray(Q); K = np.asarray(K); V = np.asarray(V)
    L, d = Q.shape[-2:]
    Qf, Kf, Vf = (x.reshape(-1, L, d) for x in (Q, K, V))
    heads = range(Qf.shape[0]) if heads is None else heads
    rows = slice(None) if rows is None else rows
    out = []
    for h in heads:
        q = Qf[h][rows].astype(np.float64)
        k = Kf[h].astype(np.float64)
        v = Vf[h].astype(np.float64)
        s = (q @ k.T) * (1.0 / np.sqrt(d))
        s -= s.max(axis=1, keepdims=True)
        p = np.exp(s)
        p /= p.sum(axis=1, keepdims=True)
        out.append(p @ v)
    return np.stack(out)


def attention_backward_f64(Q, K, V, dO, causal=False):
    """Analytic gradient of naive_attention (common/reference.py:15-21) for one head, float64:
    S = Q K^T / sqrt(d), P = softmax(S) (row-wise; causal: keys above the diagonal excluded), O = P V;
      dV = P^T dO;  dP = dO V^T;  dS = P o (dP - rowsum(dP o P));  dQ = dS K / sqrt(d);  dK = dS^T Q / sqrt(d).
    (rowsum(dP o P) = rowsum(dO o O): the Delta of the flash-attention backward.)  The reference has no backward pass
    (README.md:80-84); tests/test_oracle.py pins this against central finite differences of naive_attention_ex_f64.
    [L,d] x4 -> (dQ, dK, dV)."""
    Q, K, V, dO = (np.asarray(x, dtype=np.float64) for x in (Q, K, V, dO))
    L, d = Q.shape
    scale = 1.0 / np.sqrt(d)
    s = (Q @ K.T) * scale
    if causal:
        s = np.where(np.tril(np.ones((L, K.shape[0]), dtype=bool)), s, -np.inf)
    p = np.exp(s - s.max(axis=1, keepdims=True))
    p /= p.sum(axis=1, keepdims=True)
    dV = p.T @ dO
    dP = dO @ V.T
    dS = p * (dP - (dP * p).sum(axis=1, keepdims=True))
    return (dS @ K) * scale, (dS.T @ Q) * scale, dV
