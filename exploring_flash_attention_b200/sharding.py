"""(batch, head) sharding across the GPUs of one box, and sequence (context) sharding for long L.

Every (b,h) head is independent — the reference maps them to independent grid rows (blockIdx.y,
flash_attention_v1/CUDA/flash_attention_v1.h:170-172) and independent OpenMP iterations (common/standard.h:41-43) —
so rank r of G simply owns a contiguous slice of the flattened B*H axis and there is NO collective on the compute
path.  `gather_heads` (NCCL all_gather over NVLink on GPUs, gloo in the CPU tests) exists only so one rank can
verify the assembled output.

`ring_attention` is the sequence-sharded path: the multi-GPU generalisation of V2's split-KV
(flash_attention_v2/README.md:5-21).  Rank r owns rows [r*Ls, (r+1)*Ls) of Q, K and V; the K/V shards travel round
a ring (NCCL send/recv over NVLink, overlapped with the partial-attention kernel of the shard already present) and
the N per-shard partials (O~, LSE) are merged by the same combine kernel V2 uses.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def head_range(BH: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice [begin, end) of the flattened head axis owned by `rank`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(BH, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_heads(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """[B,H,L,d] (or [BH,L,d]) -> this rank's [1,n,L,d] view (no copy)."""
    L, d = x.shape[-2:]
    flat = x.reshape(-1, L, d)
    b, e = head_range(flat.shape[0], rank, world)
    return flat[b:e].unsqueeze(0)


def gather_heads(local: torch.Tensor, BH: int, group=None) -> torch.Tensor:
    """All ranks contribute their [1,n_r,L,d] slice; returns the assembled [BH,L,d] on every rank (verification only)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    L, d = local.shape[-2:]
    n_max = (BH + world - 1) // world
    pad = torch.zeros((n_max, L, d), dtype=local.dtype, device=local.device)
    b, e = head_range(BH, rank, world)
    pad[: e - b] = local.reshape(-1, L, d)
    out = torch.empty((world * n_max, L, d), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        rb, re = head_range(BH, r, world)
        parts.append(out[r * n_max: r * n_max + (re - rb)])
    return torch.cat(parts, dim=0)


def ring_attention(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, group=None, partial_fn=None, combine_fn=None):
    """Non-causal attention over a sequence sharded across the ranks of `group`.

    Q, K, V: this rank's [B,H,Ls,d] shards (same Ls on every rank, sequence split in rank order).  Returns this rank's
    [B,H,Ls,d] rows of softmax(Q_all K_all^T / sqrt(d)) V_all.  Step s computes the partial of the local queries
    against the shard that started on rank (r - s) mod N while that shard is already being forwarded to rank r+1.

    partial_fn / combine_fn default to the CUDA kernels (ops.flash_attention_partial, ops.flash_attention_v2_combine);
    the CPU tests inject stand-ins to exercise the ring schedule over gloo.
    """
    if partial_fn is None or combine_fn is None:
        from . import ops
        partial_fn = partial_fn or ops.flash_attention_partial
        combine_fn = combine_fn or ops.flash_attention_v2_combine
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if Q.dim() != 4 or Q.shape != K.shape or Q.shape != V.shape:
        raise ValueError("Q, K, V must be [B,H,Ls,d] shards with identical shapes")
    B, H, Ls, d = Q.shape
    kv = torch.stack([K, V]).contiguous()          # one message per hop: [2,B,H,Ls,d]
    nxt = torch.empty_like(kv) if world > 1 else None
    o_parts = torch.empty((world, B * H, Ls, d), dtype=torch.float32, device=Q.device)
    lse_parts = torch.empty((world, B * H, Ls), dtype=torch.float32, device=Q.device)
    send_to = dist.get_global_rank(group, (rank + 1) % world) if group is not None else (rank + 1) % world
    recv_from = dist.get_global_rank(group, (rank - 1) % world) if group is not None else (rank - 1) % world
    for s in range(world):
        reqs = []
        if s + 1 < world:
            # enqueued behind everything already on the current stream, so the buffer being overwritten (`nxt`, read
            # by the previous step's kernel) is quiescent; runs on the communicator's stream beside this step's kernel
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, kv, send_to, group),
                                           dist.P2POp(dist.irecv, nxt, recv_from, group)])
        partial_fn(Q, kv[0], kv[1], o_parts[s], lse_parts[s])
        for r in reqs:
            r.wait()
        if s + 1 < world:
            kv, nxt = nxt, kv
    return combine_fn(o_parts, lse_parts, Q.dtype, (B, H, Ls, d))
