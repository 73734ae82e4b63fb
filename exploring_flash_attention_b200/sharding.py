"""(batch, head) sharding across the GPUs of one box.

Every (b,h) head is independent — the reference maps them to independent grid rows (blockIdx.y,
flash_attention_v1/CUDA/flash_attention_v1.h:170-172) and independent OpenMP iterations (common/standard.h:41-43) —
so rank r of G simply owns a contiguous slice of the flattened B*H axis and there is NO collective on the compute
path.  `gather_heads` (NCCL all_gather over NVLink on GPUs, gloo in the CPU tests) exists only so one rank can
verify the assembled output.
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
