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


_peer_rings: dict = {}


def _peer_ring(shape, dtype, device, group):
    """Symmetric (peer-mapped) double buffer for the K/V shards + a copy stream, cached per (group, shape, dtype)."""
    import torch.distributed._symmetric_memory as symm_mem
    pg = group if group is not None else dist.group.WORLD
    key = (pg.group_name, tuple(shape), dtype, device.index)
    if key not in _peer_rings:
        buf = symm_mem.empty((2,) + tuple(shape), dtype=dtype, device=device)
        hdl = symm_mem.rendezvous(buf, pg)
        _peer_rings[key] = (buf, hdl, torch.cuda.Stream(device))
    return _peer_rings[key]


def zigzag_shard(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """[..., L, d] -> this rank's causal-ring shard: chunks r and 2N-1-r of the 2N equal chunks of the sequence,
    concatenated (so every rank owns one early and one late chunk and causal work is balanced)."""
    L = x.shape[-2]
    if L % (2 * world) != 0:
        raise ValueError("sequence length must be a multiple of 2 * world size")
    C = L // (2 * world)
    return torch.cat([x[..., rank * C:(rank + 1) * C, :], x[..., (2 * world - 1 - rank) * C:(2 * world - rank) * C, :]],
                     dim=-2)


def zigzag_unshard(shards) -> torch.Tensor:
    """Inverse of zigzag_shard: the per-rank [..., 2C, d] shards (in rank order) -> the [..., L, d] sequence."""
    world = len(shards)
    C = shards[0].shape[-2] // 2
    chunks = [None] * (2 * world)
    for r, sh in enumerate(shards):
        chunks[r] = sh[..., :C, :]
        chunks[2 * world - 1 - r] = sh[..., C:, :]
    return torch.cat(chunks, dim=-2)


def ring_attention(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, group=None, partial_fn=None, combine_fn=None,
                   transport: str = "auto", causal: bool = False):
    """Attention over a sequence sharded across the ranks of `group`.

    Q, K, V: this rank's [B,H,Ls,d] shards (same Ls on every rank).  Returns this rank's [B,H,Ls,d] rows of
    softmax(Q_all K_all^T / sqrt(d)) V_all.  Step s computes the partial of the local queries against the shard that
    started on rank (r - s) mod N while the next shard is already on its way.

    causal=False: the sequence is split in rank order (rank r owns rows [r*Ls, (r+1)*Ls)).
    causal=True:  zig-zag layout (`zigzag_shard`): rank r owns chunks r and 2N-1-r of 2N chunks.  Step 0 is plain
      causal attention over the local shard; a shard from a lower rank j contributes its early chunk to all local
      queries; a shard from a higher rank contributes both chunks to the late local queries only — every step is
      half a dense block, so all ranks do the same work.

    transport "nccl": send/recv pairs on the communicator's stream (works on any backend; its kernels need SMs, so
      under the persistent attention kernel the hop is mostly exposed: 10.9 ms at L=16384 on 8 GPUs).
    transport "peer": every rank PULLS the next shard out of its left neighbour's symmetric-memory buffer with a
      copy-engine transfer over NVLink on a side stream — no SMs involved, the hop hides under the partial kernel
      (4.94 ms for the same problem, 1.3 % above the kernels alone).
    transport "auto" (default): "peer" for CUDA tensors, "nccl" otherwise.

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
    if transport not in ("auto", "nccl", "peer"):
        raise ValueError("transport must be 'auto', 'nccl' or 'peer'")
    if transport == "auto":
        transport = "peer" if Q.is_cuda else "nccl"
    B, H, Ls, d = Q.shape
    o_parts = torch.empty((world, B * H, Ls, d), dtype=torch.float32, device=Q.device)
    lse_parts = torch.empty((world, B * H, Ls), dtype=torch.float32, device=Q.device)
    if causal and Ls % 2 != 0:
        raise ValueError("causal ring attention needs an even number of local rows (two zig-zag chunks)")
    C = Ls // 2

    def step(s, k, v):
        """Partial of the local queries against the shard (k, v) that started on rank (rank - s) mod world -> slot s."""
        if not causal:
            partial_fn(Q, k, v, o_parts[s], lse_parts[s])
        elif s == 0:
            partial_fn(Q, k, v, o_parts[0], lse_parts[0], causal=True)
        elif (rank - s) % world < rank:
            partial_fn(Q, k[:, :, :C], v[:, :, :C], o_parts[s], lse_parts[s])
        else:
            partial_fn(Q[:, :, C:], k, v, o_parts[s][:, C:], lse_parts[s][:, C:])
            o_parts[s][:, :C].zero_()                       # early queries see nothing of a later rank's shard:
            lse_parts[s][:, :C].fill_(float("-inf"))        # weight exp(-inf) = 0 in the merge

    if transport == "peer" and world > 1:
        buf, hdl, copy_stream = _peer_ring((2, B, H, Ls, d), Q.dtype, Q.device, group)
        main = torch.cuda.current_stream(Q.device)
        left = (rank - 1) % world
        cur = 0
        buf[0, 0].copy_(K)
        buf[0, 1].copy_(V)
        for s in range(world):
            arrived = None
            if s + 1 < world:
                copy_stream.wait_stream(main)   # my step s-1 kernel (last reader of buf[1-cur]) and the staging copies
                with torch.cuda.stream(copy_stream):
                    # after this barrier every rank's buf[cur] is complete and nobody still pulls from a buf[1-cur]
                    hdl.barrier(channel=0)
                    src = hdl.get_buffer(left, buf[cur].shape, buf.dtype, storage_offset=cur * buf[0].numel())
                    buf[1 - cur].copy_(src, non_blocking=True)
                    arrived = torch.cuda.Event()
                    arrived.record(copy_stream)
            step(s, buf[cur, 0], buf[cur, 1])
            if arrived is not None:
                main.wait_event(arrived)
                cur = 1 - cur
        hdl.barrier(channel=1)                  # no rank restages buf[0] while a neighbour still pulls from it
        return combine_fn(o_parts, lse_parts, Q.dtype, (B, H, Ls, d))

    kv = torch.stack([K, V]).contiguous()          # one message per hop: [2,B,H,Ls,d]
    nxt = torch.empty_like(kv) if world > 1 else None
    send_to = dist.get_global_rank(group, (rank + 1) % world) if group is not None else (rank + 1) % world
    recv_from = dist.get_global_rank(group, (rank - 1) % world) if group is not None else (rank - 1) % world
    for s in range(world):
        reqs = []
        if s + 1 < world:
            # enqueued behind everything already on the current stream, so the buffer being overwritten (`nxt`, read
            # by the previous step's kernel) is quiescent; runs on the communicator's stream beside this step's kernel
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, kv, send_to, group),
                                           dist.P2POp(dist.irecv, nxt, recv_from, group)])
        step(s, kv[0], kv[1])
        for r in reqs:
            r.wait()
        if s + 1 < world:
            kv, nxt = nxt, kv
    return combine_fn(o_parts, lse_parts, Q.dtype, (B, H, Ls, d))
