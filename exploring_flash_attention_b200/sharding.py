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
    """Peer-mapped staging buffer for this rank's K/V shard ([2,B,H,Ls,d] in symmetric memory), two local receive
    buffers and two copy streams (K and V travel on separate copy engines); cached per (group, shape, dtype)."""
    import torch.distributed._symmetric_memory as symm_mem
    pg = group if group is not None else dist.group.WORLD
    key = (pg.group_name, tuple(shape), dtype, device.index)
    if key not in _peer_rings:
        src = symm_mem.empty((2,) + tuple(shape), dtype=dtype, device=device)
        hdl = symm_mem.rendezvous(src, pg)
        recv = torch.empty((2, 2) + tuple(shape), dtype=dtype, device=device)
        _peer_rings[key] = (src, hdl, (torch.cuda.Stream(device), torch.cuda.Stream(device)), recv)
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
    transport "peer": every shard is staged once in its owner's symmetric memory and every rank PULLS the shard it
      needs next straight from the owner with copy-engine transfers over NVLink/NVSwitch on side streams — no SMs
      involved, the transfer hides under the partial kernel (4.94 ms for the same problem, 1.3 % above the kernels
      alone).
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
        # NVSwitch gives every GPU full bandwidth to every peer, so nothing is forwarded: each shard is staged once in
        # its owner's symmetric memory and at step s rank r pulls the shard of rank (r - s) mod N straight from there
        # (at any step the N pulls form a permutation: one reader per source).
        src, hdl, streams, recv = _peer_ring((B, H, Ls, d), Q.dtype, Q.device, group)
        main = torch.cuda.current_stream(Q.device)
        src[0].copy_(K)
        src[1].copy_(V)
        hdl.barrier(channel=0)                  # every shard is staged (and every rank has left the previous call)

        def pull(s):
            peer = hdl.get_buffer((rank - s) % world, src.shape, src.dtype)
            events = []
            for t, st in enumerate(streams):
                st.wait_stream(main)            # the kernel of step s-2, last reader of recv[s % 2], is already enqueued
                with torch.cuda.stream(st):
                    recv[s % 2, t].copy_(peer[t], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(st)
                    events.append(ev)
            return events

        pending = pull(1)
        for s in range(world):
            if s == 0:
                k, v = K, V
            else:
                for ev in pending:
                    main.wait_event(ev)
                k, v = recv[s % 2, 0], recv[s % 2, 1]
                if s + 1 < world:
                    pending = pull(s + 1)       # travels while the kernel of step s runs
            step(s, k, v)
        hdl.barrier(channel=1)                  # nobody restages its shard while a peer may still be pulling it
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


# ---------------------------------------------------------------------------------------------------------------------
# Sequence-sharded attention by head exchange (all-to-all), the NVSwitch-native alternative to the ring
# ---------------------------------------------------------------------------------------------------------------------
_a2a_state: dict = {}


def _a2a_buffers(BH, hpr, Ls, L, d, dtype, device, group):
    """Symmetric (peer-mapped) staging for the local Q, K, V rows ([3,BH,Ls,d]) and for the outputs of the heads this rank
    computes ([hpr,L,d]); two local [3,hc,L,d] assembly buffers are allocated by the caller per chunk size."""
    import torch.distributed._symmetric_memory as symm_mem
    pg = group if group is not None else dist.group.WORLD
    key = (pg.group_name, BH, Ls, d, dtype, device.index)
    if key not in _a2a_state:
        src = symm_mem.empty((3, BH, Ls, d), dtype=dtype, device=device)
        out = symm_mem.empty((hpr, L, d), dtype=dtype, device=device)
        h_src = symm_mem.rendezvous(src, pg)
        h_out = symm_mem.rendezvous(out, pg)
        streams = tuple(torch.cuda.Stream(device) for _ in range(4))   # Q, K, V pulls + the output pulls
        _a2a_state[key] = (src, out, h_src, h_out, streams, {})
    return _a2a_state[key]


def alltoall_attention(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, group=None, causal: bool = False,
                       transport: str = "auto", chunks: int = 4, attn_fn=None):
    """Attention over a sequence sharded across the ranks of `group`, by exchanging heads instead of circulating K/V.

    Q, K, V: this rank's [B,H,Ls,d] rows (rank r owns rows [r*Ls, (r+1)*Ls), also when causal); B*H must be a multiple
    of the world size.  Every rank fetches, for ITS share of the heads, the rows of all ranks (one all-to-all), runs the
    plain fused-tile kernel over the full sequence for those heads — no partials, no merge, causal balanced for free —
    and the outputs travel back by the inverse exchange.  On NVSwitch every GPU reaches every peer at full bandwidth,
    so the exchange costs 2*(N-1)/N of one Q,K,V,O pass over NVLink instead of the ring's N partial passes through HBM
    (the ring writes and re-reads N fp32 partials: 8x the output bytes on 8 GPUs); the same (b,h) independence that makes
    head sharding communication-free (flash_attention_v1/CUDA/flash_attention_v1.h:170-172) is what allows it.

    transport "peer" (CUDA default): rows are staged once in symmetric memory; each rank PULLS the blocks it needs with
      copy-engine 2-D copies (fa_copy_2d_multi_async) on side streams, in up to `chunks` head groups sized to the persistent
      kernel's round quantisation (_a2a_chunk_plan), so the pulls of group c+1 and the output pulls of group c-1 hide
      under the attention kernel of group c.
    transport "collective": two dist.all_to_all_single calls (any backend; used by the gloo CPU tests).
    attn_fn(q, k, v) -> o on [1,h,L,d] tensors defaults to ops.flash_attention_v1_ex(causal=causal).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if Q.dim() != 4 or Q.shape != K.shape or Q.shape != V.shape:
        raise ValueError("Q, K, V must be [B,H,Ls,d] shards with identical shapes")
    B, H, Ls, d = Q.shape
    BH = B * H
    if BH % world != 0:
        raise ValueError("B*H must be a multiple of the world size (heads are exchanged whole)")
    if transport not in ("auto", "peer", "collective"):
        raise ValueError("transport must be 'auto', 'peer' or 'collective'")
    if transport == "auto":
        transport = "peer" if Q.is_cuda else "collective"
    if attn_fn is None:
        from . import ops
        attn_fn = lambda q, k, v, out=None: ops.flash_attention_v1_ex(q, k, v, out, causal=causal)
    hpr = BH // world            # heads this rank computes: flat heads [rank*hpr, (rank+1)*hpr)
    L = Ls * world

    if transport == "collective" or world == 1:
        def to_heads(x):         # [B,H,Ls,d] -> my heads over the whole sequence [hpr, L, d]
            send = x.reshape(world, hpr, Ls, d).contiguous()
            recv = torch.empty_like(send)
            dist.all_to_all_single(recv, send, group=group)          # recv[p] = rank p's rows of my heads
            return recv.permute(1, 0, 2, 3).reshape(hpr, L, d).contiguous()
        q, k, v = to_heads(Q), to_heads(K), to_heads(V)
        o = attn_fn(q[None], k[None], v[None])[0]                    # [hpr, L, d]
        send = o.reshape(hpr, world, Ls, d).permute(1, 0, 2, 3).contiguous()   # [p] = my heads' rows owned by rank p
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)              # recv[p] = rank p's heads, my rows
        return recv.reshape(B, H, Ls, d)

    src, out_sym, h_src, h_out, streams, cache = _a2a_buffers(BH, hpr, Ls, L, d, Q.dtype, Q.device, group)
    main = torch.cuda.current_stream(Q.device)
    src[0].copy_(Q.reshape(BH, Ls, d))
    src[1].copy_(K.reshape(BH, Ls, d))
    src[2].copy_(V.reshape(BH, Ls, d))
    h_src.barrier(channel=0)     # every rank's rows are staged (and every rank has left the previous call)
    bounds = _a2a_chunk_plan(hpr, L, chunks, Q.device)
    n_chunks = len(bounds) - 1
    hc_max = max(bounds[c + 1] - bounds[c] for c in range(n_chunks))
    if "full" not in cache or cache["full"].shape[2] < hc_max:
        # recv: blocks as they arrive, one contiguous [hc, Ls, d] block per (tensor, peer); full: the same rows regrouped
        # per head over the whole sequence (what the kernel reads); ofull: the kernel's output before it is regrouped by
        # destination rank into the symmetric buffer
        cache["recv"] = torch.empty((2, 3, world, hc_max, Ls, d), dtype=Q.dtype, device=Q.device)
        cache["full"] = torch.empty((2, 3, hc_max, L, d), dtype=Q.dtype, device=Q.device)
        cache["ofull"] = torch.empty((2, hc_max, L, d), dtype=Q.dtype, device=Q.device)
    recv, full, ofull = cache["recv"], cache["full"], cache["ofull"]
    out_by_rank = out_sym.view(world, hpr, Ls, d)      # same bytes as [hpr, L, d]: [dest rank][head][row][d]
    # Every block address of the exchange, computed once per (buffers, plan) and handed to the library as pointer arrays:
    # one C call per (chunk, tensor) instead of one Python -> ctypes round trip per block.  All blocks are CONTIGUOUS on
    # both sides (plain cudaMemcpyAsync): strided 2-D copies did not run beside the persistent kernel (8 GPUs, C4: the
    # 1.05 ms of pulls stayed fully exposed), 1-D copies do; the regrouping is two small local copy kernels per chunk.
    import ctypes
    from . import _lib
    lib = _lib.load()
    es = Q.element_size()
    h0 = rank * hpr
    plan_key = (tuple(bounds), recv.data_ptr())
    if cache.get("plan_key") != plan_key:
        peer_ptr = [h_src.get_buffer(p, src.shape, src.dtype).data_ptr() for p in range(world)]
        out_ptr = [h_out.get_buffer(p, out_sym.shape, out_sym.dtype).data_ptr() for p in range(world)]
        order = [(rank + s) % world for s in range(world)]     # the local block first, then a different peer per rank
        arr = lambda xs: (ctypes.c_void_p * len(xs))(*xs)
        plans = []
        for c in range(n_chunks):
            hb, he = bounds[c], bounds[c + 1]
            ins = []
            for t in range(3):
                dst = [recv[c % 2, t, p].data_ptr() for p in order]                                        # [hc, Ls, d]
                srcs = [peer_ptr[p] + ((t * BH + h0 + hb) * Ls * d) * es for p in order]                   # [t, h0+hb:h0+he]
                ins.append((arr(dst), arr(srcs)))
            odst = [p * hpr + hb for p in order]                                                           # O[p*hpr+hb : p*hpr+he]
            osrc = [out_ptr[p] + ((rank * hpr + hb) * Ls * d) * es for p in order]                         # peer's [rank, hb:he]
            plans.append((ins, odst, arr(osrc), he - hb))
        cache["plan_key"], cache["plans"] = plan_key, plans
    plans = cache["plans"]
    O = torch.empty((BH, Ls, d), dtype=Q.dtype, device=Q.device)
    in_streams, out_stream = streams[:3], streams[3]
    row_bytes = Ls * d * es

    def pull(c):
        """Blocks of all ranks for head chunk c -> recv[c % 2]; Q, K, V on separate copy streams."""
        ins, _, _, nh = plans[c]
        events = []
        for t, st in enumerate(in_streams):
            st.wait_stream(main)             # the regrouping copy of chunk c-2, last reader of recv[c % 2], is already enqueued
            _lib.check(lib.fa_copy_multi_async(world, ins[t][0], ins[t][1], nh * row_bytes, st.cuda_stream))
            ev = torch.cuda.Event()
            ev.record(st)
            events.append(ev)
        return events

    pending = pull(0)
    for c in range(n_chunks):
        hb, he = bounds[c], bounds[c + 1]
        nh = he - hb
        for ev in pending:
            main.wait_event(ev)
        for t in range(3):       # [peer][head][Ls][d] -> [head][peer*Ls + row][d]
            full[c % 2, t, :nh].view(nh, world, Ls, d).copy_(recv[c % 2, t, :, :nh].permute(1, 0, 2, 3))
        if c + 1 < n_chunks:
            pending = pull(c + 1)            # travels while the kernel of chunk c runs
        q, k, v = (full[c % 2, t, :nh][None] for t in range(3))
        attn_fn(q, k, v, ofull[c % 2, :nh][None])
        out_by_rank[:, hb:he].copy_(ofull[c % 2, :nh].view(nh, world, Ls, d).permute(1, 0, 2, 3))
        # outputs of chunk c go home under the kernel of chunk c+1: the side stream waits for this rank's kernel, meets
        # the other ranks (a device-side barrier on that stream, so the main stream never blocks on a peer), then pulls
        # this rank's rows of every peer's chunk-c heads
        out_stream.wait_stream(main)
        with torch.cuda.stream(out_stream):
            h_out.barrier(channel=c)
            _, odst, osrc, _ = plans[c]
            o_base = O.data_ptr()
            dst = (ctypes.c_void_p * world)(*[o_base + h * row_bytes for h in odst])
            _lib.check(lib.fa_copy_multi_async(world, dst, osrc, nh * row_bytes, out_stream.cuda_stream))
    main.wait_stream(out_stream)
    h_out.barrier(channel=n_chunks)   # nobody restages or overwrites outputs while a peer may still be pulling them
    return O.reshape(B, H, Ls, d)


def _a2a_chunk_plan(hpr: int, L: int, max_chunks: int, device) -> list:
    sms = torch.cuda.get_device_properties(device).multi_processor_count if device.type == "cuda" else 1
    return list(_a2a_plan(hpr, -(-L // 256), max(1, min(max_chunks, 4)), sms))


import functools  # noqa: E402


@functools.lru_cache(maxsize=64)
def _a2a_plan(hpr: int, items_per_head: int, max_chunks: int, sms: int, pull_ratio: float = 0.5) -> tuple:
    """Head-chunk boundaries for the pipelined exchange, by a small time model (unit: one round of the persistent kernel).
    The attention kernel runs one CTA per SM over 256-row work items, so a chunk of h heads costs
    ceil(h * items_per_head / SMs) rounds (32 heads at L = 16384 on 148 SMs: 8+8+8+8 costs 16 rounds, 9+14+9 costs 15,
    one chunk 14).  Pulling a head's Q, K, V rows costs `pull_ratio` of its compute time (NVLink copy engines vs the
    tensor pipe; 0.5 is conservative), its output a third of that.  The pull of chunk c+1 and the output pull of chunk
    c-1 run under the kernel of chunk c; the first pull and the last output pull are exposed.  Minimises
        pull(c0) + sum_c max(rounds(c), pull(c+1)) + out_pull(c_last)."""
    head_rounds = items_per_head / sms
    rounds = lambda h: -(-h * items_per_head // sms)
    pull = lambda h: pull_ratio * h * head_rounds

    def cost(sizes):
        t = pull(sizes[0])
        for i, h in enumerate(sizes):
            t += max(rounds(h), pull(sizes[i + 1])) if i + 1 < len(sizes) else rounds(h)
        return t + pull(sizes[-1]) / 3.0

    best, best_cost = (hpr,), cost((hpr,))

    def search(prefix, left, k):
        nonlocal best, best_cost
        if k == 1 or left == 1:
            c = cost(prefix + (left,))
            if c < best_cost - 1e-9:
                best, best_cost = prefix + (left,), c
            return
        for h in range(1, left):
            search(prefix + (h,), left - h, k - 1)
        search(prefix, left, 1)

    if hpr > 1 and max_chunks > 1:
        search((), hpr, max_chunks)
    out = [0]
    for h in best:
        out.append(out[-1] + h)
    return tuple(out)
