"""CPU, world_size 2, gloo: the N>1 host logic — (b,h) sharding and the verification gather — reproduces the
unsharded result.  The per-rank compute is stood in for by the oracle (the CUDA kernel cannot run here)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, BH, L, d, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from exploring_flash_attention_b200.sharding import gather_heads, head_range, shard_heads
    from oracle import reference
    g = torch.Generator().manual_seed(42)
    Q, K, V = (torch.rand((1, BH, L, d), generator=g) * 2 - 1 for _ in range(3))     # same tensors on every rank
    qs, ks, vs = (shard_heads(x, rank, world) for x in (Q, K, V))
    b, e = head_range(BH, rank, world)
    assert qs.shape == (1, e - b, L, d)
    local = torch.from_numpy(reference.naive_attention_batched_f64(qs.numpy(), ks.numpy(), vs.numpy())).float()
    full = gather_heads(local.unsqueeze(0), BH)
    dist.barrier()
    if rank == 0:
        q.put(full.numpy())
    dist.destroy_process_group()


def test_two_rank_head_sharding_and_gather():
    BH, L, d, world = 5, 24, 16, 2          # 5 heads over 2 ranks: uneven shards (3 + 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, BH, L, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, str(ROOT))
    from oracle import reference
    g = torch.Generator().manual_seed(42)
    Q, K, V = (torch.rand((1, BH, L, d), generator=g) * 2 - 1 for _ in range(3))
    ref = reference.naive_attention_batched_f64(Q.numpy(), K.numpy(), V.numpy())
    assert full.shape == (BH, L, d)
    np.testing.assert_allclose(full, ref, atol=1e-6)


def _ring_worker(rank, world, port, B, H, L, d, causal, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from exploring_flash_attention_b200.sharding import ring_attention, zigzag_shard
    from oracle import reference
    g = torch.Generator().manual_seed(7)
    Q, K, V = (torch.rand((B, H, L, d), generator=g) * 2 - 1 for _ in range(3))     # same tensors on every rank
    Ls = L // world
    if causal:
        qs, ks, vs = (zigzag_shard(x, rank, world).contiguous() for x in (Q, K, V))
    else:
        qs, ks, vs = (x[:, :, rank * Ls:(rank + 1) * Ls].contiguous() for x in (Q, K, V))

    def partial_fn(q_, k_, v_, o_out, lse_out, causal=False):      # oracle stand-in for the CUDA partial kernel
        nq, nk = q_.shape[2], k_.shape[2]
        for i in range(B * H):
            o, lse = reference.naive_attention_ex_f64(q_.reshape(-1, nq, d)[i].numpy(), k_.reshape(-1, nk, d)[i].numpy(),
                                                      v_.reshape(-1, nk, d)[i].numpy(), causal=causal)
            o_out[i] = torch.from_numpy(o).float()
            lse_out[i] = torch.from_numpy(lse).float()

    def combine_fn(o_parts, lse_parts, dtype, shape):
        return torch.from_numpy(reference.merge_partials_f64(o_parts.numpy(), lse_parts.numpy())).to(dtype).reshape(shape)

    local = ring_attention(qs, ks, vs, partial_fn=partial_fn, combine_fn=combine_fn, causal=causal)
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local)
    if rank == 0:
        q.put(torch.stack(out).numpy())
    dist.barrier()
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("causal", [False, True])
def test_ring_attention_schedule_over_gloo(causal):
    """Sequence sharded over 3 ranks: every K/V shard must visit every rank exactly once and the merged partials
    must equal unsharded attention (contiguous shards; zig-zag shards with the chunk-level causal rules)."""
    B, H, L, d, world = 1, 2, 36, 8, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, B, H, L, d, causal, q)) for r in range(world)]
    for p in procs:
        p.start()
    shards = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, str(ROOT))
    from exploring_flash_attention_b200.sharding import zigzag_shard, zigzag_unshard
    from oracle import reference
    g = torch.Generator().manual_seed(7)
    Q, K, V = (torch.rand((B, H, L, d), generator=g) * 2 - 1 for _ in range(3))
    if causal:
        full = zigzag_unshard([torch.from_numpy(x) for x in shards]).numpy()
        ref = np.stack([reference.naive_attention_ex_f64(Q[0, h].numpy(), K[0, h].numpy(), V[0, h].numpy(), causal=True)[0]
                        for h in range(H)])
        assert torch.equal(zigzag_unshard([zigzag_shard(Q, r, world) for r in range(world)]), Q)
    else:
        full = np.concatenate(list(shards), axis=2)
        ref = reference.naive_attention_batched_f64(Q.numpy(), K.numpy(), V.numpy())
    np.testing.assert_allclose(full.reshape(B * H, L, d), ref, atol=1e-5)


def _a2a_worker(rank, world, port, B, H, L, d, causal, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from exploring_flash_attention_b200.sharding import alltoall_attention
    from oracle import reference
    g = torch.Generator().manual_seed(9)
    Q, K, V = (torch.rand((B, H, L, d), generator=g) * 2 - 1 for _ in range(3))     # same tensors on every rank
    Ls = L // world
    qs, ks, vs = (x[:, :, rank * Ls:(rank + 1) * Ls].contiguous() for x in (Q, K, V))
    seen = []

    def attn_fn(q_, k_, v_, out=None):           # oracle stand-in for the CUDA kernel: [1,h,L,d] -> [1,h,L,d]
        seen.append(tuple(q_.shape))
        o = np.stack([reference.naive_attention_ex_f64(q_[0, i].numpy(), k_[0, i].numpy(), v_[0, i].numpy(), causal=causal)[0]
                      for i in range(q_.shape[1])])
        return torch.from_numpy(o).float()[None]

    local = alltoall_attention(qs, ks, vs, causal=causal, attn_fn=attn_fn)
    assert seen == [(1, B * H // world, L, d)]    # this rank computed its share of the heads over the WHOLE sequence
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local)
    if rank == 0:
        q.put(torch.cat(out, dim=2).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("causal", [False, True])
def test_alltoall_attention_over_gloo(causal):
    """Sequence sharded over 3 ranks, heads exchanged by all-to-all: every rank sees whole sequences for a third of the
    heads and the reassembled rows equal unsharded attention."""
    B, H, L, d, world = 2, 3, 30, 8, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_a2a_worker, args=(r, world, port, B, H, L, d, causal, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, str(ROOT))
    from oracle import reference
    g = torch.Generator().manual_seed(9)
    Q, K, V = (torch.rand((B, H, L, d), generator=g) * 2 - 1 for _ in range(3))
    ref = np.stack([reference.naive_attention_ex_f64(Q.reshape(-1, L, d)[h].numpy(), K.reshape(-1, L, d)[h].numpy(),
                                                     V.reshape(-1, L, d)[h].numpy(), causal=causal)[0] for h in range(B * H)])
    np.testing.assert_allclose(full.reshape(B * H, L, d), ref, atol=1e-5)
