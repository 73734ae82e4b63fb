"""GPU, 2+ devices (skipped on a 1-GPU box): heads sharded over ranks, each rank runs the CUDA path on its slice, the
slices are all-gathered over NCCL and the assembled tensor is checked against the oracle."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from exploring_flash_attention_b200 import ops
    from exploring_flash_attention_b200.sharding import gather_heads, shard_heads
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    B, H, L, d = 1, 5, 640, 128                       # 5 heads over 2 ranks: uneven shards
    g = torch.Generator().manual_seed(42)
    Q, K, V = ((torch.rand((B, H, L, d), generator=g) * 2 - 1).bfloat16() for _ in range(3))
    qs, ks, vs = (shard_heads(x, rank, world).contiguous().cuda() for x in (Q, K, V))
    local = ops.flash_attention_v1(qs, ks, vs, sync=True)
    full = gather_heads(local, B * H)
    dist.barrier()
    if rank == 0:
        q.put((full.float().cpu().numpy(), Q.float().numpy(), K.float().numpy(), V.float().numpy()))
    dist.destroy_process_group()


def test_two_gpu_head_sharding_nccl_gather():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import reference
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, Q, K, V = _get_or_fail(q, procs, 240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = reference.naive_attention_batched_f64(Q, K, V)
    assert full.shape == ref.shape
    assert np.abs(full - ref).max() <= 2e-3


def test_tensors_on_a_non_current_device():
    """The C ABI launches on the current device; ops must switch to the tensors' device (cuda:1 while cuda:0 is current)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from exploring_flash_attention_b200 import ops
    from oracle import reference
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(5)
    Q, K, V = ((torch.rand((1, 2, 300, 64), generator=g) * 2 - 1).bfloat16().to("cuda:1") for _ in range(3))
    O = ops.flash_attention_v1(Q, K, V, sync=True)
    O2 = ops.flash_attention_v2(Q, K, V, 128, sync=True)
    torch.cuda.synchronize(1)
    assert O.device.index == 1 and torch.cuda.current_device() == 0
    ref = reference.naive_attention_batched_f64(*(x.float().cpu().numpy() for x in (Q, K, V)))
    assert np.abs(O.float().cpu().numpy().reshape(2, 300, 64) - ref).max() <= 2e-3
    assert np.abs(O2.float().cpu().numpy().reshape(2, 300, 64) - ref).max() <= 2e-3


def _get_or_fail(q, procs, timeout):
    """q.get() that gives up as soon as a worker has died with an error (instead of sitting out the whole timeout while
    the surviving rank hangs in a collective, which on a GPU box is minutes of billed time)."""
    import queue
    import time
    t_end = time.time() + timeout
    while time.time() < t_end:
        try:
            return q.get(timeout=2)
        except queue.Empty:
            dead = [p for p in procs if p.exitcode not in (None, 0)]
            if dead:
                for p in procs:
                    if p.is_alive():
                        p.kill()
                pytest.fail(f"worker exited with code {dead[0].exitcode} before producing a result")
    for p in procs:
        if p.is_alive():
            p.kill()
    pytest.fail("workers timed out")


def _ring_worker(rank, world, port, q, transport="nccl", causal=False):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from exploring_flash_attention_b200.sharding import ring_attention, zigzag_shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    B, H, L, d = 1, 3, 1024, 128
    g = torch.Generator().manual_seed(77)
    Q, K, V = ((torch.rand((B, H, L, d), generator=g) * 2 - 1).bfloat16() for _ in range(3))
    Ls = L // world
    if causal:
        qs, ks, vs = (zigzag_shard(x, rank, world).contiguous().cuda() for x in (Q, K, V))
    else:
        qs, ks, vs = (x[:, :, rank * Ls:(rank + 1) * Ls].contiguous().cuda() for x in (Q, K, V))
    local = ring_attention(qs, ks, vs, transport=transport, causal=causal)
    if transport == "peer":   # second call reuses the cached symmetric buffers (restaging must not race the last pull)
        assert torch.equal(local, ring_attention(qs, ks, vs, transport=transport, causal=causal))
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    torch.cuda.synchronize()
    if rank == 0:
        q.put((out.float().cpu().numpy(), Q.float().numpy(), K.float().numpy(), V.float().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport,causal", [("nccl", False), ("peer", False), ("peer", True), ("nccl", True)])
def test_two_gpu_ring_attention(transport, causal):
    """Sequence-sharded (context-parallel) attention: K/V shards travel round the ring (NCCL send/recv, or copy-engine
    pulls out of the neighbour's symmetric-memory buffer), partials merged per rank; causal uses the zig-zag layout."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from exploring_flash_attention_b200.sharding import zigzag_unshard
    from oracle import reference
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ring_worker, args=(r, 2, port, q, transport, causal)) for r in range(2)]
    for p in procs:
        p.start()
    shards, Q, K, V = _get_or_fail(q, procs, 240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    if causal:
        full = zigzag_unshard([torch.from_numpy(x) for x in shards]).numpy()
        ref = np.stack([reference.naive_attention_ex_f64(Q[0, h], K[0, h], V[0, h], causal=True)[0] for h in range(Q.shape[1])])
        tol = 2e-3 * max(1.0, 2 * np.abs(ref).max())      # the first causal rows are O(1)
    else:
        full = np.concatenate(list(shards), axis=2)
        ref = reference.naive_attention_batched_f64(Q, K, V)
        tol = 2e-3
    assert np.abs(full.reshape(ref.shape) - ref).max() <= tol


def _a2a_worker(rank, world, port, q, transport, causal):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from exploring_flash_attention_b200.sharding import alltoall_attention
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    B, H, L, d = 2, 5, 1024, 128            # 10 heads over 2 ranks, 5 each, in 4 pipelined chunks (2+1+1+1)
    g = torch.Generator().manual_seed(78)
    Q, K, V = ((torch.rand((B, H, L, d), generator=g) * 2 - 1).bfloat16() for _ in range(3))
    Ls = L // world
    qs, ks, vs = (x[:, :, rank * Ls:(rank + 1) * Ls].contiguous().cuda() for x in (Q, K, V))
    local = alltoall_attention(qs, ks, vs, transport=transport, causal=causal)
    assert torch.equal(local, alltoall_attention(qs, ks, vs, transport=transport, causal=causal))   # buffers are reused
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    torch.cuda.synchronize()
    if rank == 0:
        q.put((out.float().cpu().numpy(), Q.float().numpy(), K.float().numpy(), V.float().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport,causal", [("peer", False), ("peer", True), ("collective", False)])
def test_two_gpu_alltoall_attention(transport, causal):
    """Sequence-sharded attention by head exchange: copy-engine pulls out of the peers' symmetric memory (or NCCL
    all-to-all), the plain fused-tile kernel over whole sequences for this rank's heads, outputs pulled back."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import reference
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_a2a_worker, args=(r, 2, port, q, transport, causal)) for r in range(2)]
    for p in procs:
        p.start()
    shards, Q, K, V = _get_or_fail(q, procs, 240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full = np.concatenate(list(shards), axis=2)            # [B,H,L,d] from the per-rank row blocks
    B, H, L, d = Q.shape
    ref = np.stack([reference.naive_attention_ex_f64(Q.reshape(-1, L, d)[h], K.reshape(-1, L, d)[h], V.reshape(-1, L, d)[h],
                                                     causal=causal)[0] for h in range(B * H)])
    tol = 2e-3 * (max(1.0, 2 * np.abs(ref).max()) if causal else 1.0)
    assert np.abs(full.reshape(ref.shape) - ref).max() <= tol
