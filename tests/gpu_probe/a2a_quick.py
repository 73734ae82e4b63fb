"""Lean N-GPU timing of the two sequence-parallel paths on the C4 shape (run under torchrun)."""
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
import torch.distributed as dist
from exploring_flash_attention_b200 import sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
B, H, L, d = 8, 32, 16384, 128
q, k, v = ((torch.rand((B, H, L // world, d), device="cuda") * 2 - 1).bfloat16() for _ in range(3))


def timed(fn, n=3):
    for _ in range(2):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 3)


res = {"world": world,
       "alltoall_dense_ms": timed(lambda: sharding.alltoall_attention(q, k, v)),
       "alltoall_causal_ms": timed(lambda: sharding.alltoall_attention(q, k, v, causal=True)),
       "alltoall_dense_1chunk_ms": timed(lambda: sharding.alltoall_attention(q, k, v, chunks=1)),
       "ring_dense_ms": timed(lambda: sharding.ring_attention(q, k, v, transport="peer"))}
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
