"""Where the head-exchange path spends its time (run under torchrun): staging, input pulls alone, kernels alone, output
pulls alone, and the whole call, max over ranks.  python -m torch.distributed.run --nproc-per-node N tests/gpu_probe/a2a_breakdown.py"""
import ctypes
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
import torch.distributed as dist
from exploring_flash_attention_b200 import _lib, ops, sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
B, H, L, d = 8, 32, 16384, 128
Ls, BH = L // world, B * H
hpr = BH // world
q, k, v = ((torch.rand((B, H, Ls, d), device="cuda") * 2 - 1).bfloat16() for _ in range(3))


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


res = {"world": world, "whole_call_ms": timed(lambda: sharding.alltoall_attention(q, k, v))}
src, out_sym, h_src, h_out, streams, cache = sharding._a2a_buffers(BH, hpr, Ls, L, d, q.dtype, q.device, None)
res["staging_ms"] = timed(lambda: (src[0].copy_(q.reshape(BH, Ls, d)), src[1].copy_(k.reshape(BH, Ls, d)), src[2].copy_(v.reshape(BH, Ls, d))))
res["barrier_ms"] = timed(lambda: h_src.barrier(channel=0))
lib = _lib.load()
es = 2
full = cache["full"]
peers = [h_src.get_buffer(p, src.shape, src.dtype).data_ptr() for p in range(world)]
row_bytes = Ls * d * es
main = torch.cuda.current_stream()


def pull_all(n_streams):
    """All of this rank's input blocks (hpr heads x 3 tensors x world peers) with the copies spread over n_streams."""
    sts = [torch.cuda.Stream() for _ in range(n_streams)] if not hasattr(pull_all, "s") or len(pull_all.s) != n_streams else pull_all.s
    pull_all.s = sts
    big = cache.setdefault("probe_full", torch.empty((3, hpr, L, d), dtype=q.dtype, device=q.device))
    i = 0
    for t in range(3):
        for p in range(world):
            st = sts[i % n_streams]; i += 1
            st.wait_stream(main)
            dst = big[t].data_ptr() + p * Ls * d * es
            s_ = peers[p] + ((t * BH + rank * hpr) * Ls * d) * es
            _lib.check(lib.fa_copy_2d_async(dst, L * d * es, s_, row_bytes, row_bytes, hpr, st.cuda_stream))
    for st in sts:
        main.wait_stream(st)


pulled_mb = 3 * hpr * (world - 1) * Ls * d * es / 1e6
for ns in (1, 3, 6):
    ms = timed(lambda: pull_all(ns))
    res[f"input_pull_{ns}_streams_ms"] = ms
    res[f"input_pull_{ns}_streams_GBps_remote"] = round(pulled_mb / ms, 1)
big = cache["probe_full"]
o = torch.empty((1, hpr, L, d), dtype=q.dtype, device="cuda")
res["kernel_all_heads_one_launch_ms"] = timed(lambda: ops.flash_attention_v1(big[0][None], big[1][None], big[2][None], o))
bounds = sharding._a2a_chunk_plan(hpr, L, 4, q.device)
res["chunk_plan"] = list(bounds)
def chunked():
    for a, b in zip(bounds, bounds[1:]):
        ops.flash_attention_v1(big[0][None, a:b], big[1][None, a:b], big[2][None, a:b], o[:, a:b])
res["kernel_chunked_ms"] = timed(chunked)
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
