"""Sequence-sharded (ring) attention timing: torchrun --nproc-per-node N tests/gpu_probe/ring_bench.py [B H L d]
Every rank owns L/N rows of Q, K, V (bf16).  Prints, from rank 0: the ring time (max over ranks, CUDA events), the
same partial kernels + combine without any communication (the overlap target), and max-abs error of sampled rows
against a float64 torch evaluation over the gathered K/V."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from exploring_flash_attention_b200 import ops  # noqa: E402
from exploring_flash_attention_b200.sharding import ring_attention  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
B, H, L, d = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (8, 32, 16384, 128)
Ls = L // world
g = torch.Generator(device="cuda").manual_seed(1234 + rank)
Q, K, V = ((torch.rand((B, H, Ls, d), generator=g, device="cuda") * 2 - 1).bfloat16() for _ in range(3))


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


o_parts = torch.empty((world, B * H, Ls, d), dtype=torch.float32, device="cuda")
lse_parts = torch.empty((world, B * H, Ls), dtype=torch.float32, device="cuda")


def no_comm():
    for s in range(world):
        ops.flash_attention_partial(Q, K, V, o_parts[s], lse_parts[s])
    return ops.flash_attention_v2_combine(o_parts, lse_parts, Q.dtype, (B, H, Ls, d))


transport = os.environ.get("RING_TRANSPORT", "peer")
causal = os.environ.get("RING_CAUSAL", "0") == "1"     # zig-zag layout: the local shard is chunks r and 2N-1-r
ms_ring = timed(lambda: ring_attention(Q, K, V, transport=transport, causal=causal))
ms_nocomm = timed(no_comm)
O = ring_attention(Q, K, V, transport=transport, causal=causal)
# check a few rows of head 0 against float64 over the gathered keys
Kall = torch.empty((world,) + tuple(K.shape), dtype=K.dtype, device="cuda")
Vall = torch.empty_like(Kall)
dist.all_gather_into_tensor(Kall, K)
dist.all_gather_into_tensor(Vall, V)
if causal:
    from exploring_flash_attention_b200.sharding import zigzag_unshard
    kf = zigzag_unshard([Kall[r][0, 0] for r in range(world)]).double()
    vf = zigzag_unshard([Vall[r][0, 0] for r in range(world)]).double()
    rows = slice(Ls - min(Ls // 2, 256), Ls)              # late rows of the late chunk: global rows (2N-1-r)*C + ...
    C = Ls // 2
    grow = (2 * world - 1 - rank) * C + torch.arange(C - (rows.stop - rows.start), C, device="cuda")
    sc = Q[0, 0, rows].double() @ kf.T / d ** 0.5
    sc = sc.masked_fill(torch.arange(L, device="cuda")[None, :] > grow[:, None], float("-inf"))
    ref = torch.softmax(sc, dim=-1) @ vf
else:
    kf = torch.cat([Kall[r][0, 0] for r in range(world)]).double()
    vf = torch.cat([Vall[r][0, 0] for r in range(world)]).double()
    rows = slice(0, min(Ls, 256))
    ref = torch.softmax(Q[0, 0, rows].double() @ kf.T / d ** 0.5, dim=-1) @ vf
err = torch.tensor([(O[0, 0, rows].double() - ref).abs().max().item()], device="cuda")
dist.all_reduce(err, op=dist.ReduceOp.MAX)
if rank == 0:
    flops = 4.0 * B * H * L * L * d * (0.5 if causal else 1.0)
    print(json.dumps({"ring_attention": {"B": B, "H": H, "L": L, "d": d, "n_gpus": world, "transport": transport, "causal": causal, "ms": ms_ring,
                                         "tflops_total": flops / ms_ring / 1e9, "ms_same_kernels_no_comm": ms_nocomm,
                                         "kv_bytes_per_hop": 2 * B * H * Ls * d * 2, "max_abs_err": err.item()}}), flush=True)
dist.barrier()
dist.destroy_process_group()
