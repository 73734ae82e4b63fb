"""How close is the host-buffer entry point (fa_forward_host) to the PCIe floor of this box?  Times, for the C2 shape:
H2D of Q,K,V alone, D2H of O alone, both directions at once on two streams, and the pipelined library call."""
import json
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops

B, H, L, d = 32, 8, 1024, 128
host = [torch.empty((B, H, L, d), dtype=torch.bfloat16).uniform_(-1, 1).pin_memory() for _ in range(4)]
dev = [torch.empty((B, H, L, d), dtype=torch.bfloat16, device="cuda") for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def wall(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


def h2d():
    for i in range(3):
        dev[i].copy_(host[i], non_blocking=True)


def d2h():
    host[3].copy_(dev[3], non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        h2d()
    with torch.cuda.stream(s2):
        d2h()


nbytes = host[0].numel() * 2
res = {"shape": [B, H, L, d], "h2d_3_tensors_ms": wall(h2d), "d2h_1_tensor_ms": wall(d2h), "both_directions_ms": wall(both),
       "fa_forward_host_ms": wall(lambda: ops.flash_attention_host(host[0], host[1], host[2], host[3]))}
res["h2d_GBps"] = 3 * nbytes / res["h2d_3_tensors_ms"] / 1e6
res["d2h_GBps"] = nbytes / res["d2h_1_tensor_ms"] / 1e6
print(json.dumps(res))
