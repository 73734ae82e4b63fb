"""Fixed cost vs per-unit cost of the split-KV tile kernel: time at L=256 d=64 bf16 kvs=64 (8 units per head) for head
counts giving 1, 2, 4, 8, 16, 32, 64 units per SM; also the combine kernel and an empty-ish launch for scale."""
import json
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops

def timed(fn, n=100):
    """Device time per call: n calls in one CUDA graph (no per-call host launch cost), best of 5 replays."""
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return round(best, 2)

L, d, kvs = 256, 64, 64
out = {}
x = torch.zeros(1, device="cuda")
out["tiny_torch_kernel_us"] = timed(lambda: x.add_(1))
for heads in (19, 37, 74, 148, 296, 592, 1184):
    q, k, v = ((torch.rand((1, heads, L, d), device="cuda") * 2 - 1).bfloat16() for _ in range(3))
    ws = ops.v2_workspace(1, heads, L, d, kvs, q.device)
    o = torch.empty_like(q)
    out[f"heads{heads}_units_per_sm_{heads * 8 / 148:.1f}"] = {
        "splitkv_us": timed(lambda: ops.flash_attention_v2_splitkv(q, k, v, kvs, *ws)),
        "combine_us": timed(lambda: ops.flash_attention_v2_combine(ws[0], ws[1], torch.bfloat16, (1, heads, L, d), o)),
        "v1_us": timed(lambda: ops.flash_attention_v1(q, k, v, o))}
print(json.dumps(out, indent=1))
