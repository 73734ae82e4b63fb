"""Backward throughput (graph-free, long kernels): python tests/gpu_probe/bwd_bench.py"""
import json
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops

def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n

for B, H, L, d, causal in ((32, 8, 1024, 128, False), (4, 16, 4096, 128, False), (2, 16, 16384, 128, False), (4, 16, 4096, 128, True),
                           (32, 8, 1024, 64, False), (4, 16, 4096, 64, False)):
    q, k, v, do = ((torch.rand((B, H, L, d), device="cuda") * 2 - 1).bfloat16() for _ in range(4))
    o, lse = ops.flash_attention_v1_ex(q, k, v, causal=causal, return_lse=True)
    ws = torch.empty(ops.backward_workspace_bytes(B, H, L), dtype=torch.uint8, device="cuda")
    ms_f = timed(lambda: ops.flash_attention_v1_ex(q, k, v, o, causal=causal))
    ms_b = timed(lambda: ops.flash_attention_backward(q, k, v, o, do, lse, causal=causal, workspace=ws))
    fl = 4.0 * B * H * L * L * d * (0.5 if causal else 1.0)
    print(json.dumps({"shape": [B, H, L, d], "causal": causal, "fwd_ms": round(ms_f, 4), "fwd_tflops": round(fl / ms_f / 1e9, 1),
                      "bwd_ms": round(ms_b, 4), "bwd_tflops_2.5x_fwd_flops": round(2.5 * fl / ms_b / 1e9, 1)}), flush=True)
