"""Dev probe: C4-shaped runs (L=16384, d=128, bf16) at growing head counts."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops
L, d = 16384, 128
for BH in (16, 37, 74, 148, 256):
    q, k, v = (torch.randn((1, BH, L, d), device="cuda", dtype=torch.bfloat16) for _ in range(3))
    o = torch.empty_like(q)
    for _ in range(2):
        ops.flash_attention_v1(q, k, v, o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        ops.flash_attention_v1(q, k, v, o)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"BH={BH} items={BH * 64} {ms:.3f} ms {4.0 * BH * L * L * d / (ms * 1e-3) / 1e12:.0f} TF", flush=True)
    del q, k, v, o
