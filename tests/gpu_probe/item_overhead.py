"""Dev probe: per-item fixed overhead of the persistent K1 kernel.  1184 items (8 per SM) at several L; fits
t_CTA = 8 * (a + b * L/128)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import torch
from exploring_flash_attention_b200 import ops

res = []
for L, BH in ((512, 592), (1024, 296), (2048, 148), (4096, 74), (8192, 37)):
    q, k, v = (torch.randn((1, BH, L, 128), device="cuda", dtype=torch.bfloat16) for _ in range(3))
    o = torch.empty_like(q)
    for _ in range(3):
        ops.flash_attention_v1(q, k, v, o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.flash_attention_v1(q, k, v, o)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = 4.0 * BH * L * L * 128 / (ms * 1e-3) / 1e12
    res.append((L // 128, ms * 1e3 / 8))
    print(f"L={L} BH={BH} items={BH * L // 256} {ms * 1e3:.1f} us  {tf:.0f} TF  per-item {ms * 1e3 / 8:.2f} us", flush=True)
x = np.array([r[0] for r in res], float)
y = np.array([r[1] for r in res], float)
b, a = np.polyfit(x, y, 1)
print(f"fit: per-item = {a:.2f} us fixed + {b:.3f} us per KV tile (pair of Q tiles)")
