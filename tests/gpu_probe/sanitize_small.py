"""Tiny workload for compute-sanitizer (one tool per gpurun call): every kernel family once, ragged shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops

g = torch.Generator().manual_seed(0)
mk = lambda B, H, L, d, dt: tuple((torch.rand((B, H, L, d), generator=g) * 2 - 1).to(dt).cuda() for _ in range(3))
for (B, H, L, d, dt) in ((1, 3, 300, 128, torch.bfloat16), (1, 2, 200, 32, torch.float32), (1, 2, 130, 32, torch.float16),
                         (1, 150, 128, 64, torch.bfloat16)):
    Q, K, V = mk(B, H, L, d, dt)
    O = ops.flash_attention_v1(Q, K, V, sync=True)
    O2 = ops.flash_attention_v2(Q, K, V, 96, sync=True)
    print("ok", B, H, L, d, dt, float((O.float() - O2.float()).abs().max()))
Q, K, V = mk(1, 2, 200, 256, torch.bfloat16)
print("ok tiled-d", float(ops.flash_attention_v1_tiled_d(Q, K, V, sync=True).float().abs().max()))
