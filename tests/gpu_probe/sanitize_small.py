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
    O = ops.flash_attention_v1(Q, K, V, sync=True)                       # K1
    O2 = ops.flash_attention_v2(Q, K, V, 96, sync=True)                  # K3a' (single-tile splits) + K3b
    O3 = ops.flash_attention_v2(Q, K, V, 160, sync=True)                 # K1 SPLIT (longer splits) + K3b
    print("ok", B, H, L, d, dt, float((O.float() - O2.float()).abs().max()), float((O.float() - O3.float()).abs().max()))
Q, K, V = mk(1, 2, 200, 256, torch.bfloat16)
print("ok tiled-d", float(ops.flash_attention_v1_tiled_d(Q, K, V, sync=True).float().abs().max()))              # K2
print("ok tiled-d causal+lse", float(ops.flash_attention_v1_ex(Q, K, V, causal=True, return_lse=True, sync=True)[1].abs().max()))
print("ok tiled-d v2", float(ops.flash_attention_v2(Q, K, V, 64, sync=True).float().abs().max()))              # K2 split mode
Q, K, V = mk(1, 2, 200, 512, torch.bfloat16)
print("ok tiled-d pair", float(ops.flash_attention_v1_tiled_d(Q, K, V, sync=True).float().abs().max()))         # K2P
for d in (64, 128):                                                                                            # K4
    Q, K, V = mk(1, 2, 300, d, torch.bfloat16)
    dO = torch.ones_like(Q)
    for causal in (False, True):
        O, lse = ops.flash_attention_v1_ex(Q, K, V, causal=causal, return_lse=True, sync=True)
        grads = ops.flash_attention_backward(Q, K, V, O, dO, lse, causal=causal, sync=True)
        print("ok backward", d, causal, [float(x.float().abs().max()) for x in grads])
q = torch.randn((2, 100, 40), device="cuda")
print("ok naive", float(ops.naive_attention_reference(q, q, q).abs().max()))                                   # K0
