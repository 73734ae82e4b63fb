"""ncu target: a few split-KV launches on a steady-state shape: python tests/gpu_probe/prof_splitkv.py [B H L d dtype kvs]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops
a = sys.argv[1:]
B, H, L, d = (int(x) for x in a[:4]) if len(a) >= 4 else (32, 8, 1024, 64)
dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[a[4] if len(a) > 4 else "bf16"]
kvs = int(a[5]) if len(a) > 5 else 64
q, k, v = ((torch.rand((B, H, L, d), device="cuda") * 2 - 1).to(dtype) for _ in range(3))
ws = ops.v2_workspace(B, H, L, d, kvs, q.device)
for _ in range(4):
    ops.flash_attention_v2_splitkv(q, k, v, kvs, *ws)
torch.cuda.synchronize()
print("ok")
