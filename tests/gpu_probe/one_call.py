"""One library call for a profiler to capture: python tests/gpu_probe/one_call.py B H L d [bf16|f16|f32]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops

B, H, L, d = (int(x) for x in sys.argv[1:5])
dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[sys.argv[5] if len(sys.argv) > 5 else "bf16"]
g = torch.Generator().manual_seed(0)
Q, K, V = ((torch.rand((B, H, L, d), generator=g) * 2 - 1).to(dtype).cuda() for _ in range(3))
for _ in range(2):
    O = ops.flash_attention_v1(Q, K, V, sync=True)
print("ok", float(O.float().abs().mean()))
