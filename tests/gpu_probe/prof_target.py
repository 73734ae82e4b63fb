"""Small, fixed launch sequences for ncu: `python tests/gpu_probe/prof_target.py c1|c2|c3|c4|c5|bwd [launches]`.
Each runs `launches` (default 4) back-to-back calls of the library on the BASELINE.json config of that name
(c4: a B2 H32 slice of it, same per-head work) and exits; profile with -s 2 -c 1 to skip the warm launches."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
shape, dtype = {"c1": ((32, 8, 1024, 32), torch.float32), "c2": ((32, 8, 1024, 128), torch.bfloat16),
                "c3": ((32, 8, 256, 64), torch.bfloat16), "c4": ((2, 32, 16384, 128), torch.bfloat16),
                "c5": ((16, 8, 4096, 512), torch.bfloat16), "bwd": ((4, 16, 4096, 128), torch.bfloat16)}[which]
g = torch.Generator(device="cuda").manual_seed(42)
q, k, v = ((torch.rand(shape, generator=g, device="cuda") * 2 - 1).to(dtype) for _ in range(3))
o = torch.empty_like(q)
ws = ops.v2_workspace(*shape, 64, q.device) if which == "c3" else None
if which == "bwd":
    do = (torch.rand(shape, generator=g, device="cuda") * 2 - 1).to(dtype)
    o, lse = ops.flash_attention_v1_ex(q, k, v, return_lse=True)
    bws = torch.empty(ops.backward_workspace_bytes(*shape[:3]), dtype=torch.uint8, device="cuda")
for _ in range(n):
    if which == "bwd":
        ops.flash_attention_backward(q, k, v, o, do, lse, workspace=bws)
    elif which == "c3":
        ops.flash_attention_v2(q, k, v, 64, O=o, workspace=ws)
    else:
        ops.flash_attention_v1(q, k, v, o)
torch.cuda.synchronize()
print("ok", which, float(o.float().abs().mean()))
