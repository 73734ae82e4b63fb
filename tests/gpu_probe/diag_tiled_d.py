"""Dev diagnostic for the tiled-d kernel: error structure by column block / row block for a few shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops

def ref(Q, K, V):
    d = Q.shape[-1]
    S = (Q.double() @ K.double().transpose(-1, -2)) / d ** 0.5
    return torch.softmax(S, -1) @ V.double()

for d in (256, 512):
    for L in (128, 256, 384, 640):
        g = torch.Generator().manual_seed(1)
        Q, K, V = ((torch.rand((1, 1, L, d), generator=g) * 2 - 1).bfloat16().cuda() for _ in range(3))
        for name, (q, k, v) in {"rand": (Q, K, V), "uniformP": (torch.zeros_like(Q), K, V), "Vones": (Q, K, torch.ones_like(V))}.items():
            O = ops.flash_attention_v1_tiled_d(q, k, v, sync=True).double()
            E = (O - ref(q, k, v)).abs()[0, 0]
            colblk = [f"{E[:, c:c + 64].max().item():.1e}" for c in range(0, d, 64)]
            rowblk = [f"{E[r:r + 128].max().item():.1e}" for r in range(0, L, 128)]
            print(f"d={d} L={L} {name:9s} max={E.max().item():.2e} cols64={colblk} rows128={rowblk}", flush=True)
