"""Dev probe: SM-clock timeline of the fused-tile kernel's CTA 0 (needs a -DFA_TRACE build selected through FA_B200_LIB).

    FA_B200_LIB=build_variants/trace.so python tests/gpu_probe/trace_k1.py [B H L d] [out.npy]

Roles: 0 / 1 = row 0 of softmax warpgroup 0 / 1, 2 = the MMA-issuing thread.  Slots per KV tile:
  softmax: 0 s_full seen, 1 S in registers, 2 row max, 3 first-half exp done, 4 first-half P published,
           5 second-half exp done, 6 second-half P published
  MMA:     4i+0 p_full[i][0] seen, 4i+1 first-half PV issued, 4i+2 p_full[i][1] seen, 4i+3 QK_i(j+1) issued + committed
"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import torch
from exploring_flash_attention_b200 import _lib, ops

args = sys.argv[1:]
B, H, L, d = (int(x) for x in args[:4]) if len(args) >= 4 else (1, 16, 16384, 128)
out = args[4] if len(args) > 4 else "gpurun_out/trace_k1.npy"
lib = _lib.load()
set_trace = lib.fa_debug_set_trace
set_trace.argtypes = [ctypes.c_void_p]
set_trace.restype = None
q, k, v = (torch.randn((B, H, L, d), device="cuda", dtype=torch.bfloat16) for _ in range(3))
o = torch.empty_like(q)
for _ in range(2):
    ops.flash_attention_v1(q, k, v, o)
torch.cuda.synchronize()
buf = torch.zeros((3, 256, 8), device="cuda", dtype=torch.int64)
set_trace(buf.data_ptr())
ops.flash_attention_v1(q, k, v, o)
torch.cuda.synchronize()
set_trace(None)
t = buf.cpu().numpy().astype(np.float64)
np.save(out, t)
lo, hi = 24, 120   # steady-state tiles of the first item (128 KV tiles at L=16384)
hi = min(hi, L // 128 - 4)
sm = ["s_full->ld", "ld->max", "max->exp1", "exp1->pub1", "pub1->exp2", "exp2->pub2", "pub2->next s_full"]
for r in (0, 1):
    x = t[r]
    dl = [np.mean(x[lo:hi, s + 1] - x[lo:hi, s]) for s in range(6)]
    dl.append(np.mean(x[lo + 1:hi + 1, 0] - x[lo:hi, 6]))
    per = np.mean(x[lo + 1:hi + 1, 0] - x[lo:hi, 0])
    print(f"softmax{r}: period {per:.0f} cyc | " + " | ".join(f"{n} {v_:.0f}" for n, v_ in zip(sm, dl)))
m = t[2]
names = ["p0a seen->pv0a issued", "pv0a->p0b seen", "p0b seen->qk0 committed", "qk0->p1a seen", "p1a->pv1a issued",
         "pv1a->p1b seen", "p1b->qk1 committed"]
dl = [np.mean(m[lo:hi, s + 1] - m[lo:hi, s]) for s in range(7)]
print("mma: period %.0f cyc | " % np.mean(m[lo + 1:hi + 1, 0] - m[lo:hi, 0]) + " | ".join(f"{n} {v_:.0f}" for n, v_ in zip(names, dl))
      + f" | qk1->next p0a seen {np.mean(m[lo + 1:hi + 1, 0] - m[lo:hi, 7]):.0f}")
for i in (0, 1):
    x = t[i]
    print(f"tile {i}: qk committed -> s_full seen {np.mean(x[lo + 1:hi + 1, 0] - m[lo:hi, 4 * i + 3]):.0f} | "
          f"P half 1 published -> MMA saw it {np.mean(m[lo:hi, 4 * i + 0] - x[lo:hi, 4]):.0f} | "
          f"P half 2 published -> MMA saw it {np.mean(m[lo:hi, 4 * i + 2] - x[lo:hi, 6]):.0f}")
