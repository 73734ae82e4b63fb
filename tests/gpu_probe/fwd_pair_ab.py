"""A/B of the dense d=128 forward on CTA pairs (K1P, FA_B200_FWD_PAIR=1; K1Q, =2) against the single-CTA kernel (K1, =0).

    python tests/gpu_probe/fwd_pair_ab.py [--modes 2,0]   # default modes 0, 1, 0, each in a subprocess (the variable is read once)
    python tests/gpu_probe/fwd_pair_ab.py --run      # the cases in this process, with whatever FA_B200_FWD_PAIR says

Each case: max-abs error against float64 on sampled heads/rows; the big ones are also timed with CUDA events.
"""
import json
import os
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

CASES = [
    # name, B, H, L, dtype, input scale, timed
    ("small", 1, 2, 256, "bf16", 1.0, False),
    ("L100", 1, 2, 100, "bf16", 1.0, False),
    ("ragged333", 1, 3, 333, "bf16", 1.0, False),
    ("odd_blocks640", 2, 1, 640, "f16", 1.0, False),
    ("L1024", 2, 4, 1024, "bf16", 1.0, False),
    ("peaky", 1, 2, 1024, "bf16", 12.0, False),
    ("L1", 1, 1, 1, "bf16", 1.0, False),
    ("C2_full", 32, 8, 1024, "bf16", 1.0, True),
    ("C4_slice", 1, 16, 16384, "bf16", 1.0, True),
    ("C4_quarter", 2, 32, 16384, "bf16", 1.0, True),
]


def run():
    import torch
    from exploring_flash_attention_b200 import ops
    d = 128
    mode = os.environ.get("FA_B200_FWD_PAIR", "default")
    for name, B, H, L, dt, scale, timed in CASES:
        dtype = torch.float16 if dt == "f16" else torch.bfloat16
        g = torch.Generator().manual_seed(42)
        nb = 1 if timed else B
        Q, K, V = ((torch.rand((nb, H, L, d), generator=g) * 2 - 1) for _ in range(3))
        Q, K = Q * scale, K * scale
        Q, K, V = (x.to(dtype).cuda().expand(B, H, L, d).contiguous() for x in (Q, K, V))
        O, lse = ops.flash_attention_v1_ex(Q, K, V, return_lse=True)
        torch.cuda.synchronize()
        heads = sorted({0, B * H - 1, (B * H) // 2})
        rows = sorted(set(list(range(0, min(L, 96))) + list(range(max(0, L - 96), L)) + list(range(L // 2, min(L, L // 2 + 64)))))
        err = lerr = 0.0
        for h in heads:
            q = Q.reshape(B * H, L, d)[h][rows].double()
            k, v = K.reshape(B * H, L, d)[h].double(), V.reshape(B * H, L, d)[h].double()
            S = q @ k.T / d ** 0.5
            ref = torch.softmax(S, -1) @ v
            err = max(err, float((O.reshape(B * H, L, d)[h][rows].double() - ref).abs().max()))
            lerr = max(lerr, float((lse.reshape(B * H, L)[h][rows].double() - torch.logsumexp(S, -1)).abs().max()))
        out = {"mode": mode, "case": name, "max_abs_err": err, "lse_err": lerr, "nan": bool(torch.isnan(O).any())}
        if timed:
            fn = lambda: ops.flash_attention_v1(Q, K, V, O)
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20 if L <= 1024 else 5
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out["ms"] = ms
            out["tflops"] = 4.0 * B * H * L * L * d / ms / 1e9
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if "--run" in sys.argv:
        run()
    else:
        modes = sys.argv[sys.argv.index("--modes") + 1].split(",") if "--modes" in sys.argv else ("0", "1", "0")
        for mode in modes:
            env = dict(os.environ, FA_B200_FWD_PAIR=mode)
            try:
                r = subprocess.run([sys.executable, __file__, "--run"], capture_output=True, text=True, env=env, timeout=150)
                print(r.stdout.strip(), flush=True)
                if r.returncode != 0:
                    print(f"mode {mode}: rc={r.returncode} {r.stderr.strip()[-600:]}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"mode {mode}: TIMEOUT", flush=True)
