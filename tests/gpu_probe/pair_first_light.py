"""First light / A-B of the CTA-pair tiled-d kernel (fa_v1_tiled_d_pair_forward) against the routed tiled-d entry point
(fa_v1_tiled_d_forward) and float64.  The routed call is the single-CTA slab kernel K2 for d = 256 and, when the process runs
with FA_B200_TILED_D_PAIR=0, for d = 512 as well; otherwise d = 512 is the pair kernel on both sides and the two timings only
show the drift between a first and a second timing loop (power cap).  JSON keys keep the historical "slab_" prefix for the
routed call.

    python tests/gpu_probe/pair_first_light.py            # every case, each in its own subprocess with a timeout
    python tests/gpu_probe/pair_first_light.py --case N   # one case in this process
    FA_B200_LIB=build_variants/x.so python tests/gpu_probe/pair_first_light.py --cases 1,5,6   # A/B of a tuning build

A case prints one JSON line: max-abs error of both kernels vs float64, the error of the pair kernel by 32-row block
and 64-column block (a layout bug shows up as a block pattern), and CUDA-event timings.
"""
import json
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

CASES = [
    # (B, H, L, d, dtype, timed)
    (1, 1, 128, 512, "bf16", False),     # one KV tile, one pair
    (1, 1, 256, 512, "bf16", False),     # two tiles: S / P double buffering, second pair
    (1, 2, 333, 512, "bf16", False),     # ragged tail, two heads
    (1, 1, 100, 256, "f16", False),      # d = 256 instantiation, L < 128
    (2, 2, 1024, 512, "f16", False),
    (1, 1, 640, 512, "bf16-ramp", False),  # growing key norms: the lazy-rescale branch
    (16, 8, 4096, 512, "bf16", True),    # C5 full size
    (16, 8, 4096, 256, "bf16", True),
]


def run_case(i: int) -> None:
    import torch
    from exploring_flash_attention_b200 import ops
    B, H, L, d, kind, timed = CASES[i]
    dtype = torch.float16 if kind.startswith("f16") else torch.bfloat16
    g = torch.Generator().manual_seed(7)
    nb = 1 if timed else B   # full-size inputs: one batch entry of random data, repeated
    Q, K, V = ((torch.rand((nb, H, L, d), generator=g) * 2 - 1).to(dtype).cuda() for _ in range(3))
    if nb != B:
        Q, K, V = (x.expand(B, H, L, d).contiguous() for x in (Q, K, V))
    if kind.endswith("ramp"):
        K = (K.float() * torch.linspace(0.3, 8.0, L, device="cuda").view(1, 1, L, 1)).to(dtype)
    out = {"case": i, "shape": [B, H, L, d], "kind": kind}
    heads = [(0, 0), (B - 1, H - 1)]
    rows = sorted(set(list(range(0, min(L, 64))) + list(range(max(0, L - 64), L))))
    ref = {}
    for (b, h) in heads:
        q, k, v = Q[b, h, rows].double(), K[b, h].double(), V[b, h].double()
        ref[(b, h)] = torch.softmax(q @ k.T / d ** 0.5, -1) @ v
    for name, fn in (("slab", lambda: ops.flash_attention_v1_tiled_d(Q, K, V, sync=True)),
                     ("pair", lambda: ops.flash_attention_v1_tiled_d_pair(Q, K, V, sync=True))):
        O = fn()
        errs = [(O[b, h, rows].double() - ref[(b, h)]).abs() for (b, h) in heads]
        E = torch.stack(errs).amax(0)
        out[f"{name}_max_abs_err"] = float(E.max())
        out[f"{name}_nan"] = bool(torch.isnan(O).any())
        if name == "pair":
            out["pair_err_rows32"] = [f"{float(E[r:r + 32].max()):.1e}" for r in range(0, E.shape[0], 32)]
            out["pair_err_cols64"] = [f"{float(E[:, c:c + 64].max()):.1e}" for c in range(0, d, 64)]
        if timed:
            for _ in range(3):
                fn()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            n = 10
            for _ in range(n):
                fn()
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / n
            out[f"{name}_ms"] = ms
            out[f"{name}_tflops"] = 4.0 * B * H * L * L * d / ms / 1e9
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if "--case" in sys.argv:
        run_case(int(sys.argv[sys.argv.index("--case") + 1]))
    elif "--cases" in sys.argv:      # several cases in this process (A/B of builds selected through FA_B200_LIB)
        for i in sys.argv[sys.argv.index("--cases") + 1].split(","):
            run_case(int(i))
    else:
        for i in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, __file__, "--case", str(i)], capture_output=True, text=True, timeout=90)
                print(r.stdout.strip() or f"case {i}: rc={r.returncode} {r.stderr.strip()[-400:]}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"case {i}: TIMEOUT", flush=True)
                break
