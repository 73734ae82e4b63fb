"""A/B of the split-KV partial kernels at short splits: fa_splitkv_tile_kernel (default) vs the fused-tile kernel's SPLIT
mode (FA_B200_SPLITKV_TILE=0).  Prints one JSON line per mode: split-KV alone, combine alone, both (us)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

CASES = [(32, 8, 256, 64, "bf16", 64), (32, 8, 1024, 128, "bf16", 128), (32, 8, 1024, 32, "f32", 64), (8, 8, 2048, 64, "f16", 128)]


def child():
    import torch
    from exploring_flash_attention_b200 import ops
    out = {"tile_kernel": os.environ.get("FA_B200_SPLITKV_TILE", "1")}
    for B, H, L, d, dt, kvs in CASES:
        dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[dt]
        q, k, v = ((torch.rand((B, H, L, d), device="cuda") * 2 - 1).to(dtype) for _ in range(3))
        o = torch.empty_like(q)
        ws = ops.v2_workspace(B, H, L, d, kvs, q.device)

        def timed(fn, n=50):
            """Device time per call: n calls captured into one CUDA graph (no per-call host launch cost: the Python ->
            ctypes -> cudaLaunch path is 8-20 us per call, as long as these kernels), replayed 5 times, best replay."""
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(n):
                    fn()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1) / n * 1e3)
            return round(best, 2)

        out[f"B{B}H{H}L{L}d{d}{dt}_kvs{kvs}"] = {
            "splitkv_us": timed(lambda: ops.flash_attention_v2_splitkv(q, k, v, kvs, *ws)),
            "combine_us": timed(lambda: ops.flash_attention_v2_combine(ws[0], ws[1], dtype, (B, H, L, d), o)),
            "both_us": timed(lambda: ops.flash_attention_v2(q, k, v, kvs, O=o, workspace=ws)),
            "v1_us": timed(lambda: ops.flash_attention_v1(q, k, v, o))}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for mode in ("1", "0"):
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, FA_B200_SPLITKV_TILE=mode), check=False)
