"""Python mirrors of the reference's three self-checking drivers, running the B200 library instead of the sm_80
kernels.  Test programs (they call the CPU oracle), hence under tests/:

    python -m tests.drivers.driver v1        # flash_attention_v1/CUDA/driver.cu        B32 H8 L1024 d32
    python -m tests.drivers.driver tiled_d   # flash_attention_v1_tiled_d/CUDA/driver.cu  ... d128, D_TILE 32
    python -m tests.drivers.driver v2        # flash_attention_v2/CUDA/driver.cu        ... d128, KV_TILES_PER_BLOCK 4

Same flow as the reference mains: srand(42) U[-1,1] data in Q->K->V order rounded to fp16 (driver.cu:71-75,137,168-170);
OpenMP CPU reference timed with a wall clock; warm-up + timed launches (10+50 for v1/tiled-d, driver.cu:220-238;
1+10 for v2, flash_attention_v2/CUDA/driver.cu:155-166) timed on the host around a synchronising launcher, like the
reference, and additionally with CUDA events; compare_arrays metrics (driver.cu:78-133); verdict and exit code:
v1 PASS if max-abs < 1e-3 and exit 0 either way (driver.cu:275,293); tiled-d PASS if max-abs < 1e-2, exit !passed
(flash_attention_v1_tiled_d/CUDA/driver.cu:250,271); v2 PASS if max-abs < 0.1 and max-rel < 0.1, exit !passed
(flash_attention_v2/CUDA/driver.cu:204,223).  `--heads N` bounds the CPU reference to the first N heads (default: all).
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

CONFIGS = {
    "v1": dict(B=32, H=8, L=1024, d=32, warmup=10, runs=50, kernel="fa_fwd_kernel (tcgen05, sm_100a)"),
    "tiled_d": dict(B=32, H=8, L=1024, d=128, warmup=10, runs=50, d_tile_qk=32, d_tile_v=32,
                    kernel="fa_fwd_kernel via fa_v1_tiled_d_forward (tcgen05, sm_100a)"),
    "v2": dict(B=32, H=8, L=1024, d=128, warmup=1, runs=10, d_tile_qk=32, d_tile_v=32, kv_tiles_per_block=4, BK=16,
               kernel="fa_fwd_kernel<split> + fa_combine_kernel (sm_100a)"),
}


def compare_arrays(a, b, eps=1e-3):
    """max-abs, filtered max-rel (max(|a|,|b|) > eps) and unfiltered max-rel with a symmetric denominator."""
    a = a.astype(np.float32).ravel()
    b = b.astype(np.float32).ravel()
    diff = np.abs(a - b)
    mag = np.maximum(np.abs(a), np.abs(b))
    rel_all = diff / np.maximum(mag, 1e-8)
    mask = mag > eps
    rel = float((diff[mask] / mag[mask]).max()) if mask.any() else 0.0
    return float(diff.max()), rel, float(rel_all.max())


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=list(CONFIGS))
    ap.add_argument("--heads", type=int, default=0, help="CPU-reference only the first N heads (0 = all)")
    args = ap.parse_args(argv)
    cfg = CONFIGS[args.which]
    B, H, L, d = cfg["B"], cfg["H"], cfg["L"], cfg["d"]

    import torch
    from exploring_flash_attention_b200 import flash_attention_v1, flash_attention_v1_tiled_d, flash_attention_v2
    from oracle import cpu

    print("Flash Attention CUDA Test (B200 library)")
    print(f"Kernel: {cfg['kernel']}")
    print("Precision: FP16 (half)")
    extra = "".join(f", {k.upper()}={cfg[k]}" for k in ("d_tile_qk", "d_tile_v", "kv_tiles_per_block") if k in cfg)
    print(f"B={B}, H={H}, L={L}, d={d}{extra}")
    print(f"Total attention heads: {B * H}\n")

    hQ, hK, hV = cpu.driver_inputs(B, H, L, d, np.float16, seed=42)
    n_ref = B * H if args.heads <= 0 else min(args.heads, B * H)
    print("Computing reference (standard attention on CPU with OpenMP)...")
    t0 = time.perf_counter()
    O_ref = cpu.standard_attention_cpu(hQ, hK, hV, head_begin=0, head_end=n_ref)
    cpu_ms = (time.perf_counter() - t0) * 1e3 * (B * H / n_ref)
    note = "" if n_ref == B * H else f" (extrapolated from {n_ref} heads)"
    print(f"CPU time: {cpu_ms:.1f} ms on {cpu.max_threads()} threads{note}\n")

    dQ, dK, dV = (torch.from_numpy(x).cuda() for x in (hQ, hK, hV))
    dO = torch.empty_like(dQ)
    if args.which == "v1":
        launch = lambda: flash_attention_v1.flash_attention_v1(dQ, dK, dV, dO, B, H, L, d)
    elif args.which == "tiled_d":
        launch = lambda: flash_attention_v1_tiled_d.flash_attention_v1(dQ, dK, dV, dO, B, H, L, d, cfg["d_tile_qk"],
                                                                      cfg["d_tile_v"])
    else:
        from exploring_flash_attention_b200 import ops
        ws = ops.v2_workspace(B, H, L, d, cfg["BK"] * cfg["kv_tiles_per_block"], dQ.device)   # caller-owned, allocated once
        launch = lambda: flash_attention_v2.flash_attention_v2(dQ, dK, dV, dO, B, H, L, d, cfg["d_tile_qk"],
                                                               cfg["d_tile_v"], cfg["kv_tiles_per_block"], Bk=cfg["BK"],
                                                               workspace=ws)
    print("Running Flash Attention on GPU...")
    for _ in range(cfg["warmup"]):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(cfg["runs"]):
        launch()                       # synchronising launcher, like the reference's
    e1.record()
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t0) * 1e3
    avg = host_ms / cfg["runs"]
    print("Flash Attention complete.")
    print(f"Average GPU time ({cfg['runs']} runs): {avg:.4f} ms   (CUDA events: {e0.elapsed_time(e1) / cfg['runs']:.4f} ms)")
    print(f"Total GPU time: {host_ms:.3f} ms")
    print(f"Speedup: {cpu_ms / avg:.0f}x")
    print(f"Throughput: {4.0 * B * H * L * L * d / (avg * 1e-3) / 1e12:.1f} TFLOP/s\n")

    O = dO.cpu().numpy().reshape(B * H, L, d)[:n_ref]
    ref = O_ref.reshape(B * H, L, d)[:n_ref]
    max_abs, max_rel, max_rel_all = compare_arrays(ref, O)
    print("Results Comparison:")
    print(f"Max absolute difference: {max_abs:.3e}")
    print(f"Max relative difference (all values): {max_rel_all:.3e}")
    print(f"Max relative difference (for |values| > 1e-3): {max_rel:.3e}\n")
    print("First 5 output values:")
    print("Reference:", " ".join(f"{float(v):.6g}" for v in ref.ravel()[:5]))
    print("Flash:    ", " ".join(f"{float(v):.6g}" for v in O.ravel()[:5]))
    if args.which == "v1":
        passed = max_abs < 1e-3
    elif args.which == "tiled_d":
        passed = max_abs < 1e-2
    else:
        passed = max_abs < 0.1 and max_rel < 0.1
    print("\n✓ Test PASSED - Results match!" if passed else "\n✗ Test FAILED - Results differ significantly!")
    if args.which == "v1":
        return 0                       # the reference V1 driver returns 0 even on FAIL (driver.cu:293)
    return 0 if passed else 1


if __name__ == "__main__":
    sys.exit(main())
