"""First-light GPU check (dev tool, not part of the pytest suite): every instantiation of the fused-tile kernel,
split-KV and combine against a float64 torch evaluation on the same rounded inputs, plus CUDA-event timings.

Each case runs in its own subprocess with a timeout so a trapped/hung kernel cannot take the others down.
    python tests/gpu_probe/first_light.py            # all cases
    python tests/gpu_probe/first_light.py CASE_JSON  # one case (internal)
"""
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

CASES = [
    # name, variant, B, H, L, d, dtype, kv_per_split
    ("v1_bf16_d128_small", "v1", 1, 2, 256, 128, "bf16", 0),
    ("v1_bf16_d128_L1024", "v1", 2, 4, 1024, 128, "bf16", 0),
    ("v1_bf16_d64", "v1", 2, 2, 512, 64, "bf16", 0),
    ("v1_f16_d128", "v1", 1, 2, 512, 128, "f16", 0),
    ("v1_f32_d32", "v1", 2, 2, 1024, 32, "f32", 0),
    ("v1_f32_d64", "v1", 1, 2, 512, 64, "f32", 0),
    ("v1_bf16_d128_ragged", "v1", 1, 3, 333, 128, "bf16", 0),
    ("v1_bf16_d128_L100", "v1", 1, 2, 100, 128, "bf16", 0),
    ("v1_f32_d32_ragged", "v1", 1, 2, 777, 32, "f32", 0),
    ("v1_bf16_peaky", "v1", 1, 2, 1024, 128, "bf16", -1),  # large-magnitude scores: exercises the lazy rescale
    ("v2_bf16_d64_C3", "v2", 4, 8, 256, 64, "bf16", 64),
    ("v2_bf16_d128", "v2", 1, 4, 1024, 128, "bf16", 256),
    ("v2_f32_d32_ragged", "v2", 1, 2, 500, 32, "f32", 96),
    ("v1_f16_d32", "v1", 2, 2, 512, 32, "f16", 0),
    ("C1_f16_full", "v1", 32, 8, 1024, 32, "f16", 0),
    ("C2_full", "v1", 32, 8, 1024, 128, "bf16", 0),
    ("C1_full", "v1", 32, 8, 1024, 32, "f32", 0),
    ("C3_full", "v2", 32, 8, 256, 64, "bf16", 64),
    ("C4_slice", "v1", 1, 16, 16384, 128, "bf16", 0),
    ("C4_slice_causal", "causal", 1, 16, 16384, 128, "bf16", 0),
    ("C2_causal", "causal", 32, 8, 1024, 128, "bf16", 0),
    ("td_d256", "v1", 1, 2, 512, 256, "bf16", 0),
    ("td_d512", "v1", 1, 2, 512, 512, "bf16", 0),
    ("td_d256_big", "v1", 4, 8, 4096, 256, "bf16", 0),
    ("C5_full", "v1", 16, 8, 4096, 512, "bf16", 0),
    ("C2_tf32", "v1", 32, 8, 1024, 128, "f32", 0),
    ("d64_bf16_L1024", "v1", 32, 8, 1024, 64, "bf16", 0),
    ("d64_bf16_L8192", "v1", 2, 16, 8192, 64, "bf16", 0),
    ("d32_f16_L8192", "v1", 2, 16, 8192, 32, "f16", 0),
    ("d64_bf16_causal", "causal", 4, 8, 4096, 64, "bf16", 0),
    ("d64_bf16_v2", "v2", 2, 8, 4096, 64, "bf16", 1024),
]


def run_case(case):
    import torch
    from exploring_flash_attention_b200 import ops
    name, variant, B, H, L, d, dt, kvs = case
    dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[dt]
    g = torch.Generator(device="cpu").manual_seed(42)
    mk = lambda: (torch.rand((B, H, L, d), generator=g, dtype=torch.float32) * 2 - 1)
    Q, K, V = mk(), mk(), mk()
    if kvs == -1:
        Q = Q * 12.0
        K = K * 12.0
        kvs = 0
    Q, K, V = (x.to(dtype).cuda() for x in (Q, K, V))
    if variant == "causal":
        fn = lambda: ops.flash_attention_v1_ex(Q, K, V, causal=True)
    elif variant == "v1":
        fn = lambda: ops.flash_attention_v1(Q, K, V)
    else:
        ws = ops.v2_workspace(B, H, L, d, kvs, Q.device)
        fn = lambda: ops.flash_attention_v2(Q, K, V, kvs, workspace=ws)
    O = fn()
    torch.cuda.synchronize()
    # reference on a bounded number of heads, float64
    nh = min(B * H, 8)
    Qd, Kd, Vd = (x.reshape(B * H, L, d)[:nh].double() for x in (Q, K, V))
    rows = min(L, 2048)
    S = torch.einsum("hqd,hkd->hqk", Qd[:, :rows], Kd) / (d ** 0.5)
    if variant == "causal":
        S = S.masked_fill(torch.ones(rows, L, device=S.device, dtype=torch.bool).triu(1), float("-inf"))
    ref = torch.softmax(S, dim=-1) @ Vd
    got = O.reshape(B * H, L, d)[:nh, :rows].double()
    err = (got - ref).abs().max().item()
    nan = bool(torch.isnan(O).any().item())
    # timing
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tflops = 4.0 * B * H * L * L * d / (ms * 1e-3) / 1e12
    print(json.dumps({"case": name, "max_abs_err": err, "nan": nan, "ms": ms, "tflops": tflops,
                      "ref_absmax": ref.abs().max().item()}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1].startswith("["):
        run_case(json.loads(sys.argv[1]))
        sys.exit(0)
    only = sys.argv[1:] if len(sys.argv) > 1 else None
    for case in CASES:
        if only and case[0] not in only:
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, json.dumps(case)], capture_output=True, text=True, timeout=240)
            out = r.stdout.strip().splitlines()
            tail = out[-1] if out else ""
            print(f"[{case[0]}] rc={r.returncode} {time.time() - t0:.1f}s {tail}", flush=True)
            if r.returncode != 0:
                print("   stderr:", r.stderr.strip()[-1500:], flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{case[0]}] TIMEOUT", flush=True)
