"""Randomised parity stress (dev tool): random shapes / dtypes / variants vs a float64 torch evaluation, with the
persistent kernel seeing many different item counts, ragged tails and split sizes.  python tests/gpu_probe/stress.py [seconds]"""
import random
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from exploring_flash_attention_b200 import ops

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = random.Random(1234)
TOL = {torch.float32: 1e-3, torch.bfloat16: 2e-3, torch.float16: 2e-3}
t_end = time.time() + budget
n = 0
worst = {}
while time.time() < t_end:
    dtype = rng.choice([torch.bfloat16, torch.float16, torch.float32])
    d = rng.choice([32, 64, 128, 256] if dtype == torch.float32 else [32, 64, 128, 128, 256, 512])
    L = rng.choice([1, 7, 64, 127, 128, 129, 255, 256, 257, 300, 511, 640, 1000, 1024, 1500, 2048, 3000])
    big_rows = d >= 256 or (dtype == torch.float32 and d >= 128)     # rows of 512-1024 bytes: the tiled-d kernels
    if big_rows:
        L = min(L, 1024)
    BH = rng.choice([1, 2, 3, 5, 37, 149, 300]) if L <= 300 else rng.choice([1, 2, 3, 5, 19])
    if big_rows:
        variant = rng.choice(["td", "tdp", "v2", "varlen", "causal", "parts"] if dtype != torch.float32 else
                             ["td", "v2", "varlen", "causal", "parts"])
    else:
        variant = rng.choice(["v1", "v1", "v2", "v2", "varlen", "causal", "parts"] +
                             (["bwd", "bwd"] if dtype != torch.float32 and d in (64, 128) else []))
    g = torch.Generator().manual_seed(n)
    Q, K, V = ((torch.rand((1, BH, L, d), generator=g) * 2 - 1).to(dtype).cuda() for _ in range(3))
    mask = None          # [BH, Lq, Lk] bool of attended keys, when the variant masks
    if variant == "varlen":
        # heads become batch entries so every head gets its own key-padding length; K/V get their own length too
        Lk = rng.choice([1, 64, 129, 300, 1024, L])
        K, V = ((torch.rand((1, BH, Lk, d), generator=g) * 2 - 1).to(dtype).cuda() for _ in range(2))
        lens = torch.tensor([rng.randint(1, Lk) for _ in range(BH)], dtype=torch.int32, device="cuda")
        O = ops.flash_attention_varlen(Q.reshape(BH, 1, L, d), K.reshape(BH, 1, Lk, d), V.reshape(BH, 1, Lk, d), lens,
                                       sync=True).reshape(1, BH, L, d)
        mask = (torch.arange(Lk, device="cuda")[None, None, :] < lens[:, None, None]).expand(BH, L, Lk)
        tag = f"varlen/Lk{Lk}"
    elif variant == "causal":
        O = ops.flash_attention_v1_ex(Q, K, V, causal=True, sync=True)
        mask = torch.ones((L, L), dtype=torch.bool, device="cuda").tril()[None].expand(BH, L, L)
        tag = "causal"
    elif variant == "parts":
        # key shards of random sizes -> partials -> combine (what one rank of ring_attention does)
        cuts = sorted(set(rng.sample(range(1, L), min(L - 1, rng.randint(1, 4))))) if L > 1 else []
        bounds = [0] + cuts + [L]
        o_parts = torch.empty((len(bounds) - 1, BH, L, d), dtype=torch.float32, device="cuda")
        lse_parts = torch.empty((len(bounds) - 1, BH, L), dtype=torch.float32, device="cuda")
        for si in range(len(bounds) - 1):
            ops.flash_attention_partial(Q, K[:, :, bounds[si]:bounds[si + 1]].contiguous(),
                                        V[:, :, bounds[si]:bounds[si + 1]].contiguous(), o_parts[si], lse_parts[si])
        O = ops.flash_attention_v2_combine(o_parts, lse_parts, dtype, (1, BH, L, d))
        torch.cuda.synchronize()
        tag = f"parts/{len(bounds) - 1}"
    elif variant == "bwd":
        causal = rng.random() < 0.5
        dO = ((torch.rand((1, BH, L, d), generator=g) * 2 - 1).to(dtype)).cuda()
        O, lse = ops.flash_attention_v1_ex(Q, K, V, causal=causal, return_lse=True, sync=True)
        grads = ops.flash_attention_backward(Q, K, V, O, dO, lse, causal=causal, sync=True)
        nh = min(BH, 3)
        idx = torch.tensor(sorted(rng.sample(range(BH), nh)), device="cuda")
        q, k, v = (x[0, idx].double().requires_grad_(True) for x in (Q, K, V))
        sc = q @ k.transpose(-1, -2) / d ** 0.5
        if causal:
            sc = sc.masked_fill(~torch.ones((L, L), dtype=torch.bool, device="cuda").tril(), float("-inf"))
        (torch.softmax(sc, -1) @ v * dO[0, idx].double()).sum().backward()
        for name, got, ref in zip(("dQ", "dK", "dV"), grads, (q.grad, k.grad, v.grad)):
            err = (got[0, idx].double() - ref).abs().max().item()
            # P and dS enter the tensor core as 16-bit operands: 2^-9 relative rounding each for bf16 (the pytest cases sit
            # at 3-4e-3 of max|grad|, random shapes reach 5.1e-3), 2^-12 for fp16
            rel_tol = 1e-2 if dtype == torch.bfloat16 else 5e-3
            if not (err <= rel_tol * max(ref.abs().max().item(), 1e-3)) or torch.isnan(got).any():
                print(f"FAIL case {n}: bwd {name} causal={causal} dtype={dtype} BH={BH} L={L} d={d} err={err} max|ref|={ref.abs().max().item()}", flush=True)
                sys.exit(1)
        n += 1
        continue
    elif variant == "v1":
        O = ops.flash_attention_v1(Q, K, V, sync=True)
        tag = "v1"
    elif variant == "td":
        O = ops.flash_attention_v1_tiled_d(Q, K, V, sync=True)
        tag = "td"
    elif variant == "tdp":     # the CTA-pair kernel called explicitly (it also serves d = 256, which "td" routes to K2)
        O = ops.flash_attention_v1_tiled_d_pair(Q, K, V, sync=True)
        tag = "tdp"
    else:
        kvs = rng.choice([8, 16, 64, 64, 100, 128, 128, 256, 1000])
        O = ops.flash_attention_v2(Q, K, V, kvs, sync=True)
        tag = f"v2/{kvs}"
    nh = min(BH, 4)
    idx = torch.tensor(sorted(rng.sample(range(BH), nh)), device="cuda")
    q, k, v = (x[0, idx].double() for x in (Q, K, V))
    sc = q @ k.transpose(-1, -2) / d ** 0.5
    if mask is not None:
        sc = sc.masked_fill(~mask[idx], float("-inf"))
    ref = torch.softmax(sc, -1) @ v
    err = (O[0, idx].double() - ref).abs().max().item()
    key = (str(dtype), d)
    worst[key] = max(worst.get(key, 0.0), err)
    # tiny L gives O(1) outputs: allow for the storage type's own output rounding (2^-9 relative for bf16, 2^-12 fp16)
    out_round = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11, torch.float32: 2.0 ** -11}[dtype] * ref.abs().max().item()
    if not (err <= TOL[dtype] + out_round) or torch.isnan(O).any():
        print(f"FAIL case {n}: {tag} dtype={dtype} BH={BH} L={L} d={d} err={err}", flush=True)
        sys.exit(1)
    n += 1
print(f"stress OK: {n} random cases in {budget:.0f} s; worst max-abs error per (dtype, d): "
      + ", ".join(f"{k[0].split('.')[-1]}/d{k[1]}={v:.1e}" for k, v in sorted(worst.items())))
