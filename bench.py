#!/usr/bin/env python
"""bench.py — flash-attention forward throughput on B200 (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c1|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of synthetic Q,K,V (U[-1,1), seed 42, like the reference drivers,
flash_attention_v1/CUDA/driver.cu:71-75).  Default workload = BASELINE.json configs[1]: tiled-d forward B32 H8 L1024
d128 (bf16 storage, fp32 accumulation).  With N ranks every rank runs the same batch on its own GPU (heads are
independent: weak scaling, no data-path collective); value = N * 4*B*H*L^2*d / max-over-ranks time.
Prints ONE JSON line on rank 0 (contract in the task statement): value/ms_per_step (device-timed, inputs resident),
e2e (host buffers through the C ABI, H2D+D2H inside), roofline, cpu_baseline, clocks, gpu_launches.
`--impl reference` times the reference's own CPU path (common/standard.h compiled into oracle/_ref) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (B, H, L, d, dtype, description)
    "c2": (32, 8, 1024, 128, "bf16", "V1 tiled-d forward B32 H8 L1024 d128 bf16 (BASELINE.json configs[1])"),
    "c1": (32, 8, 1024, 32, "f32", "V1 forward B32 H8 L1024 d32 fp32 storage / tf32 MMA (configs[0])"),
    "c4": (8, 32, 16384, 128, "bf16", "long-sequence forward B8 H32 L16384 d128 bf16 (configs[3])"),
    "c5": (16, 8, 4096, 512, "bf16", "tiled-d forward B16 H8 L4096 d512 bf16 (configs[4])"),
}
METRIC, UNIT = "attn_fwd_tflops", "TFLOP/s"


def flops(B, H, L, d):
    return 4.0 * B * H * L * L * d


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"bf16_tflops": j["bf16_tflops"], "bf16_tflops_sustained": j.get("bf16_tflops_sustained"),
                "hbm_gbs": j["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Polls NVML for SM clock + throttle reasons while the GPU is under load."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._on = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001 - NVML missing: clocks reported as unavailable
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            if self._on.is_set():
                try:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.002)

    def start(self):
        self._on.set()

    def pause(self):
        self._on.clear()

    def result(self, note):
        self._stop.set()
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "note": note}


def nvml_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            return local_rank
    return local_rank


def time_steps(fn, steps, torch):
    """CUDA events on the launching (current) stream around exactly `steps` calls."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)  # ms


def run_ours(args):
    import torch
    import torch.distributed as dist
    from exploring_flash_attention_b200 import _lib, ops
    _lib.load()  # fail loudly if the CUDA library is missing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, H, L, d, dt, desc = WORKLOADS[args.workload]
    dtype = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[dt]
    esize = 4 if dt == "f32" else 2
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- synthetic inputs: NSETS rotating input sets, each (3 inputs + output) = 4*B*H*L*d*esize bytes
    # weak scaling (default): every rank runs the whole batch.  --scaling strong: the B*H heads are sharded by
    # contiguous slices (sharding.head_range), the north_star's (batch, head) partition; still no collective.
    strong = args.scaling == "strong" and world > 1
    total_flops = flops(B, H, L, d) * (1 if strong else world)
    if strong:
        from exploring_flash_attention_b200.sharding import head_range
        hb, he = head_range(B * H, rank, world)
        Bl, Hl = 1, he - hb
    else:
        Bl, Hl = B, H
    tensor_bytes = Bl * Hl * L * d * esize
    nsets = 3 if tensor_bytes * 4 < (1 << 30) else 1
    g = torch.Generator(device="cpu").manual_seed(42 + rank)
    sets = []
    for _ in range(nsets):
        q, k, v = ((torch.rand((Bl, Hl, L, d), generator=g, dtype=torch.float32) * 2 - 1).to(dtype).cuda() for _ in range(3))
        sets.append((q, k, v, torch.empty_like(q)))
    variant = 1 if args.workload in ("c2", "c5") else 0

    def step(i):
        q, k, v, o = sets[i % nsets]
        if variant == 1:
            ops.flash_attention_v1_tiled_d(q, k, v, o, d_tile_qk=32, d_tile_v=32)
        else:
            ops.flash_attention_v1(q, k, v, o)

    sampler = ClockSampler(nvml_index(local_rank)) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    if sampler:
        sampler.start()
    ms_total = time_steps(step, args.steps, torch)
    barrier()
    if sampler:
        sampler.pause()
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = total_flops / (ms_step * 1e-3) / 1e12
    clock_note = "sampled during the timed region"
    if sampler and len(sampler.samples) < 5:
        # timed region too short for NVML polling: replay the same step for ~0.4 s (untimed) and sample under that load
        sampler.start()
        t_end = time.time() + 0.4
        i = 0
        while time.time() < t_end:
            for _ in range(20):
                step(i)
                i += 1
            torch.cuda.synchronize()
        sampler.pause()
        clock_note = "timed region shorter than the NVML polling period; sampled during an untimed replay of the same step"

    # ---- roofline of the dominant kernel (the one fused forward kernel per step), timed live on its stream
    # one kernel launch per step, launches back to back on one stream: the kernel's average duration over the timed
    # region IS ms_step (a second timing pass after the clock-sampling replay would run power-capped and read lower)
    per_launch_ms = ms_step
    achieved = flops(Bl, Hl, L, d) / (per_launch_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "achieved": round(achieved, 1), "peak": pk["bf16_tflops"] / (2.0 if dt == "f32" else 1.0),
                "unit": "TFLOP/s", "frac": round(achieved / (pk["bf16_tflops"] / (2.0 if dt == "f32" else 1.0)), 4),
                "traffic": NCU_TRAFFIC_BYTES.get(args.workload), "kernel": "fa_fwd_kernel" if d <= 128 else ("fa_tiled_d_pair_kernel" if d == 512 and dt != "f32" else "fa_tiled_d_kernel"),
                "peak_source": pk["source"] + (", burst bf16 cuBLAS" if dt != "f32" else ", burst bf16 cuBLAS / 2 (tf32)"),
                "frac_of_sustained": round(achieved / (pk["bf16_tflops_sustained"] / (2.0 if dt == "f32" else 1.0)), 4)
                if pk["bf16_tflops_sustained"] else None,
                "algorithmic_flops_per_launch": flops(Bl, Hl, L, d)}

    # ---- end to end through the C ABI with HOST buffers (H2D x3 + kernel + D2H inside the timed region)
    qh, kh, vh = (t.cpu().pin_memory() for t in sets[0][:3])
    oh = torch.empty_like(qh).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ops.flash_attention_host(qh, kh, vh, oh, variant=variant)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ops.flash_attention_host(qh, kh, vh, oh, variant=variant)   # synchronous: returns after the D2H copy
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    barrier()
    e2e = {"value": round(total_flops / (e2e_ms * 1e-3) / 1e12, 3), "unit": UNIT,
           "h2d_bytes_per_step": 3 * tensor_bytes, "d2h_bytes_per_step": tensor_bytes, "ms_per_step": round(e2e_ms, 3),
           "steps": e2e_steps, "api": "fa_forward_host (include/fa_b200.h)"}

    out = None
    if rank == 0:
        also = {}
        if args.also and world == 1:
            also = side_measurements(torch, ops, pk)
        cpu_base = cpu_baseline(args.workload) if world == 1 else None
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": dt, "data": "synthetic U[-1,1) seed 42 (random Q,K,V; no weights on this path)",
            "config": {"workload": desc, "B": B, "H": H, "L": L, "d": d, "per_gpu_batch": f"{Bl * Hl} of {B * H} heads" if strong else f"B{B} H{H}",
                       "sharding": "independent (batch,head) work per GPU, no collective",
                       "l2": f"{nsets} rotating input sets, {4 * tensor_bytes * nsets / 1e6:.0f} MB working set > 126 MB L2, no flush"},
            "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu_base,
            "clocks": sampler.result(clock_note) if sampler else None,
        }
        if also:
            out["also"] = also
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture (profiles/)
NCU_TRAFFIC_BYTES: dict = {"c2": 243_684_096}   # profiles/r1_fwd_c2_persistent_full.txt: 201.45 MB read + 42.24 MB written


def side_measurements(torch, ops, pk):
    """Other BASELINE.json configs, reported beside the headline (not the bench value): C4 long sequence, C5 d=512,
    C1 tf32, C3 split-KV + combine with the combine kernel's HBM roofline."""
    res = {}

    def timed(fn, n, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / n

    g = torch.Generator(device="cpu").manual_seed(7)
    mk = lambda B, H, L, d, dtype: tuple((torch.rand((B, H, L, d), generator=g) * 2 - 1).to(dtype).cuda() for _ in range(3))
    try:
        B, H, L, d = 8, 32, 16384, 128
        q, k, v = mk(1, 32, L, d, torch.bfloat16)          # generate one batch row, tile it to B=8 (host RNG is slow)
        q, k, v = (x.expand(B, H, L, d).contiguous() for x in (q, k, v))
        o = torch.empty_like(q)
        ms = timed(lambda: ops.flash_attention_v1(q, k, v, o), 5, warm=2)
        tf = flops(B, H, L, d) / (ms * 1e-3) / 1e12
        res["c4_B8_H32_L16384_d128_bf16"] = {"ms": round(ms, 3), "tflops": round(tf, 1),
                                            "frac_of_measured_bf16_peak": round(tf / pk["bf16_tflops"], 4),
                                            "frac_of_nominal_2250": round(tf / 2250.0, 4)}
        del q, k, v, o
    except Exception as e:  # noqa: BLE001
        res["c4_error"] = str(e)[:200]
    try:
        B, H, L, d = 16, 8, 4096, 512                      # configs[4]: tiled-d on CTA pairs (fa_tiled_d_pair_kernel)
        q, k, v = mk(1, H, L, d, torch.bfloat16)
        q, k, v = (x.expand(B, H, L, d).contiguous() for x in (q, k, v))
        o = torch.empty_like(q)
        ms = timed(lambda: ops.flash_attention_v1_tiled_d(q, k, v, o), 5, warm=2)
        tf = flops(B, H, L, d) / (ms * 1e-3) / 1e12
        res["c5_B16_H8_L4096_d512_bf16"] = {"ms": round(ms, 3), "tflops": round(tf, 1),
                                           "frac_of_measured_bf16_peak": round(tf / pk["bf16_tflops"], 4),
                                           "bound": "shared-memory bandwidth / L2 delivery (DESIGN.md K2P)"}
        del q, k, v, o
    except Exception as e:  # noqa: BLE001
        res["c5_error"] = str(e)[:200]
    try:
        B, H, L, d = 32, 8, 1024, 32
        q, k, v = mk(B, H, L, d, torch.float32)
        o = torch.empty_like(q)
        ms = timed(lambda: ops.flash_attention_v1(q, k, v, o), 50)
        tf = flops(B, H, L, d) / (ms * 1e-3) / 1e12
        res["c1_B32_H8_L1024_d32_f32_tf32"] = {"ms": round(ms, 4), "tflops": round(tf, 1),
                                              "frac_of_tf32_peak(=bf16/2)": round(tf / (pk["bf16_tflops"] / 2), 4)}
    except Exception as e:  # noqa: BLE001
        res["c1_error"] = str(e)[:200]
    try:
        B, H, L, d, kvs = 32, 8, 256, 64, 64
        q, k, v = mk(B, H, L, d, torch.bfloat16)
        o = torch.empty_like(q)
        ws = ops.v2_workspace(B, H, L, d, kvs, q.device)
        ms_all = timed(lambda: ops.flash_attention_v2(q, k, v, kvs, O=o, workspace=ws), 50)
        ops.flash_attention_v2_splitkv(q, k, v, kvs, *ws)
        S = ws[0].shape[0]
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda").zero_()
        # (a) back to back (workspace may sit in the 126 MB L2), (b) L2 flushed before every launch by READING 512 MB
        # (a write-flush would leave 126 MB of dirty lines whose write-back competes with the timed kernel)
        ms_hot = timed(lambda: ops.flash_attention_v2_combine(ws[0], ws[1], torch.bfloat16, (B, H, L, d), o), 50)
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(20)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(20)]
        for i in range(20):
            flush.view(torch.int64).sum()
            e0[i].record()
            ops.flash_attention_v2_combine(ws[0], ws[1], torch.bfloat16, (B, H, L, d), o)
            e1[i].record()
        torch.cuda.synchronize()
        ms_cold = statistics.median(a.elapsed_time(b) for a, b in zip(e0, e1))
        rows = B * H * L
        alg_bytes = S * rows * d * 4 + S * rows * 4 + rows * d * 2
        res["c3_B32_H8_L256_d64_bf16_4splits"] = {
            "splitkv_plus_combine_ms": round(ms_all, 4),
            "combine": {"algorithmic_bytes": alg_bytes, "ms_l2_warm": round(ms_hot, 4), "ms_l2_flushed": round(ms_cold, 4),
                        "gbs_l2_warm": round(alg_bytes / (ms_hot * 1e-3) / 1e9, 1),
                        "gbs_l2_flushed": round(alg_bytes / (ms_cold * 1e-3) / 1e9, 1),
                        "frac_of_measured_hbm_l2_flushed": round(alg_bytes / (ms_cold * 1e-3) / 1e9 / pk["hbm_gbs"], 4)}}
        # a combine large enough to stream from HBM regardless of L2: C4-shaped rows, 4 splits, d=128
        rows = 8 * 32 * 16384 // 4
        oa = torch.randn((4, rows, 128), device="cuda", dtype=torch.float32)
        ls = torch.randn((4, rows), device="cuda", dtype=torch.float32)
        ob = torch.empty((1, 1, rows, 128), device="cuda", dtype=torch.bfloat16)
        ms_big = timed(lambda: ops.flash_attention_v2_combine(oa.view(4, 1, rows, 128), ls.view(4, 1, rows), torch.bfloat16,
                                                              (1, 1, rows, 128), ob), 20)
        big_bytes = 4 * rows * 128 * 4 + 4 * rows * 4 + rows * 128 * 2
        res["combine_streaming_2.2GB"] = {"algorithmic_bytes": big_bytes, "ms": round(ms_big, 4),
                                          "gbs": round(big_bytes / (ms_big * 1e-3) / 1e9, 1),
                                          "frac_of_measured_hbm": round(big_bytes / (ms_big * 1e-3) / 1e9 / pk["hbm_gbs"], 4)}
    except Exception as e:  # noqa: BLE001
        res["c3_error"] = str(e)[:200]
    return res


def cpu_sample(workload, heads, threads, kind_pref="reference"):
    """Times the CPU path on `heads` heads of the workload shape. Returns (seconds, kind, threads)."""
    import numpy as np
    from oracle import cpu
    B, H, L, d, dt, _ = WORKLOADS[workload]
    rng = np.random.default_rng(42)
    # the reference CPU path stores __half (USE_FP64=0); bf16/fp32 workloads are timed at fp16 storage (same flops)
    Q, K, V = (rng.uniform(-1, 1, (1, heads, L, d)).astype(np.float16) for _ in range(3))
    if kind_pref == "reference" and cpu.have_ref():
        t0 = time.perf_counter()
        cpu.ref_standard_attention_cpu(Q, K, V, n_threads=threads)
        return time.perf_counter() - t0, "reference", threads
    if not (ROOT / "oracle" / "liboracle.so").exists():
        cpu.build()
    t0 = time.perf_counter()
    cpu.standard_attention_cpu(Q, K, V, n_threads=threads)
    return time.perf_counter() - t0, "port", threads


def cpu_baseline(workload):
    """Reported baseline only: the reference's OpenMP standard_attention_cpu (common/standard.h:28-102, compiled
    unmodified into oracle/_ref) on a bounded sample of heads of the same workload, all host threads."""
    try:
        B, H, L, d, dt, _ = WORKLOADS[workload]
        threads = os.cpu_count() or 1
        per_head_gflop = flops(1, 1, L, d) / 1e9
        if per_head_gflop > 40:      # one head alone would take minutes on the host (C4: 137 GFLOP per head)
            return {"value": None, "unit": UNIT, "cores": threads, "kind": "unavailable",
                    "sample": f"one head of this workload is {per_head_gflop:.0f} GFLOP: too long for a bounded CPU sample"}
        # ~10-30 s of CPU work: the reference path runs ~0.03 GFLOP/s... measured 0.9 GFLOP/s per thread here
        heads = int(min(B * H, max(threads, 12.0 * threads / max(per_head_gflop, 1e-3))))
        heads = int(os.environ.get("FA_BENCH_CPU_HEADS", heads))   # test hook: shrink the CPU sample
        secs, kind, threads = cpu_sample(workload, heads, threads)
        tf = flops(1, heads, L, d) / secs / 1e12
        out = {"value": round(tf, 5), "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{heads} of {B * H} heads of the same shape (L={L}, d={d}, fp16 storage like the reference's "
                         f"DATA_TYPE=__half), 1 pass, {secs:.1f} s",
               "seconds": round(secs, 2), "extrapolated_full_batch_seconds": round(secs * B * H / heads, 1)}
        out["also"] = cpu_python_paths(B * H, L, d)
        return out
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}


def cpu_python_paths(n_heads, L, d):
    """The reference's two Python CPU paths on small stated samples of the same shape: its oracle
    naive_attention (common/reference.py:7-21; BLAS matmuls, all host threads) looped over heads, and the V1 tile loop
    of numpy_gpu_like_opt2.py:198-241 with the reference's default Bq = Bk = 8 (restated per tile in NumPy, so this port
    is far faster than the reference's per-element Python loops: 11 s per head at L=1024, d=32)."""
    res = {}
    try:
        import numpy as np
        from oracle import reference, tiled
        rng = np.random.default_rng(0)
        # bounded samples: ~0.3 TFLOP of BLAS work for the oracle, one head for the tile loop if it has <= 32k tile updates
        heads = max(1, min(n_heads, 32, int(0.3e12 / flops(1, 1, L, d))))
        heads = int(os.environ.get("FA_BENCH_CPU_HEADS", heads))
        Q, K, V = (rng.uniform(-1, 1, (heads, L, d)).astype(np.float32) for _ in range(3))
        t0 = time.perf_counter()
        for h in range(heads):
            reference.naive_attention(Q[h], K[h], V[h])
        secs = time.perf_counter() - t0
        res["naive_attention_numpy"] = {"heads": heads, "seconds": round(secs, 3), "dtype": "f32",
                                        "tflops": round(flops(1, heads, L, d) / secs / 1e12, 5),
                                        "extrapolated_full_batch_seconds": round(secs * n_heads / heads, 1)}
        if (L // 8) ** 2 <= 32768:
            O = np.zeros(L * d, dtype=np.float32)
            t0 = time.perf_counter()
            tiled.flash_attention_tiled(Q[0].reshape(-1), K[0].reshape(-1), V[0].reshape(-1), O, L, d, 8, 8)
            secs = time.perf_counter() - t0
            res["numpy_gpu_like_opt2_port_Bq8_Bk8"] = {"heads": 1, "seconds": round(secs, 3), "dtype": "f32", "threads": 1,
                                                       "tflops": round(flops(1, 1, L, d) / secs / 1e12, 6),
                                                       "extrapolated_full_batch_seconds": round(secs * n_heads, 1)}
    except Exception as e:  # noqa: BLE001
        res["error"] = str(e)[:200]
    return res


def run_reference(args):
    """The reference arm: its own CPU implementation of the path (standard_attention_cpu) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    B, H, L, d, dt, desc = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    per_head_gflop = flops(1, 1, L, d) / 1e9
    # bounded sample per step (~10 s of CPU work at ~1 GFLOP/s/thread), at most the whole batch, at least one head/thread
    heads = int(min(B * H, max(threads, 10.0 * threads / max(per_head_gflop, 1e-3))))
    heads = int(os.environ.get("FA_BENCH_CPU_HEADS", heads))       # test hook: shrink the CPU sample
    steps, warm = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    for _ in range(warm):
        cpu_sample(args.workload, heads, threads)
    secs, kind = [], "port"
    for _ in range(steps):
        s, kind, _ = cpu_sample(args.workload, heads, threads)
        secs.append(s)
    s = sum(secs) / len(secs)
    tf = flops(1, heads, L, d) / s / 1e12
    sample = (f"each step = {heads} of {B * H} heads of the workload shape (fp16 storage, the reference's DATA_TYPE), "
              f"{threads} OpenMP threads; {steps} timed step(s) after {warm} warm-up")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(tf, 5), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": round(s * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 storage / f32 math", "data": "synthetic U[-1,1) seed 42",
        "config": {"workload": desc, "B": B, "H": H, "L": L, "d": d, "sample_heads": heads},
        "cpu_baseline": {"value": round(tf, 5), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(tf, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU runs the whole batch; strong: the batch's heads are sharded across the GPUs")
    ap.add_argument("--no-also", dest="also", action="store_false", help="skip the side measurements (C4/C1/C3)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", __file__, *sys.argv[1:]]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
