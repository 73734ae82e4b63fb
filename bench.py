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


def bench_config(workload, strong=False, world=1):
    """The `config` object of the JSON line — identical in both arms (ours and --impl reference) by construction."""
    B, H, L, d, dt, desc = WORKLOADS[workload]
    tensor_bytes = B * H * L * d * (4 if dt == "f32" else 2) // (world if strong else 1)
    nsets = 3 if tensor_bytes * 4 < (1 << 30) else 1
    return {"workload": desc, "B": B, "H": H, "L": L, "d": d,
            "per_gpu_batch": f"1/{world} of the {B * H} heads" if strong else f"B{B} H{H}",
            "sharding": "independent (batch,head) work per GPU, no collective",
            "l2": f"GPU arm: {nsets} rotating input sets, {4 * tensor_bytes * nsets / 1e6:.0f} MB working set > 126 MB L2, no flush"}


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
    launches = [0]

    def step(i):
        q, k, v, o = sets[i % nsets]
        if variant == 1:
            ops.flash_attention_v1_tiled_d(q, k, v, o, d_tile_qk=32, d_tile_v=32)
        else:
            ops.flash_attention_v1(q, k, v, o)
        launches[0] += 1      # one kernel of ours per step (fa_fwd_kernel / fa_tiled_d[_pair]_kernel), nothing else

    K = args.steps
    sampler = ClockSampler(nvml_index(local_rank)) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        step(i)
    # calibration (untimed): how many blocks of K steps make the timed region at least MIN_REGION_MS long, so that the
    # clock samples come from the timed region itself and a max-over-ranks is not the jitter of a 2 ms window
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for i in range(K):
        step(i)
    c1.record()
    c1.synchronize()
    est_block_ms = max_over_ranks(c0.elapsed_time(c1))
    MIN_REGION_MS = 60.0
    R = max(1, min(2000, int(-(-MIN_REGION_MS // max(est_block_ms, 1e-3)))))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
    barrier()
    launches[0] = 0
    if sampler:
        sampler.start()
    ev[0].record()
    for r in range(R):          # R blocks of EXACTLY K steps, back to back on the launching (current) stream
        for i in range(K):
            step(r * K + i)
        ev[r + 1].record()
    ev[R].synchronize()
    barrier()
    if sampler:
        sampler.pause()
    gpu_launches = launches[0]
    block_ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(R)]
    ms_total = max_over_ranks(ev[0].elapsed_time(ev[R]))
    ms_step = ms_total / (R * K)                       # the bench value: whole timed region, max over ranks
    ms_step_median_block = max_over_ranks(statistics.median(block_ms)) / K
    value = total_flops / (ms_step * 1e-3) / 1e12
    clock_note = f"sampled during the timed region ({ms_total:.0f} ms)"

    # ---- roofline of the dominant kernel (the one fused forward kernel per step), timed live on its stream
    # one kernel launch per step, launches back to back on one stream: the kernel's average duration over the timed
    # region IS ms_step
    per_launch_ms = ms_step
    achieved = flops(Bl, Hl, L, d) / (per_launch_ms * 1e-3) / 1e12
    peak = pk["bf16_tflops"] / (2.0 if dt == "f32" else 1.0)
    traffic = NCU_TRAFFIC_BYTES.get(args.workload, {})
    roofline = {"bound": "tensor", "achieved": round(achieved, 1), "peak": peak,
                "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "traffic": traffic.get("bytes"), "traffic_source": traffic.get("source"),
                "kernel": "fa_fwd_kernel" if d <= 128 else ("fa_tiled_d_pair_kernel" if d == 512 and dt != "f32" else "fa_tiled_d_kernel"),
                "peak_source": pk["source"] + (", burst bf16 cuBLAS" if dt != "f32" else ", burst bf16 cuBLAS / 2 (tf32)"),
                "frac_of_sustained": round(achieved / (pk["bf16_tflops_sustained"] / (2.0 if dt == "f32" else 1.0)), 4)
                if pk["bf16_tflops_sustained"] else None,
                "algorithmic_flops_per_launch": flops(Bl, Hl, L, d),
                "algorithmic_bytes_per_launch": 4 * tensor_bytes}

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
    # the PCIe floor of the same step on this box, ALL ranks copying at once: the three inputs host->device on one
    # stream while the output goes device->host on another (full duplex), no kernel.  e2e_ms / pcie_floor_ms says how
    # much of the end-to-end time is the host link (shared by the ranks) rather than this library.
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    dq, dk, dv, do = sets[0]

    def copies():
        with torch.cuda.stream(s_in):
            dq.copy_(qh, non_blocking=True)
            dk.copy_(kh, non_blocking=True)
            dv.copy_(vh, non_blocking=True)
        with torch.cuda.stream(s_out):
            oh.copy_(do, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copies()
    floor_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    barrier()
    e2e = {"value": round(total_flops / (e2e_ms * 1e-3) / 1e12, 3), "unit": UNIT,
           "h2d_bytes_per_step": 3 * tensor_bytes, "d2h_bytes_per_step": tensor_bytes, "ms_per_step": round(e2e_ms, 3),
           "steps": e2e_steps, "api": "fa_forward_host (include/fa_b200.h)",
           "pcie_floor_ms": round(floor_ms, 3),
           "pcie_floor_note": f"H2D of the 3 inputs + D2H of the output on two streams, no kernel, all {world} rank(s) at once "
                              f"(max over ranks): {3 * tensor_bytes / floor_ms / 1e6:.1f} GB/s in per GPU; the library's "
                              f"step is {e2e_ms / floor_ms:.2f}x this floor"}
    del qh, kh, vh, oh, sets

    also = {}
    if args.also:
        # every rank takes part (C4 heads are sharded over the ranks; the ring needs all of them); rank 0 reports
        also["c4_strong"] = c4_strong(torch, dist, ops, pk, rank, world, max_over_ranks, barrier)
        if world > 1:
            also["ring"] = ring_checks(torch, dist, ops, rank, world, max_over_ranks, barrier)
        elif rank == 0:
            also.update(side_measurements(torch, ops, pk))

    out = None
    if rank == 0:
        cpu_base = cpu_baseline(args.workload) if world == 1 else None
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": dt, "data": "synthetic U[-1,1) seed 42 (random Q,K,V; no weights on this path)",
            "config": bench_config(args.workload, strong, world),
            "timing": {"repeats": R, "timed_steps": R * K, "timed_region_ms": round(ms_total, 3),
                       "ms_per_step_median_block": round(ms_step_median_block, 5),
                       "note": f"{R} back-to-back blocks of exactly {K} steps inside ONE barrier+synchronize bracket (>= {MIN_REGION_MS:.0f} ms so "
                               "clocks are sampled in the timed region); ms_per_step = whole region / timed_steps, max over ranks"},
            "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu_base,
            "clocks": sampler.result(clock_note) if sampler else None,
        }
        if also:
            out["also"] = also
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def c4_strong(torch, dist, ops, pk, rank, world, max_over_ranks, barrier):
    """BASELINE.json configs[3] at every N: B8 H32 L16384 d128 bf16, the 256 heads sharded over the ranks by
    sharding.head_range (the reference's independent grid rows, flash_attention_v1.h:170-172), no collective.
    Whole-job TFLOP/s = 4*B*H*L^2*d / max-over-ranks time; per-rank roofline fraction; and, outside the timed region, a
    sampled-row check of rank 0's output against the float64 oracle (the checker, oracle/reference.py)."""
    try:
        from exploring_flash_attention_b200.sharding import head_range
        B, H, L, d = 8, 32, 16384, 128
        hb, he = head_range(B * H, rank, world)
        nh = he - hb
        g = torch.Generator(device="cuda").manual_seed(4242 + rank)
        q, k, v = ((torch.rand((1, nh, L, d), generator=g, device="cuda") * 2 - 1).bfloat16() for _ in range(3))
        o = torch.empty_like(q)
        for _ in range(2):
            ops.flash_attention_v1(q, k, v, o)
        barrier()
        n = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ops.flash_attention_v1(q, k, v, o)
        e1.record()
        e1.synchronize()
        barrier()
        my_ms = e0.elapsed_time(e1) / n
        ms = max_over_ranks(my_ms)
        tf = flops(B, H, L, d) / (ms * 1e-3) / 1e12
        per_rank_tf = flops(1, nh, L, d) / (my_ms * 1e-3) / 1e12
        res = {"workload": "B8 H32 L16384 d128 bf16, heads sharded", "n_gpus": world, "heads_per_rank": nh, "steps": n,
               "ms": round(ms, 3), "tflops": round(tf, 1), "scaling": "strong",
               "rank0_tflops": round(per_rank_tf, 1), "rank0_roofline_frac": round(per_rank_tf / pk["bf16_tflops"], 4),
               "rank0_frac_of_sustained": round(per_rank_tf / pk["bf16_tflops_sustained"], 4) if pk["bf16_tflops_sustained"] else None,
               "frac_of_nominal_2250_per_gpu": round(tf / world / 2250.0, 4)}
        if rank == 0:
            import numpy as np
            from oracle import reference          # checker only, untimed
            heads = sorted({0, nh // 2, nh - 1})
            rows = np.r_[0:16, L // 2:L // 2 + 16, L - 16:L]
            f = lambda x: x[0, heads].float().cpu().numpy()
            ref = reference.naive_attention_batched_f64(f(q), f(k), f(v), rows=rows)
            got = o[0, heads][:, rows].float().cpu().numpy()
            res["max_abs_err"] = float(np.abs(got - ref).max())
            res["max_abs_err_note"] = f"{len(heads)} heads x {len(rows)} rows of rank 0's shard vs the float64 oracle; tolerance 2e-3"
        del q, k, v, o
        return res
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:300]}


def ring_checks(torch, dist, ops, rank, world, max_over_ranks, barrier):
    """N > 1 only, outside the bench value: (1) parity of the sequence-sharded path (sharding.ring_attention, both
    transports, dense and causal zig-zag) and of gather_heads against the float64 oracle on a small case; (2) the C4
    problem sequence-sharded over the ring, timed beside the head-sharded figure."""
    import numpy as np
    res = {}
    try:
        from exploring_flash_attention_b200 import sharding
        from oracle import reference              # checker only, untimed
        B, H, d = 1, max(2, world), 128          # at least one head per rank for the head-exchange path
        Ls = 256
        L = Ls * world
        g = torch.Generator(device="cpu").manual_seed(77)        # same full tensors on every rank
        Qf, Kf, Vf = ((torch.rand((B, H, L, d), generator=g) * 2 - 1).bfloat16() for _ in range(3))
        f64 = lambda x: x.float().numpy().reshape(B * H, L, d)
        errs, tols = {}, {}
        for causal in (False, True):
            ref = np.stack([reference.naive_attention_ex_f64(f64(Qf)[i], f64(Kf)[i], f64(Vf)[i], causal=causal)[0] for i in range(B * H)])
            if causal:
                mine = [sharding.zigzag_shard(x, rank, world).cuda().contiguous() for x in (Qf, Kf, Vf)]
                ref_mine = sharding.zigzag_shard(torch.from_numpy(ref), rank, world).numpy()
            else:
                mine = [x[:, :, rank * Ls:(rank + 1) * Ls].cuda().contiguous() for x in (Qf, Kf, Vf)]
                ref_mine = ref[:, rank * Ls:(rank + 1) * Ls]
            for transport in ("nccl", "peer"):
                try:
                    O = sharding.ring_attention(*mine, transport=transport, causal=causal)
                    torch.cuda.synchronize()
                    err = float(np.abs(O.float().cpu().numpy().reshape(B * H, -1, d) - ref_mine).max())
                except Exception as e:  # noqa: BLE001
                    err = float("nan")
                    res[f"error_{transport}_{'causal' if causal else 'dense'}"] = str(e)[:200]
                errs[f"{transport}_{'causal' if causal else 'dense'}"] = max_over_ranks(err)
            # same rule as tests/test_parity_gpu.py: 2e-3 on averaged rows; early causal rows are O(1) copies of V rows, whose
            # bf16 storage rounding alone is 2^-9 of their magnitude
            tols["causal" if causal else "dense"] = 2e-3 * max(1.0, 2.0 * float(np.abs(ref).max()))
            # the same rows by head exchange (sharding.alltoall_attention): rank-order row blocks also when causal
            mine_a = [x[:, :, rank * Ls:(rank + 1) * Ls].cuda().contiguous() for x in (Qf, Kf, Vf)]
            try:
                O = sharding.alltoall_attention(*mine_a, causal=causal)
                torch.cuda.synchronize()
                err = float(np.abs(O.float().cpu().numpy().reshape(B * H, -1, d) - ref[:, rank * Ls:(rank + 1) * Ls]).max())
            except Exception as e:  # noqa: BLE001
                err = float("nan")
                res[f"error_alltoall_{'causal' if causal else 'dense'}"] = str(e)[:200]
            errs[f"alltoall_{'causal' if causal else 'dense'}"] = max_over_ranks(err)
        res["max_abs_err"] = errs
        res["tolerance"] = tols
        res["parity_ok"] = all(e <= tols["causal" if k.endswith("causal") else "dense"] for k, e in errs.items())
        res["max_abs_err_note"] = f"B{B} H{H} L{L} d{d} bf16, every rank's rows vs the float64 oracle, max over ranks"
        # gather_heads: head-sharded outputs assembled over NCCL equal the single-GPU output
        Q8, K8, V8 = (x.repeat(1, 3, 1, 1).contiguous() for x in (Qf, Kf, Vf))
        local = [sharding.shard_heads(x, rank, world).cuda().contiguous() for x in (Q8, K8, V8)]
        O_local = ops.flash_attention_v1(*local)
        O_all = sharding.gather_heads(O_local, 3 * H * B)
        O_one = ops.flash_attention_v1(Q8.cuda(), K8.cuda(), V8.cuda())
        res["gather_heads_bit_equal"] = bool(torch.equal(O_all.reshape(O_one.shape), O_one))
    except Exception as e:  # noqa: BLE001
        res["parity_error"] = str(e)[:300]
    try:
        from exploring_flash_attention_b200 import sharding
        B, H, L, d = 8, 32, 16384, 128
        Ls = L // world
        g = torch.Generator(device="cuda").manual_seed(99 + rank)
        q, k, v = ((torch.rand((B, H, Ls, d), generator=g, device="cuda") * 2 - 1).bfloat16() for _ in range(3))
        timings = {}
        for causal in (False, True):
            for _ in range(2):
                sharding.ring_attention(q, k, v, transport="peer", causal=causal)
            barrier()
            n = 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                sharding.ring_attention(q, k, v, transport="peer", causal=causal)
            e1.record()
            e1.synchronize()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1) / n)
            fl = flops(B, H, L, d) * (0.5 if causal else 1.0)
            timings["causal" if causal else "dense"] = {"ms": round(ms, 3), "tflops": round(fl / (ms * 1e-3) / 1e12, 1)}
        res["c4_sequence_sharded_peer_transport"] = timings
        a2a = {}
        for causal in (False, True):
            for _ in range(2):
                sharding.alltoall_attention(q, k, v, causal=causal)
            barrier()
            n = 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                sharding.alltoall_attention(q, k, v, causal=causal)
            e1.record()
            e1.synchronize()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1) / n)
            fl = flops(B, H, L, d) * (0.5 if causal else 1.0)
            a2a["causal" if causal else "dense"] = {"ms": round(ms, 3), "tflops": round(fl / (ms * 1e-3) / 1e12, 1)}
        res["c4_sequence_sharded_alltoall"] = a2a
    except Exception as e:  # noqa: BLE001
        res["ring_c4_error"] = str(e)[:300]
    return res


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture (profiles/)
NCU_TRAFFIC_BYTES: dict = {
    "c2": {"bytes": 243_479_296, "source": "profiles/r2_fwd_c2_full.txt (201.39 MB read + 42.09 MB written; tensor pipe 59.6 %)"},
    "c1": {"bytes": 113_222_400, "source": "profiles/r2_fwd_c1_full.txt (100.71 MB read + 12.51 MB written)"},
    "c5": {"bytes": 2_130_111_648, "source": "profiles/r2_tiled_d_pair_c5_full.txt (1610.66 MB read + 519.45 MB written)"},
}


def side_measurements(torch, ops, pk):
    """Other BASELINE.json configs, reported beside the headline (not the bench value): C5 d=512, C1 tf32, C3 split-KV +
    combine with the combine kernel's HBM roofline.  (C4 is c4_strong, run at every N.)"""
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

    def graph_timed(fn, n=20, reps=5):
        """Device time per call without the per-call host launch path (Python -> ctypes -> cudaLaunch is 8-20 us, as long
        as the C3 kernels themselves): n calls captured in one CUDA graph, best of `reps` replays."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(n):
                fn()
        best = float("inf")
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / n)
        return best

    g = torch.Generator(device="cpu").manual_seed(7)
    mk = lambda B, H, L, d, dtype: tuple((torch.rand((B, H, L, d), generator=g) * 2 - 1).to(dtype).cuda() for _ in range(3))
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
        # backward on the headline shape and on a long sequence: dQ, dK, dV from (Q, K, V, O, LSE, dO)
        import numpy as np
        from oracle import reference              # checker only, untimed
        bw = {}
        for (B, H, L, d), tag in (((32, 8, 1024, 128), "c2_shape_B32_H8_L1024_d128"), ((4, 16, 8192, 128), "B4_H16_L8192_d128")):
            q, k, v = mk(B, H, L, d, torch.bfloat16) if L <= 1024 else tuple(
                (torch.rand((B, H, L, d), device="cuda") * 2 - 1).bfloat16() for _ in range(3))
            do = (torch.rand((B, H, L, d), device="cuda") * 2 - 1).bfloat16()
            o, lse = ops.flash_attention_v1_ex(q, k, v, return_lse=True)
            ws = torch.empty(ops.backward_workspace_bytes(B, H, L), dtype=torch.uint8, device="cuda")
            ms = timed(lambda: ops.flash_attention_backward(q, k, v, o, do, lse, workspace=ws), 10, warm=2)
            fl = 2.5 * flops(B, H, L, d)
            dq, dk, dv = ops.flash_attention_backward(q, k, v, o, do, lse, workspace=ws, sync=True)
            h = B * H // 2
            f = lambda x: x.reshape(B * H, L, d)[h].float().cpu().numpy()
            rq, rk, rv = reference.attention_backward_f64(f(q), f(k), f(v), f(do))
            rel = max(float(np.abs(f(g_) - r_).max() / np.abs(r_).max()) for g_, r_ in ((dq, rq), (dk, rk), (dv, rv)))
            bw[tag] = {"ms": round(ms, 4), "tflops_(2.5x_fwd_flops)": round(fl / (ms * 1e-3) / 1e12, 1),
                       "frac_of_measured_bf16_peak": round(fl / (ms * 1e-3) / 1e12 / pk["bf16_tflops"], 4),
                       "max_err_rel_to_max_grad": rel, "tolerance": 5e-3}
            del q, k, v, do, o, lse, dq, dk, dv
        bw["note"] = ("three kernels per call (row statistics, dK/dV pass, dQ pass; S and dP recomputed in both passes: 7 GEMMs "
                      "issued for the 5 counted); error = one sampled head vs the float64 gradient oracle")
        res["backward_bf16"] = bw
    except Exception as e:  # noqa: BLE001
        res["backward_error"] = str(e)[:200]
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
        ms_all_dev = graph_timed(lambda: ops.flash_attention_v2(q, k, v, kvs, O=o, workspace=ws))
        ms_split_dev = graph_timed(lambda: ops.flash_attention_v2_splitkv(q, k, v, kvs, *ws))
        ms_comb_dev = graph_timed(lambda: ops.flash_attention_v2_combine(ws[0], ws[1], torch.bfloat16, (B, H, L, d), o))
        ms_v1_dev = graph_timed(lambda: ops.flash_attention_v1(q, k, v, o))
        res["c3_B32_H8_L256_d64_bf16_4splits"] = {
            "splitkv_plus_combine_ms": round(ms_all, 4),
            "device_time_cuda_graph": {"splitkv_plus_combine_ms": round(ms_all_dev, 4), "splitkv_ms": round(ms_split_dev, 4),
                                       "combine_ms_l2_warm": round(ms_comb_dev, 4), "fused_v1_same_shape_ms": round(ms_v1_dev, 4),
                                       "note": "20 calls replayed from one CUDA graph: no per-call host launch cost "
                                               "(splitkv_plus_combine_ms above is the eager Python loop, host-bound at this size)"},
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
    # --steps / --warmup are honoured as given; each step is a bounded sample of the workload (a number of heads of the
    # same shape, at least one per thread) sized so that the whole run, warm-up included, is ~90 s of CPU work at the
    # ~1 GFLOP/s/thread the reference path reaches
    steps, warm = max(1, args.steps), max(0, args.warmup)
    per_step_s = 90.0 / (steps + warm)
    heads = int(min(B * H, max(threads, per_step_s * threads / max(per_head_gflop, 1e-3))))
    heads = int(os.environ.get("FA_BENCH_CPU_HEADS", heads))       # test hook: shrink the CPU sample
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
        "config": bench_config(args.workload, False, world),
        "cpu_baseline": {"value": round(tf, 5), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "sample_heads": heads},
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
