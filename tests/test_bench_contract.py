"""CPU: the reference arm of bench.py prints one JSON line with the contract's keys (sample shrunk via the test hook)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_json_line():
    env = dict(os.environ, FA_BENCH_CPU_HEADS="2", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "c1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "attn_fwd_tflops" and j["unit"] == "TFLOP/s"
    assert j["higher_is_better"] is True and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["value"] > 0 and "workload" in j["config"]
    sys.path.insert(0, str(ROOT))
    import bench
    assert j["config"] == bench.bench_config("c1")            # the same object the GPU arm prints


def test_reference_arm_honours_steps_and_warmup():
    env = dict(os.environ, FA_BENCH_CPU_HEADS="1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "2",
                        "--workload", "c1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert j["steps"] == 4 and j["warmup"] == 2


def test_non_rank0_reference_arm_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", FA_BENCH_CPU_HEADS="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
