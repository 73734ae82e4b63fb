"""A/B timing of tuning variants (dev tool): python tests/gpu_probe/ab.py build_variants/*.so -- CASE [CASE...]"""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
args = sys.argv[1:]
sep = args.index("--") if "--" in args else len(args)
libs, cases = args[:sep], (args[sep + 1:] or ["C2_full", "C4_slice"])
for lib in libs:
    env = dict(os.environ, FA_B200_LIB=str(Path(lib).resolve()))
    r = subprocess.run([sys.executable, str(ROOT / "tests/gpu_probe/first_light.py"), *cases], capture_output=True, text=True, env=env)
    out = []
    for line in r.stdout.splitlines():
        if "{" in line:
            j = json.loads(line[line.index("{"):])
            out.append(f"{j['case']}: {j['tflops']:.0f} TF {j['ms']:.4f} ms err {j['max_abs_err']:.1e}")
        elif "rc=" in line and "rc=0" not in line:
            out.append(line[:160])
    print(Path(lib).name, " | ".join(out), flush=True)
