"""In-tree build of libfa_b200.so (hand-written sm_100a CUDA + the C ABI in include/fa_b200.h).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
# FA_B200_LIB lets a developer point the loader at an alternative build (A/B experiments); default is the in-tree .so
LIB_PATH = Path(os.environ["FA_B200_LIB"]) if os.environ.get("FA_B200_LIB") else CSRC / "libfa_b200.so"
SOURCES = ["fa_api.cu"]
HEADERS = ["sm100_ptx.cuh", "fa_fwd_sm100.cuh", "fa_tiled_d_sm100.cuh", "fa_tiled_d_pair_sm100.cuh", "fa_combine_sm100.cuh",
           "fa_naive_sm100.cuh", "fa_bwd_sm100.cuh", "fa_splitkv_sm100.cuh",
           "../../include/fa_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "--use_fast_math",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libfa_b200.so")


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any((CSRC / f).stat().st_mtime > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: Path | None = None) -> Path:
    """Compile the library if missing or older than its sources. Returns the .so path.
    `defines` / `out` build a tuning variant (e.g. defines=("FA_P_HALVES=1",), out=Path("/tmp/x.so"))."""
    target = Path(out) if out else LIB_PATH
    if not force and not defines and out is None and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", str(target), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return target


CONSUMER_SRC = CSRC.parents[1] / "tests" / "drivers" / "driver_v1.cu"
CONSUMER_BIN = CONSUMER_SRC.with_suffix("")


def build_consumer(force: bool = False) -> Path:
    """Compile tests/drivers/driver_v1.cu — a C++ program that includes include/fa_b200.h and links -lfa_b200, the way a
    reference driver.cu would after INTEGRATION.md §1 (it proves the header compiles as C++ and the ABI links)."""
    build()
    newest = max(CONSUMER_SRC.stat().st_mtime, (CSRC / "../../include/fa_b200.h").stat().st_mtime)
    if not force and CONSUMER_BIN.exists() and CONSUMER_BIN.stat().st_mtime > newest:
        return CONSUMER_BIN
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-I", str(CSRC / "../../include"),
           str(CONSUMER_SRC), "-L", str(CSRC), "-lfa_b200", "-Xlinker", "-rpath", "-Xlinker",
           "$ORIGIN/../../exploring_flash_attention_b200/csrc", "-o", str(CONSUMER_BIN)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return CONSUMER_BIN


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
