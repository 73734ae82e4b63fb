"""CPU: the C-ABI library loads and exports every symbol include/fa_b200.h declares (no compute calls), argument
validation that does not need a device, and the host-side mirror of the reference's Python interface."""
import contextlib
import io
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "fa_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fa_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    from exploring_flash_attention_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 11
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and include/fa_b200.h disagree"
    for n in names:
        assert hasattr(lib, n), f"libfa_b200.so does not export {n}"


def test_pure_host_entry_points():
    from exploring_flash_attention_b200 import _lib
    lib = _lib.load()
    assert lib.fa_v2_num_splits(256, 64) == 4          # C3: BK_ref 16 x KV_TILES_PER_BLOCK 4 = 64 keys per split
    assert lib.fa_v2_num_splits(1000, 96) == 11
    assert lib.fa_v2_num_splits(256, 0) == 0
    # Oaccum 4*256*256*64*4 B + LSE 4*256*256*4 B (both already 256-B multiples)
    assert lib.fa_v2_workspace_bytes(32, 8, 256, 64, 64) == 4 * 256 * 256 * 64 * 4 + 4 * 256 * 256 * 4


def test_argument_validation_without_device():
    """Validation happens before any CUDA call, so these return status codes on a CPU-only box too
    (the reference launchers assert()/abort instead: flash_attention_v1.h:263-264)."""
    from exploring_flash_attention_b200 import _lib
    lib = _lib.load()
    buf = np.zeros(64, dtype=np.float32)
    p = buf.ctypes.data  # 16-byte aligned enough for the check; never dereferenced
    p -= p % 16
    assert lib.fa_v1_forward(p, p, p, p, 0, 1, 8, 32, 0, None) == -1            # FA_ERR_SHAPE
    assert b"positive" in lib.fa_last_error()
    assert lib.fa_v1_forward(p, p, p, p, 1, 1, 8, 32, 7, None) == -2            # FA_ERR_DTYPE
    assert lib.fa_v1_forward(p + 4, p, p, p, 1, 1, 8, 32, 0, None) == -3        # FA_ERR_ALIGN
    assert lib.fa_v1_forward(None, p, p, p, 1, 1, 8, 32, 0, None) == -3
    assert lib.fa_v1_tiled_d_forward(p, p, p, p, 1, 1, 8, 128, 48, 32, 1, None) == -1   # d_tile_qk must divide d
    assert lib.fa_v2_forward(p, p, p, p, 1, 1, 8, 64, 64, 1, None, 0, None) == -6       # FA_ERR_WORKSPACE
    assert lib.fa_v2_combine(p, p, p, 1, 1, 8, 48, 2, 1, None) == -4            # FA_ERR_UNSUPPORTED_D
    with pytest.raises(_lib.FlashAttentionError) as ei:
        _lib.check(lib.fa_v1_forward(p, p, p, p, 1, 1, 8, 32, 9, None))
    assert ei.value.code == -2 and "FA_ERR_DTYPE" in str(ei.value)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    from exploring_flash_attention_b200 import ops
    from exploring_flash_attention_b200.flash_attention_v1 import flash_attention_tiled
    z = np.zeros(8 * 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        flash_attention_tiled(z, z, z, z.copy(), 8, 32)
    t = torch.zeros(1, 1, 8, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.flash_attention_v1(t, t, t)


def test_product_package_never_imports_oracle():
    pkg = ROOT / "exploring_flash_attention_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
    for f in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert "oracle/" not in f.read_text()


def test_check_accuracy_matches_reference_behaviour(golden):
    from exploring_flash_attention_b200.common.reference import check_accuracy, print_comparison
    good = np.ones((4, 4))
    bad = good + 0.5
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        check_accuracy(good, good, "x")
        with pytest.raises(AssertionError) as ei:
            check_accuracy(bad, good)
    assert str(ei.value) == str(golden["check_accuracy_msg"])
    assert buf.getvalue() == str(golden["check_accuracy_stdout"])
    assert bool(golden["check_accuracy_raised"])
    with contextlib.redirect_stdout(io.StringIO()) as out:
        print_comparison(np.arange(20.0).reshape(4, 5), np.zeros((4, 5)), 2, 3)
    assert "Output shape: (4, 5)" in out.getvalue() and "First 2 rows (reference):" in out.getvalue()


def test_head_range_partitions_everything():
    from exploring_flash_attention_b200.sharding import head_range
    for BH in (1, 7, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [head_range(BH, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == BH
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        head_range(8, 2, 2)


def test_cpp_consumer_compiles_and_links_against_the_header():
    """tests/drivers/driver_v1.cu includes include/fa_b200.h from C++ and links -lfa_b200 (run on the GPU by
    tests/test_parity_round2_gpu.py); here: it builds, and its only repo dependency is the library."""
    import subprocess
    from exploring_flash_attention_b200 import _build
    exe = _build.build_consumer()
    assert exe.exists()
    needed = subprocess.run(["readelf", "-d", str(exe)], capture_output=True, text=True).stdout
    assert "libfa_b200.so" in needed


def test_host_entry_point_validates_before_touching_the_device():
    import torch
    from exploring_flash_attention_b200 import _lib, ops
    lib = _lib.load()
    buf = np.zeros(1024, dtype=np.float32)
    p = buf.ctypes.data
    assert lib.fa_forward_host(7, p, p, p, p, 1, 1, 8, 32, 0, 0) == -1           # bad variant: before any cudaMalloc
    assert lib.fa_forward_host(2, p, p, p, p, 1, 1, 8, 32, 0, 0) == -1           # V2 needs kv_per_split
    assert lib.fa_forward_host(0, p, p, None, p, 1, 1, 8, 32, 0, 0) == -3
    assert lib.fa_naive_attention_workspace_bytes(2, 100, 50, 0) == 2 * 100 * 50 * 4
    assert lib.fa_naive_attention_workspace_bytes(2, 100, 50, 3) == 2 * 100 * 50 * 8
    assert lib.fa_naive_attention_workspace_bytes(2, 100, 50, 1) == 0             # bf16 is not a naive-attention type
    assert lib.fa_naive_attention(p, p, p, p, 1, 8, 8, 4, 1, p, 1 << 20, None) == -2
    q = torch.zeros((1, 2, 8, 32))
    for bad in (torch.zeros((1, 2, 4, 32)), torch.zeros((1, 2, 8, 32), dtype=torch.float16), torch.zeros((1, 2, 32, 8)).transpose(2, 3)):
        with pytest.raises(_lib.FlashAttentionError):
            ops.flash_attention_host(q, bad, q)
        with pytest.raises(_lib.FlashAttentionError):
            ops.flash_attention_host(q, q, q, bad)


def test_alltoall_chunk_plan_partitions_the_heads():
    """sharding._a2a_plan: boundaries start at 0, end at the head count, strictly increase, respect the chunk limit; a
    single chunk when splitting cannot help (one head, or a whole number of rounds already)."""
    from exploring_flash_attention_b200.sharding import _a2a_plan
    for hpr in (1, 2, 5, 16, 32, 37):
        for items in (1, 4, 64):
            for max_chunks in (1, 2, 4):
                b = _a2a_plan(hpr, items, max_chunks, 148)
                assert b[0] == 0 and b[-1] == hpr and all(x < y for x, y in zip(b, b[1:]))
                assert len(b) - 1 <= max_chunks
    assert _a2a_plan(1, 64, 4, 148) == (0, 1)
    assert _a2a_plan(32, 64, 1, 148) == (0, 32)
    # 32 heads x 64 items on 148 SMs: one chunk is 14 rounds; the plan may add at most one round of quantisation loss
    b = _a2a_plan(32, 64, 4, 148)
    rounds = sum(-(-(y - x) * 64 // 148) for x, y in zip(b, b[1:]))
    assert rounds <= 15
