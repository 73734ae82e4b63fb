"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (imported from /root/reference, and its
common/standard.h compiled into oracle/_ref) on small seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so the vectors are committed; tests/test_oracle.py checks the oracle
restatements (oracle/*.py, oracle/standard_attention.c) against them, and tests/test_parity_gpu.py checks the CUDA
path against the same files.
"""
import importlib.util
import io
import contextlib
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
sys.path.insert(0, str(OUT))


def load(relpath, name):
    spec = importlib.util.spec_from_file_location(name, REF / relpath)
    mod = importlib.util.module_from_spec(spec)
    sys.path.insert(0, str((REF / relpath).parent))
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path.pop(0)
    return mod


from inputs import SMALL_CASES, qkv  # noqa: E402  (tests/golden/inputs.py: the seeded inputs, shared with the tests)


def main():
    ref_common = load("common/reference.py", "ref_common_reference")
    v1_opt2 = load("flash_attention_v1/numpy_gpu_like_opt2.py", "ref_v1_opt2")
    v1_basic = load("flash_attention_v1/numpy_basic.py", "ref_v1_basic")
    td_basic = load("flash_attention_v1_tiled_d/numpy_basic.py", "ref_td_basic")
    td_gpu = load("flash_attention_v1_tiled_d/numpy_gpu_like.py", "ref_td_gpu")
    v2 = load("flash_attention_v2/numpy_gpu_like.py", "ref_v2")
    out = {}

    # --- naive_attention known answers at the reference scripts' own __main__ configs ---
    for tag, (L, d, dt) in {"v1_opt2_main": (1024, 32, np.float64), "v1_basic_main": (2048, 32, np.float16),
                            "tiled_d_main": (2048, 128, np.float16), "v2_main": (256, 128, np.float16)}.items():
        Q, K, V = qkv(0, L, d, dt)
        O = ref_common.naive_attention(Q, K, V)
        out[f"naive_{tag}_head"] = np.asarray(O[:8], dtype=np.float64)        # first 8 rows
        out[f"naive_{tag}_sum"] = np.array([np.asarray(O, np.float64).sum(), np.abs(np.asarray(O, np.float64)).sum()])
        # and the float64-upcast evaluation the parity tests use
        O64 = ref_common.naive_attention(Q.astype(np.float64), K.astype(np.float64), V.astype(np.float64))
        out[f"naive_{tag}_f64_head"] = O64[:8]

    # --- small full-output cases ---
    for dt_name, dt in (("f64", np.float64), ("f16", np.float16)):
        # V1 tile loop, ragged L (100 = 12 tiles of 8 + 4)
        seed, L, d = SMALL_CASES['v1']
        Q, K, V = qkv(seed, L, d, dt)
        O = np.zeros(L * d, dtype=dt)
        v1_opt2.flash_attention_tiled(Q.flatten(), K.flatten(), V.flatten(), O, L, d, Bq=8, Bk=8)
        out[f"v1_opt2_{dt_name}_O"] = O.reshape(L, d)
        out[f"v1_basic_{dt_name}_O"] = v1_basic.flash_attention_tiled(Q, K, V, Bq=8, Bk=8)
        out[f"v1_{dt_name}_naive"] = np.asarray(ref_common.naive_attention(Q, K, V))

        # tiled-d, ragged L (5 tiles of 8 + 4), d = 4 chunks of 16
        seed, L, d = SMALL_CASES['td']
        Q, K, V = qkv(seed, L, d, dt)
        O = np.zeros(L * d, dtype=dt)
        td_gpu.flash_attention_tiled(Q.flatten(), K.flatten(), V.flatten(), O, L, d, Bq=8, Bk=8, d_tile_qk=16,
                                     d_tile_v=16)
        out[f"td_gpu_{dt_name}_O"] = O.reshape(L, d)
        out[f"td_basic_{dt_name}_O"] = td_basic.flash_attention_tiled_global(Q, K, V, Bq=8, Bk=8, d_tile_qk=16,
                                                                             d_tile_v=16)
        out[f"td_{dt_name}_naive"] = np.asarray(ref_common.naive_attention(Q, K, V))

        # V2 split-KV, ragged: L=52 -> 7 kv tiles of 8, KVTPB=4 -> 2 kv blocks
        seed, L, d = SMALL_CASES['v2']
        Q, K, V = qkv(seed, L, d, dt)
        O = np.zeros(L * d, dtype=dt)
        wO, wm, wl = {}, {}, {}
        v2.flash_attention_tiled_v2(Q.flatten(), K.flatten(), V.flatten(), O, wO, wm, wl, L, d, Bq=8, Bk=8,
                                    d_tile_qk=16, d_tile_v=16, kv_tiles_per_block=4)
        out[f"v2_{dt_name}_O"] = O.reshape(L, d)
        for key in [(0, 0), (0, 1), (6, 1)]:
            out[f"v2_{dt_name}_wsO_{key[0]}_{key[1]}"] = wO[key]
            out[f"v2_{dt_name}_wsm_{key[0]}_{key[1]}"] = wm[key]
            out[f"v2_{dt_name}_wsl_{key[0]}_{key[1]}"] = wl[key]
        out[f"v2_{dt_name}_naive"] = np.asarray(ref_common.naive_attention(Q, K, V))

    # --- the V2 script's own parity figure (README.md:76 "0.0011"): L=256, d=128, fp16, KVTPB=4 ---
    L, d = 256, 128
    Q, K, V = qkv(0, L, d, np.float16)
    O = np.zeros(L * d, dtype=np.float16)
    v2.flash_attention_tiled_v2(Q.flatten(), K.flatten(), V.flatten(), O, {}, {}, {}, L, d, Bq=8, Bk=8, d_tile_qk=16,
                                d_tile_v=16, kv_tiles_per_block=4)
    naive = ref_common.naive_attention(Q, K, V)
    out["v2_main_O"] = O.reshape(L, d)
    out["v2_main_maxabs"] = np.array([np.abs(O.reshape(L, d) - naive).max()])

    # --- check_accuracy behaviour (passes / raises) ---
    good = np.ones((4, 4)); bad = good + 0.5
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref_common.check_accuracy(good, good, "x")
        try:
            ref_common.check_accuracy(bad, good)
            raised = False
        except AssertionError as e:
            raised = True
            out["check_accuracy_msg"] = np.array(str(e))
    out["check_accuracy_raised"] = np.array(raised)
    out["check_accuracy_stdout"] = np.array(buf.getvalue())

    # --- the reference's C++ standard_attention_cpu (compiled from common/standard.h into oracle/_ref) ---
    from oracle import cpu
    cpu.build()
    rng = np.random.default_rng(42)
    for dt_name, dt in (("f16", np.float16), ("f64", np.float64)):
        Q, K, V = (rng.uniform(-1, 1, (2, 2, 72, 32)).astype(dt) for _ in range(3))
        out[f"std_{dt_name}_Q"], out[f"std_{dt_name}_K"], out[f"std_{dt_name}_V"] = Q, K, V
        out[f"std_{dt_name}_O"] = cpu.ref_standard_attention_cpu(Q, K, V)

    np.savez_compressed(OUT / "reference_golden.npz", **out)
    print("wrote", OUT / "reference_golden.npz", f"{(OUT / 'reference_golden.npz').stat().st_size / 1024:.0f} KiB,",
          len(out), "arrays; v2_main_maxabs =", out["v2_main_maxabs"][0])


if __name__ == "__main__":
    main()
