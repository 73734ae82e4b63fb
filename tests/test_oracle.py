"""CPU: the oracle (oracle/) is pinned against outputs of the reference itself (tests/golden/reference_golden.npz,
made by tests/golden/make_golden.py) and against the known-answer values recorded in SURVEY.md §4."""
import numpy as np
import pytest
from inputs import MAIN_CASES, SMALL_CASES, qkv

from oracle import cpu, reference, tiled

# naive_attention rows 0-2, cols 0-4 printed by the reference scripts' __main__ blocks
KNOWN = {
    "v1_opt2_main": [[-0.03210377, 0.00682497, -0.05486944, -0.05994513, -0.06322986],
                     [0.01009175, 0.04648276, -0.0552196, -0.06310206, -0.04191145],
                     [-0.02401624, 0.10580362, -0.110481, -0.0438595, -0.12604527]],
    "v1_basic_main": [[0.03394495, 0.03612464, 0.03398394, -0.03650765, 0.0410633],
                      [0.04161668, -0.02311628, 0.01994352, 0.00046414, 0.0409057],
                      [0.01616624, -0.05193618, 0.06023855, 0.04849177, 0.06318308]],
    "tiled_d_main": [[0.01493275, -0.01196932, 0.04407122, 0.03425278, -0.0706882],
                     [-0.04354614, 0.02632476, 0.04128716, -0.0496016, -0.03870687],
                     [-0.06336362, 0.07009426, -0.01577, 0.03467158, -0.0091555]],
    "v2_main": [[0.00875255, -0.01454427, -0.129318, -0.16537592, -0.17424585],
                [0.03296062, -0.04955971, 0.03004273, -0.06485624, -0.21110692],
                [-0.07597745, -0.08185412, -0.19403478, -0.23770748, -0.29803779]],
}


@pytest.mark.parametrize("tag", list(MAIN_CASES))
def test_naive_attention_known_answers(tag, golden):
    L, d, dt = MAIN_CASES[tag]
    Q, K, V = qkv(0, L, d, dt)
    O = np.asarray(reference.naive_attention(Q, K, V), dtype=np.float64)
    np.testing.assert_allclose(O[:3, :5], np.array(KNOWN[tag]), atol=5e-9)       # printed to 8 decimals
    assert np.array_equal(O[:8], golden[f"naive_{tag}_head"])                     # bit-equal to the reference run
    np.testing.assert_allclose([O.sum(), np.abs(O).sum()], golden[f"naive_{tag}_sum"], rtol=1e-12)
    O64 = reference.naive_attention_f64(Q, K, V)
    assert np.array_equal(O64[:8], golden[f"naive_{tag}_f64_head"])
    assert np.array_equal(reference.naive_attention_batched_f64(Q[None, None], K[None, None], V[None, None])[0][:8], O64[:8])


def test_row_subset_matches_full():
    Q, K, V = qkv(5, 300, 32, np.float32)
    full = reference.naive_attention_f64(Q, K, V)
    rows = np.array([0, 7, 128, 299])
    sub = reference.naive_attention_batched_f64(Q[None], K[None], V[None], rows=rows)[0]
    np.testing.assert_allclose(sub, full[rows], rtol=0, atol=1e-15)


@pytest.mark.parametrize("dt_name,dt,tol", [("f64", np.float64, 1e-13), ("f16", np.float16, 4e-3)])
def test_tiled_restatements_match_reference_outputs(dt_name, dt, tol, golden):
    """float64: the restatement differs from the reference only in summation order; fp16: same recurrence with fp16
    state, rounding order differs (the reference sums element by element)."""
    seed, L, d = SMALL_CASES["v1"]
    Q, K, V = qkv(seed, L, d, dt)
    O = np.zeros(L * d, dtype=dt)
    tiled.flash_attention_tiled(Q.flatten(), K.flatten(), V.flatten(), O, L, d, Bq=8, Bk=8)
    assert np.abs(O.reshape(L, d).astype(np.float64) - golden[f"v1_opt2_{dt_name}_O"]).max() <= tol
    assert np.abs(O.reshape(L, d).astype(np.float64) - golden[f"v1_basic_{dt_name}_O"]).max() <= tol

    seed, L, d = SMALL_CASES["td"]
    Q, K, V = qkv(seed, L, d, dt)
    O = np.zeros(L * d, dtype=dt)
    tiled.flash_attention_tiled_d(Q.flatten(), K.flatten(), V.flatten(), O, L, d, 8, 8, 16, 16)
    assert np.abs(O.reshape(L, d).astype(np.float64) - golden[f"td_gpu_{dt_name}_O"]).max() <= tol
    Og = tiled.flash_attention_tiled_global(Q, K, V, 8, 8, 16, 16)
    assert np.abs(Og.astype(np.float64) - golden[f"td_basic_{dt_name}_O"]).max() <= tol

    seed, L, d = SMALL_CASES["v2"]
    Q, K, V = qkv(seed, L, d, dt)
    O = np.zeros(L * d, dtype=dt)
    wO, wm, wl = {}, {}, {}
    tiled.flash_attention_tiled_v2(Q.flatten(), K.flatten(), V.flatten(), O, wO, wm, wl, L, d, 8, 8, 16, 16, 4)
    assert np.abs(O.reshape(L, d).astype(np.float64) - golden[f"v2_{dt_name}_O"]).max() <= tol
    assert sorted(wO) == [(q, k) for q in range(7) for k in range(2)]
    for key in [(0, 0), (0, 1), (6, 1)]:
        for name, ws in (("wsO", wO), ("wsm", wm), ("wsl", wl)):
            ref = golden[f"v2_{dt_name}_{name}_{key[0]}_{key[1]}"].astype(np.float64)
            got = ws[key].astype(np.float64)
            n = 4 if key[0] == 6 else 8  # the last q tile holds 52 - 48 = 4 live rows; the rest is padding
            live = n * d if name == "wsO" else n
            assert np.abs(got[:live] - ref[:live]).max() <= tol * max(1.0, np.abs(ref[:live]).max())


def test_v2_main_parity_figure(golden):
    """README.md:76 quotes max-abs 0.0011 for the V2 simulation at L=256, d=128, fp16, KVTPB=4 (reproduced 0.0011728)."""
    assert abs(float(golden["v2_main_maxabs"][0]) - 0.0011728) < 1e-6
    L, d, dt = MAIN_CASES["v2_main"]
    Q, K, V = qkv(0, L, d, dt)
    O = np.zeros(L * d, dtype=dt)
    tiled.flash_attention_tiled_v2(Q.flatten(), K.flatten(), V.flatten(), O, {}, {}, {}, L, d, 8, 8, 16, 16, 4)
    ref64 = reference.naive_attention_f64(Q, K, V)
    assert np.abs(O.reshape(L, d).astype(np.float64) - ref64).max() < 4e-3          # same error class as the reference
    assert np.abs(O.reshape(L, d).astype(np.float64) - golden["v2_main_O"]).max() < 4e-3


@pytest.mark.parametrize("dt_name", ["f16", "f64"])
def test_c_restatement_bit_equal_to_reference_cpp(dt_name, golden):
    Q, K, V = (golden[f"std_{dt_name}_{n}"] for n in "QKV")
    O = cpu.standard_attention_cpu(Q, K, V, n_threads=2)
    assert np.array_equal(O, golden[f"std_{dt_name}_O"])
    if cpu.have_ref():  # prebuilt oracle/_ref travels to the GPU box; in the build container it is rebuilt from source
        assert np.array_equal(cpu.ref_standard_attention_cpu(Q, K, V, n_threads=2), golden[f"std_{dt_name}_O"])
    ref64 = reference.naive_attention_batched_f64(Q, K, V).reshape(Q.shape)
    assert np.abs(O.astype(np.float64) - ref64).max() < (2e-4 if dt_name == "f16" else 2e-6)


def test_c_restatement_storage_types_and_head_range():
    import torch
    rng = np.random.default_rng(9)
    Q, K, V = (rng.uniform(-1, 1, (1, 3, 40, 32)).astype(np.float32) for _ in range(3))
    ref64 = reference.naive_attention_batched_f64(Q, K, V).reshape(Q.shape)
    O32 = cpu.standard_attention_cpu(Q, K, V)
    assert np.abs(O32 - ref64).max() < 2e-6
    to_bf = lambda x: torch.from_numpy(x).bfloat16().view(torch.uint16).numpy()
    from_bf = lambda x: torch.from_numpy(x).view(torch.bfloat16).float().numpy()
    Ob = from_bf(cpu.standard_attention_cpu(to_bf(Q), to_bf(K), to_bf(V), dtype_name="bfloat16"))
    refb = reference.naive_attention_batched_f64(from_bf(to_bf(Q)), from_bf(to_bf(K)), from_bf(to_bf(V))).reshape(Q.shape)
    assert np.abs(Ob - refb).max() < 2e-3
    part = cpu.standard_attention_cpu(Q, K, V, head_begin=1, head_end=2)
    assert np.array_equal(part[0, 1], O32[0, 1]) and not part[0, 0].any() and not part[0, 2].any()


@pytest.mark.parametrize("d,known", [(32, [-0.00755692, 0.00990295, -0.0185699]), (128, [0.020874, -0.0102615, 0.00642014])])
def test_driver_inputs_and_cpu_reference_known_answers(d, known):
    """the reference's CPU output O[0..2] for its drivers' own data (srand(42), U[-1,1], fp16,
    B32 H8 L1024; flash_attention_v1/CUDA/driver.cu:137-177 and the tiled-d driver with d=128)."""
    Q, K, V = cpu.driver_inputs(32, 8, 1024, d)
    O = cpu.standard_attention_cpu(Q, K, V, head_begin=0, head_end=1)
    np.testing.assert_allclose(O[0, 0, 0, :3].astype(np.float64), known, rtol=0, atol=6e-8)


def test_extended_oracle_reduces_to_reference_and_masks():
    Q, K, V = qkv(7, 50, 16, np.float64)
    O, lse = reference.naive_attention_ex_f64(Q, K, V, causal=False)
    np.testing.assert_allclose(O, reference.naive_attention(Q, K, V), atol=1e-14)
    s = (Q @ K.T) / 4.0
    np.testing.assert_allclose(lse, np.log(np.exp(s).sum(1)), atol=1e-12)
    Oc, lsec = reference.naive_attention_ex_f64(Q, K, V, causal=True)
    np.testing.assert_allclose(Oc[0], V[0], atol=1e-15)                       # row 0 sees key 0 only
    np.testing.assert_allclose(Oc[-1], O[-1], atol=1e-14)                     # last row sees everything
    np.testing.assert_allclose(lsec[3], np.log(np.exp(s[3, :4]).sum()), atol=1e-12)


def test_backward_oracle_matches_finite_differences():
    """attention_backward_f64 (the backward's oracle; the reference has no backward) against central differences of the
    forward oracle, dense and causal, on a loss sum(O * W)."""
    from oracle import reference
    rng = np.random.default_rng(5)
    L, d = 12, 8
    Q, K, V, W = (rng.standard_normal((L, d)) for _ in range(4))
    for causal in (False, True):
        dQ, dK, dV = reference.attention_backward_f64(Q, K, V, W, causal=causal)
        loss = lambda q, k, v: float((reference.naive_attention_ex_f64(q, k, v, causal=causal)[0] * W).sum())
        eps = 1e-6
        for X, dX, idx in ((Q, dQ, 0), (K, dK, 1), (V, dV, 2)):
            num = np.zeros_like(X)
            for i in range(L):
                for c in range(d):
                    Xp, Xm = X.copy(), X.copy()
                    Xp[i, c] += eps
                    Xm[i, c] -= eps
                    args_p = [Q, K, V]
                    args_m = [Q, K, V]
                    args_p[idx], args_m[idx] = Xp, Xm
                    num[i, c] = (loss(*args_p) - loss(*args_m)) / (2 * eps)
            assert np.abs(num - dX).max() <= 1e-7, (causal, idx, np.abs(num - dX).max())
