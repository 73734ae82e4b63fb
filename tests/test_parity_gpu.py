"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.

Tolerances are the ones BASELINE.json's north_star states: max-abs <= 1e-3 for fp32 storage (tf32 products, fp32
accumulate), <= 2e-3 for bf16 (and fp16) storage; inputs U[-1,1) like the reference drivers
(flash_attention_v1/CUDA/driver.cu:71-75) unless noted.  The oracle always sees the already-rounded inputs
up-cast to float64.
"""
import numpy as np
import pytest
import torch
from inputs import MAIN_CASES, SMALL_CASES, qkv

from oracle import cpu, reference

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-3, torch.bfloat16: 2e-3, torch.float16: 2e-3}


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from exploring_flash_attention_b200 import _lib, ops as _ops
    _lib.load()  # raises if libfa_b200.so is missing: there is no fallback to hide behind
    return _ops


def uniform_qkv(B, H, L, d, dtype, seed=42, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return tuple(((torch.rand((B, H, L, d), generator=g) * 2 - 1) * scale).to(dtype).cuda() for _ in range(3))


def oracle_out(Q, K, V, heads=None, rows=None):
    f = lambda x: x.float().cpu().numpy()
    return reference.naive_attention_batched_f64(f(Q), f(K), f(V), heads=heads, rows=rows)


def max_err(O, ref, heads=None, rows=None):
    L, d = O.shape[-2:]
    got = O.float().cpu().numpy().reshape(-1, L, d).astype(np.float64)
    if heads is not None:
        got = got[list(heads)]
    if rows is not None:
        got = got[:, rows]
    return np.abs(got - ref).max()


CASES = [  # B, H, L, d, dtype
    (2, 2, 256, 128, torch.bfloat16), (1, 3, 384, 64, torch.bfloat16), (1, 2, 512, 128, torch.float16),
    (1, 2, 256, 64, torch.float16), (2, 2, 512, 32, torch.float32), (1, 2, 256, 64, torch.float32),
    # ragged / edge: L not a multiple of the 128-key or 256-row tiles, single Q tile, tiny L
    (1, 3, 333, 128, torch.bfloat16), (1, 2, 100, 128, torch.bfloat16), (1, 2, 129, 64, torch.bfloat16),
    (1, 2, 1, 128, torch.bfloat16), (1, 2, 777, 32, torch.float32), (1, 1, 257, 64, torch.float32),
    (3, 1, 8, 64, torch.float16),
    # 16-bit d = 32: 64-byte rows, 64B-swizzle instantiation (the reference V1 driver's own config is fp16 d=32)
    (2, 2, 512, 32, torch.float16), (1, 3, 333, 32, torch.bfloat16), (1, 1, 100, 32, torch.float16),
]


@pytest.mark.parametrize("B,H,L,d,dtype", CASES)
def test_v1_matches_oracle(ops, B, H, L, d, dtype):
    Q, K, V = uniform_qkv(B, H, L, d, dtype)
    O = ops.flash_attention_v1(Q, K, V, sync=True)
    assert not torch.isnan(O).any()
    assert max_err(O, oracle_out(Q, K, V)) <= TOL[dtype]
    O2 = ops.flash_attention_v1_tiled_d(Q, K, V, d_tile_qk=d // 2, d_tile_v=d // 2, sync=True)
    assert torch.equal(O, O2)      # d-chunk hints change scheduling only


@pytest.mark.parametrize("B,H,L,d,dtype", [
    (1, 2, 256, 256, torch.bfloat16), (1, 2, 384, 512, torch.bfloat16), (1, 1, 512, 512, torch.float16),
    (1, 2, 333, 512, torch.bfloat16), (1, 1, 100, 256, torch.float16), (2, 1, 129, 512, torch.bfloat16),
    (1, 1, 1, 512, torch.bfloat16),
    # fp32 storage / tf32 products through the same kernel: the reference tiled-d default D=128 in its USE_FP64 mode
    (2, 2, 512, 128, torch.float32), (1, 2, 333, 128, torch.float32), (1, 2, 300, 256, torch.float32),
    (1, 1, 100, 128, torch.float32),
])
def test_tiled_d_large_head_dims_match_oracle(ops, B, H, L, d, dtype):
    """K2: d = 256 / 512 (TMEM holds one 256-wide O slab per CTA); d_tile hints are validated like the reference."""
    from exploring_flash_attention_b200 import FlashAttentionError
    Q, K, V = uniform_qkv(B, H, L, d, dtype)
    O = ops.flash_attention_v1_tiled_d(Q, K, V, d_tile_qk=32, d_tile_v=32, sync=True)
    assert not torch.isnan(O).any()
    assert max_err(O, oracle_out(Q, K, V)) <= TOL[dtype]
    assert torch.equal(O, ops.flash_attention_v1(Q, K, V, sync=True))          # V1 entry point routes d > 128 here
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_v1_tiled_d(Q, K, V, d_tile_qk=48, d_tile_v=32)      # must divide d


def test_tiled_d_rescale_path(ops):
    B, H, L, d = 1, 1, 640, 512
    Q, K, V = uniform_qkv(B, H, L, d, torch.bfloat16)
    ramp = torch.linspace(0.3, 8.0, L, device="cuda").view(1, 1, L, 1)
    K = (K.float() * ramp).bfloat16()
    O = ops.flash_attention_v1_tiled_d(Q, K, V, sync=True)
    ref = oracle_out(Q, K, V)
    assert not torch.isnan(O).any()
    assert max_err(O, ref) <= 2e-3 * max(1.0, np.abs(ref).max()) * 4


def test_c5_full_size_sampled(ops):
    """BASELINE.json configs[4]: B16 H8 L4096 d512 bf16 — sampled oracle rows + softmax-rows-sum-to-one."""
    B, H, L, d = 16, 8, 4096, 512
    g = torch.Generator(device="cpu").manual_seed(42)
    base = [((torch.rand((1, H, L, d), generator=g) * 2 - 1)).bfloat16().cuda() for _ in range(3)]
    Q, K, V = (x.expand(B, H, L, d).contiguous() for x in base)
    Q[1:] = Q[1:].roll(1, dims=2)                                               # make batches differ
    O = ops.flash_attention_v1_tiled_d(Q, K, V, sync=True)
    heads = [0, 77, B * H - 1]
    rows = np.r_[0:32, L // 2:L // 2 + 32, L - 32:L]
    assert max_err(O, oracle_out(Q, K, V, heads=heads, rows=rows), heads=heads, rows=rows) <= 2e-3
    ones = torch.ones_like(V)
    assert (ops.flash_attention_v1_tiled_d(Q, K, ones, sync=True).float() - 1).abs().max().item() <= 4e-3


@pytest.mark.parametrize("B,H,L,d,dtype,kvs", [
    (4, 8, 256, 64, torch.bfloat16, 64),       # C3 geometry: 4 splits of 64 keys
    (1, 4, 1024, 128, torch.bfloat16, 256), (1, 2, 500, 32, torch.float32, 96), (1, 2, 300, 64, torch.float16, 300),
    (1, 2, 200, 128, torch.bfloat16, 8),       # 25 splits of 8 keys (one reference tile each)
    (2, 2, 384, 32, torch.float16, 64),
])
def test_v2_splitkv_and_combine_match_oracle(ops, B, H, L, d, dtype, kvs):
    Q, K, V = uniform_qkv(B, H, L, d, dtype)
    ref = oracle_out(Q, K, V)
    O = ops.flash_attention_v2(Q, K, V, kvs, sync=True)
    assert max_err(O, ref) <= TOL[dtype]
    # the workspace itself: each split's normalised partial + LSE against the oracle restricted to that key range
    Oacc, LSE = ops.flash_attention_v2_splitkv(Q, K, V, kvs)
    torch.cuda.synchronize()
    S = Oacc.shape[0]
    assert S == -(-L // kvs)
    f = lambda x: x.float().cpu().numpy().reshape(-1, L, d).astype(np.float64)
    q, k, v = f(Q), f(K), f(V)
    for s in (0, S - 1):
        ks = slice(s * kvs, min(L, (s + 1) * kvs))
        sc = np.einsum("hqd,hkd->hqk", q, k[:, ks]) / np.sqrt(d)
        lse = np.log(np.exp(sc - sc.max(-1, keepdims=True)).sum(-1)) + sc.max(-1)
        p = np.exp(sc - lse[..., None])
        part = p @ v[:, ks]
        assert np.abs(Oacc[s].cpu().numpy() - part).max() <= TOL[dtype] * 2
        assert np.abs(LSE[s].cpu().numpy() - lse).max() <= 2e-3
    # combine alone is exact fp32 arithmetic on the workspace: re-derive it on the host in float64
    w = torch.softmax(LSE.double(), dim=0)
    merged = (w[..., None] * Oacc.double()).sum(0).reshape(B, H, L, d)
    O3 = ops.flash_attention_v2_combine(Oacc, LSE, torch.float32, (B, H, L, d))
    assert (O3.double() - merged).abs().max().item() <= 2e-6


def test_lazy_rescale_path_large_scores(ops):
    """Scores with a growing running max force the O-rescale branch (threshold 2^8) on every later KV tile."""
    B, H, L, d = 1, 2, 1024, 128
    Q, K, V = uniform_qkv(B, H, L, d, torch.bfloat16, scale=1.0)
    ramp = torch.linspace(0.5, 14.0, L, device="cuda").view(1, 1, L, 1)
    K = (K.float() * ramp).bfloat16()
    Q = (Q.float() * 3).bfloat16()
    O = ops.flash_attention_v1(Q, K, V, sync=True)
    ref = oracle_out(Q, K, V)
    assert not torch.isnan(O).any()
    # peaked softmax: outputs are O(1), so bf16 output rounding alone is 2^-9; compare relative to magnitude
    assert max_err(O, ref) <= 2e-3 * max(1.0, np.abs(ref).max()) * 4


@pytest.mark.parametrize("name", ["v1", "td", "v2"])
def test_against_reference_golden_outputs(ops, golden, name):
    """Same inputs the reference itself was run on (tests/golden/make_golden.py): N(0,1), one head, ragged L.
    fp16 storage vs the reference's fp16 run; fp32(tf32) vs its float64 run.  N(0,1) inputs give O(1) outputs, so the
    fp16 comparison carries the reference's own fp16 rounding of S/P/O (its error vs float64 is ~1e-3, README.md:76)."""
    from exploring_flash_attention_b200 import flash_attention_v1, flash_attention_v1_tiled_d, flash_attention_v2
    seed, L, d = SMALL_CASES[name]
    for dt_name, dt, tol in (("f16", np.float16, 6e-3), ("f64", np.float64, 4e-3)):
        Q, K, V = qkv(seed, L, d, dt)
        O = np.zeros(L * d, dtype=dt)
        if name == "v1":
            flash_attention_v1.flash_attention_tiled(Q.flatten(), K.flatten(), V.flatten(), O, L, d, Bq=8, Bk=8)
            ref = golden[f"v1_opt2_{dt_name}_O"]
        elif name == "td":
            flash_attention_v1_tiled_d.flash_attention_tiled(Q.flatten(), K.flatten(), V.flatten(), O, L, d, 8, 8, 16, 16)
            ref = golden[f"td_gpu_{dt_name}_O"]
        else:
            wO, wm, wl = {}, {}, {}
            flash_attention_v2.flash_attention_tiled_v2(Q.flatten(), K.flatten(), V.flatten(), O, wO, wm, wl, L, d, 8, 8,
                                                        16, 16, 4)
            ref = golden[f"v2_{dt_name}_O"]
            assert sorted(wO) == [(q, k) for q in range(7) for k in range(2)]
            # the reference's own merge formula applied to our workspace triple reproduces the output
            from oracle import tiled
            O_chk = np.zeros(L * d, dtype=np.float64)
            for qt in range(7):
                tiled.reduction_kernel(wO, wm, wl, O_chk, qt, 2, L, d, 8)
            assert np.abs(O_chk.reshape(L, d) - O.reshape(L, d).astype(np.float64)).max() <= 2e-3
        assert np.abs(O.reshape(L, d).astype(np.float64) - ref.astype(np.float64)).max() <= tol
        naive = golden[f"{name}_{dt_name}_naive"].astype(np.float64)
        assert np.abs(O.reshape(L, d).astype(np.float64) - naive).max() <= tol


def test_reference_script_main_config_v2(ops, golden):
    """flash_attention_v2/numpy_gpu_like.py __main__: L=256, d=128, fp16, KVTPB=4 -> the reference reports 0.0011."""
    from exploring_flash_attention_b200.common.reference import check_accuracy, naive_attention
    from exploring_flash_attention_b200.flash_attention_v2 import flash_attention_tiled_v2
    L, d, dt = MAIN_CASES["v2_main"]
    Q, K, V = qkv(0, L, d, dt)
    O = np.zeros(L * d, dtype=dt)
    flash_attention_tiled_v2(Q.flatten(), K.flatten(), V.flatten(), O, {}, {}, {}, L, d, 8, 8, 16, 16, 4)
    ref64 = reference.naive_attention_f64(Q, K, V)
    err = np.abs(O.reshape(L, d).astype(np.float64) - ref64).max()
    assert err <= 2e-3, err                                  # the reference's own simulation sits at 1.17e-3
    check_accuracy(O.reshape(L, d).astype(np.float64), ref64, "B200 V2")      # reference tolerances, must not raise
    assert np.abs(naive_attention(Q, K, V).astype(np.float64) - ref64).max() <= 2e-3


def test_against_c_oracle_fp16_driver_inputs(ops):
    """What the reference drivers check: GPU vs standard_attention_cpu on U[-1,1] fp16 data, PASS if max-abs < 1e-3
    (flash_attention_v1/CUDA/driver.cu:275)."""
    B, H, L, d = 2, 2, 256, 64
    Q, K, V = uniform_qkv(B, H, L, d, torch.float16)
    O = ops.flash_attention_v1(Q, K, V, sync=True)
    f = lambda x: x.cpu().numpy()
    ref = cpu.standard_attention_cpu(f(Q), f(K), f(V))
    assert np.abs(O.cpu().numpy().astype(np.float32) - ref.astype(np.float32)).max() < 1e-3


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties + sampled oracle rows
# ---------------------------------------------------------------------------------------------------------------
FULL = {"C1": (32, 8, 1024, 32, torch.float32), "C2": (32, 8, 1024, 128, torch.bfloat16),
        "C4_slice": (1, 8, 16384, 128, torch.bfloat16)}


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_sampled_rows_and_properties(ops, name):
    B, H, L, d, dtype = FULL[name]
    tol = TOL[dtype]
    Q, K, V = uniform_qkv(B, H, L, d, dtype)
    O = ops.flash_attention_v1(Q, K, V, sync=True)
    assert not torch.isnan(O).any()
    heads = [0, B * H // 2, B * H - 1]
    rows = np.r_[0:64, L // 2 - 32:L // 2 + 32, L - 64:L]
    assert max_err(O, oracle_out(Q, K, V, heads=heads, rows=rows), heads=heads, rows=rows) <= tol
    # (1) softmax rows sum to one: V = 1 -> O = 1
    ones = torch.ones_like(V)
    assert (ops.flash_attention_v1(Q, K, ones, sync=True).float() - 1).abs().max().item() <= 4e-3
    # (2) permuting the keys (K and V rows together) leaves O unchanged
    perm = torch.randperm(L, generator=torch.Generator().manual_seed(1)).cuda()
    Op = ops.flash_attention_v1(Q, K[:, :, perm].contiguous(), V[:, :, perm].contiguous(), sync=True)
    assert (Op.float() - O.float()).abs().max().item() <= tol
    # (3) linearity in V
    V2 = torch.roll(V, 1, dims=2)
    Osum = ops.flash_attention_v1(Q, K, (0.5 * V.float() + 0.25 * V2.float()).to(dtype), sync=True)
    lin = 0.5 * O.float() + 0.25 * ops.flash_attention_v1(Q, K, V2, sync=True).float()
    assert (Osum.float() - lin).abs().max().item() <= 2 * tol
    # (4) split-KV + combine agrees with the fused path (V2 vs V1)
    Ov2 = ops.flash_attention_v2(Q, K, V, max(64, L // 4), sync=True)
    assert (Ov2.float() - O.float()).abs().max().item() <= tol


def test_c3_full_v2(ops):
    B, H, L, d = 32, 8, 256, 64
    Q, K, V = uniform_qkv(B, H, L, d, torch.bfloat16)
    O = ops.flash_attention_v2(Q, K, V, 64, sync=True)
    assert max_err(O, oracle_out(Q, K, V, heads=range(0, 256, 37)), heads=range(0, 256, 37)) <= 2e-3


def test_host_buffer_entry_point(ops):
    B, H, L, d = 2, 4, 512, 128
    g = torch.Generator().manual_seed(3)
    Qh, Kh, Vh = ((torch.rand((B, H, L, d), generator=g) * 2 - 1).bfloat16().pin_memory() for _ in range(3))
    Oh = ops.flash_attention_host(Qh, Kh, Vh, variant=0)
    Od = ops.flash_attention_v1(Qh.cuda(), Kh.cuda(), Vh.cuda(), sync=True)
    assert torch.equal(Oh, Od.cpu())
    Oh2 = ops.flash_attention_host(Qh, Kh, Vh, variant=2, kv_per_split=128)
    assert (Oh2.float() - Oh.float()).abs().max().item() <= 2e-3


def test_error_behaviour_on_device(ops):
    from exploring_flash_attention_b200 import FlashAttentionError
    Q, K, V = uniform_qkv(1, 1, 64, 48, torch.bfloat16)
    with pytest.raises(FlashAttentionError) as ei:
        ops.flash_attention_v1(Q, K, V)
    assert ei.value.code == -4
    Q, K, V = uniform_qkv(1, 1, 64, 64, torch.bfloat16)
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_v1(Q, K[:, :, :32], V)


@pytest.mark.parametrize("which", ["v1", "tiled_d", "v2"])
def test_reference_driver_mirrors_pass(ops, which, capsys):
    """The three reference drivers' flows (full B32 H8 L1024 shapes on the GPU; CPU reference bounded to 8 heads)."""
    from tests.drivers import driver
    if which == "v1":
        pytest.skip("fp16 d=32 needs the 64-byte-swizzle instantiation") if not _fp16_d32_supported(ops) else None
    rc = driver.main([which, "--heads", "8"])
    out = capsys.readouterr().out
    assert rc == 0 and "Test PASSED" in out, out[-800:]


def _fp16_d32_supported(ops):
    from exploring_flash_attention_b200 import FlashAttentionError
    try:
        Q, K, V = uniform_qkv(1, 1, 128, 32, torch.float16)
        ops.flash_attention_v1(Q, K, V, sync=True)
        return True
    except FlashAttentionError:
        return False


@pytest.mark.parametrize("B,H,L,d,dtype", [
    (1, 2, 512, 128, torch.bfloat16), (2, 2, 333, 64, torch.bfloat16), (1, 3, 1000, 128, torch.float16),
    (1, 2, 700, 32, torch.float32), (1, 2, 129, 64, torch.float32), (1, 2, 100, 32, torch.float16), (1, 1, 1, 128, torch.bfloat16),
    (1, 5, 2048, 128, torch.bfloat16),
])
def test_causal_and_lse_match_extended_oracle(ops, B, H, L, d, dtype):
    """causal mask + LSE output of the fused-tile kernel (fa_v1_forward_ex)."""
    Q, K, V = uniform_qkv(B, H, L, d, dtype)
    f = lambda x: x.float().cpu().numpy().reshape(-1, L, d)
    q, k, v = f(Q), f(K), f(V)
    for causal in (False, True):
        O, lse = ops.flash_attention_v1_ex(Q, K, V, causal=causal, return_lse=True, sync=True)
        assert not torch.isnan(O).any() and not torch.isnan(lse).any()
        for h in range(B * H):
            ref_o, ref_lse = reference.naive_attention_ex_f64(q[h], k[h], v[h], causal=causal)
            got = O.float().cpu().numpy().reshape(-1, L, d)[h].astype(np.float64)
            scale = max(1.0, np.abs(ref_o).max() * 2)          # early causal rows are O(1): allow their storage rounding
            assert np.abs(got - ref_o).max() <= TOL[dtype] * scale
            assert np.abs(lse.cpu().numpy().reshape(-1, L)[h] - ref_lse).max() <= 2e-3
        if not causal:
            assert torch.equal(O, ops.flash_attention_v1(Q, K, V, sync=True))   # same kernel, extras off
    O = ops.flash_attention_v1_ex(Q, K, V, causal=True, sync=True)
    # row 0 attends to key 0 only: O[0] = V[0] (tf32 truncates V's mantissa to 10 bits, 16-bit storage is exact)
    assert (O[:, :, 0].float() - V[:, :, 0].float()).abs().max().item() <= (1e-3 if dtype == torch.float32 else 1e-6)


def test_unsupported_head_dims_are_refused_everywhere(ops):
    """d outside {32,64,128,256,512} (and fp32 d = 512) has no kernel: every entry point says so, none falls back."""
    from exploring_flash_attention_b200 import FlashAttentionError
    for d, dtype in ((48, torch.bfloat16), (1024, torch.bfloat16), (512, torch.float32)):
        Q, K, V = uniform_qkv(1, 1, 128, d, dtype)
        for call in (lambda: ops.flash_attention_v1_ex(Q, K, V, causal=True), lambda: ops.flash_attention_varlen(Q, K, V),
                     lambda: ops.flash_attention_partial(Q, K, V), lambda: ops.flash_attention_v2(Q, K, V, 64)):
            with pytest.raises(FlashAttentionError) as ei:
                call()
            assert ei.value.code == -4


def test_v2_reference_signature_kernels_driver_loop(ops, golden):
    """The reference's two-kernel driver loop (flash_attention_v2/numpy_gpu_like.py:378-402) written against our
    reference-signature partial_attention_kernel / reduction_kernel, on the golden V2 case (L=52 ragged, d=64, fp16)."""
    from exploring_flash_attention_b200.flash_attention_v2 import partial_attention_kernel, reduction_kernel
    seed, L, d = SMALL_CASES["v2"]
    Bq = Bk = 8
    tiles = 4
    Q, K, V = qkv(seed, L, d, np.float16)
    wO, wm, wl = {}, {}, {}
    nq, nkt = -(-L // Bq), -(-L // Bk)
    nkb = -(-nkt // tiles)
    for qt in range(nq):
        for kb in range(nkb):
            partial_attention_kernel(Q.flatten(), K.flatten(), V.flatten(), wO, wm, wl, qt, kb, L, d, Bq, Bk, 16, 16,
                                     kb * tiles, min(kb * tiles + tiles, nkt))
    O = np.zeros(L * d, dtype=np.float16)
    for qt in range(nq):
        reduction_kernel(wO, wm, wl, O, qt, nkb, L, d, Bq)
    ref64 = reference.naive_attention_f64(Q, K, V)
    assert np.abs(O.reshape(L, d).astype(np.float64) - ref64).max() <= 4e-3
    assert np.abs(O.reshape(L, d).astype(np.float64) - golden["v2_f16_O"].astype(np.float64)).max() <= 6e-3
    # the GPU reduction also merges the REFERENCE's own (un-normalised O, m, l) workspace triples
    from oracle import tiled
    rO, rm, rl = {}, {}, {}
    Od = np.zeros(L * d, dtype=np.float64)
    Q64, K64, V64 = (x.astype(np.float64) for x in (Q, K, V))
    tiled.flash_attention_tiled_v2(Q64.flatten(), K64.flatten(), V64.flatten(), Od, rO, rm, rl, L, d, Bq, Bk, 16, 16, tiles)
    O2 = np.zeros(L * d, dtype=np.float64)
    for qt in range(nq):
        reduction_kernel(rO, rm, rl, O2, qt, nkb, L, d, Bq)
    assert np.abs(O2 - Od).max() <= 1e-5


@pytest.mark.parametrize("B,H,Lq,Lk,d,dtype,causal", [
    (3, 2, 512, 512, 128, torch.bfloat16, False), (3, 2, 512, 512, 128, torch.bfloat16, True),
    (2, 3, 300, 777, 64, torch.bfloat16, False), (4, 1, 1000, 130, 32, torch.float16, False),
    (2, 2, 129, 640, 32, torch.float32, False), (3, 2, 700, 700, 64, torch.float32, True),
    (2, 2, 64, 2048, 128, torch.float16, False),
])
def test_key_padding_and_rectangular_match_extended_oracle(ops, B, H, Lq, Lk, d, dtype, causal):
    """SURVEY.md §8(f)-1 "dynamic sequence lengths": per-batch key-padding lengths and Lq != Lk (fa_v1_forward_varlen)."""
    g = torch.Generator().manual_seed(11)
    Q = ((torch.rand((B, H, Lq, d), generator=g) * 2 - 1).to(dtype)).cuda()
    K, V = (((torch.rand((B, H, Lk, d), generator=g) * 2 - 1).to(dtype)).cuda() for _ in range(2))
    lens = [Lk, 1, max(1, Lk // 2 + 3), 128][:B]
    kv_lens = torch.tensor(lens, dtype=torch.int32, device="cuda")
    for use_lens in (False, True):
        O, lse = ops.flash_attention_varlen(Q, K, V, kv_lens if use_lens else None, causal=causal, return_lse=True, sync=True)
        assert not torch.isnan(O).any() and not torch.isnan(lse).any()
        for b in range(B):
            for h in range(H):
                f = lambda x: x[b, h].float().cpu().numpy()
                ref_o, ref_lse = reference.naive_attention_ex_f64(f(Q), f(K), f(V), causal=causal,
                                                                  kv_len=lens[b] if use_lens else None)
                scale = max(1.0, np.abs(ref_o).max() * 2)
                assert np.abs(O[b, h].float().cpu().numpy() - ref_o).max() <= TOL[dtype] * scale
                assert np.abs(lse[b, h].cpu().numpy() - ref_lse).max() <= 2e-3
    if Lq == Lk:   # no mask, square: bit-identical to the plain entry point
        assert torch.equal(ops.flash_attention_varlen(Q, K, V, sync=True), ops.flash_attention_v1_ex(Q, K, V, sync=True))
    if not causal:  # out-of-range lengths are clamped to [1, Lk]
        wild = torch.tensor([10 ** 6, -5, 0, Lk + 1][:B], dtype=torch.int32, device="cuda")
        Ow = ops.flash_attention_varlen(Q, K, V, wild, sync=True)
        clamped = torch.tensor([Lk, 1, 1, Lk][:B], dtype=torch.int32, device="cuda")
        assert torch.equal(Ow, ops.flash_attention_varlen(Q, K, V, clamped, sync=True))


@pytest.mark.parametrize("B,H,Lq,d,dtype,shards", [
    (1, 3, 384, 128, torch.bfloat16, [256, 128, 333]), (2, 2, 200, 64, torch.float16, [64, 1, 500, 129]),
    (1, 2, 256, 32, torch.float32, [100, 700]),
])
def test_partials_over_key_shards_merge_to_full_attention(ops, B, H, Lq, d, dtype, shards):
    """SURVEY.md §8(f)-2 building block: fa_partial_forward per K/V shard + fa_v2_combine == attention over all keys
    (what each rank of ring_attention computes)."""
    g = torch.Generator().manual_seed(13)
    Lk = sum(shards)
    Q = ((torch.rand((B, H, Lq, d), generator=g) * 2 - 1).to(dtype)).cuda()
    K, V = (((torch.rand((B, H, Lk, d), generator=g) * 2 - 1).to(dtype)).cuda() for _ in range(2))
    o_parts = torch.empty((len(shards), B * H, Lq, d), dtype=torch.float32, device="cuda")
    lse_parts = torch.empty((len(shards), B * H, Lq), dtype=torch.float32, device="cuda")
    off = 0
    f = lambda x: x.float().cpu().numpy()
    for s, n in enumerate(shards):
        ks, vs = K[:, :, off:off + n].contiguous(), V[:, :, off:off + n].contiguous()
        ops.flash_attention_partial(Q, ks, vs, o_parts[s], lse_parts[s])
        for i in range(B * H):      # each partial against the extended oracle on that shard alone
            ref_o, ref_lse = reference.naive_attention_ex_f64(f(Q).reshape(-1, Lq, d)[i], f(ks).reshape(-1, n, d)[i],
                                                              f(vs).reshape(-1, n, d)[i])
            assert np.abs(o_parts[s, i].cpu().numpy() - ref_o).max() <= TOL[dtype] * max(1.0, np.abs(ref_o).max() * 2)
            assert np.abs(lse_parts[s, i].cpu().numpy() - ref_lse).max() <= 2e-3
        off += n
    O = ops.flash_attention_v2_combine(o_parts, lse_parts, dtype, (B, H, Lq, d))
    torch.cuda.synchronize()
    assert not torch.isnan(O).any()
    for i in range(B * H):
        ref_o, _ = reference.naive_attention_ex_f64(f(Q).reshape(-1, Lq, d)[i], f(K).reshape(-1, Lk, d)[i], f(V).reshape(-1, Lk, d)[i])
        assert np.abs(f(O).reshape(-1, Lq, d)[i] - ref_o).max() <= TOL[dtype]
    # and the oracle's own merge of the GPU partials agrees with the GPU combine
    merged = reference.merge_partials_f64(o_parts.cpu().numpy(), lse_parts.cpu().numpy())
    assert np.abs(f(O).reshape(-1, Lq, d) - merged).max() <= TOL[dtype]


@pytest.mark.parametrize("dtype,d", [(torch.bfloat16, 128), (torch.float16, 64), (torch.float32, 32)])
def test_partial_causal_and_row_windows(ops, dtype, d):
    """The zig-zag ring's building blocks: a causal partial (the diagonal block) and partials whose Q / K,V / outputs
    are row windows of taller tensors, merged into causal attention over the whole local sequence."""
    B, H, C = 2, 2, 192
    g = torch.Generator().manual_seed(17)
    Q, K, V = (((torch.rand((B, H, 2 * C, d), generator=g) * 2 - 1).to(dtype)).cuda() for _ in range(3))
    f = lambda x: x.float().cpu().numpy().reshape(B * H, -1, d)
    # (1) causal partial over the whole local sequence == extended oracle
    o_c, lse_c = ops.flash_attention_partial(Q, K, V, causal=True)
    torch.cuda.synchronize()
    for i in range(B * H):
        ref_o, ref_lse = reference.naive_attention_ex_f64(f(Q)[i], f(K)[i], f(V)[i], causal=True)
        assert np.abs(o_c[i].cpu().numpy() - ref_o).max() <= TOL[dtype] * max(1.0, 2 * np.abs(ref_o).max())
        assert np.abs(lse_c[i].cpu().numpy() - ref_lse).max() <= 2e-3
    # (2) the same thing assembled from windows: [early q x early k causal] [late q x early k full] [late q x late k causal]
    o_parts = torch.zeros((2, B * H, 2 * C, d), dtype=torch.float32, device="cuda")
    lse_parts = torch.full((2, B * H, 2 * C), float("-inf"), dtype=torch.float32, device="cuda")
    ops.flash_attention_partial(Q[:, :, :C], K[:, :, :C], V[:, :, :C], o_parts[0][:, :C], lse_parts[0][:, :C], causal=True)
    ops.flash_attention_partial(Q[:, :, C:], K[:, :, :C], V[:, :, :C], o_parts[0][:, C:], lse_parts[0][:, C:])
    ops.flash_attention_partial(Q[:, :, C:], K[:, :, C:], V[:, :, C:], o_parts[1][:, C:], lse_parts[1][:, C:], causal=True)
    O = ops.flash_attention_v2_combine(o_parts, lse_parts, dtype, (B, H, 2 * C, d))
    Oc = ops.flash_attention_v1_ex(Q, K, V, causal=True, sync=True)
    torch.cuda.synchronize()
    assert not torch.isnan(O).any()
    for i in range(B * H):
        ref_o, _ = reference.naive_attention_ex_f64(f(Q)[i], f(K)[i], f(V)[i], causal=True)
        assert np.abs(f(O)[i] - ref_o).max() <= TOL[dtype] * max(1.0, 2 * np.abs(ref_o).max())
    assert (O.float() - Oc.float()).abs().max().item() <= 2 * TOL[dtype]
    # a transposed (non-window) layout is refused rather than misread
    from exploring_flash_attention_b200 import FlashAttentionError
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_partial(Q.transpose(1, 2), K.transpose(1, 2), V.transpose(1, 2))


def test_varlen_rejects_bad_arguments(ops):
    from exploring_flash_attention_b200 import FlashAttentionError
    q = torch.zeros((1, 1, 128, 64), dtype=torch.bfloat16, device="cuda")
    k = torch.zeros((1, 1, 256, 64), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_varlen(q, k, k, causal=True)          # causal needs Lq == Lk
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_varlen(q, k, k, torch.ones(3, dtype=torch.int32, device="cuda"))   # kv_lens must have B entries
