"""GPU (-m gpu): round-2 parity cases — the generic (> 32 splits) combine kernel and the d = 256 / 512 combine
instantiations, the slab tiled-d kernel's split / causal / LSE / key-padding / partial modes (rows of 512-1024 bytes:
16-bit d = 256, 512 and fp32 d = 128, 256), caller-buffer validation in the torch wrappers, the per-device host staging
and the tensor-map cache.  Oracle and tolerances as in test_parity_gpu.py."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import reference
from tests.test_parity_gpu import TOL, max_err, oracle_out, uniform_qkv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from exploring_flash_attention_b200 import _lib, ops as _ops
    _lib.load()
    return _ops


# ---------------------------------------------------------------------------------------------------------------
# combine: > 32 splits takes fa_combine_generic_kernel; d = 256 / 512 instantiations (flash_attention_v2.h:356-435)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,rows,d,out_dtype", [
    (33, 300, 64, torch.bfloat16), (64, 257, 128, torch.float16), (100, 130, 32, torch.float32),
    (40, 64, 256, torch.bfloat16), (33, 50, 512, torch.float32),
    (4, 1000, 256, torch.bfloat16), (7, 333, 512, torch.bfloat16), (2, 129, 256, torch.float32), (32, 200, 512, torch.float16),
])
def test_combine_many_splits_and_large_d_match_oracle_merge(ops, S, rows, d, out_dtype):
    g = torch.Generator().manual_seed(S * 1000 + d)
    Oacc = (torch.rand((S, 1, rows, d), generator=g) * 2 - 1).cuda()
    LSE = (torch.randn((S, 1, rows), generator=g) * 3).cuda()
    LSE[S // 2, :, ::7] = float("-inf")                          # an empty split on some rows (nothing attended)
    O = ops.flash_attention_v2_combine(Oacc, LSE, out_dtype, (1, 1, rows, d))
    torch.cuda.synchronize()
    assert not torch.isnan(O).any()
    ref = reference.merge_partials_f64(Oacc.cpu().numpy(), LSE.cpu().numpy())
    tol = 2e-6 if out_dtype == torch.float32 else 4e-3           # storage rounding of an O(1) value (2^-9 bf16, 2^-12 fp16)
    assert np.abs(O.float().cpu().numpy().astype(np.float64) - ref).max() <= tol


@pytest.mark.parametrize("L,kvs,d,dtype", [(264, 8, 64, torch.bfloat16), (512, 8, 32, torch.float32), (800, 8, 128, torch.float16)])
def test_v2_with_more_than_32_splits(ops, L, kvs, d, dtype):
    """33 / 64 / 100 splits of 8 keys (one reference BK tile each) through split-KV + the generic combine."""
    Q, K, V = uniform_qkv(1, 2, L, d, dtype)
    assert ops.v2_num_splits(L, kvs) in (33, 64, 100)
    O = ops.flash_attention_v2(Q, K, V, kvs, sync=True)
    assert max_err(O, oracle_out(Q, K, V)) <= TOL[dtype]


# ---------------------------------------------------------------------------------------------------------------
# slab tiled-d kernel: splits, causal, LSE, key padding, partials (d = 256 / 512 16-bit, d = 128 / 256 fp32)
# ---------------------------------------------------------------------------------------------------------------
BIG = [(1, 2, 384, 256, torch.bfloat16), (1, 2, 300, 512, torch.bfloat16), (2, 1, 512, 512, torch.float16),
       (1, 2, 333, 128, torch.float32), (1, 1, 260, 256, torch.float32), (1, 1, 1, 512, torch.bfloat16)]


@pytest.mark.parametrize("B,H,L,d,dtype", BIG)
def test_large_rows_causal_and_lse(ops, B, H, L, d, dtype):
    Q, K, V = uniform_qkv(B, H, L, d, dtype)
    f = lambda x: x.float().cpu().numpy().reshape(-1, L, d)
    q, k, v = f(Q), f(K), f(V)
    for causal in (False, True):
        O, lse = ops.flash_attention_v1_ex(Q, K, V, causal=causal, return_lse=True, sync=True)
        assert not torch.isnan(O).any() and not torch.isnan(lse).any()
        for h in range(B * H):
            ref_o, ref_lse = reference.naive_attention_ex_f64(q[h], k[h], v[h], causal=causal)
            scale = max(1.0, np.abs(ref_o).max() * 2)
            assert np.abs(f(O)[h] - ref_o).max() <= TOL[dtype] * scale
            assert np.abs(lse.cpu().numpy().reshape(-1, L)[h] - ref_lse).max() <= 2e-3
    O = ops.flash_attention_v1_ex(Q, K, V, causal=True, sync=True)
    assert (O[:, :, 0].float() - V[:, :, 0].float()).abs().max().item() <= (1e-3 if dtype == torch.float32 else 1e-6)
    # extras off: the dense entry point (CTA-pair kernel at 16-bit d = 512) agrees within storage rounding
    assert (ops.flash_attention_v1_ex(Q, K, V, sync=True).float() - ops.flash_attention_v1(Q, K, V, sync=True).float()).abs().max().item() <= TOL[dtype]


@pytest.mark.parametrize("B,H,L,d,dtype,kvs", [
    (1, 2, 512, 256, torch.bfloat16, 128), (1, 2, 300, 512, torch.bfloat16, 64), (1, 1, 640, 512, torch.float16, 200),
    (2, 2, 512, 128, torch.float32, 128),      # the reference V2's own default D = 128 in its USE_FP64 mode
    (1, 2, 333, 256, torch.float32, 100), (1, 1, 264, 256, torch.bfloat16, 8),     # 33 splits -> generic combine, d = 256
])
def test_v2_large_rows(ops, B, H, L, d, dtype, kvs):
    Q, K, V = uniform_qkv(B, H, L, d, dtype)
    ref = oracle_out(Q, K, V)
    O = ops.flash_attention_v2(Q, K, V, kvs, sync=True)
    assert not torch.isnan(O).any()
    assert max_err(O, ref) <= TOL[dtype]
    Oacc, LSE = ops.flash_attention_v2_splitkv(Q, K, V, kvs)
    torch.cuda.synchronize()
    S = Oacc.shape[0]
    f = lambda x: x.float().cpu().numpy().reshape(-1, L, d).astype(np.float64)
    q, k, v = f(Q), f(K), f(V)
    for s in (0, S - 1):
        ks = slice(s * kvs, min(L, (s + 1) * kvs))
        for h in range(B * H):
            ref_o, ref_lse = reference.naive_attention_ex_f64(q[h], k[h][ks], v[h][ks])
            assert np.abs(Oacc[s, h].cpu().numpy() - ref_o).max() <= TOL[dtype] * 2
            assert np.abs(LSE[s, h].cpu().numpy() - ref_lse).max() <= 2e-3
    merged = reference.merge_partials_f64(Oacc.cpu().numpy(), LSE.cpu().numpy()).reshape(B, H, L, d)
    assert np.abs(O.float().cpu().numpy() - merged).max() <= TOL[dtype]


@pytest.mark.parametrize("B,H,Lq,Lk,d,dtype,causal", [
    (3, 1, 256, 256, 256, torch.bfloat16, True), (2, 2, 130, 400, 512, torch.bfloat16, False),
    (2, 1, 300, 300, 128, torch.float32, True), (2, 1, 64, 700, 256, torch.float16, False),
])
def test_large_rows_key_padding_and_rectangular(ops, B, H, Lq, Lk, d, dtype, causal):
    g = torch.Generator().manual_seed(23)
    Q = ((torch.rand((B, H, Lq, d), generator=g) * 2 - 1).to(dtype)).cuda()
    K, V = (((torch.rand((B, H, Lk, d), generator=g) * 2 - 1).to(dtype)).cuda() for _ in range(2))
    lens = [Lk, 1, max(1, Lk // 2 + 3)][:B]
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


def test_large_rows_partials_merge(ops):
    B, H, Lq, d, dtype, shards = 1, 2, 200, 256, torch.bfloat16, [128, 300, 1]
    g = torch.Generator().manual_seed(29)
    Lk = sum(shards)
    Q = ((torch.rand((B, H, Lq, d), generator=g) * 2 - 1).to(dtype)).cuda()
    K, V = (((torch.rand((B, H, Lk, d), generator=g) * 2 - 1).to(dtype)).cuda() for _ in range(2))
    o_parts = torch.empty((len(shards), B * H, Lq, d), dtype=torch.float32, device="cuda")
    lse_parts = torch.empty((len(shards), B * H, Lq), dtype=torch.float32, device="cuda")
    off = 0
    for s, n in enumerate(shards):      # K/V shards addressed as row windows of the full tensors
        ops.flash_attention_partial(Q, K[:, :, off:off + n], V[:, :, off:off + n], o_parts[s], lse_parts[s])
        off += n
    O = ops.flash_attention_v2_combine(o_parts, lse_parts, dtype, (B, H, Lq, d))
    torch.cuda.synchronize()
    assert max_err(O, np.stack([reference.naive_attention_ex_f64(*(x.float().cpu().numpy().reshape(B * H, -1, d)[i] for x in (Q, K, V)))[0]
                                for i in range(B * H)])) <= TOL[dtype]


# ---------------------------------------------------------------------------------------------------------------
# torch wrappers: caller-supplied buffers are validated before their raw pointers reach the C ABI
# ---------------------------------------------------------------------------------------------------------------
def test_caller_buffers_are_validated(ops):
    from exploring_flash_attention_b200 import FlashAttentionError
    Q, K, V = uniform_qkv(1, 2, 256, 64, torch.bfloat16)
    ws_small = ops.v2_workspace(1, 2, 256, 64, 128, Q.device)            # 2 splits
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_v2(Q, K, V, 64, workspace=ws_small)          # needs 4 splits: would write out of bounds
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_v2_splitkv(Q, K, V, 64, ws_small[0], ws_small[1])
    ws = ops.v2_workspace(1, 2, 256, 64, 64, Q.device)
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_v2_splitkv(Q, K, V, 64, ws[0].double(), ws[1])
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_v2_splitkv(Q, K, V, 64, ws[0].transpose(2, 3), ws[1])
    for bad_O in (torch.empty((1, 2, 128, 64), dtype=torch.bfloat16, device="cuda"), torch.empty_like(Q, dtype=torch.float16),
                  torch.empty((1, 2, 64, 256), dtype=torch.bfloat16, device="cuda").transpose(2, 3), torch.empty(Q.shape, dtype=Q.dtype)):
        for call in (lambda o: ops.flash_attention_v1(Q, K, V, o), lambda o: ops.flash_attention_v1_ex(Q, K, V, o),
                     lambda o: ops.flash_attention_varlen(Q, K, V, O=o), lambda o: ops.flash_attention_v1_tiled_d(Q, K, V, o),
                     lambda o: ops.flash_attention_v2(Q, K, V, 64, O=o),
                     lambda o: ops.flash_attention_v2_combine(ws[0], ws[1], torch.bfloat16, (1, 2, 256, 64), o)):
            with pytest.raises(FlashAttentionError):
                call(bad_O)
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_v2_combine(ws[0], ws[1][:2], torch.bfloat16, (1, 2, 256, 64))
    # a correct reused workspace and output still work
    O = torch.empty_like(Q)
    assert ops.flash_attention_v2(Q, K, V, 64, O=O, workspace=ws, sync=True) is O
    assert max_err(O, oracle_out(Q, K, V)) <= 2e-3


def test_tensor_maps_are_cached_across_launches(ops):
    from exploring_flash_attention_b200 import _lib
    lib = _lib.load()
    Q, K, V = uniform_qkv(1, 2, 512, 128, torch.bfloat16)
    O = torch.empty_like(Q)
    ops.flash_attention_v1(Q, K, V, O, sync=True)
    h0, m0, h1, m1 = (ctypes.c_ulonglong() for _ in range(4))
    lib.fa_debug_map_cache_stats(ctypes.byref(h0), ctypes.byref(m0))
    for _ in range(10):
        ops.flash_attention_v1(Q, K, V, O)
    torch.cuda.synchronize()
    lib.fa_debug_map_cache_stats(ctypes.byref(h1), ctypes.byref(m1))
    assert m1.value == m0.value and h1.value - h0.value == 40          # 4 operands x 10 launches, none re-encoded
    # a recycled address with another shape must not hit the stale entry
    Q2, K2, V2 = (x[:, :, :256].contiguous() for x in (Q, K, V))
    O2 = ops.flash_attention_v1(Q2, K2, V2, sync=True)
    assert max_err(O2, oracle_out(Q2, K2, V2)) <= 2e-3
    assert max_err(ops.flash_attention_v1(Q, K, V, sync=True), oracle_out(Q, K, V)) <= 2e-3


def test_full_c4_tensor_sampled_rows(ops):
    """BASELINE.json configs[3] at FULL size (B8 H32 L16384 d128 bf16, 4.3 GB of tensors): sampled heads x sampled
    row blocks against the oracle (row-subset naive_attention, float64) + rows of softmax sum to one."""
    B, H, L, d = 8, 32, 16384, 128
    g = torch.Generator(device="cuda").manual_seed(42)
    Q, K, V = ((torch.rand((B, H, L, d), generator=g, device="cuda") * 2 - 1).bfloat16() for _ in range(3))
    O = ops.flash_attention_v1(Q, K, V, sync=True)
    heads = [0, 101, B * H - 1]
    rows = np.r_[0:32, L // 2:L // 2 + 32, L - 32:L]
    f = lambda x: x.reshape(B * H, L, d)[heads].float().cpu().numpy()
    ref = reference.naive_attention_batched_f64(f(Q), f(K), f(V), rows=rows)
    got = O.reshape(B * H, L, d)[heads][:, rows].float().cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-3
    V.fill_(1.0)
    O = ops.flash_attention_v1(Q, K, V, sync=True)
    assert (O.float() - 1).abs().max().item() <= 4e-3


# ---------------------------------------------------------------------------------------------------------------
# the package's independent evaluation (drop-in for common/reference.py naive_attention) and the C++ consumer
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,Lq,Lk,d,dtype,tol", [
    (3, 200, 200, 64, torch.float32, 2e-6), (2, 130, 333, 48, torch.float32, 2e-6), (5, 64, 1000, 20, torch.float64, 1e-13),
    (1, 1, 1, 1, torch.float64, 1e-15), (2, 257, 257, 128, torch.float64, 1e-13),
])
def test_independent_naive_attention_matches_oracle(ops, n, Lq, Lk, d, dtype, tol):
    g = torch.Generator().manual_seed(31)
    Q = torch.randn((n, Lq, d), generator=g, dtype=torch.float64).to(dtype).cuda()
    K, V = (torch.randn((n, Lk, d), generator=g, dtype=torch.float64).to(dtype).cuda() for _ in range(2))
    O = ops.naive_attention_reference(Q, K, V)
    O_small_ws = ops.naive_attention_reference(Q, K, V, max_workspace_bytes=1)      # one head's scores at a time
    assert torch.equal(O, O_small_ws)
    for h in range(n):
        ref, _ = reference.naive_attention_ex_f64(Q[h].cpu().numpy(), K[h].cpu().numpy(), V[h].cpu().numpy())
        assert np.abs(O[h].double().cpu().numpy() - ref).max() <= tol * max(1.0, np.abs(ref).max())


def test_dropin_naive_attention_is_not_the_kernel_under_test(ops):
    """common.reference.naive_attention (the name the reference scripts compare everything with) agrees with the float64
    oracle far more tightly than the tensor-core kernels can, for any head dim, and in the caller's dtype."""
    from exploring_flash_attention_b200.common.reference import check_accuracy, naive_attention
    from exploring_flash_attention_b200.flash_attention_v1 import flash_attention_tiled
    rng = np.random.default_rng(0)
    for L, d, dt, tol in ((256, 128, np.float64, 1e-12), (100, 24, np.float64, 1e-12), (300, 64, np.float32, 2e-6), (128, 32, np.float16, 1e-3)):
        Q, K, V = (rng.standard_normal((L, d)).astype(dt) for _ in range(3))
        out = naive_attention(Q, K, V)
        assert out.dtype == dt and out.shape == (L, d)
        ref = reference.naive_attention_f64(Q, K, V)
        assert np.abs(out.astype(np.float64) - ref).max() <= tol * max(1.0, np.abs(ref).max())
    # the reference scripts' own check (numpy_gpu_like_opt2.py __main__): tiled kernel vs naive_attention, float64 buffers
    L, d = 64, 32
    Q, K, V = (rng.standard_normal((L, d)) for _ in range(3))
    O = np.zeros(L * d)
    flash_attention_tiled(Q.flatten(), K.flatten(), V.flatten(), O, L, d, Bq=8, Bk=8)
    naive = naive_attention(Q, K, V)
    diff = np.abs(O.reshape(L, d) - naive).max()
    assert 0 < diff <= 4e-3          # tf32 products vs fp64: close, and visibly NOT the same computation
    check_accuracy(O.reshape(L, d), naive, "drop-in V1 vs drop-in oracle")


def test_cpp_consumer_runs_and_passes():
    """The compiled C++ driver (tests/drivers/driver_v1.cu) on the reference V1 driver's own config."""
    import subprocess
    from exploring_flash_attention_b200 import _build
    exe = _build.build_consumer()
    res = subprocess.run([str(exe), "3"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    assert "Test PASSED" in res.stdout


def test_host_staging_on_a_second_device_in_the_same_process(ops):
    """fa_forward_host caches one staging set PER DEVICE: after a call on cuda:0 a call on cuda:1 must not reuse
    cuda:0's buffers, streams or events."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = torch.Generator().manual_seed(5)
    Qh, Kh, Vh = ((torch.rand((1, 4, 512, 128), generator=g) * 2 - 1).bfloat16().pin_memory() for _ in range(3))
    with torch.cuda.device(0):
        O0 = ops.flash_attention_host(Qh, Kh, Vh, variant=0).clone()
    with torch.cuda.device(1):
        O1 = ops.flash_attention_host(Qh, Kh, Vh, variant=0).clone()
        O1v2 = ops.flash_attention_host(Qh, Kh, Vh, variant=2, kv_per_split=128).clone()
    with torch.cuda.device(0):
        O0b = ops.flash_attention_host(Qh, Kh, Vh, variant=0).clone()
    assert torch.equal(O0, O1) and torch.equal(O0, O0b)
    assert (O1v2.float() - O0.float()).abs().max().item() <= 2e-3
    assert max_err(O0.cuda(), oracle_out(Qh, Kh, Vh)) <= 2e-3


# ---------------------------------------------------------------------------------------------------------------
# backward: dQ, dK, dV against the float64 analytic gradient (itself pinned to finite differences)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,L,d,dtype", [
    (1, 2, 256, 128, torch.bfloat16), (2, 2, 384, 64, torch.bfloat16), (1, 2, 512, 128, torch.float16),
    (1, 3, 333, 128, torch.bfloat16), (1, 2, 100, 64, torch.float16), (1, 1, 129, 128, torch.bfloat16),
    (1, 1, 1, 64, torch.bfloat16), (1, 2, 1000, 64, torch.bfloat16),
])
def test_backward_matches_analytic_gradient(ops, B, H, L, d, dtype):
    """Parity bar for the backward: max-abs error <= 5e-3 of the gradient's own magnitude (max |grad|), per tensor."""
    Q, K, V = uniform_qkv(B, H, L, d, dtype)
    g = torch.Generator().manual_seed(99)
    dO = ((torch.rand((B, H, L, d), generator=g) * 2 - 1)).to(dtype).cuda()
    f = lambda x: x.float().cpu().numpy().reshape(B * H, L, d).astype(np.float64)
    for causal in (False, True):
        O, lse = ops.flash_attention_v1_ex(Q, K, V, causal=causal, return_lse=True, sync=True)
        dQ, dK, dV = ops.flash_attention_backward(Q, K, V, O, dO, lse, causal=causal, sync=True)
        for t in (dQ, dK, dV):
            assert not torch.isnan(t).any() and not torch.isinf(t).any()
        for h in range(B * H):
            rQ, rK, rV = reference.attention_backward_f64(f(Q)[h], f(K)[h], f(V)[h], f(dO)[h], causal=causal)
            for name, got, ref in (("dQ", f(dQ)[h], rQ), ("dK", f(dK)[h], rK), ("dV", f(dV)[h], rV)):
                err = np.abs(got - ref).max()
                assert err <= 5e-3 * max(np.abs(ref).max(), 1e-3), (name, causal, h, err, np.abs(ref).max())


def test_backward_c2_shape_sampled_heads(ops):
    """BASELINE.json configs[1] shape (B32 H8 L1024 d128 bf16): sampled heads against the float64 gradient."""
    B, H, L, d = 32, 8, 1024, 128
    Q, K, V = uniform_qkv(B, H, L, d, torch.bfloat16)
    dO = (torch.rand((B, H, L, d), generator=torch.Generator().manual_seed(7)) * 2 - 1).bfloat16().cuda()
    O, lse = ops.flash_attention_v1_ex(Q, K, V, return_lse=True, sync=True)
    ws = torch.empty(ops.backward_workspace_bytes(B, H, L), dtype=torch.uint8, device="cuda")
    dQ, dK, dV = ops.flash_attention_backward(Q, K, V, O, dO, lse, workspace=ws, sync=True)
    f = lambda x, h: x.reshape(B * H, L, d)[h].float().cpu().numpy().astype(np.float64)
    for h in (0, 131, B * H - 1):
        rQ, rK, rV = reference.attention_backward_f64(f(Q, h), f(K, h), f(V, h), f(dO, h))
        for got, ref in ((f(dQ, h), rQ), (f(dK, h), rK), (f(dV, h), rV)):
            assert np.abs(got - ref).max() <= 5e-3 * np.abs(ref).max()
    from exploring_flash_attention_b200 import FlashAttentionError
    with pytest.raises(FlashAttentionError):
        ops.flash_attention_backward(Q, K, V, O, dO, lse, workspace=ws[:100])
    q32 = torch.zeros((1, 1, 128, 32), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(FlashAttentionError) as ei:
        ops.flash_attention_backward(q32, q32, q32, q32, q32, torch.zeros((1, 1, 128), device="cuda"))
    assert ei.value.code == -4


# ---------------------------------------------------------------------------------------------------------------
# memory safety without compute-sanitizer (closed on this GPU pool): every output buffer sits between canary bands
# ---------------------------------------------------------------------------------------------------------------
class _Canary:
    """Carves tensors out of one big sentinel-filled allocation, 4 KB of canary on both sides of each, and checks the
    canaries afterwards."""

    def __init__(self, nbytes=1 << 28):
        self.buf = torch.full((nbytes,), 0xA5, dtype=torch.uint8, device="cuda")
        self.off = 4096
        self.spans = []

    def empty(self, shape, dtype):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        start = self.off
        self.spans.append((start, start + n))
        self.off = (start + n + 4096 + 255) // 256 * 256
        assert self.off < self.buf.numel()
        return self.buf[start:start + n].view(dtype).reshape(shape)

    def check(self):
        torch.cuda.synchronize()
        mask = torch.ones(self.off, dtype=torch.bool, device="cuda")
        for a, b in self.spans:
            mask[a:b] = False
        assert bool((self.buf[:self.off][mask] == 0xA5).all()), "a kernel wrote outside its output buffer"


@pytest.mark.parametrize("B,H,L,d,dtype", [
    (1, 3, 333, 128, torch.bfloat16), (2, 2, 129, 64, torch.float16), (1, 2, 257, 32, torch.float32), (1, 5, 1, 128, torch.bfloat16),
    (1, 2, 200, 256, torch.bfloat16), (1, 2, 130, 512, torch.bfloat16), (1, 2, 333, 128, torch.float32),
])
def test_kernels_never_write_outside_their_output_buffers(ops, B, H, L, d, dtype):
    Q, K, V = uniform_qkv(B, H, L, d, dtype)
    c = _Canary()
    ref = oracle_out(Q, K, V)
    O = c.empty((B, H, L, d), dtype)
    ops.flash_attention_v1(Q, K, V, O)
    c.check()
    assert max_err(O, ref) <= TOL[dtype]
    O2 = c.empty((B, H, L, d), dtype)
    lse_dummy = ops.flash_attention_v1_ex(Q, K, V, O2, causal=True, return_lse=True)[1]
    c.check()
    for kvs in (8 if L <= 200 else 24, 96, 160):
        S = ops.v2_num_splits(L, kvs)
        Oacc, LSEacc = c.empty((S, B * H, L, d), torch.float32), c.empty((S, B * H, L), torch.float32)
        O3 = c.empty((B, H, L, d), dtype)
        ops.flash_attention_v2(Q, K, V, kvs, O=O3, workspace=(Oacc, LSEacc))
        c.check()
        assert max_err(O3, ref) <= TOL[dtype]
    Op, Lp = c.empty((B * H, L, d), torch.float32), c.empty((B * H, L), torch.float32)
    ops.flash_attention_partial(Q, K, V, Op, Lp)
    c.check()
    if dtype != torch.float32 and d in (64, 128):
        dO = torch.ones_like(Q)
        Of, lse = ops.flash_attention_v1_ex(Q, K, V, return_lse=True, sync=True)
        ws = c.empty((ops.backward_workspace_bytes(B, H, L),), torch.uint8)
        grads = ops.flash_attention_backward(Q, K, V, Of, dO, lse, workspace=ws)      # dQ, dK, dV are allocated inside ...
        c.check()                                                                     # ... the workspace is ours
        assert all(not torch.isnan(g_).any() for g_ in grads)
