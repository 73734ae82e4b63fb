"""Torch-facing wrappers over the C ABI: device tensors in, device tensors out, on the current stream.

These are the B200 drop-ins for the reference host launchers
(flash_attention_v1/CUDA/flash_attention_v1.h:251, flash_attention_v1_tiled_d/CUDA/flash_attention_v1.h:312,
flash_attention_v2/CUDA/flash_attention_v2.h:438).  torch is used for memory, streams and nothing else.
"""
from __future__ import annotations

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.FA_DTYPE_F32, torch.bfloat16: _lib.FA_DTYPE_BF16, torch.float16: _lib.FA_DTYPE_F16}


def _prep(Q, K, V):
    if not (Q.is_cuda and K.is_cuda and V.is_cuda):
        raise RuntimeError("flash-attention B200 path needs CUDA tensors: there is no CPU fallback")
    if Q.dtype not in _DTYPES:
        raise _lib.FlashAttentionError(-2, f"unsupported dtype {Q.dtype}")
    if not (Q.dtype == K.dtype == V.dtype):
        raise _lib.FlashAttentionError(-2, "Q, K, V must share a dtype")
    if Q.dim() != 4 or Q.shape != K.shape or Q.shape != V.shape:
        raise _lib.FlashAttentionError(-1, "Q, K, V must be [B,H,L,d] with identical shapes")
    return Q.contiguous(), K.contiguous(), V.contiguous()


def _check_buffer(t, shape, dtype, device, what):
    """Caller-supplied output / workspace buffers go to the C ABI as raw pointers: refuse anything the kernels would
    write out of bounds (wrong shape, dtype, device, or a non-contiguous view)."""
    if not isinstance(t, torch.Tensor):
        raise _lib.FlashAttentionError(-1, f"{what} must be a torch tensor")
    if t.device != device:
        raise _lib.FlashAttentionError(-3, f"{what} is on {t.device}, expected {device}")
    if t.dtype != dtype:
        raise _lib.FlashAttentionError(-2, f"{what} has dtype {t.dtype}, expected {dtype}")
    if tuple(t.shape) != tuple(shape):
        raise _lib.FlashAttentionError(-1, f"{what} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    if not t.is_contiguous():
        raise _lib.FlashAttentionError(-3, f"{what} must be contiguous")
    return t


def _out_like(O, Q, what="O"):
    return torch.empty_like(Q) if O is None else _check_buffer(O, Q.shape, Q.dtype, Q.device, what)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_device_of(fn):
    """The C ABI launches on the CURRENT device and stream: make the tensors' device current for the call."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = next((a for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)          # already current: skip the (several us) device-guard round trip
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper


@_on_device_of
def flash_attention_v1(Q, K, V, O=None, sync: bool = False):
    """O = softmax(Q K^T / sqrt(d)) V for [B,H,L,d] tensors (fused-tile kernel, d <= 128)."""
    Q, K, V = _prep(Q, K, V)
    B, H, L, d = Q.shape
    O = _out_like(O, Q)
    lib = _lib.load()
    _lib.check(lib.fa_v1_forward(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), B, H, L, d, _DTYPES[Q.dtype],
                                 _stream()))
    if sync:
        torch.cuda.current_stream().synchronize()
    return O


@_on_device_of
def flash_attention_v1_ex(Q, K, V, O=None, causal: bool = False, return_lse: bool = False, sync: bool = False):
    """Fused-tile kernel with the extras the reference lists as future work: causal masking and the per-row
    log-sum-exp (natural log of sum_j exp(q.k_j/sqrt(d))). Returns O, or (O, LSE [B,H,L] fp32)."""
    Q, K, V = _prep(Q, K, V)
    B, H, L, d = Q.shape
    O = _out_like(O, Q)
    lse = torch.empty((B, H, L), dtype=torch.float32, device=Q.device) if return_lse else None
    lib = _lib.load()
    _lib.check(lib.fa_v1_forward_ex(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(),
                                    lse.data_ptr() if return_lse else None, B, H, L, d, _DTYPES[Q.dtype],
                                    1 if causal else 0, _stream()))
    if sync:
        torch.cuda.current_stream().synchronize()
    return (O, lse) if return_lse else O


def _prep_rect(Q, K, V):
    """Like _prep, but K/V may hold a different number of rows than Q ([B,H,Lq,d] vs [B,H,Lk,d])."""
    if not (Q.is_cuda and K.is_cuda and V.is_cuda):
        raise RuntimeError("flash-attention B200 path needs CUDA tensors: there is no CPU fallback")
    if Q.dtype not in _DTYPES:
        raise _lib.FlashAttentionError(-2, f"unsupported dtype {Q.dtype}")
    if not (Q.dtype == K.dtype == V.dtype):
        raise _lib.FlashAttentionError(-2, "Q, K, V must share a dtype")
    if Q.dim() != 4 or K.dim() != 4 or K.shape != V.shape or Q.shape[:2] != K.shape[:2] or Q.shape[3] != K.shape[3]:
        raise _lib.FlashAttentionError(-1, "Q must be [B,H,Lq,d] and K, V [B,H,Lk,d]")
    return Q.contiguous(), K.contiguous(), V.contiguous()


@_on_device_of
def flash_attention_varlen(Q, K, V, kv_lens=None, O=None, causal: bool = False, return_lse: bool = False,
                           sync: bool = False):
    """Fused-tile kernel with a key-padding mask and/or Lq != Lk ("dynamic sequence lengths", which the reference
    lists as future work, flash_attention_v1/README_v1.md:169).  kv_lens: int32 [B] on the device — batch entry b
    attends to its first kv_lens[b] keys (clamped to [1, Lk])."""
    Q, K, V = _prep_rect(Q, K, V)
    B, H, Lq, d = Q.shape
    Lk = K.shape[2]
    O = _out_like(O, Q)
    if kv_lens is not None:
        if kv_lens.dtype != torch.int32 or kv_lens.numel() != B or kv_lens.device != Q.device:
            raise _lib.FlashAttentionError(-1, "kv_lens must be an int32 tensor of B entries on Q's device")
        kv_lens = kv_lens.contiguous()
    lse = torch.empty((B, H, Lq), dtype=torch.float32, device=Q.device) if return_lse else None
    lib = _lib.load()
    _lib.check(lib.fa_v1_forward_varlen(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(),
                                        lse.data_ptr() if return_lse else None,
                                        kv_lens.data_ptr() if kv_lens is not None else None,
                                        B, H, Lq, Lk, d, _DTYPES[Q.dtype], 1 if causal else 0, _stream()))
    if sync:
        torch.cuda.current_stream().synchronize()
    return (O, lse) if return_lse else O


def _head_rows(x, what):
    """x: [B,H,L,d] (or [BH,L,d]) whose rows are dense and whose heads are evenly spaced (a contiguous tensor or a
    window x[..., r0:r1, :] of one).  Returns the spacing of consecutive heads in rows."""
    L, d = x.shape[-2:]
    st = x.stride()
    ok = st[-1] == 1 and st[-2] == d and st[-3] % d == 0 and st[-3] >= L * d
    if ok and x.dim() == 4 and x.shape[0] > 1:
        ok = st[0] == x.shape[1] * st[1]
    if not ok or x.dim() not in (3, 4):
        raise _lib.FlashAttentionError(-3, f"{what} must be contiguous or a row window of a contiguous [B,H,L,d] tensor")
    return st[-3] // d


@_on_device_of
def flash_attention_partial(Q, K, V, Opartial=None, LSEpartial=None, causal: bool = False):
    """One un-merged attention partial: Q [B,H,Lq,d] against one key/value shard [B,H,Lk,d].
    Returns (Opartial [B*H,Lq,d] fp32 normalised by the shard's own row sums, LSEpartial [B*H,Lq] fp32);
    partials stacked on a leading axis are flash_attention_v2_combine's input.  Q, K, V and the outputs may be row
    windows (x[:, :, r0:r1]) of taller contiguous tensors; causal=True (Lq == Lk) masks keys above the diagonal."""
    if not (Q.is_cuda and K.is_cuda and V.is_cuda):
        raise RuntimeError("flash-attention B200 path needs CUDA tensors: there is no CPU fallback")
    if Q.dtype not in _DTYPES or not (Q.dtype == K.dtype == V.dtype):
        raise _lib.FlashAttentionError(-2, "Q, K, V must share a supported dtype")
    if Q.dim() != 4 or K.dim() != 4 or K.shape != V.shape or Q.shape[:2] != K.shape[:2] or Q.shape[3] != K.shape[3]:
        raise _lib.FlashAttentionError(-1, "Q must be [B,H,Lq,d] and K, V [B,H,Lk,d]")
    B, H, Lq, d = Q.shape
    Lk = K.shape[2]
    if Opartial is None:
        Opartial = torch.empty((B * H, Lq, d), dtype=torch.float32, device=Q.device)
    if LSEpartial is None:
        LSEpartial = torch.empty((B * H, Lq), dtype=torch.float32, device=Q.device)
    q_rows, k_rows, v_rows, o_rows = (_head_rows(x, n) for x, n in ((Q, "Q"), (K, "K"), (V, "V"), (Opartial, "Opartial")))
    lse_ok = LSEpartial.stride(-1) == 1 and LSEpartial.dtype == torch.float32 and \
        (LSEpartial.numel() == LSEpartial.shape[-1] or LSEpartial.stride(-2) == o_rows)
    if k_rows != v_rows or not lse_ok or Opartial.dtype != torch.float32:
        raise _lib.FlashAttentionError(-3, "K and V must share a layout; LSEpartial must be fp32 with Opartial's head spacing")
    lib = _lib.load()
    _lib.check(lib.fa_partial_forward(Q.data_ptr(), K.data_ptr(), V.data_ptr(), Opartial.data_ptr(),
                                      LSEpartial.data_ptr(), B, H, Lq, Lk, d, _DTYPES[Q.dtype], q_rows, k_rows, o_rows,
                                      1 if causal else 0, _stream()))
    return Opartial, LSEpartial


@_on_device_of
def flash_attention_v1_tiled_d(Q, K, V, O=None, d_tile_qk: int = 32, d_tile_v: int = 32, sync: bool = False):
    """Tiled-d variant (head dims up to 512); d_tile_* are validated streaming hints."""
    Q, K, V = _prep(Q, K, V)
    B, H, L, d = Q.shape
    O = _out_like(O, Q)
    lib = _lib.load()
    _lib.check(lib.fa_v1_tiled_d_forward(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), B, H, L, d,
                                         d_tile_qk, d_tile_v, _DTYPES[Q.dtype], _stream()))
    if sync:
        torch.cuda.current_stream().synchronize()
    return O


@_on_device_of
def flash_attention_v1_tiled_d_pair(Q, K, V, O=None, sync: bool = False):
    """Tiled-d on CTA pairs (fa_v1_tiled_d_pair_forward): bf16/fp16, d in {256, 512}."""
    Q, K, V = _prep(Q, K, V)
    B, H, L, d = Q.shape
    O = _out_like(O, Q)
    lib = _lib.load()
    _lib.check(lib.fa_v1_tiled_d_pair_forward(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), B, H, L, d,
                                              _DTYPES[Q.dtype], _stream()))
    if sync:
        torch.cuda.current_stream().synchronize()
    return O


def v2_num_splits(L: int, kv_per_split: int) -> int:
    return _lib.load().fa_v2_num_splits(L, kv_per_split)


def v2_workspace(B, H, L, d, kv_per_split, device):
    """Caller-owned workspace: (Oaccum [S,B*H,L,d] fp32, LSEaccum [S,B*H,L] fp32)."""
    S = v2_num_splits(L, kv_per_split)
    if S <= 0:
        raise _lib.FlashAttentionError(-1, "kv_per_split must be positive")
    return (torch.empty((S, B * H, L, d), dtype=torch.float32, device=device),
            torch.empty((S, B * H, L), dtype=torch.float32, device=device))


@_on_device_of
def flash_attention_v2_splitkv(Q, K, V, kv_per_split: int, Oaccum=None, LSEaccum=None):
    Q, K, V = _prep(Q, K, V)
    B, H, L, d = Q.shape
    if Oaccum is None or LSEaccum is None:
        Oaccum, LSEaccum = v2_workspace(B, H, L, d, kv_per_split, Q.device)
    else:
        # a reused workspace must hold exactly this call's splits: the kernel writes n_splits*B*H*L rows into it
        S = v2_num_splits(L, kv_per_split)
        if S <= 0:
            raise _lib.FlashAttentionError(-1, "kv_per_split must be positive")
        _check_buffer(Oaccum, (S, B * H, L, d), torch.float32, Q.device, "Oaccum")
        _check_buffer(LSEaccum, (S, B * H, L), torch.float32, Q.device, "LSEaccum")
    lib = _lib.load()
    _lib.check(lib.fa_v2_splitkv_forward(Q.data_ptr(), K.data_ptr(), V.data_ptr(), Oaccum.data_ptr(),
                                         LSEaccum.data_ptr(), B, H, L, d, kv_per_split, _DTYPES[Q.dtype], _stream()))
    return Oaccum, LSEaccum


@_on_device_of
def flash_attention_v2_combine(Oaccum, LSEaccum, out_dtype, shape, O=None):
    B, H, L, d = shape
    if not (isinstance(Oaccum, torch.Tensor) and Oaccum.is_cuda and Oaccum.dim() >= 1):
        raise RuntimeError("flash-attention B200 path needs CUDA tensors: there is no CPU fallback")
    if out_dtype not in _DTYPES:
        raise _lib.FlashAttentionError(-2, f"unsupported dtype {out_dtype}")
    S = Oaccum.shape[0]
    # [S,B*H,L,d] or any contiguous view with the same element count per split (e.g. [S,1,B*H*L,d])
    if Oaccum.dtype != torch.float32 or not Oaccum.is_contiguous() or Oaccum.numel() != S * B * H * L * d:
        raise _lib.FlashAttentionError(-1, f"Oaccum must be contiguous fp32 with {S}*{B * H * L * d} elements")
    if (not isinstance(LSEaccum, torch.Tensor) or LSEaccum.dtype != torch.float32 or not LSEaccum.is_contiguous()
            or LSEaccum.numel() != S * B * H * L or LSEaccum.device != Oaccum.device):
        raise _lib.FlashAttentionError(-1, f"LSEaccum must be contiguous fp32 with {S}*{B * H * L} elements on Oaccum's device")
    if O is None:
        O = torch.empty((B, H, L, d), dtype=out_dtype, device=Oaccum.device)
    else:
        _check_buffer(O, (B, H, L, d), out_dtype, Oaccum.device, "O")
    lib = _lib.load()
    _lib.check(lib.fa_v2_combine(Oaccum.data_ptr(), LSEaccum.data_ptr(), O.data_ptr(), B, H, L, d, S,
                                 _DTYPES[out_dtype], _stream()))
    return O


@_on_device_of
def flash_attention_v2(Q, K, V, kv_per_split: int, O=None, workspace=None, sync: bool = False):
    """Split-KV forward + combine. `workspace` = (Oaccum, LSEaccum) from v2_workspace(), reused across calls.
    (Validates once and makes the two library calls itself: at the reference's C3 size the two kernels take 25 us and
    every avoidable microsecond of host work per call shows.)"""
    Q, K, V = _prep(Q, K, V)
    B, H, L, d = Q.shape
    S = -(-L // kv_per_split) if kv_per_split > 0 else 0          # = fa_v2_num_splits(L, kv_per_split)
    if S <= 0:
        raise _lib.FlashAttentionError(-1, "kv_per_split must be positive")
    if workspace is None:
        Oaccum = torch.empty((S, B * H, L, d), dtype=torch.float32, device=Q.device)
        LSEaccum = torch.empty((S, B * H, L), dtype=torch.float32, device=Q.device)
    else:
        # a reused workspace must hold exactly this call's splits: the kernel writes n_splits*B*H*L rows into it
        Oaccum, LSEaccum = workspace
        _check_buffer(Oaccum, (S, B * H, L, d), torch.float32, Q.device, "Oaccum")
        _check_buffer(LSEaccum, (S, B * H, L), torch.float32, Q.device, "LSEaccum")
    O = _out_like(O, Q)
    lib = _lib.load()
    dt, st = _DTYPES[Q.dtype], _stream()
    _lib.check(lib.fa_v2_splitkv_forward(Q.data_ptr(), K.data_ptr(), V.data_ptr(), Oaccum.data_ptr(), LSEaccum.data_ptr(),
                                         B, H, L, d, kv_per_split, dt, st))
    _lib.check(lib.fa_v2_combine(Oaccum.data_ptr(), LSEaccum.data_ptr(), O.data_ptr(), B, H, L, d, S, dt, st))
    if sync:
        torch.cuda.current_stream().synchronize()
    return O


@_on_device_of
def flash_attention_backward(Q, K, V, O, dO, LSE, causal: bool = False, workspace=None, sync: bool = False):
    """Gradients (dQ, dK, dV) of O = softmax(Q K^T / sqrt(d)) V for [B,H,L,d] bf16 / fp16 tensors, d in {64, 128}.
    O and LSE come from flash_attention_v1_ex(..., causal=causal, return_lse=True).  `workspace`: optional uint8 CUDA
    tensor of backward_workspace_bytes(B, H, L) bytes, reused across calls."""
    Q, K, V = _prep(Q, K, V)
    B, H, L, d = Q.shape
    O = _check_buffer(O, Q.shape, Q.dtype, Q.device, "O")
    dO = _check_buffer(dO.contiguous(), Q.shape, Q.dtype, Q.device, "dO")
    LSE = _check_buffer(LSE, (B, H, L), torch.float32, Q.device, "LSE")
    lib = _lib.load()
    need = lib.fa_v1_backward_workspace_bytes(B, H, L)
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=Q.device)
    elif (not isinstance(workspace, torch.Tensor) or workspace.device != Q.device or workspace.dtype != torch.uint8
          or workspace.numel() < need or not workspace.is_contiguous()):
        raise _lib.FlashAttentionError(-6, f"workspace must be a contiguous uint8 tensor of at least {need} bytes on {Q.device}")
    dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
    _lib.check(lib.fa_v1_backward(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), dO.data_ptr(), LSE.data_ptr(),
                                  dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(), B, H, L, d, _DTYPES[Q.dtype],
                                  1 if causal else 0, workspace.data_ptr(), workspace.numel(), _stream()))
    if sync:
        torch.cuda.current_stream().synchronize()
    return dQ, dK, dV


def backward_workspace_bytes(B: int, H: int, L: int) -> int:
    return _lib.load().fa_v1_backward_workspace_bytes(B, H, L)


def naive_attention_reference(Q, K, V, max_workspace_bytes: int = 1 << 30):
    """The package's independent evaluation (fa_naive_attention): materialised scores, CUDA-core fp32 / fp64 math.
    Q [...,Lq,d], K, V [...,Lk,d] float32 or float64 CUDA tensors (any d) -> O like Q."""
    if not (Q.is_cuda and K.is_cuda and V.is_cuda):
        raise RuntimeError("flash-attention B200 path needs CUDA tensors: there is no CPU fallback")
    if Q.dtype not in (torch.float32, torch.float64) or not (Q.dtype == K.dtype == V.dtype):
        raise _lib.FlashAttentionError(-2, "naive_attention_reference computes in float32 or float64")
    if Q.dim() < 2 or K.shape != V.shape or Q.shape[:-2] != K.shape[:-2] or Q.shape[-1] != K.shape[-1]:
        raise _lib.FlashAttentionError(-1, "Q must be [...,Lq,d] and K, V [...,Lk,d]")
    Q, K, V = Q.contiguous(), K.contiguous(), V.contiguous()
    Lq, d = Q.shape[-2:]
    Lk = K.shape[-2]
    n_heads = Q.numel() // (Lq * d)
    dt = _lib.FA_DTYPE_F64 if Q.dtype == torch.float64 else _lib.FA_DTYPE_F32
    lib = _lib.load()
    with torch.cuda.device(Q.device):
        one = lib.fa_naive_attention_workspace_bytes(1, Lq, Lk, dt)
        heads_at_once = max(1, min(n_heads, max_workspace_bytes // one))
        ws = torch.empty(one * heads_at_once, dtype=torch.uint8, device=Q.device)
        O = torch.empty_like(Q)
        _lib.check(lib.fa_naive_attention(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), n_heads, Lq, Lk, d, dt,
                                          ws.data_ptr(), ws.numel(), _stream()))
        torch.cuda.current_stream().synchronize()
    return O


def flash_attention_host(Qh, Kh, Vh, Oh=None, variant: int = 0, kv_per_split: int = 0):
    """Host-buffer path (H2D x3 + kernel + D2H inside the library), like the reference drivers do around their
    launchers (flash_attention_v1/CUDA/driver.cu:184-247). Tensors are CPU tensors, ideally pinned."""
    if Qh.is_cuda:
        raise RuntimeError("flash_attention_host takes host tensors")
    if Qh.dim() != 4 or Qh.dtype not in _DTYPES:
        raise _lib.FlashAttentionError(-1, "Qh must be a [B,H,L,d] host tensor of a supported dtype")
    B, H, L, d = Qh.shape
    cpu = torch.device("cpu")
    for t, what in ((Qh, "Qh"), (Kh, "Kh"), (Vh, "Vh")):
        _check_buffer(t, Qh.shape, Qh.dtype, cpu, what)
    if Oh is None:
        Oh = torch.empty_like(Qh).pin_memory()
    else:
        _check_buffer(Oh, Qh.shape, Qh.dtype, cpu, "Oh")
    lib = _lib.load()
    _lib.check(lib.fa_forward_host(variant, Qh.data_ptr(), Kh.data_ptr(), Vh.data_ptr(), Oh.data_ptr(), B, H, L, d,
                                   kv_per_split, _DTYPES[Qh.dtype]))
    return Oh
