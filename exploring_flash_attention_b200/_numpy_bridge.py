"""NumPy <-> device bridge for the reference-named Python entry points.

The reference scripts pass NumPy buffers (float16 or float64; 1-D flattened [L*d] or 2-D [L,d]) for ONE head.
Here the same calls run on the GPU through the C ABI: the buffers are staged to device tensors, the sm_100a kernel
runs, and the result is copied back into the caller's output buffer in its own dtype.  float64 / float32 buffers are
computed as FA_DTYPE_F32 (tf32 tensor-core products, fp32 accumulation): B200 has no fp64 tensor path.
"""
from __future__ import annotations

import numpy as np
import torch

_NP2TORCH = {np.dtype(np.float16): torch.float16, np.dtype(np.float32): torch.float32,
             np.dtype(np.float64): torch.float32}


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("exploring_flash_attention_b200 runs on a CUDA device only (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def to_device_head(x, L, d):
    """NumPy [L*d] / [L,d] (or torch tensor) -> contiguous CUDA tensor [1,1,L,d] in the compute dtype."""
    dev = require_cuda()
    if isinstance(x, torch.Tensor):
        t = x.reshape(1, 1, L, d)
        if t.dtype == torch.float64:
            t = t.float()
        return t.to(dev).contiguous()
    a = np.asarray(x)
    if a.dtype not in _NP2TORCH:
        raise TypeError(f"unsupported buffer dtype {a.dtype}")
    return torch.from_numpy(np.ascontiguousarray(a.reshape(1, 1, L, d))).to(dev, dtype=_NP2TORCH[a.dtype])


def store_head(dst, O_dev, L, d):
    """Write a device result [1,1,L,d] into the caller's NumPy / torch buffer in place (any shape holding L*d)."""
    if isinstance(dst, torch.Tensor):
        dst.reshape(L, d).copy_(O_dev.reshape(L, d).to(dst.dtype))
        return
    flat = dst.reshape(-1)
    if flat.base is None and flat is not dst and not np.shares_memory(flat, dst):
        raise ValueError("output buffer must be contiguous so it can be written in place")
    flat[: L * d] = O_dev.reshape(-1).float().cpu().numpy().astype(dst.dtype)
