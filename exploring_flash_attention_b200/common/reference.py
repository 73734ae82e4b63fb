"""Drop-in for the reference's common/reference.py: same three names, same arguments, same prints and
AssertionError behaviour (common/reference.py:7-21, :24-78, :81-96).

`check_accuracy` / `print_comparison` are host-side validation logic restated from the reference.
`naive_attention` keeps its (Q,K,V)->O contract for one [L,d] head and stays what it is in the reference — the thing
the kernels are compared WITH: it runs fa_naive_attention (csrc/fa_naive_sm100.cuh), which materialises the [L,L] score
matrix and evaluates reference.py:16-21 step by step on the CUDA cores in fp32 (float16 / float32 buffers) or fp64
(float64 buffers, like the reference's own float64 runs).  It shares no code, no tensor-core path and no online-softmax
recurrence with the fused kernels, and accepts any head dim.  This package has no CPU compute path; the float64 CPU
oracle the tests judge against lives in oracle/, outside the package.
"""
from __future__ import annotations

import numpy as np


def naive_attention(Q, K, V):
    """softmax(Q K^T / sqrt(d)) V for [L,d] NumPy arrays; returns [L,d] in Q's dtype. Runs on the current CUDA device,
    independently of the fused kernels (materialised scores, fp32 / fp64 CUDA-core math)."""
    import torch

    from .. import ops
    from .._numpy_bridge import require_cuda
    dev = require_cuda()
    Qa = np.asarray(Q)
    compute = torch.float64 if Qa.dtype == np.float64 else torch.float32
    q, k, v = (torch.from_numpy(np.ascontiguousarray(np.asarray(x))).to(dev, dtype=compute) for x in (Q, K, V))
    O = ops.naive_attention_reference(q, k, v)
    return O.cpu().numpy().astype(Qa.dtype)


def check_accuracy(output, reference, config_str="", max_abs_tol=1e-2, max_rel_tol=0.5, mean_rel_tol=0.05):
    """Prints max-abs, filtered max-rel (|reference| > 1e-3) and mean-rel error; raises AssertionError when a
    tolerance is exceeded (common/reference.py:24-78, defaults identical)."""
    output = np.asarray(output)
    reference = np.asarray(reference)
    errors = []
    diff = np.abs(output - reference).max()
    if config_str:
        print(f"\nMax absolute difference ({config_str}):", diff)
    else:
        print("\nMax absolute difference:", diff)
    if diff > max_abs_tol:
        errors.append(f"Max absolute difference {diff:.6f} exceeds tolerance {max_abs_tol}")
    mask = np.abs(reference) > 1e-3
    if mask.any():
        rel = np.abs(output[mask] - reference[mask]) / np.abs(reference[mask])
        print("Max relative difference (|reference| > 1e-3):", rel.max())
        if rel.max() > max_rel_tol:
            errors.append(f"Max relative difference {rel.max():.6f} exceeds tolerance {max_rel_tol}")
        print("Mean relative error (|reference| > 1e-3):", rel.mean())
        if rel.mean() > mean_rel_tol:
            errors.append(f"Mean relative error {rel.mean():.6f} exceeds tolerance {mean_rel_tol}")
    else:
        print("Warning: No values > 1e-3 for relative error calculation")
    if not errors:
        print("✓ PASSED - All error metrics within tolerances")
        return
    print("✗ FAILED - Error tolerance(s) exceeded:")
    for e in errors:
        print(f"  - {e}")
    raise AssertionError(f"Accuracy check failed: {'; '.join(errors)}")


def print_comparison(output, reference, num_rows=3, num_cols=5):
    """Side-by-side corner print (common/reference.py:81-96)."""
    print("Output shape:", output.shape)
    print(f"First {num_rows} rows (output):")
    print(output[:num_rows, :num_cols])
    print(f"\nFirst {num_rows} rows (reference):")
    print(reference[:num_rows, :num_cols])
