"""Drop-in for the reference's common/reference.py: same three names, same arguments, same prints and
AssertionError behaviour (common/reference.py:7-21, :24-78, :81-96).

`check_accuracy` / `print_comparison` are host-side validation logic restated from the reference.
`naive_attention` keeps its (Q,K,V)->O contract for one [L,d] head but is evaluated on the GPU by the fused-tile
kernel through the C ABI — this package has no CPU compute path.  (The float64 CPU oracle the tests judge against
lives in oracle/, outside the package.)
"""
from __future__ import annotations

import numpy as np


def naive_attention(Q, K, V):
    """softmax(Q K^T / sqrt(d)) V for [L,d] NumPy arrays; returns [L,d] in Q's dtype. Runs on the current CUDA device."""
    from .. import ops
    from .._numpy_bridge import to_device_head
    L, d = Q.shape
    q, k, v = (to_device_head(x, L, d) for x in (Q, K, V))
    O = ops.flash_attention_v1(q, k, v, sync=True)
    return O.reshape(L, d).float().cpu().numpy().astype(np.asarray(Q).dtype)


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
