"""GPU (-m gpu): the dense d = 128 forward on CTA pairs — K1P (csrc/fa_fwd_pair_sm100.cuh, FA_B200_FWD_PAIR=1) and the
experimental K1Q (csrc/fa_fwd_pair2_sm100.cuh, =2) — against the float64 oracle through the reference-shaped entry
points, next to K1 (=0).  The variable is read once per process (compiled default: FA_FWD_PAIR_DEFAULT in
csrc/fa_api.cu), so each check runs in a subprocess; whatever needs masks, lengths or a split must keep going to K1 in
the same process.  Tolerance: 2e-3 (bf16 / fp16, BASELINE.json).
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

_CHECK = """
import sys
sys.path.insert(0, {root!r})
import numpy as np, torch
from exploring_flash_attention_b200 import ops
from oracle import reference
f = lambda x: x.float().cpu().numpy()
g = torch.Generator().manual_seed(11)
worst = 0.0
for (B, H, L, dt) in ((1, 2, 256, torch.bfloat16), (1, 3, 333, torch.bfloat16), (1, 2, 100, torch.float16),
                      (2, 1, 640, torch.float16), (2, 4, 1024, torch.bfloat16), (1, 1, 1, torch.bfloat16),
                      (3, 5, 777, torch.bfloat16)):
    Q, K, V = ((torch.rand((B, H, L, 128), generator=g) * 2 - 1).to(dt).cuda() for _ in range(3))
    O, lse = ops.flash_attention_v1_ex(Q, K, V, return_lse=True, sync=True)
    assert not torch.isnan(O).any()
    ref = reference.naive_attention_batched_f64(f(Q), f(K), f(V)).reshape(B, H, L, 128)
    worst = max(worst, float(np.abs(f(O) - ref).max()))
    S = torch.einsum("bhqd,bhkd->bhqk", Q.double(), K.double()) / 128 ** 0.5
    assert float((lse.double() - torch.logsumexp(S, -1)).abs().max()) <= 1e-4
    assert torch.equal(O, ops.flash_attention_v1(Q, K, V, sync=True))
    assert torch.equal(O, ops.flash_attention_v1_tiled_d(Q, K, V, sync=True))
    # masks, lengths and splits are not served by the pair kernel: these must still be right (K1)
    Oc = ops.flash_attention_v1_ex(Q, K, V, causal=True, sync=True)
    refc = np.stack([reference.naive_attention_ex_f64(q, k, v, causal=True)[0] for q, k, v in
                     zip(f(Q).reshape(-1, L, 128), f(K).reshape(-1, L, 128), f(V).reshape(-1, L, 128))])
    # early causal rows average over few keys and are O(1): allow their storage rounding (as test_parity_gpu.py does)
    worst = max(worst, float(np.abs(f(Oc).reshape(-1, L, 128) - refc).max()) / max(1.0, 2 * float(np.abs(refc).max())))
    O2 = ops.flash_attention_v2(Q, K, V, 128, sync=True)
    worst = max(worst, float(np.abs(f(O2) - ref).max()))
print("WORST", worst)
assert worst <= 2e-3, worst
"""


@pytest.mark.parametrize("mode", ["1", "2", "0"])
def test_dense_d128_forward_on_cta_pairs(mode):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    root = str(Path(__file__).resolve().parents[1])
    env = dict(os.environ, FA_B200_FWD_PAIR=mode)
    r = subprocess.run([sys.executable, "-c", _CHECK.format(root=root)], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
