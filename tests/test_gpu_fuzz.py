"""Seeded sweep over shapes, decay regimes, initial states and entry points: the tensor-core route (incl. the
time-segmented one and the per-stream exact fallback) against the exact SIMT kernels.  `pytest -m gpu`."""
import random

import pytest
import torch

from tests.util import make_inputs, relrms

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(rng):
    H = rng.choice([1, 2, 3, 5])
    kind = rng.choice(["short", "short", "ragged", "long", "segmented"])
    if kind == "short":
        B, T = rng.choice([1, 2, 4]), rng.randint(1, 200)
    elif kind == "ragged":
        B, T = rng.choice([1, 3]), 64 * rng.randint(1, 12) + rng.choice([1, 17, 31, 63])
    elif kind == "long":
        B, T = 1, 64 * rng.randint(20, 60)
    else:
        B, T = rng.choice([1, 2]), 64 * rng.randint(32, 80) + rng.choice([0, 0, 5, 40])      # T >= 2048: segments
    decay = rng.choice(["model", "model", "randn"])
    state = rng.choice([None, "shared", "batched"])
    hazard = rng.random() < 0.25 and T > 40
    return B, T, H, decay, state, hazard


@pytest.mark.parametrize("seed", list(range(36)))
def test_fuzz_tc_vs_simt(seed):
    import rwkv_lm_ext_b200 as M
    rng = random.Random(1000 + seed)
    B, T, H, decay, state, hazard = _case(rng)
    C = H * 64
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=seed, decay=decay, device=DEV)
    if hazard:
        t0 = rng.randint(0, T - 33)
        hh = rng.randrange(H)
        w[rng.randrange(B), t0:t0 + 32, 64 * hh:64 * hh + 8] = 3.5
    g = torch.Generator().manual_seed(seed)
    s0 = None
    if state == "shared":
        s0 = (torch.randn(H, 64, 64, generator=g) * 0.3).bfloat16().to(DEV)
    elif state == "batched":
        s0 = (torch.randn(B, H, 64, 64, generator=g) * 0.3).bfloat16().to(DEV)

    def run(impl):
        M.set_impl(impl)
        try:
            leaves = [t.clone().requires_grad_(True) for t in (r, k, v, w, u)]
            s = None if s0 is None else s0.clone().requires_grad_(True)
            if s is None:
                y, sT = M.RUN_CUDA_RWKV6(B, T, C, H, *leaves), None
            elif state == "shared":
                y, sT = M.WKV_6STATE.apply(B, T, C, H, *leaves, s), None
            else:
                y, sT = M.WKV_6STATE_INFCTX.apply(B, T, C, H, *leaves, s.clone())
            y.backward(gy)
            return y.detach(), [t.grad for t in leaves], (None if s is None else s.grad), sT
        finally:
            M.set_impl("auto")

    y, grads, gs, sT = run("auto")
    ys, grads_s, gs_s, sTs = run("simt")
    tag = f"B={B} T={T} H={H} {decay} state={state} hazard={hazard}"
    assert torch.isfinite(y.float()).all(), tag
    assert relrms(y, ys) < 8e-3, tag
    for name, a, b_ in zip(("gr", "gk", "gv", "gw", "gu"), grads, grads_s):
        assert torch.isfinite(a.float()).all(), f"{tag} {name}"
        assert relrms(a, b_) < 1.2e-2 or b_.float().norm().item() < 1e-3, f"{tag} {name} {relrms(a, b_):.3g}"
    if gs is not None:
        assert relrms(gs, gs_s) < 1.2e-2, tag + " gs"
    if sT is not None:
        assert relrms(sT, sTs) < 8e-3, tag + " sT"
