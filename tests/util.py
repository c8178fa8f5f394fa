"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated tolerances: bf16 outputs / gradients against the fp64 oracle
BF16_MAXABS_REL = 2.0 ** -7     # max|err| <= 2^-7 * max|ref| + 1e-2
BF16_MAXABS_ABS = 1e-2
BF16_RELRMS = 1e-2              # ||err|| / ||ref|| <= 1e-2
STATE_RELRMS = 1e-3             # fp32 carried states


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def relrms(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def assert_bf16_close(got, ref, what="", relrms_tol=BF16_RELRMS, maxabs_rel=BF16_MAXABS_REL):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    err = (got - ref).abs().max().item()
    bound = maxabs_rel * ref.abs().max().item() + BF16_MAXABS_ABS
    rr = relrms(got, ref)
    assert err <= bound, f"{what}: max abs err {err:.4g} > {bound:.4g} (relrms {rr:.3g})"
    assert rr <= relrms_tol or ref.norm().item() < 1e-6, f"{what}: rel-RMS {rr:.4g} > {relrms_tol}"


from rwkv_lm_ext_b200.synthetic import make_inputs  # noqa: E402,F401
