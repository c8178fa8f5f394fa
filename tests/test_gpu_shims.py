"""The drop-in `cuda/*_op.cpp` modules (shim/cuda/, built the way the reference's own `load()` builds them) called
with the reference's pybind signatures, against (a) this package's operators, (b) the fp64 oracle and (c) the
reference's OWN CUDA kernels (oracle/_ref, compiled from /root/reference/cuda/*.cu) for the forwards whose reference
implementation is well defined: wkv6, wkv6state, wkv6infctx, wkv6_bi (with the reference's quirk flag) and rwkv6.
`pytest -m gpu`."""
import os
import sys

import pytest
import torch

from tests.util import STATE_RELRMS, assert_bf16_close, make_inputs, relrms

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shim"))


@pytest.fixture(scope="module")
def shims():
    import build_shims
    os.environ.setdefault("WKV6_B200_LIB", os.path.join(ROOT, "rwkv_lm_ext_b200", "libwkv6_b200.so"))
    return {n: build_shims.load_shim(n) for n in build_shims.MODULES}      # prebuilt by __graft_entry__.build(): import only


@pytest.fixture(scope="module")
def O():
    from oracle import wkv6_oracle
    return wkv6_oracle


@pytest.fixture(scope="module")
def R():
    from oracle import ref_cuda
    return ref_cuda


def test_wkv6_module_forward_backward(shims, O):
    """wkv6.forward / backward exactly as src/model.py:205-231 calls them (fp32 ew, pre-allocated outputs, gu [B,C])."""
    B, T, H = 2, 200, 2
    C = H * 64
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=31, decay="model")
    ref = O.wkv6_backward(r, k, v, w, u, gy)
    r, k, v, w, u, gy = (t.to(DEV) for t in (r, k, v, w, u, gy))
    ew = (-(w.float().exp())).contiguous()
    y = torch.empty((B, T, C), device=DEV, dtype=torch.bfloat16)
    shims["wkv6"].forward(B, T, C, H, r, k, v, ew, u, y)
    gr, gk, gv, gw = (torch.empty((B, T, C), device=DEV, dtype=torch.bfloat16) for _ in range(4))
    gu = torch.empty((B, C), device=DEV, dtype=torch.bfloat16)
    shims["wkv6"].backward(B, T, C, H, r, k, v, ew, u, gy, gr, gk, gv, gw, gu)
    assert_bf16_close(y, ref["y"], "shim y")
    for a, key in zip((gr, gk, gv, gw), ("gr", "gk", "gv", "gw")):
        assert_bf16_close(a, ref[key], f"shim {key}")
    assert_bf16_close(torch.sum(gu, 0).view(H, 64), ref["gu"], "shim gu")


@pytest.mark.parametrize("name", ["wkv6state", "wkv6infctx"])
def test_state_modules_against_the_reference_cuda_forward(shims, O, R, name):
    if not R.available(name):
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    B, T, H = 2, 192, 3
    C = H * 64
    r, k, v, w, u, _ = make_inputs(B, T, H, seed=41, decay="model", device=DEV)
    g = torch.Generator().manual_seed(7)
    shape = (H, 64, 64) if name == "wkv6state" else (B, H, 64, 64)
    s0 = (torch.randn(*shape, generator=g) * 0.5).bfloat16().to(DEV)
    s_ref, s_our = s0.clone(), s0.clone()
    y_ref = R.state_forward(name, r, k, v, w, u, s_ref)                      # the reference's own kernel
    torch.cuda.synchronize()
    y = torch.empty_like(r)
    shims[name].forward(B, T, C, H, r, k, v, w, u, s_our, y)
    assert_bf16_close(y, y_ref.float(), f"{name} y vs reference CUDA", relrms_tol=6e-3)
    fn = O.wkv6state_forward if name == "wkv6state" else O.wkv6infctx_forward
    out = fn(*(t.cpu() for t in (r, k, v, w, u)), s0.cpu())
    y_or = out[0] if isinstance(out, tuple) else out
    assert_bf16_close(y, y_or, f"{name} y vs oracle")
    if name == "wkv6infctx":                                                  # final state written back in place, bf16
        assert_bf16_close(s_our, s_ref.float(), "infctx final state vs reference CUDA", relrms_tol=6e-3)
        assert_bf16_close(s_our, out[1], "infctx final state vs oracle")
    else:
        assert torch.equal(s_our, s0)                                         # the shared state is read-only


def test_bi_module_against_the_reference_cuda_forward(shims, O, R):
    if not R.available("wkv6_bi"):
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    B, T, H = 3, 64, 2
    C = H * 64
    r, k, v, w, u, _ = make_inputs(B, T, H, seed=51, decay="model", device=DEV)
    mask = torch.ones(B, T, dtype=torch.int32, device=DEV)
    mask[0, 60:] = 0                       # cuda/wkv6_bi.py:66-67
    mask[1, 40:] = 0                       # row 2 has no zero: the reference runs no reverse pass there (its quirk)
    y_ref = R.bi_forward(mask, r, k, v, w, u)
    torch.cuda.synchronize()
    ew = (-(w.float().exp())).contiguous()
    y = torch.empty_like(r)
    shims["wkv6_bi"].forward(B, T, C, H, mask, r, k, v, ew, u, y)
    # rows WITH padding: same function as the reference's kernel (up to the first masked position)
    for b, p in ((0, 60), (1, 40)):
        assert_bf16_close(y[b, :p + 1], y_ref[b, :p + 1].float(), f"bi row {b} vs reference CUDA", relrms_tol=8e-3)
        assert y[b, p + 1:].abs().max().item() == 0.0
    y_or = O.wkv6_bi_forward(mask.cpu(), *(t.cpu() for t in (r, k, v, w, u)))
    assert_bf16_close(y, y_or, "bi y vs oracle")
    # the row without a zero: the reference's result equals the oracle with ref_quirks (causal pass only)
    y_q = O.wkv6_bi_forward(mask.cpu(), *(t.cpu() for t in (r, k, v, w, u)), ref_quirks=True)
    assert_bf16_close(y_ref[2].float(), y_q[2], "reference CUDA row without padding == quirk", relrms_tol=8e-3)


@pytest.mark.parametrize("T", [1, 300])
def test_rwkv6_module_against_the_reference_cuda_forward(shims, R, T):
    if not R.available("rwkv6"):
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    H = 4
    C = H * 64
    r, k, v, w, u, _ = make_inputs(1, T, H, seed=61, decay="model", device=DEV)
    r, k, v, w = (t[0].contiguous() for t in (r, k, v, w))
    g = torch.Generator().manual_seed(3)
    st0 = (torch.randn(H, 64, 64, generator=g) * 0.3).to(DEV)
    decay = torch.exp(-torch.exp(w.float())).contiguous()                    # src/model_run.py:64
    st_ref, st_our = st0.clone(), st0.clone()
    y_ref = R.rwkv6_forward_bf16(st_ref, r, k, v, decay, u)
    torch.cuda.synchronize()
    y = torch.empty_like(r)
    shims["rwkv6"].forward_bf16(1, T, C, H, st_our, r, k, v, decay, u, y)
    assert_bf16_close(y, y_ref.float(), "rwkv6 y vs reference CUDA", relrms_tol=6e-3)
    assert relrms(st_our, st_ref) < STATE_RELRMS, relrms(st_our, st_ref)
