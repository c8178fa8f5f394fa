"""The C port of the oracle (oracle/wkv6_oracle.c) against the fp64 PyTorch oracle.  CPU only."""
import torch

from oracle import c_oracle as CO
from oracle import wkv6_oracle as O


def _inputs(B, T, H, seed, decay="randn"):
    g = torch.Generator().manual_seed(seed)
    C = H * 64
    r, k, v, gy = (torch.randn(B, T, C, generator=g) for _ in range(4))
    if decay == "randn":
        w = torch.randn(B, T, C, generator=g)
    else:
        w = torch.rand(B, T, C, generator=g) * 5 - 6
    u = torch.randn(H, 64, generator=g) * 0.3
    return r, k, v, w, u, gy


def _relerr(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def test_c_forward_backward_zero_state():
    r, k, v, w, u, gy = _inputs(2, 37, 2, 0)
    ref = O.wkv6_backward(r, k, v, w, u, gy)
    assert _relerr(CO.forward(r, k, v, w, u), ref["y"]) < 1e-5
    g = CO.backward(r, k, v, w, u, gy)
    for key in ("gr", "gk", "gv", "gw"):
        assert _relerr(g[key], ref[key]) < 1e-4, key
    assert _relerr(g["gu"].sum(0).view(2, 64), ref["gu"]) < 1e-4
    assert g["gw"][:, 0].abs().max() == 0 and g["gw"][:, -1].abs().max() == 0


def test_c_w_kinds_agree():
    r, k, v, w, u, gy = _inputs(1, 20, 1, 1, "decay")
    y0 = CO.forward(r, k, v, w, u, w_kind=0)
    y1 = CO.forward(r, k, v, -torch.exp(w), u, w_kind=1)
    y2 = CO.forward(r, k, v, torch.exp(-torch.exp(w)), u, w_kind=2)
    assert _relerr(y1, y0) < 1e-6 and _relerr(y2, y0) < 1e-5


def test_c_state_variants():
    r, k, v, w, u, gy = _inputs(2, 25, 2, 2, "decay")
    g = torch.Generator().manual_seed(9)
    s = torch.randn(2, 2, 64, 64, generator=g) * 0.5
    ref = O.wkv6_backward(r, k, v, w, u, gy, s=s, s_layout="infctx")
    y, sT = CO.forward(r, k, v, w, u, s0=s, want_state=True)
    yo, so = O.wkv6infctx_forward(r, k, v, w, u, s)
    assert _relerr(y, yo) < 1e-5 and _relerr(sT, so) < 1e-5
    gc = CO.backward(r, k, v, w, u, gy, s0=s)
    for key in ("gr", "gk", "gv", "gw", "gs"):
        assert _relerr(gc[key], ref[key]) < 1e-4, key
    # shared [H,64,64] state: gs partials sum over the batch (src/model.py:182)
    ref2 = O.wkv6_backward(r, k, v, w, u, gy, s=s[0], s_layout="state")
    gc2 = CO.backward(r, k, v, w, u, gy, s0=s[0])
    assert _relerr(gc2["gs"].sum(0), ref2["gs"]) < 1e-4
    assert _relerr(gc2["gw"], ref2["gw"]) < 1e-4
