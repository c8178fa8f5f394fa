"""The whole time-mix layer (rwkv_lm_ext_b200.tmix) against an fp64 restatement of
RWKV_Tmix_x060.forward (src/model.py:434-477, infctx :738-781, state tuning :560-584) built from the
oracle's pieces: output and every parameter / input gradient.  `pytest -m gpu`."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import relrms

pytestmark = pytest.mark.gpu
DEV = "cuda"


def assert_chain_close(got, ref, what, relrms_tol):
    """A dozen bf16 ops deep: rel-RMS bound plus a loose max-abs bound (5 % of the largest value)."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all(), what
    rr = relrms(got, ref)
    assert rr <= relrms_tol, f"{what}: rel-RMS {rr:.4g} > {relrms_tol}"
    err, top = (got - ref).abs().max().item(), ref.abs().max().item()
    assert err <= 0.05 * top + 1e-2, f"{what}: max abs err {err:.4g} vs max |ref| {top:.4g}"


def make_layer(M, C, H, seed, state_tuning=False):
    g = torch.Generator().manual_seed(seed)
    layer = M.Tmix_x060(C, H, state_tuning=state_tuning)
    with torch.no_grad():
        for name in ("time_maa_x", "time_maa_w", "time_maa_k", "time_maa_v", "time_maa_r", "time_maa_g"):
            getattr(layer, name).copy_(torch.rand(1, 1, C, generator=g))
        layer.time_maa_w1.copy_(torch.randn(C, 160, generator=g) * 0.05)
        layer.time_maa_w2.copy_(torch.randn(5, 32, C, generator=g) * 0.05)
        layer.time_decay.copy_(-6 + 5 * torch.rand(1, 1, C, generator=g))
        layer.time_decay_w1.copy_(torch.randn(C, 64, generator=g) * 0.05)
        layer.time_decay_w2.copy_(torch.randn(64, C, generator=g) * 0.05)
        layer.time_faaaa.copy_(torch.randn(H, 64, generator=g) * 0.3)
        if state_tuning:
            layer.time_state.copy_(torch.randn(H, 64, 64, generator=g) * 0.2)
        for lin in (layer.receptance, layer.key, layer.value, layer.gate, layer.output):
            lin.weight.copy_(torch.randn(C, C, generator=g) / C ** 0.5)
        layer.ln_x.weight.copy_(0.5 + torch.rand(C, generator=g))
        layer.ln_x.bias.copy_(torch.randn(C, generator=g) * 0.1)
    return layer.bfloat16()


def reference_forward(O, p, x, H, shift_state=None, wkv_state=None):
    """fp64, differentiable; p = dict of fp64 leaves with the layer's parameter names."""
    B, T, C = x.shape
    maa5 = torch.cat([p[n].view(1, C) for n in ("time_maa_w", "time_maa_k", "time_maa_v", "time_maa_r", "time_maa_g")])
    xw, xk, xv, xr, xg = O.tmix_ddlerp(x, p["time_maa_x"].view(C), maa5, p["time_maa_w1"], p["time_maa_w2"], shift_state)
    r, k, v = xr @ p["receptance.weight"].T, xk @ p["key.weight"].T, xv @ p["value.weight"].T
    g = F.silu(xg @ p["gate.weight"].T)
    w = O.tmix_decay(xw, p["time_decay"], p["time_decay_w1"], p["time_decay_w2"])
    s0 = None
    if "time_state" in p:
        s0 = p["time_state"].transpose(-1, -2).unsqueeze(0).expand(B, -1, -1, -1)     # [H,val,key] -> [B,H,key,val]
    if wkv_state is not None:
        s0 = wkv_state.transpose(-1, -2)
    y, sT = O.wkv6_recurrence(r, k, v, w, p["time_faaaa"], s0)
    out = O.groupnorm_gate(y, g, p["ln_x.weight"], p["ln_x.bias"], H, 1e-5 * 64) @ p["output.weight"].T
    return out, sT.transpose(-1, -2)


@pytest.mark.parametrize("flavour", ["plain", "states", "infctx"])
def test_tmix_layer_forward_backward(flavour):
    import rwkv_lm_ext_b200 as M
    from oracle import wkv6_oracle as O
    B, T, H = 2, 70, 2
    C = H * 64
    layer = make_layer(M, C, H, 5, state_tuning=flavour == "states").to(DEV)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, T, C, generator=g).bfloat16()
    gout = torch.randn(B, T, C, generator=g).bfloat16()
    shift = torch.randn(B, C, generator=g).bfloat16() if flavour == "infctx" else None
    wkv = (torch.randn(B, H, 64, 64, generator=g) * 0.2).bfloat16() if flavour == "infctx" else None

    p = {n: t.detach().cpu().double().requires_grad_(True) for n, t in layer.state_dict().items()}
    x64 = x.double().requires_grad_(True)
    ref, ref_state = reference_forward(O, p, x64, H, None if shift is None else shift.double(),
                                       None if wkv is None else wkv.double())
    (ref * gout.double()).sum().backward()

    xd = x.to(DEV).requires_grad_(True)
    if flavour == "infctx":
        out, (last_x, new_state) = layer(xd, (shift.to(DEV), wkv.to(DEV)))
        assert torch.equal(last_x, xd[:, -1])
        assert relrms(new_state, ref_state) < 1e-2
    else:
        out = layer(xd)
    assert_chain_close(out, ref, f"tmix {flavour} out", 2e-2)
    out.backward(gout.to(DEV))
    assert_chain_close(xd.grad, x64.grad, f"tmix {flavour} gx", 3e-2)
    for name, prm in layer.named_parameters():
        want = p[name].grad
        assert prm.grad is not None, name
        rr = relrms(prm.grad, want)
        assert rr < 4e-2, f"tmix {flavour} grad {name}: rel-RMS {rr:.3g}"


def test_tmix_binds_onto_a_foreign_module():
    """INTEGRATION: tmix_x060_forward bound as a method of a module that only shares parameter names."""
    import types
    import rwkv_lm_ext_b200 as M
    H, C = 2, 128
    src = make_layer(M, C, H, 9).to(DEV)

    class Foreign(torch.nn.Module):
        pass
    f = Foreign()
    for n, prm in src.named_parameters(recurse=False):
        setattr(f, n, prm)
    for n, mod in src.named_children():
        setattr(f, n, mod)
    f.forward = types.MethodType(M.tmix_x060_forward, f)
    x = torch.randn(2, 33, C, device=DEV).bfloat16()
    with torch.no_grad():
        assert torch.equal(f(x), src(x))


def test_tmix_infctx_chunks_equal_one_long_pass():
    """Two infctx calls chained through BlockStateList (fp32 carry) == one plain call on the whole sequence."""
    import rwkv_lm_ext_b200 as M
    B, T, H = 2, 256, 2
    C = H * 64
    layer = make_layer(M, C, H, 21).to(DEV)
    x = torch.randn(B, T, C, generator=torch.Generator().manual_seed(22)).bfloat16().to(DEV)
    states = M.BlockStateList.create(1, B, C, H, DEV, torch.bfloat16, wkv_dtype=torch.float32)
    with torch.no_grad():
        full = layer(x)
        outs = []
        for lo in (0, 128):
            out, tms = layer(x[:, lo:lo + 128].contiguous(), states[0].time_mix_state)
            assert isinstance(tms, M.TimeMixState)
            states[0] = M.BlockState(tms, states[0].channel_mix_state)
            outs.append(out)
    assert relrms(torch.cat(outs, 1), full) < 6e-3
