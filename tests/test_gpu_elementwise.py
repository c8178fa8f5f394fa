"""Memory-bound neighbours of the recurrence on the GPU against the oracle / torch expressions.
Integer results are bit-exact.  `pytest -m gpu`."""
import pytest
import torch

from tests.util import assert_bf16_close, load_golden, relrms

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def M():
    import rwkv_lm_ext_b200 as M
    M.load()
    return M


@pytest.fixture(scope="module")
def O():
    from oracle import wkv6_oracle
    return wkv6_oracle


def test_golden_mask_rev_idx_eos(M):
    c = load_golden("mask_rev_idx")
    idx = c["idx"].to(DEV)
    mask, rev = M.create_mask_and_rev_idx(idx, emb_id=1, pad_id=0)
    assert mask.dtype == torch.int32 and torch.equal(mask.cpu(), c["mask"])
    assert rev.dtype == torch.int64 and torch.equal(rev.cpu(), c["rev_idx"])
    assert torch.equal(M.eos_index(idx, 1).cpu(), c["eos_pos"])


@pytest.mark.parametrize("B,T", [(1, 1), (3, 7), (64, 512), (5, 4096), (2, 100000)])
def test_eos_index_bit_exact(M, B, T):
    g = torch.Generator().manual_seed(T)
    idx = torch.randint(2, 65536, (B, T), generator=g)
    for b in range(B):
        if b % 3 == 0:
            idx[b, T - 1] = 1                       # eos last
        elif b % 3 == 1:
            idx[b, int(torch.randint(0, T, (1,), generator=g))] = 1
            idx[b, T // 2] = 1                      # duplicates: first wins
        # else: absent -> 0
    want = torch.eq(idx, 1).int().argmax(-1)
    got = M.eos_index(idx.to(DEV), 1)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), want)


def test_gather_and_reverse_bit_exact(M):
    g = torch.Generator().manual_seed(0)
    B, T, D = 7, 33, 768
    x = torch.randn(B, T, D, generator=g).bfloat16().to(DEV)
    idx = torch.randint(2, 100, (B, T), generator=g)
    idx[:, -1] = 1
    idx[2, 5] = 1
    idx[2, 6:] = 0
    idx = idx.to(DEV)
    rows, pos = M.eos_gather(x, idx, 1)
    assert torch.equal(rows, x[torch.arange(B, device=DEV), pos])
    mask, rev = M.create_mask_and_rev_idx(idx)
    want = torch.gather(x, 1, rev.unsqueeze(-1).expand(-1, -1, D))
    got = M.reverse_x(x, rev)
    assert torch.equal(got, want)
    assert torch.equal(M.reverse_x(got, rev), x)        # the permutation is an involution
    # both directions stacked as one batch of 2B rows (the bidirectional encoders' input), one kernel
    from rwkv_lm_ext_b200 import heads
    assert torch.equal(heads.stack_reversed(x, rev), torch.cat([x, want], 0))
    xg = x.clone().requires_grad_(True)                  # under autograd: the differentiable pieces
    heads.stack_reversed(xg, rev).sum().backward()
    assert torch.equal(xg.grad, torch.full_like(x, 2.0))
    # odd D falls back to scalar copies
    x2 = torch.randn(2, 5, 13, generator=g).bfloat16().to(DEV)
    p2 = torch.tensor([4, 0], device=DEV)
    assert torch.equal(M.gather_rows(x2, p2), x2[torch.arange(2, device=DEV), p2])


def test_pooling_golden_and_random(M, O):
    c = load_golden("pooling")
    x = c["x"].bfloat16().to(DEV)
    L = c["actual_len"].to(DEV)
    for variant, kinds in (("train", ("weightedmean", "lasttoken", "avg")), ("infer", ("weightedmean", "lasttoken"))):
        for kind in kinds:
            got = M.pooling(x, L, kind, variant)
            ref = c[f"{variant}_{kind}"]
            if kind == "lasttoken":
                assert torch.equal(got.float().cpu(), ref)
            elif variant == "train":
                assert_bf16_close(got, ref, f"pool {variant} {kind}")
            else:
                assert relrms(got, ref) < 1e-6
    g = torch.Generator().manual_seed(1)
    B, T, D = 16, 512, 2048
    x = torch.randn(B, T, D, generator=g).bfloat16()
    L = torch.randint(1, T, (B,), generator=g)
    L[0] = T - 1
    for kind in ("weightedmean", "avg"):
        ref = O.pooling(x, L, kind, "train")
        assert_bf16_close(M.pooling(x.to(DEV), L.to(DEV), kind, "train"), ref.float(), kind)
    ref = O.pooling(x, L, "weightedmean", "infer")
    assert relrms(M.pooling(x.to(DEV), L.to(DEV), "weightedmean", "infer"), ref) < 1e-6


def test_ddlerp_matches_eager_bf16_chain(M, O):
    """Bit-exact against the reference's eager bf16 op chain (src/model.py:437-448) run by torch on
    the same device, and close to the fp64 oracle."""
    g = torch.Generator().manual_seed(2)
    B, T, C, R = 2, 37, 256, 32
    x = torch.randn(B, T, C, generator=g).bfloat16().to(DEV)
    maa_x = torch.rand(C, generator=g).bfloat16().to(DEV)
    maa = torch.rand(5, C, generator=g).bfloat16().to(DEV)
    w1 = (torch.randn(C, 5 * R, generator=g) * 0.05).bfloat16().to(DEV)
    w2 = (torch.randn(5, R, C, generator=g) * 0.05).bfloat16().to(DEV)
    shift = torch.nn.ZeroPad2d((0, 0, 1, -1))
    xx = shift(x) - x
    xxx_ref = x + xx * maa_x
    assert torch.equal(M.tmix_shift_lerp(x, maa_x), xxx_ref)
    m = torch.bmm(torch.tanh(xxx_ref @ w1).view(B * T, 5, -1).transpose(0, 1), w2).view(5, B, T, C)
    want = torch.stack([x + xx * (maa[n] + m[n]) for n in range(5)])
    got = M.tmix_ddlerp_mix(x, maa, m)
    assert torch.equal(got, want)
    ref = O.tmix_ddlerp(x.cpu(), maa_x.cpu(), maa.cpu(), w1.cpu(), w2.cpu())
    for n in range(5):
        assert_bf16_close(got[n], ref[n], f"ddlerp {n}", relrms_tol=2e-2)
    # infctx variant: previous token comes from shift_state (src/model.py:738-745)
    st = torch.randn(B, C, generator=g).bfloat16().to(DEV)
    xx2 = torch.cat([st.unsqueeze(1), x[:, :-1]], 1) - x
    assert torch.equal(M.tmix_shift_lerp(x, maa_x, st), x + xx2 * maa_x)


def test_groupnorm_gate(M, O):
    g = torch.Generator().manual_seed(3)
    B, T, H = 3, 21, 4
    C = H * 64
    y = (torch.randn(B, T, C, generator=g) * 3).bfloat16()
    gate = torch.randn(B, T, C, generator=g).bfloat16()
    lw = torch.rand(C, generator=g).bfloat16()
    lb = (torch.randn(C, generator=g) * 0.1).bfloat16()
    eps = 1e-5 * 8 ** 2
    ref = O.groupnorm_gate(y, gate, lw, lb, H, eps)
    got = M.groupnorm_gate(y.to(DEV), gate.to(DEV), lw.to(DEV), lb.to(DEV), H, eps)
    assert_bf16_close(got, ref, "gn*gate")
    ln = torch.nn.GroupNorm(H, C, eps=eps).to(DEV).bfloat16()
    ln.weight.data.copy_(lw)
    ln.bias.data.copy_(lb)
    eager = ln(y.to(DEV).view(B * T, C)).view(B, T, C) * gate.to(DEV)
    assert relrms(got, eager) < 4e-3


# --------------------------------------------------------------------------------------------------
# gradients of the memory-bound pieces: fp64 autograd of the oracle expressions is the reference
# --------------------------------------------------------------------------------------------------
def _leaf64(t):
    return t.detach().cpu().double().requires_grad_(True)


@pytest.mark.parametrize("B,T,C,with_state", [(2, 37, 256, False), (3, 130, 768, True), (1, 1, 64, True), (5, 64, 320, False)])
def test_ddlerp_backward(M, B, T, C, with_state):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, T, C, generator=g).bfloat16().to(DEV).requires_grad_(True)
    maa_x = torch.rand(1, 1, C, generator=g).bfloat16().to(DEV).requires_grad_(True)
    maa = torch.rand(5, C, generator=g).bfloat16().to(DEV).requires_grad_(True)
    m = (torch.randn(5, B, T, C, generator=g) * 0.3).bfloat16().to(DEV).requires_grad_(True)
    st = torch.randn(B, C, generator=g).bfloat16().to(DEV).requires_grad_(True) if with_state else None
    go1 = torch.randn(B, T, C, generator=g).bfloat16().to(DEV)
    go5 = torch.randn(5, B, T, C, generator=g).bfloat16().to(DEV)

    def ref(x, maa_x, maa, m, st):
        prev = torch.zeros(B, 1, C, dtype=x.dtype) if st is None else st.unsqueeze(1)
        xx = torch.cat([prev, x[:, :-1]], 1) - x
        return x + xx * maa_x, torch.stack([x + xx * (maa[n] + m[n]) for n in range(5)])

    leaves = [_leaf64(t) if t is not None else None for t in (x, maa_x, maa, m, st)]
    r1, r5 = ref(*leaves)
    (r1 * go1.cpu().double()).sum().backward()
    want1 = [l.grad.clone() if l is not None and l.grad is not None else None for l in leaves]
    for l in leaves:
        if l is not None:
            l.grad = None
    (r5 * go5.cpu().double()).sum().backward()
    want5 = [l.grad if l is not None else None for l in leaves]

    out1 = M.tmix_shift_lerp(x, maa_x, st)
    out1.backward(go1)
    assert_bf16_close(x.grad, want1[0], "shift_lerp gx")
    assert_bf16_close(maa_x.grad, want1[1], "shift_lerp gmaa_x")
    if with_state:
        assert_bf16_close(st.grad, want1[4], "shift_lerp gshift")
        st.grad = None
    x.grad = None
    out5 = M.tmix_ddlerp_mix(x, maa, m, st)          # five [B,T,C] outputs when a gradient is needed
    assert len(out5) == 5
    torch.autograd.backward(list(out5), [go5[n] for n in range(5)])
    assert_bf16_close(x.grad, want5[0], "ddlerp gx")
    assert_bf16_close(maa.grad, want5[2], "ddlerp gmaa")
    assert_bf16_close(m.grad, want5[3], "ddlerp gm")
    if with_state:
        assert_bf16_close(st.grad, want5[4], "ddlerp gshift")


@pytest.mark.parametrize("act", [None, "silu"])
@pytest.mark.parametrize("B,T,H", [(3, 21, 4), (2, 300, 12), (1, 1, 1)])
def test_groupnorm_gate_backward(M, O, B, T, H, act):
    g = torch.Generator().manual_seed(13)
    C = H * 64
    eps = 1e-5 * 8 ** 2
    y = (torch.randn(B, T, C, generator=g) * 3).bfloat16().to(DEV).requires_grad_(True)
    gate = torch.randn(B, T, C, generator=g).bfloat16().to(DEV).requires_grad_(True)
    lw = (torch.rand(C, generator=g) + 0.5).bfloat16().to(DEV).requires_grad_(True)
    lb = (torch.randn(C, generator=g) * 0.1).bfloat16().to(DEV).requires_grad_(True)
    go = torch.randn(B, T, C, generator=g).bfloat16().to(DEV)
    leaves = [_leaf64(t) for t in (y, gate, lw, lb)]
    gate64 = torch.nn.functional.silu(leaves[1]) if act == "silu" else leaves[1]
    ref = O.groupnorm_gate(leaves[0], gate64, leaves[2], leaves[3], H, eps)
    (ref * go.cpu().double()).sum().backward()
    out = M.groupnorm_gate(y, gate, lw, lb, H, eps, gate_act=act)
    assert_bf16_close(out, ref, "gn*gate fwd")
    out.backward(go)
    for name, t, l in zip(("gy", "gg", "gln_w", "gln_b"), (y, gate, lw, lb), leaves):
        assert_bf16_close(t.grad, l.grad, "gn*gate " + name)


def test_heads_backward(M, O):
    """pooling / eos gather / reverse gather inside autograd (training heads, src/model_ext.py:1708-1769)."""
    g = torch.Generator().manual_seed(17)
    B, T, D = 5, 40, 136
    xb = torch.randn(B, T, D, generator=g).bfloat16()
    L = torch.tensor([1, 7, 39, 20, 12])
    go = torch.randn(B, D, generator=g).bfloat16()
    for kind, variant in (("weightedmean", "train"), ("avg", "train"), ("lasttoken", "train"), ("weightedmean", "infer")):
        x = xb.to(DEV).requires_grad_(True)
        out = M.pooling(x, L.to(DEV), kind, variant)
        out.backward(go.to(DEV).to(out.dtype))
        x64 = xb.double().requires_grad_(True)
        if kind == "lasttoken":
            ref = x64[torch.arange(B), L]
        else:                                # the fp64 form of oracle.pooling (which rounds its result to bf16)
            Lp = (L + 1 if variant == "infer" else L).unsqueeze(1)
            if kind == "weightedmean":
                wts = (torch.arange(1, T + 1).double() / Lp) * (torch.arange(T) <= Lp)
            else:
                wts = (torch.arange(T).unsqueeze(0) < Lp).double()
            ref = (x64 * wts.unsqueeze(-1)).sum(1) / Lp
            assert relrms(out, O.pooling(xb, L, kind, variant)) < 1e-2
        (ref * go.double()).sum().backward()
        assert_bf16_close(x.grad, x64.grad, f"pooling {kind}/{variant} gx")
    # eos gather: gradient lands on the gathered row only, bit-exact
    idx = torch.randint(2, 100, (B, T), generator=g)
    idx[0, 5] = 1
    idx[2, 0] = 1
    idx[3, T - 1] = 1
    x = xb.to(DEV).requires_grad_(True)
    rows, pos = M.eos_gather(x, idx.to(DEV), 1)
    rows.backward(go.to(DEV))
    want = torch.zeros(B, T, D, dtype=torch.bfloat16)
    want[torch.arange(B), pos.cpu()] = go
    assert torch.equal(x.grad.cpu(), want)
    # reverse gather
    mask, rev = M.create_mask_and_rev_idx(idx.to(DEV), 1, 0)
    x = xb.to(DEV).requires_grad_(True)
    gor = torch.randn(B, T, D, generator=g).bfloat16().to(DEV)
    M.reverse_x(x, rev).backward(gor)
    x2 = xb.to(DEV).requires_grad_(True)
    torch.gather(x2, 1, rev.unsqueeze(-1).expand(-1, -1, D)).backward(gor)
    assert torch.equal(x.grad, x2.grad)


@pytest.mark.parametrize("B,T,C,with_state", [(2, 37, 256, False), (3, 130, 768, True), (1, 1, 64, True), (2, 300, 2048, False)])
def test_ddlerp_lora_fused(M, B, T, C, with_state):
    """tmix_ddlerp_lora (LoRA product on the tensor cores, m never materialised) == bmm + tmix_ddlerp_mix up to
    the accumulation order of the K=32 product; gradients through the recompute path match too."""
    g = torch.Generator().manual_seed(23)
    R = 32
    x = torch.randn(B, T, C, generator=g).bfloat16().to(DEV)
    maa = torch.rand(5, C, generator=g).bfloat16().to(DEV)
    h = torch.tanh(torch.randn(B * T, 5 * R, generator=g)).bfloat16().to(DEV)
    w2 = (torch.randn(5, R, C, generator=g) * 0.1).bfloat16().to(DEV)
    st = torch.randn(B, C, generator=g).bfloat16().to(DEV) if with_state else None
    m = torch.bmm(h.view(B * T, 5, R).transpose(0, 1), w2).view(5, B, T, C)
    want = M.tmix_ddlerp_mix(x, maa, m, st)
    got = M.tmix_ddlerp_lora(x, maa, h, w2, st)
    assert got.shape == want.shape
    # identical except where the fp32 sum of 32 products rounds to a neighbouring bf16 value
    diff = (got.float() - want.float()).abs()
    assert (diff > 0).float().mean().item() < 0.02
    assert relrms(got, want) < 2e-3
    # fp64 reference of the whole expression
    xx = (torch.cat([(st if with_state else torch.zeros(B, C, device=DEV, dtype=torch.bfloat16)).unsqueeze(1), x[:, :-1]], 1).double() - x.double())
    m64 = torch.bmm(h.double().view(B * T, 5, R).transpose(0, 1), w2.double()).view(5, B, T, C)
    ref = x.double() + xx * (maa.double().view(5, 1, 1, C) + m64)
    assert relrms(got, ref) < 1e-2           # four bf16 roundings deep, like the eager chain itself
    assert abs(relrms(got, ref) - relrms(want, ref)) < 5e-4
    # gradients
    leaves = [t.clone().requires_grad_(True) for t in (x, maa, h, w2)] + ([st.clone().requires_grad_(True)] if with_state else [None])
    outs = M.tmix_ddlerp_lora(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4])
    go = [torch.randn(B, T, C, generator=g).bfloat16().to(DEV) for _ in range(5)]
    torch.autograd.backward(list(outs), go)
    l64 = [t.detach().double().requires_grad_(True) if t is not None else None for t in leaves]
    prev = torch.zeros(B, 1, C, device=DEV, dtype=torch.float64) if l64[4] is None else l64[4].unsqueeze(1)
    xx64 = torch.cat([prev, l64[0][:, :-1]], 1) - l64[0]
    m64 = torch.bmm(l64[2].view(B * T, 5, R).transpose(0, 1), l64[3]).view(5, B, T, C)
    ref = l64[0] + xx64 * (l64[1].view(5, 1, 1, C) + m64)
    (ref * torch.stack(go).double()).sum().backward()
    for name, a, b_ in zip(("gx", "gmaa", "gh", "gw2", "gshift"), leaves, l64):
        if a is not None:
            assert relrms(a.grad, b_.grad) < 1.5e-2, name


def test_groupnorm_gate_pair(M):
    """Reverse gather + average inside the GroupNorm*gate kernel == the three separate ops."""
    g = torch.Generator().manual_seed(29)
    B, T, H = 3, 50, 2
    C = H * 64
    y = torch.randn(B, T, C, generator=g).bfloat16().to(DEV)
    yr = torch.randn(B, T, C, generator=g).bfloat16().to(DEV)
    gate = torch.randn(B, T, C, generator=g).bfloat16().to(DEV)
    lw = (torch.rand(C, generator=g) + 0.5).bfloat16().to(DEV)
    lb = (torch.randn(C, generator=g) * 0.1).bfloat16().to(DEV)
    idx = torch.randint(2, 50, (B, T), generator=g)
    for b, n in enumerate((49, 20, 3)):
        idx[b, n] = 1
        idx[b, n + 1:] = 0
    _, rev = M.create_mask_and_rev_idx(idx.to(DEV), 1, 0)
    for act in (None, "silu"):
        want = M.groupnorm_gate((y + M.reverse_x(yr, rev)) / 2, gate, lw, lb, H, 64e-5, gate_act=act)
        got = M.groupnorm_gate_pair(y, yr, rev, gate, lw, lb, H, 64e-5, gate_act=act)
        assert torch.equal(got, want)


def test_no_writes_outside_the_outputs(M):
    """Ragged shapes (T not a multiple of the 8- / 128-row tiles, C not a multiple of 256): every output of the
    TMA-fed / tcgen05 kernels sits between sentinel regions that must stay untouched."""
    from rwkv_lm_ext_b200 import _lib
    lib = _lib.load()
    p = _lib.ptr
    g = torch.Generator().manual_seed(31)
    B, T, C, R = 3, 77, 320, 32
    PAD = 4096
    st = torch.cuda.current_stream().cuda_stream

    def guarded(*shape, dtype=torch.bfloat16):
        n = 1
        for s_ in shape:
            n *= s_
        buf = torch.full((n + 2 * PAD,), 7.0, dtype=dtype, device=DEV)
        return buf, buf[PAD:PAD + n].view(*shape)

    def intact(buf):
        return bool((buf[:PAD] == 7.0).all() and (buf[-PAD:] == 7.0).all())

    x = torch.randn(B, T, C, generator=g).bfloat16().to(DEV)
    maa = torch.rand(5, C, generator=g).bfloat16().to(DEV)
    h = torch.tanh(torch.randn(B * T, 5 * R, generator=g)).bfloat16().to(DEV)
    w2 = (torch.randn(5, R, C, generator=g) * 0.1).bfloat16().to(DEV)
    m = torch.randn(5, B, T, C, generator=g).bfloat16().to(DEV)
    gos = [torch.randn(B, T, C, generator=g).bfloat16().to(DEV) for _ in range(5)]
    # LoRA-fused forward (tcgen05 + TMA stores, 128-row tiles)
    obuf, out = guarded(5, B, T, C)
    _lib.check(lib.tmix_ddlerp_lora_bf16(B, T, C, R, p(x), None, p(maa), p(h), p(w2), p(out), st), "lora")
    torch.cuda.synchronize()
    assert intact(obuf)
    mm = torch.bmm(h.view(B * T, 5, R).transpose(0, 1), w2).view(5, B, T, C)
    assert relrms(out, M.tmix_ddlerp_mix(x, maa, mm)) < 2e-3
    # TMA-fed backward (8-row tiles, 256-channel column blocks)
    gxb, gx = guarded(B, T, C)
    gmb, gm = guarded(5, B, T, C)
    gab, ga = guarded(5, C, dtype=torch.float32)
    ws = torch.empty(lib.elementwise_backward_workspace_bytes(B, T, C, 5), dtype=torch.uint8, device=DEV)
    _lib.check(lib.tmix_ddlerp_mix_backward_bf16(B, T, C, p(x), None, p(maa), p(m), *[p(t) for t in gos], p(gx), p(gm), p(ga),
                                                 None, p(ws), ws.numel(), st), "ddlerp bwd")
    torch.cuda.synchronize()
    assert intact(gxb) and intact(gmb) and intact(gab)
    assert torch.isfinite(gx.float()).all() and torch.isfinite(gm.float()).all() and torch.isfinite(ga).all()
    # channel-mix two-output shift-lerp and its gradient
    o2b, o2 = guarded(2, B, T, C)
    _lib.check(lib.cmix_shift_lerp2_bf16(B, T, C, p(x), None, p(maa[:2].contiguous()), p(o2), st), "cmix fwd")
    g2b, g2 = guarded(B, T, C)
    a2b, a2 = guarded(2, C, dtype=torch.float32)
    ws3 = torch.empty(lib.elementwise_backward_workspace_bytes(B, T, C, 3), dtype=torch.uint8, device=DEV)
    _lib.check(lib.cmix_shift_lerp2_backward_bf16(B, T, C, p(x), None, p(maa[:2].contiguous()), p(gos[0]), p(gos[1]), p(g2), p(a2),
                                                  None, p(ws3), ws3.numel(), st), "cmix bwd")
    torch.cuda.synchronize()
    assert intact(o2b) and intact(g2b) and intact(a2b)


@pytest.mark.parametrize("D", [768, 2048, 2560, 4096])
@pytest.mark.parametrize("with_delta", [True, False])
def test_add_layernorm_matches_the_eager_chain(D, with_delta):
    """x_new = x + delta, y = LayerNorm(x_new) (the residual glue of Block.forward, src/model.py:904-933) in one kernel:
    forward equal to torch's bf16 add + layer_norm up to one bf16 ulp of y, backward against fp32 autograd, including the
    affine-parameter gradients and the case where they are frozen."""
    import rwkv_lm_ext_b200 as M
    torch.manual_seed(D)
    B, T = 3, 37
    x = torch.randn(B, T, D, device=DEV).bfloat16()
    delta = (torch.randn(B, T, D, device=DEV) * 0.5).bfloat16() if with_delta else None
    w = (1 + 0.3 * torch.randn(D, device=DEV)).bfloat16()
    b = (0.2 * torch.randn(D, device=DEV)).bfloat16()
    xs = x + delta if with_delta else x
    y_ref = torch.nn.functional.layer_norm(xs, (D,), w, b, 1e-5)
    x_new, y = M.add_layernorm(x, delta, w, b, 1e-5)
    assert torch.equal(x_new, xs)
    assert (y.float() - y_ref.float()).abs().max().item() <= 2.0 ** -7 * y_ref.float().abs().max().item()
    assert relrms(y, y_ref) < 3e-3
    # backward
    go_x = torch.randn(B, T, D, device=DEV).bfloat16()
    go_y = torch.randn(B, T, D, device=DEV).bfloat16()
    leaves32 = [t.float().detach().requires_grad_(True) for t in ([x, delta] if with_delta else [x]) + [w, b]]
    xs32 = leaves32[0] + leaves32[1] if with_delta else leaves32[0]
    y32 = torch.nn.functional.layer_norm(xs32, (D,), leaves32[-2], leaves32[-1], 1e-5)
    (y32 * go_y.float()).sum().backward() if not with_delta else ((y32 * go_y.float()).sum() + (xs32 * go_x.float()).sum()).backward()
    leaves = [t.detach().clone().requires_grad_(True) for t in ([x, delta] if with_delta else [x]) + [w, b]]
    xn, yy = M.add_layernorm(leaves[0], leaves[1] if with_delta else None, leaves[-2], leaves[-1], 1e-5)
    ((yy.float() * go_y.float()).sum() if not with_delta else (yy.float() * go_y.float()).sum() + (xn.float() * go_x.float()).sum()).backward()
    for a, r_, name in zip(leaves, leaves32, (["x", "delta"] if with_delta else ["x"]) + ["w", "b"]):
        assert relrms(a.grad, r_.grad) < 6e-3, (name, relrms(a.grad, r_.grad))
    # frozen affine parameters (LoRA / state tuning): no column pass, same input gradient
    xf = x.detach().clone().requires_grad_(True)
    _, y2 = M.add_layernorm(xf, delta, w, b, 1e-5)
    (y2.float() * go_y.float()).sum().backward()
    ref_gx = torch.autograd.grad((torch.nn.functional.layer_norm((xf.float() + (delta.float() if with_delta else 0)), (D,), w.float(), b.float(), 1e-5)
                                  * go_y.float()).sum(), xf)[0]
    assert relrms(xf.grad, ref_gx) < 6e-3
