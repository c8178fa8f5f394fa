"""BASELINE config 4 in miniature: a 2-layer RWKV-6 language model trained for a few AdamW steps on random
tokens, (a) with the fused time-mix / channel-mix forwards and the tensor-core WKV6 kernels, (b) with the
reference's eager elementwise chain (src/model.py:434-468, :635-644 restated) around the exact SIMT WKV6 kernels.
Same initial weights, same batches: the two loss curves must stay together and go down.  `pytest -m gpu`."""
import copy

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


class CMix(torch.nn.Module):
    def __init__(self, D, FF):
        super().__init__()
        self.time_maa_k = torch.nn.Parameter(torch.rand(1, 1, D))
        self.time_maa_r = torch.nn.Parameter(torch.rand(1, 1, D))
        self.key = torch.nn.Linear(D, FF, bias=False)
        self.receptance = torch.nn.Linear(D, D, bias=False)
        self.value = torch.nn.Linear(FF, D, bias=False)


def eager_tmix(M, l, x):
    B, T, C = x.shape
    H = l.time_faaaa.shape[0]
    xx = F.pad(x, (0, 0, 1, -1)) - x
    xxx = x + xx * l.time_maa_x
    xxx = torch.tanh(xxx @ l.time_maa_w1).view(B * T, 5, -1).transpose(0, 1)
    mw, mk, mv, mr, mg = torch.bmm(xxx, l.time_maa_w2).view(5, B, T, -1).unbind(0)
    xw, xk, xv = x + xx * (l.time_maa_w + mw), x + xx * (l.time_maa_k + mk), x + xx * (l.time_maa_v + mv)
    xr, xg = x + xx * (l.time_maa_r + mr), x + xx * (l.time_maa_g + mg)
    r, k, v, g = l.receptance(xr), l.key(xk), l.value(xv), F.silu(l.gate(xg))
    w = l.time_decay + torch.tanh(xw @ l.time_decay_w1) @ l.time_decay_w2
    y = M.RUN_CUDA_RWKV6(B, T, C, H, r, k, v, w, l.time_faaaa)
    return l.output(l.ln_x(y.view(B * T, C)).view(B, T, C) * g)


def eager_cmix(l, x):
    xx = F.pad(x, (0, 0, 1, -1)) - x
    k = torch.relu(l.key(x + xx * l.time_maa_k)) ** 2
    return torch.sigmoid(l.receptance(x + xx * l.time_maa_r)) * l.value(k)


def build(M, V, D, H, L):
    torch.manual_seed(3)
    m = torch.nn.Module()
    m.emb = torch.nn.Embedding(V, D)
    m.blocks = torch.nn.ModuleList()
    for _ in range(L):
        b = torch.nn.Module()
        b.ln1, b.ln2 = torch.nn.LayerNorm(D), torch.nn.LayerNorm(D)
        b.att = M.Tmix_x060(D, H)
        with torch.no_grad():
            for n, p in b.att.named_parameters():
                if n in ("time_maa_w1", "time_maa_w2", "time_decay_w1", "time_decay_w2"):
                    p.uniform_(-0.02, 0.02)
                elif n == "time_decay":
                    p.copy_(-5 + 4 * torch.rand_like(p))
                elif n.startswith("time_maa"):
                    p.uniform_(0, 1)
                elif n == "time_faaaa":
                    p.normal_(0, 0.3)
        b.ffn = CMix(D, 4 * D)
        m.blocks.append(b)
    m.ln_out = torch.nn.LayerNorm(D)
    m.head = torch.nn.Linear(D, V, bias=False)
    return m


def forward(M, m, idx, fused):
    x = m.emb(idx)
    for b in m.blocks:
        h = b.ln1(x)
        x = x + (M.tmix_x060_forward(b.att, h) if fused else eager_tmix(M, b.att, h))
        h = b.ln2(x)
        x = x + (M.cmix_x060_forward(b.ffn, h) if fused else eager_cmix(b.ffn, h))
    return m.head(m.ln_out(x))


def test_training_curves_agree():
    import rwkv_lm_ext_b200 as M
    V, D, H, L, B, T, STEPS = 96, 128, 2, 2, 4, 96, 12
    base = build(M, V, D, H, L)
    g = torch.Generator().manual_seed(4)
    # a learnable stream: the next token is a fixed function of the current one, plus noise tokens
    perm = torch.randperm(V, generator=g)
    batches = []
    for _ in range(STEPS):
        x0 = torch.randint(0, V, (B, 1), generator=g)
        seq = [x0]
        for _t in range(T):
            seq.append(perm[seq[-1]])
        batches.append(torch.cat(seq, 1).to(DEV))
    curves = {}
    for fused in (True, False):
        m = copy.deepcopy(base).bfloat16().to(DEV)
        master = [p.detach().float().clone().requires_grad_(True) for p in m.parameters()]      # fp32 master weights
        opt = torch.optim.AdamW(master, lr=3e-3, weight_decay=0.0)
        M.set_impl("auto" if fused else "simt")
        try:
            losses = []
            for data in batches:
                logits = forward(M, m, data[:, :-1], fused)
                loss = F.cross_entropy(logits.float().view(-1, V), data[:, 1:].reshape(-1))
                m.zero_grad(set_to_none=True)
                loss.backward()
                for mp, p in zip(master, m.parameters()):
                    mp.grad = p.grad.float()
                opt.step()
                with torch.no_grad():
                    for mp, p in zip(master, m.parameters()):
                        p.copy_(mp)
                losses.append(loss.item())
        finally:
            M.set_impl("auto")
        curves[fused] = losses
    a, b = curves[True], curves[False]
    assert a[-1] < 0.7 * a[0] and b[-1] < 0.7 * b[0], (a, b)            # both learn
    for i, (x, y) in enumerate(zip(a, b)):
        assert abs(x - y) <= 0.05 * max(x, y) + 0.02, (i, a, b)         # and stay together
