"""Whole bi-directional encoder on the fused kernels against embeddings produced by the reference's own
model code (tests/golden/make_encoder_golden.py ran RwkvEncoder.encode_sentence of
src/model_encoder_run.py on its CPU path).  SURVEY.md 8(d) config 3 accuracy gate: cosine >= 0.999 per
passage.  `pytest -m gpu`."""
import os

import numpy as np
import pytest
import torch

from tests.util import relrms

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder_2x128.npz")


class _CMix(torch.nn.Module):          # same attribute names and math as BiRWKV_CMix_x060 (src/model_encoder_run.py:212-238)
    def __init__(self, D, F):
        super().__init__()
        self.time_maa_k = torch.nn.Parameter(torch.zeros(1, 1, D))
        self.time_maa_r = torch.nn.Parameter(torch.zeros(1, 1, D))
        self.key = torch.nn.Linear(D, F, bias=False)
        self.receptance = torch.nn.Linear(D, D, bias=False)
        self.value = torch.nn.Linear(F, D, bias=False)

    def forward(self, x):
        xx = torch.nn.functional.pad(x, (0, 0, 1, -1)) - x
        k = torch.relu(self.key(x + xx * self.time_maa_k)) ** 2
        return torch.sigmoid(self.receptance(x + xx * self.time_maa_r)) * self.value(k)


def build_model(M, c):
    L, D, F = int(c["n_layer"]), int(c["n_embd"]), int(c["dim_ffn"])
    H = D // 64
    m = torch.nn.Module()
    m.emb = torch.nn.Embedding(c["w:emb.weight"].shape[0], D)
    m.blocks = torch.nn.ModuleList()
    for i in range(L):
        b = torch.nn.Module()
        if i == 0:
            b.ln0 = torch.nn.LayerNorm(D)
        b.ln1, b.ln2 = torch.nn.LayerNorm(D), torch.nn.LayerNorm(D)
        b.att = M.Tmix_x060(D, H)
        b.ffn = _CMix(D, F)
        m.blocks.append(b)
    m.ln_out = torch.nn.LayerNorm(D)
    m.emb_id, m.pad_id = int(c["emb_id"]), int(c["pad_id"])
    sd = {k[2:]: torch.from_numpy(c[k].astype(np.float32)) for k in c.files if k.startswith("w:")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing, missing          # every parameter of the rebuilt model comes from the reference's state_dict
    assert all(u.endswith(("copy_mask", "tiny_mask")) for u in unexpected), unexpected
    return m


def test_bi_encoder_matches_reference_model():
    import rwkv_lm_ext_b200 as M
    c = np.load(GOLD)
    model = build_model(M, c).bfloat16().to(DEV).eval()
    idx = torch.from_numpy(c["idx"]).to(DEV)
    with torch.no_grad():
        hidden = M.bi_encoder_hidden(model, idx)
        emb = M.bi_encoder_encode(model, idx)
    ref_emb, ref_hidden = torch.from_numpy(c["emb"]), torch.from_numpy(c["hidden"])
    cos = torch.nn.functional.cosine_similarity(emb.float().cpu(), ref_emb, dim=-1)
    assert cos.min().item() >= 0.999, cos
    assert relrms(emb, ref_emb) < 3e-2
    # every non-padded position of the last hidden state, not only the gathered token
    valid = torch.from_numpy(c["idx"]) != int(c["pad_id"])
    assert relrms(hidden.float().cpu()[valid], ref_hidden[valid]) < 3e-2


def test_bi_encoder_is_differentiable():
    import rwkv_lm_ext_b200 as M
    c = np.load(GOLD)
    model = build_model(M, c).bfloat16().to(DEV).train()
    idx = torch.from_numpy(c["idx"]).to(DEV)
    emb = M.bi_encoder_encode(model, idx)
    emb.float().pow(2).sum().backward()
    for name, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name


def test_graphed_forward_replays_the_eager_forward():
    """GraphedForward: the whole bi-encoder forward captured once per input shape; replays on new ids (device or pinned
    host) give bit-identical embeddings to the eager call, and a weight update is seen by later replays."""
    import rwkv_lm_ext_b200 as M
    from rwkv_lm_ext_b200.synthetic import make_bi_encoder, make_passages
    model = make_bi_encoder(2, 128, 2, 448, 512, seed=0, device=DEV)
    a = make_passages(6, T=96, vocab=512, seed=1, min_len=20).to(DEV)
    b = make_passages(6, T=96, vocab=512, seed=2, min_len=20)
    enc = M.GraphedForward(M.bi_encoder_encode, model, a)
    with torch.no_grad():
        want_a, want_b = M.bi_encoder_encode(model, a), M.bi_encoder_encode(model, b.to(DEV))
    assert torch.equal(enc(a), want_a)
    assert torch.equal(enc(b.pin_memory()), want_b)                  # ids straight from pinned host memory
    assert torch.equal(enc(a).clone(), want_a)
    with torch.no_grad():
        model.blocks[0].att.time_maa_k.add_(0.25)                    # a frozen-parameter cache (tmix._maa5) must notice
        model.blocks[1].ffn.value.weight.mul_(0.5)
        want2 = M.bi_encoder_encode(model, a)
    assert not torch.equal(want2, want_a)
    assert torch.equal(enc(a), want2)
    with pytest.raises(AssertionError):
        enc(a[:, :64])
