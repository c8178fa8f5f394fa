"""The composed task models on the fused kernels against outputs produced by the REFERENCE's own classes
(tests/golden/make_task_golden.py ran `RwkvForSequenceEmbedding` and `RwkvForClassification` of src/model_ext.py
around `RWKV` of src/model.py on the CPU): three pooling types, the dense + tanh head, classification /
cross-encoder logits, the hidden states, and the bucketed corpus encoder against row-by-row calls.  `pytest -m gpu`."""
import os

import numpy as np
import pytest
import torch

from tests.util import relrms

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "task_models_2x128.npz")


class _CMix(torch.nn.Module):          # attribute names of RWKV_CMix_x060 (src/model.py:616-644)
    def __init__(self, D, F):
        super().__init__()
        self.time_maa_k = torch.nn.Parameter(torch.zeros(1, 1, D))
        self.time_maa_r = torch.nn.Parameter(torch.zeros(1, 1, D))
        self.key = torch.nn.Linear(D, F, bias=False)
        self.receptance = torch.nn.Linear(D, D, bias=False)
        self.value = torch.nn.Linear(F, D, bias=False)


def build(M, c):
    L, D, F = int(c["n_layer"]), int(c["n_embd"]), int(c["dim_ffn"])
    base = torch.nn.Module()
    base.emb = torch.nn.Embedding(c["w:rwkvModel.emb.weight"].shape[0], D)
    base.blocks = torch.nn.ModuleList()
    for i in range(L):
        b = torch.nn.Module()
        if i == 0:
            b.ln0 = torch.nn.LayerNorm(D)
        b.ln1, b.ln2 = torch.nn.LayerNorm(D), torch.nn.LayerNorm(D)
        b.att = M.Tmix_x060(D, D // 64)
        b.ffn = _CMix(D, F)
        base.blocks.append(b)
    base.ln_out = torch.nn.LayerNorm(D)
    base.head = torch.nn.Linear(D, c["w:rwkvModel.head.weight"].shape[0], bias=False)
    sd = {k[len("w:rwkvModel."):]: torch.from_numpy(c[k].astype(np.float32)) for k in c.files if k.startswith("w:rwkvModel.")}
    missing, unexpected = base.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)     # the reference's state_dict, name for name
    w = torch.nn.Module()
    w.rwkvModel = base
    w.embedding_id = w.class_id = 1
    w.pad_id = 0
    w.dense = torch.nn.Linear(D, c["w:dense.weight"].shape[0])
    w.dense.load_state_dict({"weight": torch.from_numpy(c["w:dense.weight"].astype(np.float32)),
                             "bias": torch.from_numpy(c["w:dense.bias"].astype(np.float32))})
    w.activation = torch.nn.Tanh()
    w.score = torch.nn.Linear(D, c["w:score.weight"].shape[0], bias=False)
    w.score.load_state_dict({"weight": torch.from_numpy(c["w:score.weight"].astype(np.float32))})
    return w.bfloat16().to(DEV).eval()


def _close(got, ref, what, cos_min=0.999, rr=3e-2):
    got, ref = got.float().cpu(), torch.from_numpy(ref)
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=-1)
    assert cos.min().item() >= cos_min, (what, cos)
    assert relrms(got, ref) < rr, (what, relrms(got, ref))


def test_task_models_match_the_reference_classes():
    import rwkv_lm_ext_b200 as M
    c = np.load(GOLD)
    w = build(M, c)
    idx = torch.from_numpy(c["idx"]).to(DEV)
    with torch.no_grad():
        hidden = M.causal_hidden(w, idx)
        valid = torch.from_numpy(c["idx"]) != 0
        assert relrms(hidden.float().cpu()[valid], torch.from_numpy(c["hidden"])[valid]) < 3e-2
        for pool in ("weightedmean", "lasttoken", "avg"):
            w.pooling_type, w.add_mlp = pool, False
            _close(M.sequence_embedding(w, idx), c["emb_" + pool], pool)
        w.pooling_type, w.add_mlp = "lasttoken", True
        _close(M.sequence_embedding(w, idx), c["emb_mlp"], "dense + tanh head")
        _close(M.classification_logits(w, idx), c["logits"], "classification / cross-encoder logits", cos_min=0.998, rr=5e-2)


def test_bucketed_corpus_encoder_equals_row_by_row():
    import rwkv_lm_ext_b200 as M
    c = np.load(GOLD)
    w = build(M, c)
    w.pooling_type, w.add_mlp = "lasttoken", False
    g = torch.Generator().manual_seed(5)
    lens = [3, 70, 17, 64, 128, 1, 33, 200, 65]
    seqs = [torch.randint(4, 500, (n,), generator=g).tolist() for n in lens]
    with torch.no_grad():
        got = M.encode_corpus(w, seqs, max_tokens=512)
        assert got.shape == (len(seqs), 128)
        for i, s in enumerate(seqs):                              # the reference's way: one sentence at a time
            one = M.sequence_embedding(w, torch.tensor([s + [1]], device=DEV))
            assert relrms(got[i], one[0]) < 2e-2, (i, len(s))
    # cross-encoder rows: query [sep] document [cls] + padding, truncation keeps the class token
    rows = M.cross_encoder_rows([[5, 6], [7] * 40], [[8, 9, 10], [11] * 40], max_len=16)
    assert rows[0].tolist() == [5, 6, 2, 8, 9, 10, 1] + [0] * 9
    assert rows[1].tolist() == [7] * 15 + [1]
    with torch.no_grad():
        sc = M.classification_logits(w, rows.to(DEV))
    assert sc.shape == (2, 3) and torch.isfinite(sc.float()).all()
