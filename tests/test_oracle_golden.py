"""The oracle against the reference's own outputs (tests/golden/*.npz, made by make_golden.py).
CPU only.  Tolerances: the reference numbers are fp32, the oracle fp64."""
import os

import numpy as np
import torch

from oracle import wkv6_oracle as O


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _close(a, b, atol, rtol=1e-4):
    a, b = a.double(), b.double()
    assert a.shape == b.shape
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, atol=atol, rtol=rtol), f"max err {err}"


def test_forward_matches_reference_cpu_paths(golden_dir):
    c = _load(golden_dir, "wkv6_2x10x256_randn")
    y = O.wkv6_forward(c["r"], c["k"], c["v"], c["w"], c["u"])
    _close(y, c["y_run_rwkv6_forward"], 2e-4)     # src/model_encoder_run.py:31-62
    _close(y, c["y_pytorch_forward"], 2e-4)       # tests/test_cpu.py:190-231
    _close(y, c["y_fla"], 2e-4)                   # fla/ops/rwkv6/recurrent_naive.py:8-36


def test_forward_realistic_decay(golden_dir):
    c = _load(golden_dir, "wkv6_2x150x128_decay")
    y = O.wkv6_forward(c["r"], c["k"], c["v"], c["w"], c["u"])
    _close(y, c["y_run_rwkv6_forward"], 1e-3)
    _close(y, c["y_fla"], 1e-3)


def test_gradients_match_reference_naive_autograd(golden_dir):
    for name in ("wkv6_2x10x256_randn", "wkv6_2x150x128_decay"):
        c = _load(golden_dir, name)
        g = O.wkv6_backward(c["r"], c["k"], c["v"], c["w"], c["u"], c["gy"])
        for key in ("gr", "gk", "gv", "gw", "gu"):
            scale = c[key].abs().max().item()
            _close(g[key], c[key], 2e-5 * max(scale, 1.0), rtol=2e-4)


def test_gw_edges_are_zero(golden_dir):
    # cuda/wkv6_cuda.cu:201,226 write exact zeros at t=0 and t=T-1; mathematically they are zero.
    c = _load(golden_dir, "wkv6_2x10x256_randn")
    g = O.wkv6_backward(c["r"], c["k"], c["v"], c["w"], c["u"], c["gy"])
    assert g["gw"][:, 0].abs().max().item() < 1e-12
    assert g["gw"][:, -1].abs().max().item() == 0.0


def test_initial_state_forms(golden_dir):
    c = _load(golden_dir, "wkv6state_2x70x128")
    s_vk = c["s0_kv"].transpose(-1, -2).contiguous()             # CUDA ops hold [value,key]
    y, s_out = O.wkv6infctx_forward(c["r"], c["k"], c["v"], c["w"], c["u"], s_vk)
    _close(y, c["y_fla"], 1e-3)
    g = O.wkv6_backward(c["r"], c["k"], c["v"], c["w"], c["u"], c["gy"], s=s_vk, s_layout="infctx")
    for key in ("gr", "gk", "gv", "gw", "gu"):
        scale = c[key].abs().max().item()
        _close(g[key], c[key], 2e-5 * max(scale, 1.0), rtol=2e-4)
    _close(g["gs"].transpose(-1, -2), c["gs_kv"], 2e-5 * c["gs_kv"].abs().max().item(), rtol=2e-4)
    # shared-state ("states") form == infctx form with the state broadcast over the batch
    s_h = s_vk[0]
    y2 = O.wkv6state_forward(c["r"], c["k"], c["v"], c["w"], c["u"], s_h)
    y3, _ = O.wkv6infctx_forward(c["r"], c["k"], c["v"], c["w"], c["u"], s_h.expand(2, -1, -1, -1))
    _close(y2, y3, 0.0)
    # chunked carry == one long call (the infctx contract, src/model.py:1167-1190)
    ya, sa = O.wkv6infctx_forward(c["r"][:, :33], c["k"][:, :33], c["v"][:, :33], c["w"][:, :33], c["u"], s_vk)
    yb, sb = O.wkv6infctx_forward(c["r"][:, 33:], c["k"][:, 33:], c["v"][:, 33:], c["w"][:, 33:], c["u"], sa)
    _close(torch.cat([ya, yb], 1), y, 1e-12)
    _close(sb, s_out, 1e-12)


def test_inference_form_equals_training_form(golden_dir):
    c = _load(golden_dir, "wkv6state_2x70x128")
    s_vk = c["s0_kv"].transpose(-1, -2).contiguous()
    y, s_out = O.wkv6infctx_forward(c["r"], c["k"], c["v"], c["w"], c["u"], s_vk)
    decay = torch.exp(-torch.exp(c["w"].double()))
    yi, si = O.rwkv6_inference_forward(s_vk[1], c["r"][1], c["k"][1], c["v"][1], decay[1], c["u"])
    _close(yi, y[1], 1e-10)
    _close(si, s_out[1], 1e-10)


def test_mask_and_reverse_index_bit_exact(golden_dir):
    c = _load(golden_dir, "mask_rev_idx")
    mask = O.create_mask(c["idx"])
    assert mask.dtype == torch.int32 and torch.equal(mask, c["mask"])
    assert torch.equal(O.reverse_x_idx(mask, c["idx"].size(1)), c["rev_idx"])
    assert torch.equal(O.eos_index(c["idx"], 1), c["eos_pos"])


def test_pooling_matches_reference(golden_dir):
    c = _load(golden_dir, "pooling")
    x = c["x"].bfloat16()
    L = c["actual_len"]
    for kind in ("weightedmean", "lasttoken", "avg"):
        out = O.pooling(x, L, kind, "train")
        assert torch.equal(out.float(), c[f"train_{kind}"]), kind
    for kind in ("weightedmean", "lasttoken"):
        out = O.pooling(x, L, kind, "infer")
        assert torch.equal(out.float(), c[f"infer_{kind}"]), kind


def test_bi_semantics():
    g = torch.Generator().manual_seed(0)
    B, T, C, H = 2, 12, 64, 1
    r, k, v, w = (torch.randn(B, T, C, generator=g, dtype=torch.float64) for _ in range(4))
    u = torch.randn(H, 64, generator=g, dtype=torch.float64)
    mask = torch.ones(B, T, dtype=torch.int32)
    mask[1, 7:] = 0
    y = O.wkv6_bi_forward(mask, r, k, v, w, u)
    assert y[1, 8:].abs().max().item() == 0.0
    # brute-force definition of the reverse exclusive pass for row 0, position t
    d = torch.exp(-torch.exp(w))
    yf = O.wkv6_forward(r, k, v, w, u)
    for b, p in ((0, T - 1), (1, 7)):
        for t in (0, 3, p):
            acc = torch.zeros(64, dtype=torch.float64)
            for s in range(t + 1, p + 1):
                dec = torch.ones(64, dtype=torch.float64)
                for m in range(t + 1, s):
                    dec = dec * d[b, m]
                acc += (r[b, t] * dec * k[b, s]).sum() * v[b, s]
            _close(y[b, t], yf[b, t] + acc, 1e-10)
    # quirk mode: rows without a zero lose the reverse pass
    yq = O.wkv6_bi_forward(mask, r, k, v, w, u, ref_quirks=True)
    _close(yq[0], yf[0], 1e-12)
    _close(yq[1], y[1], 0.0)


def test_oracle_pieces_compose_to_the_reference_encoder(golden_dir):
    """tests/golden/encoder_2x128.npz holds RwkvEncoder.encode_sentence (src/model_encoder_run.py) run on
    its CPU path.  Rebuilding that forward from the oracle's pieces alone (mask / reverse index / reverse
    gather, tmix_ddlerp, tmix_decay, the recurrence, groupnorm_gate, eos gather) pins those restatements
    against the reference's own model code."""
    z = np.load(os.path.join(golden_dir, "encoder_2x128.npz"))
    P = {k[2:]: torch.from_numpy(z[k].astype(np.float64)) for k in z.files if k.startswith("w:")}
    idx = torch.from_numpy(z["idx"])
    L, D = int(z["n_layer"]), int(z["n_embd"])
    H = D // 64
    ln = lambda x, p: torch.nn.functional.layer_norm(x, (D,), P[p + ".weight"], P[p + ".bias"])
    mask = O.create_mask(idx, int(z["emb_id"]), int(z["pad_id"]))
    rev = O.reverse_x_idx(mask, idx.shape[1])
    x = P["emb.weight"][idx]
    for i in range(L):
        pre = f"blocks.{i}."
        if i == 0:
            x = ln(x, pre + "ln0")
        a = pre + "att."
        maa5 = torch.cat([P[a + "time_maa_" + n].view(1, D) for n in "wkvrg"])

        def project(xin):
            xw, xk, xv, xr, xg = O.tmix_ddlerp(xin, P[a + "time_maa_x"].view(D), maa5, P[a + "time_maa_w1"], P[a + "time_maa_w2"])
            r, k, v = xr @ P[a + "receptance.weight"].T, xk @ P[a + "key.weight"].T, xv @ P[a + "value.weight"].T
            g = torch.nn.functional.silu(xg @ P[a + "gate.weight"].T)
            return r, k, v, g, O.tmix_decay(xw, P[a + "time_decay"], P[a + "time_decay_w1"], P[a + "time_decay_w2"])
        h = ln(x, pre + "ln1")
        r, k, v, g, w = project(h)
        rr, rk, rv, _, rw = project(O.reverse_x(h, rev))
        y = O.wkv6_forward(r, k, v, w, P[a + "time_faaaa"])
        ry = O.reverse_x(O.wkv6_forward(rr, rk, rv, rw, P[a + "time_faaaa"]), rev)
        att = O.groupnorm_gate((y + ry) / 2, g, P[a + "ln_x.weight"], P[a + "ln_x.bias"], H, 1e-5 * 64) @ P[a + "output.weight"].T
        x = x + att
        f = pre + "ffn."
        h = ln(x, pre + "ln2")
        xx = torch.nn.functional.pad(h, (0, 0, 1, -1)) - h
        kk = torch.relu((h + xx * P[f + "time_maa_k"]) @ P[f + "key.weight"].T) ** 2
        x = x + torch.sigmoid((h + xx * P[f + "time_maa_r"]) @ P[f + "receptance.weight"].T) * (kk @ P[f + "value.weight"].T)
    hidden = ln(x, "ln_out")
    _close(hidden, torch.from_numpy(z["hidden"]), 2e-4, rtol=1e-3)
    emb, pos = O.eos_gather(hidden, idx, int(z["emb_id"]))
    _close(emb, torch.from_numpy(z["emb"]), 2e-4, rtol=1e-3)
    assert pos.tolist() == [47, 30, 12, 1]
