"""Batched, GPU-resident forward of the reference's bi-directional encoder on the fused kernels
(SURVEY.md section 8(f) rank 3).

`bi_encoder_hidden(model, idx)` / `bi_encoder_encode(model, idx)` compute `RwkvEncoder.forward(idx, True)[1]`
and `RwkvEncoder.encode_sentence(idx)` (src/model_encoder_run.py:296-350; the trainer twin is
`src/model_ext.py:398-437`) on any module tree with the reference's attribute names

    emb, blocks[i].{ln0 (i == 0), ln1, ln2, att, ffn}, ln_out, args.emb_id / args.pad_id (or emb_id, pad_id)

Per layer ("semantics 2": a causal pass over the tokens and one over each row's reversed prefix, averaged):

    mask, rev_idx           one kernel on the device (the reference loops over the batch on the host)
    r,k,v,g,w  (x2)         tmix_x060_project: shift-lerp + ddlerp kernels around the cuBLAS Linears
    reverse gathers         128-bit gather kernel
    WKV6 (x2)               tcgen05 / TMA chunked kernel
    GroupNorm * silu(gate)  one kernel
    channel mix             cmix_x060_forward: 3 fused elementwise kernels around the two GEMMs
"""
import torch

from . import cmix, heads, ops, tmix


def _ids(model):
    a = getattr(model, "args", None)
    emb_id = getattr(a, "emb_id", getattr(model, "emb_id", 1))
    pad_id = getattr(a, "pad_id", getattr(model, "pad_id", 0))
    return int(emb_id), int(pad_id)


def bi_tmix_forward(att, x, rev_idx):
    """bi_att_forward_batch (src/model_encoder_run.py:64-75, :77-93) on the fused kernels."""
    B, T, C = x.shape
    H = att.time_faaaa.shape[0]
    r, k, v, g, w = tmix.tmix_x060_project(att, x)
    rr, rk, rv, _, rw = tmix.tmix_x060_project(att, heads.reverse_x(x, rev_idx))
    y = ops.RUN_CUDA_RWKV6(B, T, C, H, r, k, v, w, att.time_faaaa)
    ry = ops.RUN_CUDA_RWKV6(B, T, C, H, rr, rk, rv, rw, att.time_faaaa)       # still in reversed token order
    z = heads.groupnorm_gate_pair(y, ry, rev_idx, g, att.ln_x.weight, att.ln_x.bias, H, att.ln_x.eps, gate_act="silu")
    return att.output(z)


def bi_encoder_hidden(model, idx):
    """ln_out(blocks(emb(idx)))  [B,T,D] in the model's dtype (bf16 weights expected)."""
    emb_id, pad_id = _ids(model)
    idx = idx.contiguous()
    _, rev_idx = heads.create_mask_and_rev_idx(idx, emb_id, pad_id)
    x = model.emb(idx)
    for i, blk in enumerate(model.blocks):
        if i == 0 and hasattr(blk, "ln0"):
            x = blk.ln0(x)
        x = x + bi_tmix_forward(blk.att, blk.ln1(x), rev_idx)
        ffn = blk.ffn
        if all(hasattr(ffn, n) for n in ("time_maa_k", "time_maa_r", "key", "receptance", "value")):
            x = x + cmix.cmix_x060_forward(ffn, blk.ln2(x))          # x060 channel mix on the fused kernels
        else:
            x = x + ffn(blk.ln2(x))
    return model.ln_out(x)


def bi_encoder_encode(model, idx):
    """encode_sentence: the hidden state at each row's first `emb_id` token.  [B,D]."""
    emb_id, _ = _ids(model)
    hidden = bi_encoder_hidden(model, idx)
    return heads.eos_gather(hidden.contiguous(), idx, emb_id)[0]
