"""Memory-bound pieces around the recurrence, with the reference's semantics:
eos / class-token gather (src/model_ext.py:209-211,1765), pooling (src/model_ext.py:1708-1738,
src/model_run.py:777-797), mask / reverse index / reverse gather (src/model_ext.py:398-419),
token-shift ddlerp (src/model.py:437-449) and GroupNorm*gate (src/model.py:461-467).
All arithmetic is in libwkv6_b200.so; these wrappers only allocate outputs."""
import torch

from . import _lib
from ._lib import check, ptr, stream_of


def _cuda(t):
    if not t.is_cuda:
        raise _lib.Wkv6B200Error("rwkv_lm_ext_b200 kernels run on CUDA tensors only (no CPU fallback)")


def eos_index(idx: torch.Tensor, token_id: int) -> torch.Tensor:
    """== torch.eq(idx, token_id).int().argmax(-1): first occurrence, 0 if absent.  int64 [B]."""
    _cuda(idx)
    assert idx.dtype == torch.int64 and idx.dim() == 2
    idx = idx.contiguous()
    B, T = idx.shape
    pos = torch.empty(B, dtype=torch.int64, device=idx.device)
    check(_lib.load().eos_index_i64(B, T, ptr(idx), int(token_id), ptr(pos), stream_of(idx)), "eos_index_i64")
    return pos


def gather_rows(x: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """== x[torch.arange(B), pos]  for x bf16 [B,T,D]."""
    _cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 3 and pos.dtype == torch.int64
    x = x.contiguous()
    B, T, D = x.shape
    out = torch.empty(B, D, dtype=torch.bfloat16, device=x.device)
    check(_lib.load().gather_rows_bf16(B, T, D, ptr(x), ptr(pos.contiguous()), ptr(out), stream_of(x)), "gather_rows_bf16")
    return out


def eos_gather(x: torch.Tensor, idx: torch.Tensor, token_id: int):
    """The classification / embedding head gather: (x[b, first eos], pos)."""
    pos = eos_index(idx, token_id)
    return gather_rows(x, pos), pos


_KIND = {"weightedmean": 0, "lasttoken": 1, "avg": 2}


def pooling(x: torch.Tensor, actual_len: torch.Tensor, pooling_type: str, variant: str = "train") -> torch.Tensor:
    """variant="train": src/model_ext.py:1708-1738 (bf16 result for weightedmean / avg);
    variant="infer": src/model_run.py:777-797 (L = actual_len + 1, fp32 result)."""
    _cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 3
    if pooling_type == "lasttoken":
        return gather_rows(x, actual_len.to(torch.int64))
    if pooling_type not in _KIND or (variant == "infer" and pooling_type == "avg"):
        raise ValueError(pooling_type)
    x = x.contiguous()
    B, T, D = x.shape
    out = torch.empty(B, D, dtype=torch.float32, device=x.device)
    check(_lib.load().pooling_bf16(_KIND[pooling_type], 1 if variant == "infer" else 0, B, T, D, ptr(x),
                                   ptr(actual_len.to(torch.int64).contiguous()), ptr(out), stream_of(x)), "pooling_bf16")
    return out.bfloat16() if variant == "train" else out


def create_mask_and_rev_idx(idx: torch.Tensor, emb_id: int = 1, pad_id: int = 0):
    """create_mask + reverse_x_idx (src/model_ext.py:398-417) in one launch, on the device
    (the reference builds rev_idx with a Python loop over the batch on the host)."""
    _cuda(idx)
    assert idx.dtype == torch.int64 and idx.dim() == 2
    idx = idx.contiguous()
    B, T = idx.shape
    mask = torch.empty(B, T, dtype=torch.int32, device=idx.device)
    rev = torch.empty(B, T, dtype=torch.int64, device=idx.device)
    check(_lib.load().create_mask_rev_idx(B, T, ptr(idx), int(emb_id), int(pad_id), ptr(mask), ptr(rev), stream_of(idx)),
          "create_mask_rev_idx")
    return mask, rev


def reverse_x(x: torch.Tensor, rev_idx: torch.Tensor) -> torch.Tensor:
    """== torch.gather(x, 1, rev_idx[..., None].expand(-1, -1, D))  (src/model_ext.py:418-419)."""
    _cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 3 and rev_idx.dtype == torch.int64
    x = x.contiguous()
    B, T, D = x.shape
    out = torch.empty_like(x)
    check(_lib.load().gather_tokens_bf16(B, T, D, ptr(x), ptr(rev_idx.contiguous()), ptr(out), stream_of(x)),
          "gather_tokens_bf16")
    return out


def tmix_shift_lerp(x, maa_x, shift_state=None):
    """xxx = x + (time_shift(x) - x) * time_maa_x  (src/model.py:437-439)."""
    _cuda(x)
    assert x.dtype == torch.bfloat16
    x = x.contiguous()
    B, T, C = x.shape
    out = torch.empty_like(x)
    check(_lib.load().tmix_shift_lerp_bf16(B, T, C, ptr(x), ptr(shift_state), ptr(maa_x.contiguous().view(-1)), ptr(out),
                                           stream_of(x)), "tmix_shift_lerp_bf16")
    return out


def tmix_ddlerp_mix(x, maa_wkvrg, m, shift_state=None):
    """xw,xk,xv,xr,xg = x + xx * (time_maa_n + m_n)  (src/model.py:444-448).
    maa_wkvrg bf16 [5,C]; m bf16 [5,B,T,C] (the LoRA bmm output).  Returns a [5,B,T,C] tensor."""
    _cuda(x)
    assert x.dtype == torch.bfloat16 and m.dtype == torch.bfloat16
    x = x.contiguous()
    B, T, C = x.shape
    out = torch.empty(5, B, T, C, dtype=torch.bfloat16, device=x.device)
    check(_lib.load().tmix_ddlerp_mix_bf16(B, T, C, ptr(x), ptr(shift_state), ptr(maa_wkvrg.contiguous()),
                                           ptr(m.contiguous()), ptr(out), stream_of(x)), "tmix_ddlerp_mix_bf16")
    return out


def groupnorm_gate(y, g, ln_w, ln_b, H, eps):
    """ln_x(y.view(B*T, C)).view(B,T,C) * g  (src/model.py:461-467, without the output Linear)."""
    _cuda(y)
    assert y.dtype == torch.bfloat16 and g.dtype == torch.bfloat16
    y = y.contiguous()
    B, T, C = y.shape
    out = torch.empty_like(y)
    check(_lib.load().groupnorm_gate_bf16(B * T, C, H, float(eps), ptr(y), ptr(g.contiguous()), ptr(ln_w.contiguous()),
                                          ptr(ln_b.contiguous()), ptr(out), stream_of(y)), "groupnorm_gate_bf16")
    return out
