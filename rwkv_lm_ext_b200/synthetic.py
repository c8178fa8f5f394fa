"""Synthetic inputs of the benchmark / test shapes (there are no datasets or checkpoints offline)."""
import torch


def make_inputs(B, T, H, seed, decay="randn", device="cpu", u_scale=0.3):
    """bf16 r,k,v,w [B,T,H*64], u [H,64], gy [B,T,H*64].  decay: "randn" (the reference tests'
    w ~ N(0,1), tests/test_cpu.py:266) or "model" (time_decay-like U(-6,-1)+0.3N(0,1),
    src/model.py:407-411)."""
    g = torch.Generator().manual_seed(seed)
    C = H * 64
    r, k, v, gy = (torch.randn(B, T, C, generator=g).bfloat16() for _ in range(4))
    if decay == "randn":
        w = torch.randn(B, T, C, generator=g).bfloat16()
    else:
        w = (torch.rand(B, T, C, generator=g) * 5 - 6 + 0.3 * torch.randn(B, T, C, generator=g)).bfloat16()
    u = (torch.randn(H, 64, generator=g) * u_scale).bfloat16()
    return tuple(t.to(device) for t in (r, k, v, w, u, gy))
