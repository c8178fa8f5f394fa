"""Time-mix layer, fwd+bwd, at the 1B6 shape (B=8, T=4096, C=2048, H=32): the fused layer
(rwkv_lm_ext_b200.tmix) against the reference's eager chain (src/model.py:434-468 restated with torch ops)
around the SAME WKV6 operator, so the difference is the elementwise chain only.  CUDA events.
usage: python profiles/bench_tmix.py [B T]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import rwkv_lm_ext_b200 as M

M.load()
dev = "cuda"
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 4096)
H = 32
C = H * 64
torch.manual_seed(0)
layer = M.Tmix_x060(C, H)
with torch.no_grad():
    for n, p in layer.named_parameters():
        if n.startswith("time_maa_w1") or n.startswith("time_maa_w2") or n.startswith("time_decay_w"):
            p.uniform_(-1e-2, 1e-2)
        elif n == "time_decay":
            p.copy_(-6 + 5 * torch.rand_like(p))
        elif n.startswith("time_maa"):
            p.uniform_(0, 1)
        elif n == "time_faaaa":
            p.normal_(0, 0.3)
layer = layer.bfloat16().to(dev)
x = torch.randn(B, T, C, device=dev).bfloat16().requires_grad_(True)
gout = torch.randn(B, T, C, device=dev).bfloat16()
shift = torch.nn.ZeroPad2d((0, 0, 1, -1))


def eager(l, x):
    xx = shift(x) - x
    xxx = x + xx * l.time_maa_x
    xxx = torch.tanh(xxx @ l.time_maa_w1).view(B * T, 5, -1).transpose(0, 1)
    mw, mk, mv, mr, mg = torch.bmm(xxx, l.time_maa_w2).view(5, B, T, -1).unbind(0)
    xw = x + xx * (l.time_maa_w + mw)
    xk = x + xx * (l.time_maa_k + mk)
    xv = x + xx * (l.time_maa_v + mv)
    xr = x + xx * (l.time_maa_r + mr)
    xg = x + xx * (l.time_maa_g + mg)
    r, k, v, g = l.receptance(xr), l.key(xk), l.value(xv), F.silu(l.gate(xg))
    w = l.time_decay + torch.tanh(xw @ l.time_decay_w1) @ l.time_decay_w2
    y = M.RUN_CUDA_RWKV6(B, T, C, H, r, k, v, w, l.time_faaaa)
    return l.output(l.ln_x(y.view(B * T, C)).view(B, T, C) * g)


def step(fn):
    x.grad = None
    layer.zero_grad(set_to_none=True)
    fn(layer, x).backward(gout)


def timeit(fn, n=10):
    for _ in range(3):
        step(fn)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        step(fn)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def gemm_only():
    # the 5 CxC Linears + LoRA GEMMs, fwd + 2x bwd: what cuBLAS alone costs (lower bound for both)
    a = torch.randn(B * T, C, device=dev).bfloat16()
    wt = torch.randn(C, C, device=dev).bfloat16()
    for _ in range(3):
        a @ wt
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(15):
        a @ wt
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)       # 15 GEMMs = 5 Linears x (fwd + dgrad + wgrad)


t_fused, t_eager, t_gemm = timeit(M.tmix_x060_forward), timeit(eager), gemm_only()
print(json.dumps({"shape": [B, T, C], "fused_layer_ms": round(t_fused, 3), "eager_chain_ms": round(t_eager, 3),
                  "speedup": round(t_eager / t_fused, 2), "five_linears_fwd_bwd_gemm_ms": round(t_gemm, 3),
                  "tokens_per_s_fused": round(B * T / t_fused * 1e3)}))
