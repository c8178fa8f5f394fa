"""One forward+backward of the fused time-mix layer at 8x4096x2048 for an ncu launch list (per-kernel times).
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/tmix_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M

M.load()
B, T, C, H = 8, 4096, 2048, 32
dev = torch.device("cuda")
torch.manual_seed(0)
layer = M.Tmix_x060(C, H)
with torch.no_grad():
    for n, p in layer.named_parameters():
        if p.dim() >= 2 and "lora" not in n and "w1" not in n and "w2" not in n:
            p.normal_(0, 0.02)
layer = layer.bfloat16().to(dev)
x = (torch.randn(B, T, C, device=dev) * 0.5).bfloat16().requires_grad_(True)
gout = torch.randn(B, T, C, device=dev).bfloat16()
for it in range(3):
    x.grad = None
    layer.zero_grad(set_to_none=True)
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("measured")
    M.tmix_x060_forward(layer, x).backward(gout)
torch.cuda.synchronize()
print("ok")
