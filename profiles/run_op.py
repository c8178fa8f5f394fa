"""Tiny driver for ncu: runs the WKV6 op a few times at a given shape.
usage: python profiles/run_op.py [fwd|fwdbwd] B T H [impl]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_inputs

mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B, T, H = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (8, 4096, 32)
impl = sys.argv[5] if len(sys.argv) > 5 else "auto"
decay = sys.argv[6] if len(sys.argv) > 6 else "model"
M.load(); M.set_impl(impl)
r, k, v, w, u, gy = make_inputs(B, T, H, seed=0, decay=decay, device="cuda")
C = H * 64
for it in range(4):
    leaves = [t.detach().requires_grad_(mode == "fwdbwd") for t in (r, k, v, w, u)]
    y = M.RUN_CUDA_RWKV6(B, T, C, H, *leaves)
    if mode == "fwdbwd":
        y.backward(gy)
torch.cuda.synchronize()
print("ok", mode, B, T, H, impl, float(y.float().abs().mean()))
