"""add_layernorm forward at [32768, 2048] (8 B/element: x, delta in; x_new, y out).  usage: python profiles/bench_add_ln.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200 import heads
M.load()
R, D = 32768, 2048
x = torch.randn(R, D, device="cuda").bfloat16(); d = torch.randn(R, D, device="cuda").bfloat16()
w = torch.randn(D, device="cuda").bfloat16(); b = torch.randn(D, device="cuda").bfloat16()
with torch.no_grad():
    for _ in range(5):
        heads.add_layernorm(x, d, w, b, 1e-5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50):
        heads.add_layernorm(x, d, w, b, 1e-5)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(json.dumps({"kernel": "add_layernorm forward", "rows": R, "D": D, "ms": round(ms, 4), "GBs": round(R * D * 8 / ms / 1e6, 1), "frac_of_6557.8": round(R * D * 8 / ms / 1e6 / 6557.8, 3)}))
