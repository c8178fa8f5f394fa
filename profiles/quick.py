"""Quick A/B timer: WKV6 forward and backward at B T H (default 8 4096 32) with CUDA events, plus a parity check
against the exact SIMT kernels.  usage: python profiles/quick.py [B T H] [decay]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_inputs

B, T, H = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 4096, 32)
decay = sys.argv[4] if len(sys.argv) > 4 else "model"
C = H * 64
M.load()
r, k, v, w, u, gy = make_inputs(B, T, H, seed=0, decay=decay, device="cuda")


def loop(n, bwd):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        leaves = [t.detach().requires_grad_(bwd) for t in (r, k, v, w, u)]
        y = M.RUN_CUDA_RWKV6(B, T, C, H, *leaves)
        if bwd:
            y.backward(gy)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, y.detach(), [t.grad for t in leaves]


def run(n):
    tf, _, _ = loop(n, False)
    ts, y, g = loop(n, True)
    return tf, ts - tf, y, g


run(5)
tf, tb, y, g = run(40)
res = {"shape": [B, T, H], "decay": decay, "fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4), "step_ms": round(tf + tb, 4)}
if B * T * H <= 8 * 4096 * 32:
    M.set_impl("simt")
    _, _, y2, g2 = run(1)
    M.set_impl("auto")
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
    res["relrms_vs_simt"] = {n: round(rel(a, b), 5) for n, a, b in zip(("y", "gr", "gk", "gv", "gw", "gu"), [y] + g, [y2] + g2)}
print(json.dumps(res))
