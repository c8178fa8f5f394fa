"""Wave quantisation of the stream-per-CTA mapping: fwd+bwd time of the WKV6 op at H=32, T=4096 for batch
sizes around the benchmark's B=8 (256 streams on 148 SMs x 2 resident CTAs = 296 slots).  CUDA events.
usage: python profiles/bench_waves.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_inputs

M.load()
PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
T, H = 4096, 32
C = H * 64
for B in (4, 8, 9, 16, 18, 37):
    r, k, v, w, u, gy = make_inputs(B, T, H, 0, decay="model", device="cuda")
    ts = [t.detach().requires_grad_(True) for t in (r, k, v, w, u)]

    def step():
        for t in ts:
            t.grad = None
        M.RUN_CUDA_RWKV6(B, T, C, H, *ts).backward(gy)
    for _ in range(5):
        step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    gbs = B * T * C * 28 / ms / 1e6
    print(json.dumps({"B": B, "streams": B * H, "slots_296_waves": round(B * H / 296, 2), "ms": round(ms, 4),
                      "tokens_per_s": round(B * T / ms * 1e3), "hbm_frac": round(gbs / PEAK, 3)}), flush=True)
    del r, k, v, w, u, gy, ts
