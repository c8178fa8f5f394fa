"""Inference surface: `RUN_RWKV_6` (src/model_run.py:75, B = 1, fp32 state in place) at the 1B6 shape,
T = 1 (token-by-token decode, the src/model_run.py:316 use) and T = 1024 (one prefill chunk), against the reference's
own rwkv6.cu kernel (oracle/_ref) called the way `RWKV_6.forward` calls it, i.e. behind its eager
exp(-exp(w.float())).  GPU time per call from CUDA events over back-to-back calls (eager: includes launch gaps) and from
a CUDA-graph replay (kernels only).  usage: python profiles/bench_inference.py  -> one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_inputs
from oracle import ref_cuda

M.load()
H = 32
C = H * 64
out = {"shape": "B=1 H=32 N=64 (1B6), bf16, fp32 state"}


def timed(fn, n):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3          # us


def graphed(fn, n):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timed(g.replay, n)


for T in (1, 1024):
    r, k, v, w, u, _ = make_inputs(1, T, H, seed=0, decay="model", device="cuda")
    r, k, v, w = (t[0].contiguous() for t in (r, k, v, w))
    st = torch.zeros(H, 64, 64, device="cuda")
    ours = lambda: M.RUN_RWKV_6(1, T, C, H, st, r, k, v, w, u)
    res = {"ours_eager_us": round(timed(ours, 200), 2), "ours_graph_us": round(graphed(ours, 200), 2)}
    if ref_cuda.available("rwkv6"):
        st2 = torch.zeros(H, 64, 64, device="cuda")
        ref = lambda: ref_cuda.rwkv6_forward_bf16(st2, r, k, v, torch.exp(-torch.exp(w.float())).contiguous(), u)
        res["reference_eager_us"] = round(timed(ref, 200), 2)
        y1, _ = M.RUN_RWKV_6(1, T, C, H, torch.zeros(H, 64, 64, device="cuda"), r, k, v, w, u)
        y2 = ref_cuda.rwkv6_forward_bf16(torch.zeros(H, 64, 64, device="cuda"), r, k, v, torch.exp(-torch.exp(w.float())).contiguous(), u)
        torch.cuda.synchronize()
        res["relrms_vs_reference"] = float(((y1[0].float() - y2.float()).norm() / y2.float().norm()).item())
    out[f"T={T}"] = res
print(json.dumps(out))
