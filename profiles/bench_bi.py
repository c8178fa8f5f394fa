"""wkv6_bi forward and forward+backward at the bi-encoder shape of BASELINE config 3 (64 passages x 512 tokens per GPU, 1B6 heads):
tensor-core composition vs the SIMT kernels.  usage: python profiles/bench_bi.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_inputs

M.load()
B, T, H = 64, 512, 32
C = H * 64
r, k, v, w, u, gy = make_inputs(B, T, H, seed=0, decay="model", device="cuda")
mask = torch.ones(B, T, dtype=torch.int32, device="cuda")
lens = torch.randint(128, 513, (B,), generator=torch.Generator().manual_seed(0))
for b, n in enumerate(lens.tolist()):
    mask[b, n - 1:] = 0
res = {}
with torch.no_grad():
    for impl in ("auto", "simt"):
        M.set_impl(impl)
        for _ in range(3):
            y = M.RUN_CUDA_RWKV6_BI(B, T, C, H, mask, r, k, v, w, u)
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            y = M.RUN_CUDA_RWKV6_BI(B, T, C, H, mask, r, k, v, w, u)
        b_.record()
        torch.cuda.synchronize()
        res[impl] = (a.elapsed_time(b_) / 10, y)
    M.set_impl("auto")
# forward + backward through autograd
ts = [t.clone().requires_grad_(True) for t in (r, k, v, w, u)]
fb = {}
for impl in ("auto", "simt"):
    M.set_impl(impl)

    def step():
        for t in ts:
            t.grad = None
        M.RUN_CUDA_RWKV6_BI(B, T, C, H, mask, *ts).backward(gy)
    for _ in range(3):
        step()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        step()
    b_.record()
    torch.cuda.synchronize()
    fb[impl] = (a.elapsed_time(b_) / 10, [t.grad.clone() for t in ts])
M.set_impl("auto")
grel = max(((x.float() - y_.float()).norm() / y_.float().norm()).item() for x, y_ in zip(fb["auto"][1], fb["simt"][1]))
rel = ((res["auto"][1].float() - res["simt"][1].float()).norm() / res["simt"][1].float().norm()).item()
print(json.dumps({"shape": [B, T, H], "tc_ms": round(res["auto"][0], 3), "simt_ms": round(res["simt"][0], 3),
                  "passages_per_s_tc": round(B / (res["auto"][0] * 1e-3)), "relrms_tc_vs_simt": rel,
                  "fwd_bwd_tc_ms": round(fb["auto"][0], 3), "fwd_bwd_simt_ms": round(fb["simt"][0], 3),
                  "grad_relrms_tc_vs_simt": grel}))
