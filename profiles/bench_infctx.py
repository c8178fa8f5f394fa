"""BASELINE config 5 at the operator level: 3B shape (H = 40, C = 2560), B = 1, a 64k-token context as 16 chunks
of 4096 tokens with the WKV state carried between the calls (RUN_CUDA_RWKV6_STATE, src/model.py:780).
Reports tokens/s (forward + backward per chunk, truncated BPTT like the reference) and the drift of the final
state against the exact fp32 SIMT kernels for the two ways of carrying it: bf16 at every chunk boundary (what
the reference does, src/infctx_module.py:36-38) and fp32 (the extension).  CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_inputs

M.load()
B, T, H, NCH = 1, 4096, 40, 16
C = H * 64
chunks = [make_inputs(B, T, H, seed=100 + i, decay="model", device="cuda") for i in range(NCH)]


def run(state_dtype, impl="auto", grad=True):
    M.set_impl(impl)
    try:
        s = torch.zeros(B, H, 64, 64, device="cuda", dtype=state_dtype)
        for (r, k, v, w, u, gy) in chunks:
            if grad:
                leaves = [t.detach().requires_grad_(True) for t in (r, k, v, w, u)]
                y, s = M.RUN_CUDA_RWKV6_STATE(B, T, C, H, *leaves, s.detach().clone())
                y.backward(gy)
            else:
                with torch.no_grad():
                    y, s = M.RUN_CUDA_RWKV6_STATE(B, T, C, H, r, k, v, w, u, s.clone())
        return s
    finally:
        M.set_impl("auto")


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


ms = timeit(lambda: run(torch.bfloat16))
ms32 = timeit(lambda: run(torch.float32))
ms_fwd = timeit(lambda: run(torch.float32, grad=False))
ref = run(torch.float32, impl="simt", grad=False).double()
rel = lambda s: ((s.double() - ref).norm() / ref.norm()).item()
print(json.dumps({"shape": {"B": B, "H": H, "chunk": T, "chunks": NCH}, "fwd_bwd_ms_per_64k_tokens_bf16_state": round(ms, 3),
                  "fwd_bwd_ms_per_64k_tokens_fp32_state": round(ms32, 3), "fwd_only_ms_per_64k_tokens": round(ms_fwd, 3),
                  "tokens_per_s_fwd_bwd": round(NCH * T / ms * 1e3), "tokens_per_s_fwd_only": round(NCH * T / ms_fwd * 1e3),
                  "final_state_relrms_bf16_carry": rel(run(torch.bfloat16, grad=False)),
                  "final_state_relrms_fp32_carry": rel(run(torch.float32, grad=False))}))
