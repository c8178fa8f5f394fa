"""Launch-bound shapes (SFT buckets, B*T = 2048 tokens, H = 32): fwd+bwd of the operator eagerly and
replayed from a CUDA graph.  CUDA events around 50 iterations."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_inputs

M.load()
H = 32
C = H * 64
for (B, T) in ((32, 64), (8, 256), (2, 1024), (1, 2048)):
    r, k, v, w, u, gy = make_inputs(B, T, H, 0, decay="model", device="cuda")
    ts = [t.clone().requires_grad_(True) for t in (r, k, v, w, u)]

    def step():
        for t in ts:
            t.grad = None
        M.RUN_CUDA_RWKV6(B, T, C, H, *ts).backward(gy)

    def timeit(fn, n=50):
        for _ in range(5):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    eager = timeit(step)
    for t in ts:
        t.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        M.RUN_CUDA_RWKV6(B, T, C, H, *ts).backward(gy)
    graphed = timeit(g.replay)
    print(json.dumps({"B": B, "T": T, "H": H, "eager_ms": round(eager, 4), "graph_ms": round(graphed, 4),
                      "tokens_per_s_graph": round(B * T / graphed * 1e3)}), flush=True)
