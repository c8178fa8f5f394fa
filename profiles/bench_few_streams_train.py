"""Training pair (forward + backward) for calls with few streams, GPU time only: the step is captured in a CUDA
graph and replayed, so host launch latency does not hide the kernels.  Time-axis segmentation (default for
B*H <= 74, T >= 2048) against the plain one-CTA-per-stream launch (WKV6B200_NO_SEG=1).
usage: python profiles/bench_few_streams_train.py"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import rwkv_lm_ext_b200 as M
    from rwkv_lm_ext_b200.synthetic import make_inputs
    M.load()
    out = {}
    for (B, T, H) in ((1, 4096, 40), (1, 2048, 32), (1, 4096, 32), (2, 4096, 32)):
        r, k, v, w, u, gy = make_inputs(B, T, H, 0, decay="model", device="cuda")
        C = H * 64
        ts = [t.clone().requires_grad_(True) for t in (r, k, v, w, u)]

        def step():
            for t in ts:
                t.grad = None
            M.RUN_CUDA_RWKV6(B, T, C, H, *ts).backward(gy)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        for t in ts:
            t.grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            M.RUN_CUDA_RWKV6(B, T, C, H, *ts).backward(gy)
        for _ in range(3):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(30):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        out[f"B{B}_T{T}_H{H}"] = round(a.elapsed_time(b) / 30, 4)
    print(json.dumps(out))
else:
    rows = {}
    for tag, env in (("segmented", {}), ("plain", {"WKV6B200_NO_SEG": "1"})):
        o = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env={**os.environ, **env},
                           capture_output=True, text=True)
        if o.returncode != 0:
            print(o.stderr[-3000:])
            sys.exit(1)
        rows[tag] = json.loads(o.stdout.strip().splitlines()[-1])
    for shape in rows["plain"]:
        p, s = rows["plain"][shape], rows["segmented"][shape]
        print(json.dumps({"shape": shape, "plain_fwd_bwd_ms": p, "segmented_fwd_bwd_ms": s, "speedup": round(p / s, 2)}))
