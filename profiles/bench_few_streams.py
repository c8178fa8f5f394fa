"""Calls with few (b,h) streams: forward time with the time-axis segmentation of csrc/seg_scan.cu
(used for T >= 8192) against the plain one-CTA-per-stream launch (WKV6B200_NO_SEG=1); the fwd+bwd
column is not segmented in either run and is there as a control.  CUDA events.
usage: python profiles/bench_few_streams.py"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import rwkv_lm_ext_b200 as M
    from rwkv_lm_ext_b200.synthetic import make_inputs
    M.load()
    out = {}
    for (B, T, H) in ((1, 4096, 40), (1, 16384, 40), (1, 65536, 32), (2, 16384, 32)):
        r, k, v, w, u, gy = make_inputs(B, T, H, 0, decay="model", device="cuda")
        C = H * 64
        ts = [t.detach().requires_grad_(True) for t in (r, k, v, w, u)]

        def fwd():
            with torch.no_grad():
                M.RUN_CUDA_RWKV6(B, T, C, H, r, k, v, w, u)

        def fb():
            for t in ts:
                t.grad = None
            M.RUN_CUDA_RWKV6(B, T, C, H, *ts).backward(gy)
        res = []
        for fn in (fwd, fb):
            for _ in range(5):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(30):
                fn()
            b.record()
            torch.cuda.synchronize()
            res.append(round(a.elapsed_time(b) / 30, 4))
        out[f"B{B}_T{T}_H{H}"] = {"fwd_ms": res[0], "fwd_bwd_ms": res[1]}
    print(json.dumps(out))
else:
    rows = {}
    for tag, env in (("segmented", {}), ("plain", {"WKV6B200_NO_SEG": "1"})):
        o = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env={**os.environ, **env},
                           capture_output=True, text=True)
        if o.returncode != 0:
            print(o.stderr[-2000:])
            sys.exit(1)
        rows[tag] = json.loads(o.stdout.strip().splitlines()[-1])
    for shape in rows["plain"]:
        p, s = rows["plain"][shape], rows["segmented"][shape]
        print(json.dumps({"shape": shape, "plain": p, "segmented": s,
                          "fwd_speedup": round(p["fwd_ms"] / s["fwd_ms"], 2),
                          "fwd_bwd_speedup": round(p["fwd_bwd_ms"] / s["fwd_bwd_ms"], 2)}))
