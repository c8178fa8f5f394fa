"""GPU debugging aid (not a pytest module): forward-only check of the role-uniform kernel against the
fp64 oracle.  Usage on the GPU box:  WKV6_B200_FWD3=1 python -m tests.debug_fwd3 [--time]"""
import sys

import torch

from oracle import wkv6_oracle as O
from tests.util import BF16_MAXABS_ABS, BF16_MAXABS_REL, make_inputs, relrms


def check(name, got, ref):
    got, ref = got.detach().double().cpu(), ref.double()
    err = (got - ref).abs()
    bound = BF16_MAXABS_REL * ref.abs().max().item() + BF16_MAXABS_ABS
    bad = not (err.max().item() <= bound and relrms(got, ref) <= 1e-2) or not torch.isfinite(got).all()
    loc = ""
    if bad and got.dim() == 3:
        per_t = err.amax(dim=(0, 2))
        nz = (per_t > bound).nonzero().flatten().tolist()
        loc = f" bad_t={nz[:16]}{'...' if len(nz) > 16 else ''} nan={int((~torch.isfinite(got)).sum())}"
    print(f"  {name}: relrms {relrms(got, ref):.2e} max/bound {err.max().item() / bound:.2f}{' <-- FAIL' + loc if bad else ''}", flush=True)


def main():
    import rwkv_lm_ext_b200 as M
    M.load()
    dev = "cuda"
    with torch.no_grad():
        for (B, T, H, decay) in ((1, 1, 1, "model"), (1, 17, 1, "model"), (1, 64, 1, "model"), (2, 64, 2, "model"),
                                 (1, 65, 1, "model"), (1, 130, 3, "model"), (2, 257, 2, "model"), (1, 1024, 2, "model"),
                                 (1, 17, 1, "randn"), (2, 64, 2, "randn"), (1, 300, 2, "randn")):
            r, k, v, w, u, gy = make_inputs(B, T, H, seed=B * 1000 + T, decay=decay)
            print(f"B{B} T{T} H{H} {decay}")
            y = M.RUN_CUDA_RWKV6(B, T, H * 64, H, *(t.to(dev) for t in (r, k, v, w, u)))
            check("y", y, O.wkv6_forward(r, k, v, w, u))
        B, T, H = 2, 200, 2
        r, k, v, w, u, gy = make_inputs(B, T, H, seed=7, decay="model")
        s0 = (torch.randn(B, H, 64, 64, generator=torch.Generator().manual_seed(1)) * 0.5).bfloat16()
        y_ref, s_ref = O.wkv6infctx_forward(r, k, v, w, u, s0)
        for dt in (torch.bfloat16, torch.float32):
            s = s0.to(dev).to(dt)
            y, s_out = M.RUN_CUDA_RWKV6_STATE(B, T, H * 64, H, *(t.to(dev) for t in (r, k, v, w, u)), s)
            print(f"infctx state {dt}")
            check("y", y, y_ref)
            check("sT", s_out, s_ref)
    if "--time" in sys.argv:
        B, T, H = 8, 4096, 32
        r, k, v, w, u, gy = make_inputs(B, T, H, seed=1, decay="model", device=dev)
        with torch.no_grad():
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for it in range(5):
                ev[0].record()
                y = M.RUN_CUDA_RWKV6(B, T, H * 64, H, r, k, v, w, u)
                ev[1].record()
                torch.cuda.synchronize()
            print(f"fwd (no grad): {ev[0].elapsed_time(ev[1]):.3f} ms", flush=True)
            M.set_impl("simt")
            ys = M.RUN_CUDA_RWKV6(B, T, H * 64, H, r, k, v, w, u)
            M.set_impl("auto")
            print("full shape vs simt relrms", relrms(y, ys))


if __name__ == "__main__":
    main()
