"""GPU debugging aid (not a pytest module): per-gradient error of the tensor-core backward against the
fp64 oracle on a few shapes, with the error located by chunk / token, plus a quick timing at the
1B6 shape.  Usage on the GPU box:  python -m profiles.debug_parity [--time]"""
import sys

import torch

from oracle import wkv6_oracle as O
from tests.util import BF16_MAXABS_ABS, BF16_MAXABS_REL, make_inputs, relrms


def run(M, B, T, H, decay, seed, state=None, impl="auto"):
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=seed, decay=decay)
    C = H * 64
    dev = "cuda"
    if state is None:
        ref = O.wkv6_backward(r, k, v, w, u, gy)
    else:
        ref = O.wkv6_backward(r, k, v, w, u, gy, s=state, s_layout="infctx")
    M.set_impl(impl)
    leaves = [t.clone().to(dev).requires_grad_(True) for t in (r, k, v, w, u)]
    if state is None:
        y = M.RUN_CUDA_RWKV6(B, T, C, H, *leaves)
        names = ("gr", "gk", "gv", "gw", "gu")
        grads_of = leaves
    else:
        s_leaf = state.clone().to(dev).requires_grad_(True)
        y, _ = M.RUN_CUDA_RWKV6_STATE(B, T, C, H, *leaves, s_leaf.clone())
        names = ("gr", "gk", "gv", "gw", "gu", "gs")
        grads_of = leaves + [s_leaf]
    y.backward(gy.to(dev))
    torch.cuda.synchronize()
    M.set_impl("auto")
    out = [f"{impl:5s} B{B} T{T} H{H} {decay:5s}" + (" state" if state is not None else "")]
    for name, got_t in [("y", y)] + list(zip(names, [t.grad for t in grads_of])):
        got = got_t.detach().double().cpu()
        rf = ref[name].double()
        err = (got - rf).abs()
        bound = BF16_MAXABS_REL * rf.abs().max().item() + BF16_MAXABS_ABS
        flag = "" if (err.max().item() <= bound and relrms(got, rf) <= 1e-2) else " <-- FAIL"
        loc = ""
        if flag and got.dim() == 3:
            per_t = err.amax(dim=(0, 2))
            worst = per_t.argmax().item()
            nz = (per_t > bound).nonzero().flatten().tolist()
            loc = f" worst t={worst} bad_t={nz[:12]}{'...' if len(nz) > 12 else ''}"
        out.append(f"  {name}: relrms {relrms(got, rf):.2e} max/bound {err.max().item() / bound:.2f}{flag}{loc}")
    print("\n".join(out), flush=True)


def timing(M):
    B, T, H = 8, 4096, 32
    C = H * 64
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=1, decay="model", device="cuda")
    for impl, decay in (("tc", "model"), ("tc", "randn"), ("simt", "model")):
        M.set_impl(impl)
        r, k, v, w, u, gy = make_inputs(B, T, H, seed=1, decay=decay, device="cuda")
        leaves = [t.clone().requires_grad_(True) for t in (r, k, v, w, u)]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for it in range(4):
            for t in leaves:
                t.grad = None
            ev[0].record()
            y = M.RUN_CUDA_RWKV6(B, T, C, H, *leaves)
            ev[1].record()
            y.backward(gy)
            ev[2].record()
            torch.cuda.synchronize()
        print(f"{impl} {decay}: fwd {ev[0].elapsed_time(ev[1]):.3f} ms  bwd {ev[1].elapsed_time(ev[2]):.3f} ms", flush=True)
    M.set_impl("auto")


def main():
    import rwkv_lm_ext_b200 as M
    M.load()
    for (B, T, H, decay) in ((1, 17, 1, "model"), (1, 64, 1, "model"), (2, 64, 2, "model"), (1, 65, 1, "model"),
                             (1, 130, 3, "model"), (1, 257, 1, "model"), (1, 1024, 2, "model"),
                             (1, 17, 1, "randn"), (2, 64, 2, "randn"), (1, 257, 1, "randn")):
        run(M, B, T, H, decay, seed=B * 1000 + T, impl="tc")
    s0 = (torch.randn(2, 2, 64, 64, generator=torch.Generator().manual_seed(1)) * 0.5).bfloat16()
    run(M, 2, 96, 2, "model", 77, state=s0)
    run(M, 2, 200, 2, "model", 78, state=s0)
    if "--time" in sys.argv:
        timing(M)


if __name__ == "__main__":
    main()
