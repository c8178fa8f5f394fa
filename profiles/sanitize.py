"""Small calls through every kernel family, for compute-sanitizer (the reference has no race /
memory checking; this build runs its kernels under `compute-sanitizer --tool memcheck` on ragged shapes).
usage: compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/sanitize.py [part ...]
parts: wkv state infctx bi seg infer layer loss encoder   (default: all)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200 import sft
from rwkv_lm_ext_b200.synthetic import make_inputs, make_bi_encoder, make_passages

parts = set(sys.argv[1:]) or {"wkv", "state", "infctx", "bi", "seg", "infer", "layer", "loss", "encoder"}
dev = "cuda"
M.load()


def leaves(*ts):
    return [t.detach().clone().requires_grad_(True) for t in ts]


if "wkv" in parts:            # ragged T (last chunk of 44 tokens), T < 64, T = 1; both kernel routes
    for impl in ("auto", "simt"):
        M.set_impl(impl)
        for B, T, H in ((2, 300, 2), (1, 37, 3), (3, 1, 1), (1, 64, 2)):
            r, k, v, w, u, gy = make_inputs(B, T, H, seed=1, decay="randn", device=dev)
            ls = leaves(r, k, v, w, u)
            M.RUN_CUDA_RWKV6(B, T, H * 64, H, *ls).backward(gy)
    M.set_impl("auto")
    print("wkv ok")
if "state" in parts:
    B, T, H = 2, 130, 2
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=2, decay="model", device=dev)
    s = (torch.randn(H, 64, 64, device=dev) * 0.1).bfloat16()
    ls = leaves(r, k, v, w, u, s)
    M.WKV_6STATE.apply(B, T, H * 64, H, *ls).backward(gy)
    print("state ok")
if "infctx" in parts:
    B, T, H = 2, 200, 2
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=3, decay="model", device=dev)
    for dt in (torch.bfloat16, torch.float32):
        s = (torch.randn(B, H, 64, 64, device=dev) * 0.1).to(dt)
        ls = leaves(r, k, v, w, u, s)
        y, sT = M.WKV_6STATE_INFCTX.apply(B, T, H * 64, H, *ls)
        (y.float().sum() + sT.float().sum()).backward()
    print("infctx ok")
if "bi" in parts:
    B, T, H = 4, 200, 2
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=4, decay="model", device=dev)
    mask = torch.ones(B, T, dtype=torch.int32, device=dev)
    for b, p in enumerate((0, 63, 130, T)):
        mask[b, p:] = 0
    ls = leaves(r, k, v, w, u)
    M.RUN_CUDA_RWKV6_BI(B, T, H * 64, H, mask, *ls).backward(gy)
    print("bi ok")
if "seg" in parts:            # few streams: time-axis segmentation (training pair and forward only), ragged last segment
    B, T, H = 1, 4096 + 72, 2
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=5, decay="model", device=dev)
    ls = leaves(r, k, v, w, u)
    M.RUN_CUDA_RWKV6(B, T, H * 64, H, *ls).backward(gy)
    with torch.no_grad():
        M.RUN_CUDA_RWKV6(B, T, H * 64, H, r, k, v, w, u)
    print("seg ok")
if "infer" in parts:
    B, T, H = 1, 70, 2
    r, k, v, w, u, _ = make_inputs(B, T, H, seed=6, decay="model", device=dev)
    for dt in (torch.bfloat16, torch.float16, torch.float32):
        st = torch.zeros(B, H, 64, 64, device=dev)
        M.RUN_RWKV_6(B, T, H * 64, H, st, r.to(dt), k.to(dt), v.to(dt), w.to(dt), u.to(dt))
    print("infer ok")
if "layer" in parts:          # the fused time-mix / channel-mix / add+LayerNorm kernels, trainable and frozen parameters
    C, H = 256, 4
    for frozen in (False, True):
        layer = M.Tmix_x060(C, H).to(dev).bfloat16()
        with torch.no_grad():
            for p in layer.parameters():
                p.normal_(0, 0.05)
            layer.ln_x.weight.fill_(1.0)
        for p in layer.parameters():
            p.requires_grad = not frozen
        x = torch.randn(2, 100, C, device=dev).bfloat16().requires_grad_()
        ln_w, ln_b = torch.ones(C, device=dev).bfloat16(), torch.zeros(C, device=dev).bfloat16()
        xn, h = M.add_layernorm(x, torch.randn_like(x), ln_w, ln_b)
        out = layer(h)
        (out.float().sum() + xn.float().sum()).backward()
        ffn = sft._Cmix(C, 448, None).to(dev).bfloat16()
        for p in ffn.parameters():
            p.requires_grad = not frozen
        x2 = torch.randn(2, 100, C, device=dev).bfloat16().requires_grad_()
        ffn(x2).float().sum().backward()
    hid = torch.randn(3, 50, C, device=dev).bfloat16().requires_grad_()
    pos = torch.tensor([0, 20, 49], device=dev)
    for kind in ("weightedmean", "lasttoken", "avg"):
        M.pooling(hid, pos, kind).float().sum().backward()
    print("layer ok")
if "loss" in parts:
    for V in (1000, 65536):
        logits = torch.randn(2, 9, V, device=dev).bfloat16().requires_grad_()
        tgt = torch.randint(0, V, (2, 9), device=dev)
        tgt[0, :3] = -100
        sft.sft_loss(logits, tgt).backward()
    print("loss ok")
if "encoder" in parts:
    model = make_bi_encoder(2, 128, 2, 448, 512, seed=0, device=dev)
    idx = make_passages(4, T=96, vocab=512, seed=0, min_len=20).to(dev)
    with torch.no_grad():
        M.bi_encoder_encode(model, idx)
    m2 = sft.init_like_reference(sft.RwkvSft(layers=2, D=128, H=2, ffn=448, vocab=512, lora_r=4, lora_alpha=16)).to(dev).bfloat16()
    tr = sft.SftTrainer(m2)
    tr.step(torch.randint(2, 512, (4, 64), device=dev), torch.randint(2, 512, (4, 64), device=dev))
    print("encoder ok")
torch.cuda.synchronize()
print("done")
