"""Achieved HBM bandwidth of the memory-bound kernels at the 1B6 shape (B=8, T=4096, C=2048), against the
measured copy peak (MEASURED_PEAKS.json).  CUDA events on the launching stream, inputs larger than L2.
usage: python profiles/bench_elementwise.py  ->  one JSON line per kernel"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200 import heads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
M.load()
dev = "cuda"
B, T, C, H = 8, 4096, 2048, 32
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, T, C, device=dev, generator=g).bfloat16()
y = torch.randn(B, T, C, device=dev, generator=g).bfloat16()
gate = torch.randn(B, T, C, device=dev, generator=g).bfloat16()
m = torch.randn(5, B, T, C, device=dev, generator=g).bfloat16()
maa = torch.randn(5, C, device=dev, generator=g).bfloat16()
maa_x = torch.randn(C, device=dev, generator=g).bfloat16()
ln_w = torch.randn(C, device=dev, generator=g).bfloat16()
ln_b = torch.randn(C, device=dev, generator=g).bfloat16()
idx = torch.randint(2, 65536, (B * 64, 512), device=dev, generator=g)
idx[:, -1] = 1
xe = torch.randn(B * 64, 512, 1024, device=dev, generator=g).bfloat16()       # 512 passages x 512 tokens x 1024: 537 MB
lens = torch.full((B * 64,), 511, device=dev, dtype=torch.int64)
mask, rev = heads.create_mask_and_rev_idx(idx, 1, 0)
E = B * T * C


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


# backward kernels: called through the C ABI with preallocated outputs
from rwkv_lm_ext_b200 import _lib
from rwkv_lm_ext_b200._lib import ptr
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
go5 = torch.randn(5, B, T, C, device=dev, generator=g).bfloat16()
gx, gm = torch.empty_like(x), torch.empty_like(m)
gy, gg = torch.empty_like(y), torch.empty_like(y)
gmaa = torch.empty(5, C, device=dev)
glw, glb = torch.empty(C, device=dev), torch.empty(C, device=dev)
ws = torch.empty(lib.elementwise_backward_workspace_bytes(B, T, C, 5), dtype=torch.uint8, device=dev)
gxe = torch.empty_like(xe)
goe = torch.randn(B * 64, 1024, device=dev, generator=g)


def ddlerp_bwd():
    assert lib.tmix_ddlerp_mix_backward_bf16(B, T, C, ptr(x), None, ptr(maa), ptr(m), *[ptr(go5[i]) for i in range(5)], ptr(gx),
                                             ptr(gm), ptr(gmaa), None, ptr(ws), ws.numel(), st) == 0


def shift_bwd():
    assert lib.tmix_shift_lerp_backward_bf16(B, T, C, ptr(x), None, ptr(maa_x), ptr(go5), ptr(gx), ptr(gmaa), None,
                                             ptr(ws), ws.numel(), st) == 0


def gn_bwd():
    assert lib.groupnorm_gate_backward_bf16(B * T, C, H, 64e-5, 1, ptr(y), ptr(gate), ptr(ln_w), ptr(ln_b), ptr(go5), ptr(gy),
                                            ptr(gg), ptr(glw), ptr(glb), ptr(ws), ws.numel(), st) == 0


def pool_bwd():
    assert lib.pooling_backward_bf16(0, 1, B * 64, 512, 1024, ptr(lens), ptr(goe), ptr(gxe), st) == 0


hl = torch.tanh(torch.randn(B * T, 160, device=dev, generator=g)).bfloat16()
w2 = (torch.randn(5, 32, C, device=dev, generator=g) * 0.1).bfloat16()


def lora_unfused():
    mm = torch.bmm(hl.view(B * T, 5, 32).transpose(0, 1), w2).view(5, B, T, C)
    return heads.tmix_ddlerp_mix(x, maa, mm)


cases = [
    ("tmix_ddlerp_lora (LoRA product fused, tcgen05)", lambda: heads.tmix_ddlerp_lora(x, maa, hl, w2), E * 12 + B * T * 160 * 2),
    ("bmm + tmix_ddlerp_mix (unfused equivalent)", lora_unfused, E * 12 + B * T * 160 * 2),
    ("tmix_ddlerp_mix backward", ddlerp_bwd, E * 34),          # x 2 + gout 10 + m 10 -> gx 2 + gm 10 B/elem
    ("tmix_shift_lerp backward", shift_bwd, E * 6),            # x, gout -> gx
    ("groupnorm_gate backward", gn_bwd, E * 10),               # y, g, gout -> gy, gg
    ("pooling weightedmean backward", pool_bwd, xe.numel() * 2),
    ("tmix_ddlerp_mix", lambda: heads.tmix_ddlerp_mix(x, maa, m), E * 22),          # x 2 + m 10 + out 10 B/elem
    ("tmix_shift_lerp", lambda: heads.tmix_shift_lerp(x, maa_x), E * 4),
    ("groupnorm_gate", lambda: heads.groupnorm_gate(y, gate, ln_w, ln_b, H, 64e-5), E * 6),
    ("pooling weightedmean", lambda: heads.pooling(xe, lens, "weightedmean", "infer"), xe.numel() * 2),
    ("reverse_x (gather_tokens)", lambda: heads.reverse_x(xe, rev), xe.numel() * 4),
    ("create_mask_rev_idx", lambda: heads.create_mask_and_rev_idx(idx, 1, 0), idx.numel() * 20),   # idx 8 + mask 4 + rev 8
    ("eos_index", lambda: heads.eos_index(idx, 1), idx.numel() * 8),
]
for name, fn, nbytes in cases:
    ms = timeit(fn)
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": nbytes, "achieved_gbs": round(gbs, 1),
                      "peak_gbs": PEAK, "frac": round(gbs / PEAK, 3)}), flush=True)
