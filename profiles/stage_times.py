"""Stage timeline of the backward kernel from in-kernel clock64 stamps (profiling aid).
usage: python profiles/stage_times.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200 import _lib
from rwkv_lm_ext_b200.synthetic import make_inputs

B, T, H = 8, 4096, 32
lib = M.load()
lib = _lib.load()
fn = lib.wkv6b200_debug_stamps
fn.argtypes = [ctypes.c_void_p]
r, k, v, w, u, gy = make_inputs(B, T, H, seed=0, decay="model", device="cuda")
NC = T // 64
buf = torch.zeros(2, B * H, NC, 8, dtype=torch.int64, device="cuda")
for it in range(3):
    leaves = [t.detach().requires_grad_(True) for t in (r, k, v, w, u)]
    y = M.RUN_CUDA_RWKV6(B, T, H * 64, H, *leaves)
    if it == 2:
        fn(buf.data_ptr())
    y.backward(gy)
    torch.cuda.synchronize()
fn(None)
s = buf[0].cpu().double()
si = buf[1].cpu().double()
names = ["P", "wait M1", "T1", "wait M2", "T2", "wait M3", "T3", "to next P"]
d = torch.stack([s[:, :, 1] - s[:, :, 0], s[:, :, 2] - s[:, :, 1], s[:, :, 3] - s[:, :, 2], s[:, :, 4] - s[:, :, 3],
                 s[:, :, 5] - s[:, :, 4], s[:, :, 6] - s[:, :, 5], s[:, :, 7] - s[:, :, 6]], -1)[:, 2:-2]
nxt = (s[:, 1:, 0] - s[:, :-1, 7])[:, 2:-2]
print("mean cycles per stage over all CTAs / chunks (warp 0):")
for i, n in enumerate(names[:-1]):
    print(f"  {n:10s} {d[..., i].mean():8.0f}   (min {d[..., i].min():6.0f}  max {d[..., i].max():7.0f})")
print(f"  {names[-1]:10s} {nxt.mean():8.0f}   (min {nxt.min():6.0f}  max {nxt.max():7.0f})")
per_chunk = (s[:, 1:, 0] - s[:, :-1, 0])[:, 2:-2]
print(f"  chunk period {per_chunk.mean():8.0f}")

# issuer warp: M1 = [PREP wake, operands landed, MMAs issued + commit, commit observed]; M3 = [T2 wake, issued, observed]
di = torch.stack([si[:, :, 1] - si[:, :, 0], si[:, :, 2] - si[:, :, 1], si[:, :, 3] - si[:, :, 2],
                  si[:, :, 5] - si[:, :, 4], si[:, :, 6] - si[:, :, 5]], -1)[:, 2:-2]
for i, n in enumerate(["M1: wait TMA (v,gy,S_in)", "M1: issue 16 MMAs", "M1: commit -> observed", "M3: stores + issue 10 MMAs", "M3: commit -> observed"]):
    print(f"  issuer {n:28s} {di[..., i].mean():8.0f}   (min {di[..., i].min():6.0f}  max {di[..., i].max():7.0f})")
# compute side: PREP arrive (stamp 1) -> issuer wake (istamp 0); issuer observed (istamp 3) -> compute resumes (stamp 2)
print(f"  hop compute->issuer (B_PREP) {(si[:, :, 0] - s[:, :, 1])[:, 2:-2].mean():8.0f}")
print(f"  hop issuer->compute (B_M1)   {(s[:, :, 2] - si[:, :, 3])[:, 2:-2].mean():8.0f}")
