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
buf = torch.zeros(B * H, NC, 8, dtype=torch.int64, device="cuda")
for it in range(3):
    leaves = [t.detach().requires_grad_(True) for t in (r, k, v, w, u)]
    y = M.RUN_CUDA_RWKV6(B, T, H * 64, H, *leaves)
    if it == 2:
        fn(buf.data_ptr())
    y.backward(gy)
    torch.cuda.synchronize()
fn(None)
s = buf.cpu().double()
names = ["P", "wait Bm", "T1 (a+b)", "wait Dr", "T2a", "T2b + wait M3", "T3", "to next P"]
d = torch.stack([s[:, :, 1] - s[:, :, 0], s[:, :, 2] - s[:, :, 1], s[:, :, 3] - s[:, :, 2], s[:, :, 4] - s[:, :, 3],
                 s[:, :, 5] - s[:, :, 4], s[:, :, 6] - s[:, :, 5], s[:, :, 7] - s[:, :, 6]], -1)[:, 2:-2]
nxt = (s[:, 1:, 0] - s[:, :-1, 7])[:, 2:-2]
print("mean cycles per stage over all CTAs / chunks (warp 0):")
for i, n in enumerate(names[:-1]):
    print(f"  {n:10s} {d[..., i].mean():8.0f}   (min {d[..., i].min():6.0f}  max {d[..., i].max():7.0f})")
print(f"  {names[-1]:10s} {nxt.mean():8.0f}   (min {nxt.min():6.0f}  max {nxt.max():7.0f})")
per_chunk = (s[:, 1:, 0] - s[:, :-1, 0])[:, 2:-2]
print(f"  chunk period {per_chunk.mean():8.0f}")

