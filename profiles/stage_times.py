"""Stage timeline of the backward kernel from in-kernel clock64 stamps (profiling aid).
usage: python profiles/stage_times.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200 import _lib
from rwkv_lm_ext_b200.synthetic import make_inputs

# needs a library built with -DWKV6_FINE_STAMPS:  python -c "from rwkv_lm_ext_b200.build import build_library as b; b(defines=['WKV6_FINE_STAMPS'], out='rwkv_lm_ext_b200/libwkv6_b200_stamps.so')"
# and WKV6_B200_LIB pointing at it (the product library has no stamps)
FINE = True
_a = [a for a in sys.argv[1:] if not a.startswith("--")]
B, T, H = int(_a[0]) if _a else 8, 4096, 32
NS = 32 if FINE else 8
lib = M.load()
lib = _lib.load()
fn = lib.wkv6b200_debug_stamps
fn.argtypes = [ctypes.c_void_p]
r, k, v, w, u, gy = make_inputs(B, T, H, seed=0, decay="model", device="cuda")
NC = T // 64
buf = torch.zeros(B * H, NC, NS, dtype=torch.int64, device="cuda")
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


if FINE:
    order = [0, 8, 9, 10, 24, 11, 12, 13, 1, 2, 14, 25, 26, 15, 16, 17, 18, 19, 3, 4, 5, 6, 20, 21, 22, 23, 7]
    label = {0: "P start", 8: "P scan done", 9: "P scan barrier", 10: "P free barrier", 24: "P r,k loaded + parked", 11: "P math+versions done",
             12: "P parks+diag issued", 13: "P wait::st", 1: "P end", 2: "T1 start (Bm ready)", 14: "T1 Bm,G loaded", 25: "T1 dA converted", 26: "T1 dA fence.proxy", 15: "T1 dA stored",
             16: "T1 G done", 17: "T1 M1 barrier", 18: "T1 A^T loaded", 19: "T1 P^T stored", 3: "T1 end", 4: "T2 start (Dr ready)",
             5: "T2a end", 6: "T3 start (M3 ready)", 20: "T3 loop done", 21: "T3 gs/wait::st", 22: "T3 scan barrier", 23: "T3 parks loaded", 7: "T3 end"}
    print("fine timeline (mean cycles since the previous point):")
    for a, b in zip(order[:-1], order[1:]):
        dd = (s[:, :, b] - s[:, :, a])[:, 2:-2]
        print(f"  -> {label[b]:28s} {dd.mean():8.0f}")

    # per-CTA view: which SM, how many CTAs share it, when it started and how long its 64 chunks took (global timer)
    smid = s[:, 0, 31].long()
    t0, t1 = s[:, 0, 30], s[:, -1, 30]
    dur = (t1 - t0) / 1e3                                     # us from the first to the last chunk start
    start = (t0 - t0.min()) / 1e3
    share = torch.bincount(smid, minlength=148)[smid]
    for n in (1, 2):
        m = share == n
        if m.any():
            d_ = dur[m]
            print(f"CTAs on SMs with {n} CTA(s): {int(m.sum())}   first->last chunk start: mean {d_.mean():.1f} us  min {d_.min():.1f}  "
                  f"p50 {d_.median():.1f}  p90 {d_.quantile(0.9):.1f}  max {d_.max():.1f};  start offset max {start[m].max():.1f} us")
    q = dur[share == 2]
    if q.numel():
        slow = torch.argsort(dur, descending=True)[:8]
        print("slowest CTAs (blockIdx, sm, us):", [(int(i), int(smid[i]), round(float(dur[i]), 1)) for i in slow])
        per_sm = {}
        for i in range(len(dur)):
            per_sm.setdefault(int(smid[i]), []).append(round(float(dur[i]), 1))
        pairs = sorted(per_sm.items(), key=lambda kv: -max(kv[1]))[:6]
        print("slowest SMs:", pairs)
        print("fastest dual SMs:", sorted([kv for kv in per_sm.items() if len(kv[1]) == 2], key=lambda kv: max(kv[1]))[:4])
