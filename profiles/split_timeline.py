"""Per-piece timeline of the tail-split backward from the global-timer stamps of the profiling build
(-DWKV6_FINE_STAMPS, WKV6_B200_LIB).  usage: python profiles/split_timeline.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200 import _lib
from rwkv_lm_ext_b200.synthetic import make_inputs

B, T, H = 8, 4096, 32
lib = _lib.load()
fn = lib.wkv6b200_debug_stamps
fn.argtypes = [ctypes.c_void_p]
r, k, v, w, u, gy = make_inputs(B, T, H, seed=0, decay="model", device="cuda")
NC = T // 64
buf = torch.zeros(B * H, NC, 32, dtype=torch.int64, device="cuda")
for it in range(3):
    leaves = [t.detach().requires_grad_(True) for t in (r, k, v, w, u)]
    y = M.RUN_CUDA_RWKV6(B, T, H * 64, H, *leaves)
    if it == 2:
        fn(buf.data_ptr())
    y.backward(gy)
    torch.cuda.synchronize()
fn(None)
s = buf.cpu()
gt, sm = s[:, :, 30].double() / 1e3, s[:, :, 31]
t0 = gt[gt > 0].min()
gt = gt - t0
n = int(os.environ.get("WKV6_B200_SPLIT", "9"))
if n == 0:
    print("no split: stamps are indexed by iteration; stream finish times (us):", "min %.1f p50 %.1f max %.1f" % (gt.max(1).values.min(), gt.max(1).values.median(), gt.max(1).values.max()))
    sys.exit()
sb = NC - n
long_start, long_end = gt[:, sb - 1], gt[:, 0]              # chunk sb-1 is processed first, chunk 0 last (start-of-chunk stamps)
short_start, short_end = gt[:, NC - 1], gt[:, sb]
print(f"long pieces ({sb} chunks): start min {long_start.min():.1f} max {long_start.max():.1f};  last chunk starts at min {long_end.min():.1f} p50 {long_end.median():.1f} max {long_end.max():.1f} us")
d = long_end - long_start
print(f"   duration: min {d.min():.1f} p10 {d.quantile(0.1):.1f} p50 {d.median():.1f} p90 {d.quantile(0.9):.1f} max {d.max():.1f}")
print(f"short pieces ({n} chunks): start min {short_start.min():.1f} p50 {short_start.median():.1f} max {short_start.max():.1f};  duration p50 {(short_end - short_start).median():.1f} max {(short_end - short_start).max():.1f}")
# which SMs ran what
smL = sm[:, 0]
cnt = torch.bincount(smL, minlength=148)
print("long pieces per SM: ", {int(c): int((cnt == c).sum()) for c in cnt.unique()})
order = torch.argsort(short_start)
print("short piece start times (every 16th):", [round(float(short_start[i]), 1) for i in order[::16]])
slow = torch.argsort(long_end, descending=True)[:6]
print("slowest long pieces (stream, sm, start, end):", [(int(i), int(smL[i]), round(float(long_start[i]), 1), round(float(long_end[i]), 1)) for i in slow])
smS = sm[:, NC - 1]
print("short pieces by start time: (stream, sm, start, end)")
for i in order[::8]:
    print("  ", int(i), int(smS[i]), round(float(short_start[i]), 1), round(float(short_end[i]), 1))
# per SM: list of pieces that ran there with start times
per = {}
for i in range(B * H):
    per.setdefault(int(smL[i]), []).append(("L", i, round(float(long_start[i]), 1), round(float(long_end[i]), 1)))
    per.setdefault(int(smS[i]), []).append(("S", i, round(float(short_start[i]), 1), round(float(short_end[i]), 1)))
for smid in (0, 1, 74, 147):
    print("SM", smid, sorted(per.get(smid, []), key=lambda x: x[2]))

entry, alloc = s[:, NC - 1, 28].double() / 1e3 - t0, s[:, NC - 1, 29].double() / 1e3 - t0
print("short pieces: (stream, sm, CTA entry, TMEM allocated, first chunk, end)")
for i in order[::8]:
    print("  ", int(i), int(smS[i]), round(float(entry[i]), 1), round(float(alloc[i]), 1), round(float(short_start[i]), 1), round(float(short_end[i]), 1))
