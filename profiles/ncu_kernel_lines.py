"""Per-source-line instruction counts / stall samples of ONE kernel of an .ncu-rep.
usage: python profiles/ncu_kernel_lines.py rep kernel-regex [topN]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; items = []; tot = 0
src_cache = {}
def src_line(f, ln):
    import glob, os
    if f not in src_cache:
        c = glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rwkv_lm_ext_b200", "csrc", f))
        src_cache[f] = open(c[0]).read().split("\n") if c else []
    L = src_cache[f]
    return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < 10 or r[0] == "": continue
    try:
        ne = int(r[hdr["Instructions Executed"]]); ns = int(r[hdr["Warp Stall Sampling (All Samples)"]])
    except Exception:
        continue
    items.append((ne, ns, cur, int(r[0]))); tot += ne
print("total warp instructions", tot)
for ne, ns, f, ln in sorted(items, reverse=True)[:topn]:
    print(f"{100 * ne / tot:5.1f}% {ne:9d} smp {ns:5d}  {f}:{ln}  {src_line(f, ln)}")
