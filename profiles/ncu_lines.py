"""Per-CUDA-source-line instruction counts and stall samples of an .ncu-rep (needs -lineinfo).
usage: python profiles/ncu_lines.py rep [topN] [kernel-regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, items = None, None, []
tot_i = tot_s = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or r[0] in ("Function Name", "File Name") or len(r) < 10:
        continue
    if r[0] == "":
        continue      # SASS rows
    try:
        ne = int(r[hdr["Instructions Executed"]])
        ns = int(r[hdr["Warp Stall Sampling (All Samples)"]])
    except (ValueError, KeyError):
        continue
    items.append((ne, ns, cur_file, r[0], r[1].strip()))
    tot_i += ne
    tot_s += ns
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
print("--- by instructions executed")
for ne, ns, f, ln, src in sorted(items, reverse=True)[:topn]:
    print(f"{100*ne/tot_i:5.1f}% {ne:11d}  smp {100*ns/max(tot_s,1):5.1f}%  {f}:{ln}  {src[:90]}")
print("--- by stall samples")
for ne, ns, f, ln, src in sorted(items, key=lambda x: -x[1])[:topn // 2]:
    print(f"{100*ns/max(tot_s,1):5.1f}% {ns:8d}  ins {100*ne/tot_i:5.1f}%  {f}:{ln}  {src[:90]}")
