"""Summarise an .ncu-rep: headline metrics + SASS hot spots.  usage: python profiles/ncu_summary.py rep [topN]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:100] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, r):
        if h in keys:
            print(f"  {h:85s} {v:>16s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
byop, sampop, items = collections.Counter(), collections.Counter(), []
ti = ts = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        ne, ns = int(r[ix["Instructions Executed"]]), int(r[ix["Warp Stall Sampling (All Samples)"]])
    except ValueError:
        continue
    s = r[ix["Source"]]
    toks = s.split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    byop[op] += ne; sampop[op] += ns; ti += ne; ts += ns
    items.append((ns, ne, s))
print(f"instructions executed (warp-level): {ti}   stall samples: {ts}")
for op, n in byop.most_common(topn):
    print(f"  {op:12s} {n:12d} {100*n/ti:5.1f}%   samples {100*sampop[op]/max(ts,1):5.1f}%")
items.sort(reverse=True)
print("top SASS lines by stall samples:")
for ns, ne, s in items[:topn]:
    print(f"  {ns:7d} {ne:10d}  {s[:100]}")
