"""Host<->device copy bandwidth of the box (pinned memory, 134 MB tensors as in the e2e leg of bench.py):
what bounds the end-to-end number.
usage: python profiles/pcie_bandwidth.py                      one process, cuda:0
       python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/pcie_bandwidth.py
                                                              N processes copying CONCURRENTLY (one GPU each): the
                                                              aggregate is what N ranks of the e2e leg can get at best"""
import json, os, torch
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
n = 8 * 4096 * 2048
h_in = [torch.empty(n, dtype=torch.bfloat16).pin_memory() for _ in range(5)]
h_out = [torch.empty(n, dtype=torch.bfloat16).pin_memory() for _ in range(5)]
d_in = [torch.empty(n, dtype=torch.bfloat16, device="cuda") for _ in range(5)]
d_out = [torch.empty(n, dtype=torch.bfloat16, device="cuda") for _ in range(5)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
nbytes = 5 * n * 2

def run(do_in, do_out, reps=5):
    if world > 1:
        dist.barrier()                                   # all ranks copy at the same time
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for _ in range(reps):
        if do_in:
            with torch.cuda.stream(s1):
                for d, h in zip(d_in, h_in):
                    d.copy_(h, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                for h, d in zip(h_out, d_out):
                    h.copy_(d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    if world > 1:                                        # the slowest rank bounds a synchronous job
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms
run(True, True, 1)
res = {"h2d_only_GBs": round(nbytes / run(True, False) / 1e6, 1), "d2h_only_GBs": round(nbytes / run(False, True) / 1e6, 1)}
t = run(True, True)
res["both_ms_per_671MB_each_way"] = round(t, 2)
res["both_GBs_each_way"] = round(nbytes / t / 1e6, 1)
res["n_processes"] = world
if world > 1:
    res = {k: (round(v * world, 1) if k.endswith("GBs") or k.endswith("each_way") and "GBs" in k else v) for k, v in res.items()}
    res["note"] = "GB/s figures are the AGGREGATE over all processes (per-process rate of the slowest rank x N)"
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps(res))
