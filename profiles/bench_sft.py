"""BASELINE config 4: RWKV-6 1B6 LoRA SFT (r = 8, alpha = 32 on att + ffn), length buckets 64..2048 with B = 2048 / T
cycled round-robin like MyBatchSampler, forward + backward + NCCL gradient all-reduce + AdamW, one rank per GPU.
usage: python profiles/bench_sft.py [--graphs] [--layers L]
       python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/bench_sft.py
Prints one JSON line (rank 0): tokens/s total and per GPU, ms per bucket, the all-reduce time measured alone."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rwkv_lm_ext_b200 import sft


def run(layers=24, graphs=False, steps_per_bucket=3, D=2048, H=32, ffn=7168, vocab=65536):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    own_pg = world > 1 and not dist.is_initialized()
    if own_pg:
        dist.init_process_group("nccl", device_id=dev)
    with torch.device("meta"):
        model = sft.RwkvSft(layers, D, H, ffn, vocab, lora_r=8, lora_alpha=32)
    model = model.to_empty(device=dev).bfloat16()
    sft.init_like_reference(model, seed=0)
    tr = sft.SftTrainer(model, graphs=graphs)
    lengths = [64, 128, 256, 512, 1024, 2048]
    bss = sft.bucket_batch_sizes(lengths)
    n_train = sum(p.numel() for p in tr.params)
    data = sft.SyntheticSftBuckets(lengths, per_bucket=max(bss) * world * (steps_per_bucket + 3), vocab=vocab, seed=rank)
    sampler = sft.BucketBatchSampler(data.cumulative_sizes, bss, rank, world)
    batches = []
    for b in sampler:                                           # round-robin over the buckets, this rank's slices
        idx, tgt = sft.pad_only_according_data([data[i] for i in b])
        batches.append((idx.pin_memory(), tgt.pin_memory()))
        if len(batches) >= len(lengths) * (steps_per_bucket + 2):
            break

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = 2 * len(lengths)
    for idx, tgt in batches[:warm]:
        tr.step(idx.to(dev, non_blocking=True), tgt.to(dev, non_blocking=True))
    barrier()
    per_bucket = {t: [] for t in lengths}
    evs = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for idx, tgt in batches[warm:]:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = tr.step(idx.to(dev, non_blocking=True), tgt.to(dev, non_blocking=True))    # ids in from pinned host memory
        b.record()
        evs.append((idx.shape[1], a, b))
    last = float(loss.item())                                    # the step's result comes back to the host
    t1.record()
    barrier()
    ms_total = t0.elapsed_time(t1)
    for t, a, b in evs:
        per_bucket[t].append(a.elapsed_time(b))
    n_steps = len(batches) - warm
    if world > 1:
        tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    # the gradient all-reduce alone (what is hidden under backward when it overlaps)
    ar_ms = 0.0
    if world > 1:
        buf = torch.zeros_like(tr.grads.flat)
        for _ in range(3):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            dist.all_reduce(buf)
        b.record()
        torch.cuda.synchronize()
        ar_ms = a.elapsed_time(b) / 10
    tokens = n_steps * 2048
    res = {"config": f"RWKV-6 1B6 shape L{layers} D{D} H{H} FFN{ffn} LoRA r=8 alpha=32 att+ffn, buckets {lengths} x {bss}, "
                     f"{'CUDA-graph replay' if graphs else 'eager'}",
           "n_gpus": world, "trainable_params": n_train, "grad_bytes": tr.grads.bytes, "grad_buckets": len(tr.grads.buckets),
           "steps": n_steps, "tokens_per_s_total": world * tokens / (ms_total * 1e-3),
           "tokens_per_s_per_gpu": tokens / (ms_total * 1e-3), "ms_per_step": ms_total / n_steps,
           "ms_per_step_by_bucket": {str(t): round(sum(v) / len(v), 3) for t, v in per_bucket.items() if v},
           "allreduce_alone_ms": round(ar_ms, 4), "last_loss": last}
    del tr, model
    torch.cuda.empty_cache()
    if own_pg:
        dist.barrier()
        dist.destroy_process_group()
    return res, rank


if __name__ == "__main__":
    L = int(sys.argv[sys.argv.index("--layers") + 1]) if "--layers" in sys.argv else 24
    res, rank = run(layers=L, graphs="--graphs" in sys.argv)
    if rank == 0:
        print(json.dumps(res))
