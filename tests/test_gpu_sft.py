"""BASELINE config 4 on the GPU: LoRA SFT of a small RWKV-6 LM through `rwkv_lm_ext_b200.sft` (fused time-mix /
channel-mix kernels + tensor-core WKV6): the loss goes down over bucketed batches, only `lora_` parameters move, a
CUDA-graph replayed step equals the eager step, and -- with two GPUs -- two ranks with half the batch each end up
with the parameters one rank gets from the whole batch.  `pytest -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(sft, seed=0, **kw):
    torch.manual_seed(seed)                        # lora_A keeps its kaiming init (global RNG), like the reference's
    m = sft.RwkvSft(layers=2, D=128, H=2, ffn=448, vocab=512, lora_r=4, lora_alpha=16, **kw)
    return sft.init_like_reference(m, seed).to(DEV).bfloat16()


def test_lora_sft_loss_goes_down_over_length_buckets():
    from rwkv_lm_ext_b200 import sft
    model = _model(sft)
    frozen = {n: p.detach().clone() for n, p in model.named_parameters() if "lora_" not in n}
    tr = sft.SftTrainer(model, lr=3e-3)
    lengths = [64, 128, 256]
    bss = sft.bucket_batch_sizes(lengths, tokens_per_batch=512)
    ds = sft.SyntheticSftBuckets(lengths, per_bucket=32, vocab=512, seed=1)
    losses = {t: [] for t in lengths}
    for epoch in range(6):                                       # the same 3 x 32 samples again and again: must be learnt
        for batch in sft.BucketBatchSampler(ds.cumulative_sizes, bss):
            idx, tgt = sft.pad_only_according_data([ds[i] for i in batch])
            loss = tr.step(idx.to(DEV), tgt.to(DEV))
            losses[idx.shape[1]].append(loss.item())
    for t, ls in losses.items():
        n = len(ls) // 6
        first, last = sum(ls[:n]) / n, sum(ls[-n:]) / n
        assert all(map(lambda v: v == v, ls)), "NaN loss"
        assert last < first - 0.05, (t, first, last)
    for n, p in model.named_parameters():
        if "lora_" not in n:
            assert torch.equal(p, frozen[n]), n                  # base weights are frozen


def test_cuda_graph_step_equals_eager_step():
    from rwkv_lm_ext_b200 import sft
    torch.manual_seed(0)
    idx = torch.randint(2, 512, (8, 64), device=DEV)
    tgt = torch.randint(2, 512, (8, 64), device=DEV)
    tgt[:, :20] = -100
    res = []
    for graphs in (False, True):
        model = _model(sft, seed=3)
        tr = sft.SftTrainer(model, lr=1e-3, graphs=graphs)
        n = 5
        if graphs:
            n -= 2                                                # the capture warms up with two real steps on the same batch
        for _ in range(n):
            loss = tr.step(idx, tgt)
        torch.cuda.synchronize()
        res.append((loss.item(), {k: v.detach().float().clone() for k, v in model.named_parameters() if "lora_" in k}))
    (l0, p0), (l1, p1) = res
    assert abs(l0 - l1) < 2e-2 * max(1.0, abs(l0)), (l0, l1)
    for k in p0:
        assert torch.allclose(p0[k], p1[k], atol=2e-2, rtol=2e-2), k


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rwkv_lm_ext_b200 import sft
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
torch.manual_seed(11)                  # same lora_A (kaiming, global RNG) in the 1-rank and the 2-rank run
m = sft.init_like_reference(sft.RwkvSft(layers=2, D=128, H=2, ffn=448, vocab=512, lora_r=4, lora_alpha=16), 5).cuda().bfloat16()
tr = sft.SftTrainer(m, lr=1e-3, n_buckets=3)
g = torch.Generator().manual_seed(9)
idx = torch.randint(2, 512, (8, 96), generator=g); tgt = torch.randint(2, 512, (8, 96), generator=g)
per = 8 // world
for _ in range(3):
    tr.step(idx[rank * per:(rank + 1) * per].cuda(), tgt[rank * per:(rank + 1) * per].cuda())
torch.cuda.synchronize()
if rank == 0:
    torch.save({k: v.detach().float().cpu() for k, v in m.named_parameters() if "lora_" in k}, sys.argv[2])
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_equal_one_rank_full_batch(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    outs = []
    for world in (1, 2):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        out = str(tmp_path / f"p{world}.pt")
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                               "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT, out], timeout=600)
        outs.append(torch.load(out))
    for k in outs[0]:                                             # mean of the two half-batch losses == the full-batch loss
        assert torch.allclose(outs[0][k], outs[1][k], atol=3e-2, rtol=3e-2), k
