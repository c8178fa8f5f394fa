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


@pytest.mark.parametrize("shape", [(2, 37, 65536), (3, 5, 1000), (1, 1, 8)])
def test_fused_loss_equals_the_eager_chain(shape):
    """csrc/cross_entropy.cu against F.cross_entropy on the fp32 copy + L2Wrap (src/model.py:960-974, 1244-1283): loss
    to fp32 accuracy, gradient to bf16 rounding; ignored rows, an upstream gradient != 1, a tied maximum."""
    import torch.nn.functional as F
    from rwkv_lm_ext_b200 import sft
    B, T, V = shape
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + V)
    logits = (torch.randn(B, T, V, generator=g) * 2).to(torch.bfloat16)
    top = torch.randint(0, V, (B, T), generator=g)
    logits.scatter_(-1, top.unsqueeze(-1), 12.0)                 # a unique maximum per row (bf16 ties are likely otherwise)
    tgt = torch.randint(0, V, (B, T), generator=g)
    if T > 2:
        tgt[0, :2] = -100
    logits, tgt = logits.to(DEV), tgt.to(DEV)
    x1, x2 = logits.clone().requires_grad_(), logits.clone().requires_grad_()
    loss = sft.sft_loss(x1, tgt)
    ref = sft._L2Wrap.apply(F.cross_entropy(x2.view(-1, V).float(), tgt.view(-1)), x2)
    assert loss.dtype == torch.float32 and abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item()) + 1e-6
    (loss * 0.5).backward()
    (ref * 0.5).backward()
    # reference gradient in fp32 (the eager chain rounds the two terms to bf16 separately)
    p = torch.softmax(logits.float().view(-1, V), -1)
    valid = (tgt.view(-1) != -100)
    onehot = torch.zeros_like(p).scatter_(-1, tgt.view(-1).clamp(min=0).unsqueeze(-1), 1.0)
    want = (p - onehot) * (valid.float() * 0.5 / valid.sum()).unsqueeze(-1)
    want.scatter_add_(-1, top.to(DEV).view(-1, 1), torch.full((B * T, 1), 12.0 * 1e-4 / (B * T), device=DEV))
    got = x1.grad.float().view(-1, V)
    assert (got - want).abs().max().item() <= 2.0 ** -8 * want.abs().max().item()
    assert ((got - want).abs() <= 2.0 ** -8 * want.abs() + 1e-12).all()
    assert (got - x2.grad.float().view(-1, V)).abs().max().item() <= 2.0 ** -6 * want.abs().max().item()   # the eager bf16 chain
    assert (got[~valid] != 0).sum().item() == (~valid).sum().item()                                         # ignored rows: the L2Wrap term only
    # a tied maximum goes to the lowest index; every label ignored -> nan like torch
    tie = torch.zeros(1, 2, V, dtype=torch.bfloat16, device=DEV)
    tie[0, 0, 3:5] = 4.0
    tie[0, 1, V - 1] = 4.0
    tie.requires_grad_()
    sft.sft_loss(tie, torch.tensor([[1, 0]], device=DEV)).backward()
    l2 = 4.0 * 1e-4 / 2
    gt = tie.grad.float()
    pmax = torch.softmax(tie.detach().float(), -1)
    assert abs(gt[0, 0, 3].item() - (pmax[0, 0, 3].item() / 2 + l2)) <= 2.0 ** -8 * (pmax[0, 0, 3].item() / 2 + l2)
    assert abs(gt[0, 0, 4].item() - pmax[0, 0, 4].item() / 2) <= 2.0 ** -8 * pmax[0, 0, 4].item()
    none = sft.sft_loss(logits, torch.full_like(tgt, -100))
    assert none.isnan().item()


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
