"""BASELINE metric (ii), one GPU: bi-encoder passages/s at the 1B6 shape (L24, D2048, H32, FFN 7168, random
init bf16), 64 passages x 512 tokens per micro-batch, forward only.
Three variants of the same model and weights:
  fused  : rwkv_lm_ext_b200.bi_encoder_encode (fused elementwise kernels + tcgen05 WKV6)
  eager  : the reference's eager time-mix chain (src/model_encoder_run.py:150-190 restated) around the same
           tcgen05 WKV6 operator, mask / reverse index built on the host like the reference does
  eager+simt : as above with the exact SIMT WKV6 kernels (what a straight port of the reference's
           one-thread-per-channel kernel design gives)
usage: python profiles/bench_bi_encoder.py [layers]
       python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/bench_bi_encoder.py
       (each rank encodes its own micro-batches -- the batch-sharded data path of MyBatchSampler,
        data/custom_datasets.py:54 -- no collective on the data path; time = max over ranks)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import rwkv_lm_ext_b200 as M

M.load()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
dev = "cuda"
L = int(sys.argv[1]) if len(sys.argv) > 1 else 24
D, H, FFN, V = 2048, 32, 7168, 65536
B, T = 64, 512


class CMix(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.time_maa_k = torch.nn.Parameter(torch.rand(1, 1, D))
        self.time_maa_r = torch.nn.Parameter(torch.rand(1, 1, D))
        self.key = torch.nn.Linear(D, FFN, bias=False)
        self.receptance = torch.nn.Linear(D, D, bias=False)
        self.value = torch.nn.Linear(FFN, D, bias=False)

    def forward(self, x):
        xx = F.pad(x, (0, 0, 1, -1)) - x
        k = torch.relu(self.key(x + xx * self.time_maa_k)) ** 2
        return torch.sigmoid(self.receptance(x + xx * self.time_maa_r)) * self.value(k)


torch.manual_seed(0)
with torch.device(dev):
    model = torch.nn.Module()
    model.emb = torch.nn.Embedding(V, D)
    model.blocks = torch.nn.ModuleList()
    for i in range(L):
        b = torch.nn.Module()
        if i == 0:
            b.ln0 = torch.nn.LayerNorm(D)
        b.ln1, b.ln2 = torch.nn.LayerNorm(D), torch.nn.LayerNorm(D)
        b.att = M.Tmix_x060(D, H)
        with torch.no_grad():
            for n, p in b.att.named_parameters():
                if n in ("time_maa_w1", "time_maa_w2", "time_decay_w1", "time_decay_w2"):
                    p.uniform_(-1e-2, 1e-2)
                elif n == "time_decay":
                    p.copy_(-6 + 5 * torch.rand_like(p))
                elif n.startswith("time_maa"):
                    p.uniform_(0, 1)
                elif n == "time_faaaa":
                    p.normal_(0, 0.3)
        b.ffn = CMix()
        model.blocks.append(b)
    model.ln_out = torch.nn.LayerNorm(D)
model.emb_id, model.pad_id = 1, 0
model = model.bfloat16().eval()
g = torch.Generator().manual_seed(0)
idx = torch.randint(2, V, (B, T), generator=g)
lens = torch.randint(128, T, (B,), generator=g)
for b_, n in enumerate(lens.tolist()):
    idx[b_, n] = 1
    idx[b_, n + 1:] = 0
idx = idx.to(dev)
shift = torch.nn.ZeroPad2d((0, 0, 1, -1))


def eager_project(l, x):
    Bx, Tx, C = x.shape
    xx = shift(x) - x
    xxx = x + xx * l.time_maa_x
    xxx = torch.tanh(xxx @ l.time_maa_w1).view(Bx * Tx, 5, -1).transpose(0, 1)
    mw, mk, mv, mr, mg = torch.bmm(xxx, l.time_maa_w2).view(5, Bx, Tx, -1).unbind(0)
    xw = x + xx * (l.time_maa_w + mw)
    xk = x + xx * (l.time_maa_k + mk)
    xv = x + xx * (l.time_maa_v + mv)
    xr = x + xx * (l.time_maa_r + mr)
    xg = x + xx * (l.time_maa_g + mg)
    w = l.time_decay + torch.tanh(xw @ l.time_decay_w1) @ l.time_decay_w2
    return l.receptance(xr), l.key(xk), l.value(xv), F.silu(l.gate(xg)), w


def eager_encode(model, idx):
    # host-side mask / reverse index like src/model_encoder_run.py:7-26, then H2D
    ic = idx.cpu()
    mask = ((ic != 0) & (ic != 1)).int()
    rev = []
    for n in mask.sum(1):
        rev.append(torch.cat([torch.arange(0, n).flip(0), torch.arange(n, T)]))
    rev = torch.stack(rev).to(dev)
    gat = lambda t: torch.gather(t, 1, rev.unsqueeze(-1).expand(-1, -1, t.size(-1)))
    x = model.emb(idx)
    for i, blk in enumerate(model.blocks):
        if i == 0:
            x = blk.ln0(x)
        h = blk.ln1(x)
        l = blk.att
        r, k, v, gg, w = eager_project(l, h)
        rr, rk, rv, _, rw = eager_project(l, gat(h))
        y = M.RUN_CUDA_RWKV6(B, T, D, H, r, k, v, w, l.time_faaaa)
        ry = gat(M.RUN_CUDA_RWKV6(B, T, D, H, rr, rk, rv, rw, l.time_faaaa))
        y = (y + ry) / 2
        x = x + l.output(l.ln_x(y.view(B * T, D)).view(B, T, D) * gg)
        x = x + blk.ffn(blk.ln2(x))
    x = model.ln_out(x)
    pos = torch.eq(idx, 1).int().argmax(-1)
    return x[torch.arange(B, device=dev), pos]


def timeit(fn, n=5):
    with torch.no_grad():
        for _ in range(2):
            out = fn()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            out = fn()
        b_.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b_) / n, out


def max_over_ranks(ms):
    if world == 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if world > 1:
    dist.barrier()
t_f, e_f = timeit(lambda: M.bi_encoder_encode(model, idx))
t_f = max_over_ranks(t_f)
if world > 1:
    if rank == 0:
        print(json.dumps({"n_gpus": world, "layers": L, "micro_batch_per_gpu": [B, T], "fused_ms": round(t_f, 2),
                          "passages_per_s_fused_total": round(world * B / t_f * 1e3, 1)}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
t_e, e_e = timeit(lambda: eager_encode(model, idx))
M.set_impl("simt")
t_s, e_s = timeit(lambda: eager_encode(model, idx))
M.set_impl("auto")
cos = F.cosine_similarity(e_f.float(), e_s.float(), dim=-1).min().item()
flops = 2 * B * T * L * (2 * 5 * D * D + 2 * D * FFN + D * D) + 0.0     # Linears only (time-mix ones run twice)
print(json.dumps({"layers": L, "micro_batch": [B, T], "fused_ms": round(t_f, 2), "eager_ms": round(t_e, 2),
                  "eager_simt_wkv_ms": round(t_s, 2), "passages_per_s_fused": round(B / t_f * 1e3, 1),
                  "passages_per_s_eager": round(B / t_e * 1e3, 1), "passages_per_s_eager_simt": round(B / t_s * 1e3, 1),
                  "linear_tflops_fused": round(flops / t_f / 1e9, 1), "min_cosine_fused_vs_eager_simt": round(cos, 5)}))
