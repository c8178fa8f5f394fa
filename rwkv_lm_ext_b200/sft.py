"""Data-parallel LoRA / PiSSA / state-tuning SFT around the fused kernels (BASELINE.json configs[3]).

The reference trains through PyTorch Lightning + DeepSpeed (peft_train/peft_train_sft.py:404-419); what that
stack does ON the hot path is small and is restated here without it:

  * `LoraLinear`                -- src/rwkvLinear.py:41-97 (same parameter names `weight`, `lora_A`, `lora_B`, same
                                   forward incl. the PiSSA flavour), so adapters written by either side load in the other
  * `RwkvSft`                   -- the RWKV-6 language model (src/model.py:904-933, 1201-1242) with every time-mix /
                                   channel-mix layer running `tmix_x060_forward` / `cmix_x060_forward`
  * `sft_loss`                  -- src/model.py:1244-1283: token cross entropy (labels -100 ignored) + L2Wrap
  * `BucketBatchSampler`        -- data/custom_datasets.py:19-78 (`MyBatchSampler`): one batch size per length bucket
                                   (README.md:80: B = 2048 / T), buckets visited round-robin, rank r takes the r-th slice
  * `GradBuckets`               -- the NCCL gradient all-reduce: trainable gradients live in ONE flat buffer, cut into a
                                   few buckets; a bucket's all-reduce is launched (async) by the hook of its last
                                   gradient, so it overlaps the rest of the backward pass
  * `save_trainable` / `load_trainable` / `pissa_init_all` -- peft_train/Callbacks.py:7-28, peft_train_sft.py:182-208:
                                   `{name: tensor}` .pth of the trainable parameters, `init_pissa.pth`
  * `SftTrainer`                -- fwd + bwd + all-reduce + AdamW step; small buckets can replay a CUDA graph

Nothing is sharded inside the recurrence: ranks own disjoint batch slices.
"""
import math
import os

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import cmix, tmix


# ------------------------------------------------------------------------------------------------
# adapters
# ------------------------------------------------------------------------------------------------
class LoraLinear(nn.Module):
    """src/rwkvLinear.py:41-97 without the bitsandbytes quantisation branch."""

    def __init__(self, in_features, out_features, r, alpha, dropout=0.0):
        super().__init__()
        self.weight = nn.Parameter(torch.empty((out_features, in_features)))
        self.lora_A = nn.Parameter(torch.empty(r, in_features))
        self.lora_B = nn.Parameter(torch.empty(out_features, r))
        self.lora_dropout = nn.Dropout(dropout)
        self.scaling = alpha / r
        self.r = r
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B)
        self.pissa = False

    def pissa_load(self, init_A, init_B):          # src/rwkvLinear.py:61-63
        self.pissa = True
        self.weight.data = self.weight.data - (init_B.to(self.weight) @ init_A.to(self.weight))

    def pissa_init(self, svd_niter):               # src/rwkvLinear.py:66-75
        self.pissa = True
        w = self.weight.data.float()
        Ur, Sr, Vr = torch.svd_lowrank(w, self.r, niter=svd_niter)
        lora_A = torch.diag(torch.sqrt(Sr)) @ Vr.t()
        lora_B = Ur @ torch.diag(torch.sqrt(Sr))
        self.lora_A.data = lora_A.to(self.weight.dtype)
        self.lora_B.data = lora_B.to(self.weight.dtype)
        self.weight.data = (w - lora_B @ lora_A).to(self.weight.dtype)

    def forward(self, x):
        if self.training and self.lora_dropout.p > 0:          # dropout on the adapter input: the eager chain
            if self.pissa:
                return F.linear(x, self.weight) + F.linear(F.linear(x, self.lora_A), self.lora_B)
            return F.linear(x, self.weight) + self.scaling * F.linear(F.linear(self.lora_dropout(x), self.lora_A), self.lora_B)
        return _LoraFn.apply(x, self.weight, self.lora_A, self.lora_B, 1.0 if self.pissa else self.scaling)


# id(parameter) -> callable(parameter), registered by GradBuckets: a backward pass that finds its parameter here adds
# the gradient straight into the flat buffer (GEMM with beta = 1) and reports it, instead of returning a tensor that
# autograd would add with one more elementwise kernel per parameter
_GRAD_SINKS = {}


def _param_grad(p, m1, m2, alpha):
    """alpha * m1 @ m2 as the gradient of parameter p: accumulated in place when p's gradient lives in a GradBuckets
    buffer (returns None: autograd has nothing left to do), a new tensor otherwise."""
    sink = _GRAD_SINKS.get(id(p))
    if sink is not None and p.grad is not None:
        p.grad.addmm_(m1, m2, alpha=alpha)
        sink(p)
        return None
    return torch.addmm(p, m1, m2, beta=0, alpha=alpha)


class _LoraFn(torch.autograd.Function):
    """x W^T + s (x A^T) B^T (src/rwkvLinear.py:94-96) with the scaling and both additions folded into the GEMMs:
    forward 3 launches (eager: 3 GEMMs + mul + add), backward 5 (eager: 5 GEMMs + mul + add + 2 accumulations)."""

    @staticmethod
    def forward(ctx, x, weight, A, B, scaling):
        x2 = x.reshape(-1, x.shape[-1])
        xa = x2 @ A.t()                                          # [N, r]
        out = x2 @ weight.t()
        out.addmm_(xa, B.t(), alpha=scaling)
        ctx.save_for_backward(x2, weight, A, B, xa)
        ctx.scaling, ctx.x_shape = scaling, x.shape
        return out.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, weight, A, B, xa = ctx.saved_tensors
        s = ctx.scaling
        need_x, need_w, need_a, need_b = ctx.needs_input_grad[:4]
        gy2 = gy.reshape(-1, gy.shape[-1])
        gx = gw = ga = gb = None
        gxa = gy2 @ B if (need_x or need_a) else None            # [N, r], unscaled
        if need_b:
            gb = _param_grad(B, gy2.t(), xa, s)
        if need_a:
            ga = _param_grad(A, gxa.t(), x2, s)
        if need_x:
            gx = gy2 @ weight
            gx.addmm_(gxa, A, alpha=s)
            gx = gx.view(ctx.x_shape)
        if need_w:
            gw = gy2.t() @ x2
        return gx, gw, ga, gb, None


def _linear(i, o, lora):
    return LoraLinear(i, o, *lora) if lora else nn.Linear(i, o, bias=False)


class _Tmix(tmix.Tmix_x060):
    def __init__(self, C, H, lora, state_tuning):
        super().__init__(C, H, state_tuning=state_tuning)
        if lora:                                    # make_linear_att: receptance, key, value, output, gate (src/model.py:427-431)
            for n in ("receptance", "key", "value", "output", "gate"):
                setattr(self, n, _linear(C, C, lora))


class _Cmix(nn.Module):
    """Parameter names of RWKV_CMix_x060 (src/model.py:616-644); forward = cmix_x060_forward."""

    def __init__(self, C, ffn, lora):
        super().__init__()
        self.time_maa_k = nn.Parameter(torch.zeros(1, 1, C))
        self.time_maa_r = nn.Parameter(torch.zeros(1, 1, C))
        self.key = _linear(C, ffn, lora)            # make_linear_ffn (src/model.py:602-604)
        self.receptance = _linear(C, C, lora)
        self.value = _linear(ffn, C, lora)

    def forward(self, x):
        return cmix.cmix_x060_forward(self, x)


class _Block(nn.Module):
    def __init__(self, i, C, H, ffn, lora_att, lora_ffn, state_tuning):
        super().__init__()
        if i == 0:
            self.ln0 = nn.LayerNorm(C)
        self.ln1, self.ln2 = nn.LayerNorm(C), nn.LayerNorm(C)
        self.att = _Tmix(C, H, lora_att, state_tuning)
        self.ffn = _Cmix(C, ffn, lora_ffn)

    def forward(self, x):                            # src/model.py:904-933
        if hasattr(self, "ln0"):
            x = self.ln0(x)
        x = x + self.att(self.ln1(x))
        return x + self.ffn(self.ln2(x))


class RwkvSft(nn.Module):
    """RWKV-6 LM with the reference's module / parameter names (emb, blocks.N.{ln0,ln1,ln2,att,ffn}, ln_out, head)."""

    def __init__(self, layers=24, D=2048, H=32, ffn=7168, vocab=65536, lora_r=8, lora_alpha=32, lora_dropout=0.0,
                 parts=("att", "ffn"), train_type="lora"):
        super().__init__()
        lora = (lora_r, lora_alpha, lora_dropout) if train_type in ("lora", "pissa") and lora_r > 0 else None
        self.emb = nn.Embedding(vocab, D)
        self.blocks = nn.ModuleList([_Block(i, D, H, ffn, lora if "att" in parts else None, lora if "ffn" in parts else None,
                                            train_type == "state") for i in range(layers)])
        self.ln_out = nn.LayerNorm(D)
        self.head = nn.Linear(D, vocab, bias=False)
        self.train_type = train_type

    def forward(self, idx):
        from .encoders import blocks_forward      # residual adds fused with the LayerNorms that follow them
        h = blocks_forward(self.blocks, self.emb(idx), self.ln_out, lambda blk, x: blk.att(x))
        return self.head(h)

    def mark_trainable(self):
        """peft_train_sft.py:330-387: LoRA / PiSSA train the `lora_` parameters, state tuning the `state` ones."""
        key = "state" if self.train_type == "state" else "lora_"
        for n, p in self.named_parameters():
            p.requires_grad = key in n
        return [p for p in self.parameters() if p.requires_grad]


def init_like_reference(model, seed=0):
    """Parameter ranges of the reference's initialisation (src/model.py:375-432, 1291-1340): there are no checkpoints
    offline, so benchmarks and tests train from this."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = len(model.blocks)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "lora_B" in name or "time_state" in name:
                p.zero_()
            elif name.endswith("time_decay"):
                C = p.numel()
                li = int(name.split(".")[1])
                n = torch.arange(C, dtype=torch.float32)
                p.copy_((-6 + 5 * (n / max(C - 1, 1)) ** (0.7 + 1.3 * li / max(L - 1, 1))).view_as(p))
            elif "time_maa_w" in name and name[-1] in "12" or "time_decay_w" in name:
                p.copy_((torch.rand(p.shape, generator=g) * 2e-2 - 1e-2))
            elif "time_maa" in name:
                p.copy_(torch.rand(p.shape, generator=g))
            elif "time_faaaa" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            elif p.dim() == 2 and "lora_A" not in name:
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 / math.sqrt(p.shape[1])))
    return model


class _L2Wrap(torch.autograd.Function):             # src/model.py:960-974
    @staticmethod
    def forward(ctx, loss, y):
        ctx.save_for_backward(y)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        (y,) = ctx.saved_tensors
        factor = 1e-4 / (y.shape[0] * y.shape[1])
        maxx, ids = torch.max(y, -1, keepdim=True)
        gy = torch.zeros_like(y)
        gy.scatter_(-1, ids, maxx * factor)
        return grad_output, gy


class _FusedLoss(torch.autograd.Function):
    """Cross entropy + L2Wrap on bf16 logits in two kernels (csrc/cross_entropy.cu): the forward reads the [B*T, V]
    matrix once, the backward reads it once and writes the gradient once (the eager chain moves it nine times)."""

    @staticmethod
    def forward(ctx, logits, targets):
        from . import _lib
        from ._lib import check, ptr, stream_of
        V = logits.shape[-1]
        rows = logits.numel() // V
        dev = logits.device
        stats = torch.empty(3, rows, dtype=torch.float32, device=dev)
        amax = torch.empty(rows, dtype=torch.int32, device=dev)
        mean = torch.empty(2, dtype=torch.float32, device=dev)
        check(_lib.load().cross_entropy_l2wrap_bf16(rows, V, ptr(logits), ptr(targets), -100, ptr(stats), ptr(amax), ptr(mean),
                                                    stream_of(logits)), "cross_entropy_l2wrap_bf16")
        ctx.save_for_backward(logits, targets, stats, amax, mean)
        return mean[0]

    @staticmethod
    def backward(ctx, gloss):
        from . import _lib
        from ._lib import check, ptr, stream_of
        logits, targets, stats, amax, mean = ctx.saved_tensors
        V = logits.shape[-1]
        rows = logits.numel() // V
        g = torch.empty_like(logits)
        gloss = gloss.to(torch.float32).contiguous()
        check(_lib.load().cross_entropy_l2wrap_backward_bf16(rows, V, ptr(logits), ptr(targets), -100, ptr(stats), ptr(amax), ptr(mean),
                                                             ptr(gloss), 1e-4 / rows, ptr(g), stream_of(logits)),
              "cross_entropy_l2wrap_backward_bf16")
        return g, None


def sft_loss(logits, targets):
    """src/model.py:1244-1283 (my_qa_mask == 0 branch): mean cross entropy over the tokens whose label is not -100,
    wrapped in L2Wrap (src/model.py:960-974).  bf16 logits on the GPU take the fused kernels (fp32 arithmetic inside)."""
    V = logits.size(-1)
    if (logits.is_cuda and logits.dtype == torch.bfloat16 and logits.dim() == 3 and V % 8 == 0 and V <= 65536
            and targets.device == logits.device and targets.dtype == torch.long):
        return _FusedLoss.apply(logits.contiguous(), targets.reshape(-1).contiguous())
    loss = F.cross_entropy(logits.view(-1, V).float(), targets.reshape(-1))
    return _L2Wrap.apply(loss, logits)


# ------------------------------------------------------------------------------------------------
# bucketed data path
# ------------------------------------------------------------------------------------------------
def bucket_batch_sizes(train_lengths, tokens_per_batch=2048):
    """README.md:80: --train_lengths 64 ... 2048 --train_batch_sizes 32 ... 1, i.e. B * T = const."""
    return [max(1, tokens_per_batch // t) for t in train_lengths]


class BucketBatchSampler:
    """`MyBatchSampler` (data/custom_datasets.py:19-78).  The dataset is the concatenation of one fixed-length dataset
    per bucket (`cumulative_sizes` = running totals); a step takes `batch_sizes[b] * world_size` consecutive samples of
    bucket b and rank r keeps the r-th slice of `batch_sizes[b]`; buckets are visited round-robin, skipping exhausted
    ones; incomplete global batches are dropped.  `skipped_batches` resumes mid-epoch."""

    def __init__(self, cumulative_sizes, batch_sizes, rank=0, world_size=1, skipped_batches=0):
        assert len(cumulative_sizes) == len(batch_sizes)
        self.cumulative_sizes, self.batch_sizes = list(cumulative_sizes), list(batch_sizes)
        self.rank, self.world_size, self.skipped_batches = rank, world_size, skipped_batches

    def _batches_per_bucket(self):
        out, prev = [], 0
        for c, b in zip(self.cumulative_sizes, self.batch_sizes):
            out.append((c - prev) // (b * self.world_size))
            prev = c
        return out

    def __len__(self):
        return sum(self._batches_per_bucket()) - self.skipped_batches

    def __iter__(self):
        rest = self._batches_per_bucket()
        nb, cur, skipped = len(rest), 0, 0
        while sum(rest) > 0:
            while rest[cur] == 0:
                cur = (cur + 1) % nb
            if skipped < self.skipped_batches:
                skipped += 1
                rest[cur] -= 1
                continue
            b = self.batch_sizes[cur]
            first = self.cumulative_sizes[cur] - rest[cur] * b * self.world_size + self.rank * b
            yield list(range(first, first + b))
            rest[cur] -= 1
            cur = (cur + 1) % nb

    def bucket_of(self, index):
        for b, c in enumerate(self.cumulative_sizes):
            if index < c:
                return b
        raise IndexError(index)


def pad_only_according_data(features, pad_token_id=0):
    """data/custom_datasets.py:83-89: right-pad input_ids with pad and labels with -100 to the bucket's length."""
    max_len = features[0]["fixed_len"]
    ids = [f["input_ids"] + [pad_token_id] * (max_len - len(f["input_ids"])) for f in features]
    lab = [f["labels"] + [-100] * (max_len - len(f["labels"])) for f in features]
    return torch.tensor(ids, dtype=torch.long), torch.tensor(lab, dtype=torch.long)


class SyntheticSftBuckets(torch.utils.data.Dataset):
    """Synthetic stand-in for the tokenised SFT datasets (data/SftUtilities.py:58-89): `per_bucket` samples for every
    length in `train_lengths`, each a prompt of random length (labels -100, shifted as tokenize_fn_no_chunk does) followed
    by an answer and the eos token."""

    def __init__(self, train_lengths, per_bucket, vocab=65536, seed=0):
        self.lengths, self.per_bucket, self.vocab, self.seed = list(train_lengths), per_bucket, vocab, seed
        self.cumulative_sizes = [per_bucket * (i + 1) for i in range(len(self.lengths))]

    def __len__(self):
        return self.cumulative_sizes[-1]

    def __getitem__(self, i):
        b = i // self.per_bucket
        T = self.lengths[b]
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        n = int(torch.randint(T // 2 + 1, T + 1, (1,), generator=g))           # falls into this bucket (bisect_left)
        n_in = max(1, n // 2)
        ids = torch.randint(2, self.vocab, (n,), generator=g).tolist()
        labels = [-100] * (n_in - 1) + ids[n_in:] + [1]
        return {"input_ids": ids, "labels": labels, "fixed_len": T}


# ------------------------------------------------------------------------------------------------
# adapter checkpoints
# ------------------------------------------------------------------------------------------------
def save_trainable(model, out_dir, model_filename):
    """peft_train/Callbacks.py:7-28: `{name: tensor}` of the parameters with requires_grad, as
    `<out_dir>/<basename(model_filename)>.pth`."""
    os.makedirs(out_dir, exist_ok=True)
    sd = {n: p.data for n, p in model.named_parameters() if p.requires_grad}
    if not sd:
        return None
    path = os.path.join(out_dir, os.path.basename(model_filename) + ".pth")
    torch.save(sd, path)
    return path


def load_trainable(model, path):
    """peft_train_sft.py:183-186: `model.load_state_dict(torch.load(path), strict=False)`."""
    return model.load_state_dict(torch.load(path, map_location="cpu"), strict=False)


def pissa_init_all(model, svd_niter=4, init_file=None):
    """peft_train_sft.py:188-208: initialise every LoRA pair from the top singular directions of its weight and keep
    `{module}.init_lora_A / init_lora_B` (needed to merge the adapter later) -- or, when `init_file` exists, subtract the
    saved initial product from the weights instead."""
    if init_file is not None and os.path.exists(init_file):
        init = torch.load(init_file, map_location="cpu")
        for name, m in model.named_modules():
            if callable(getattr(m, "pissa_load", None)):
                m.pissa_load(init[f"{name}.init_lora_A"], init[f"{name}.init_lora_B"])
        return init
    init = {}
    for name, m in model.named_modules():
        if callable(getattr(m, "pissa_init", None)):
            m.pissa_init(svd_niter)
            init[f"{name}.init_lora_A"] = m.lora_A.data.clone()
            init[f"{name}.init_lora_B"] = m.lora_B.data.clone()
    if init_file is not None:
        torch.save(init, init_file)
    return init


# ------------------------------------------------------------------------------------------------
# data-parallel gradient all-reduce
# ------------------------------------------------------------------------------------------------
class GradBuckets:
    """All trainable gradients in one flat buffer (each `p.grad` is a view of it), cut into `n_buckets` contiguous
    buckets in the order the backward pass finishes them (last parameters first).  The post-accumulate hook of a
    bucket's last gradient launches that bucket's all-reduce asynchronously, so the transfers run under the rest
    of backward; `finish()` waits and turns the sums into means.  With world_size 1 nothing is launched."""

    def __init__(self, params, n_buckets=4, group=None):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        dt, dev = self.params[0].dtype, self.params[0].device
        assert all(p.dtype == dt and p.device == dev for p in self.params)
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        # reverse registration order ~ the order in which gradients become final
        order = list(reversed(self.params))
        per = math.ceil(total / max(1, n_buckets))
        self.buckets, off, cur_lo, cur_n, cur_cnt = [], 0, 0, 0, 0
        self._bucket_of = {}
        for p in order:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._bucket_of[id(p)] = len(self.buckets)
            off += n
            cur_n += n
            cur_cnt += 1
            if cur_n >= per:
                self.buckets.append((cur_lo, off, cur_cnt))
                cur_lo, cur_n, cur_cnt = off, 0, 0
        if cur_cnt:
            self.buckets.append((cur_lo, off, cur_cnt))
        self._pending = [c for _, _, c in self.buckets]
        self._handles = []
        self._seen = set()
        self.bytes = total * self.flat.element_size()
        self._hooks = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        for p in self.params:                                  # backward passes that write into the buffer themselves
            _GRAD_SINKS[id(p)] = self._hook

    def _hook(self, p):
        if id(p) in self._seen:                        # reported by the backward pass itself (_param_grad) and by autograd
            return
        self._seen.add(id(p))
        b = self._bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0 and self.world > 1:
            lo, hi, _ = self.buckets[b]
            self._handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Call after backward: waits for the all-reduces and averages.  Returns the flat mean gradient."""
        if self.world > 1:
            for b, left in enumerate(self._pending):           # parameters that received no gradient this step
                if left > 0:
                    lo, hi, _ = self.buckets[b]
                    self._handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for h in self._handles:
                h.wait()
            self.flat.div_(self.world)
        self._handles = []
        self._pending = [c for _, _, c in self.buckets]
        self._seen = set()
        return self.flat

    def zero(self):
        self.flat.zero_()

    def remove(self):
        for h in self._hooks:
            h.remove()
        for p in self.params:
            if _GRAD_SINKS.get(id(p)) == self._hook:
                del _GRAD_SINKS[id(p)]


class SftTrainer:
    """One optimisation step = forward, loss, backward (bucketed all-reduce under it), AdamW.
    `graphs=True`: the whole step of a bucket shape is captured into a CUDA graph at its first use and replayed
    afterwards (the small buckets, B*T = 2048 tokens in 32 rows of 64, are launch-bound otherwise)."""

    def __init__(self, model, lr=3e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.001, n_buckets=4, graphs=False):
        self.model = model
        self.params = model.mark_trainable() if hasattr(model, "mark_trainable") else [p for p in model.parameters() if p.requires_grad]
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            for p in self.params:                      # every rank starts from rank 0's adapter (what DDP does at wrap time)
                dist.broadcast(p.data, src=0)
        self.grads = GradBuckets(self.params, n_buckets)
        dev = self.params[0].device
        self.opt = torch.optim.AdamW(self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                     capturable=graphs and dev.type == "cuda", foreach=True)
        self.graphs = graphs
        self._graph = {}

    def _step(self, idx, targets):
        loss = sft_loss(self.model(idx), targets)
        loss.backward()
        self.grads.finish()
        self.opt.step()
        self.grads.zero()
        return loss.detach()

    def step(self, idx, targets):
        if not self.graphs:
            return self._step(idx, targets)
        key = tuple(idx.shape)
        if key not in self._graph:
            s_idx, s_tgt = idx.clone(), targets.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                     # warm-up outside the capture (allocator, NCCL, optimiser state)
                for _ in range(2):
                    self._step(s_idx, s_tgt)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                s_loss = self._step(s_idx, s_tgt)
            self._graph[key] = (g, s_idx, s_tgt, s_loss)
        g, s_idx, s_tgt, s_loss = self._graph[key]
        s_idx.copy_(idx, non_blocking=True)
        s_tgt.copy_(targets, non_blocking=True)
        g.replay()
        return s_loss
