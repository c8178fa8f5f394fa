"""Golden fixtures of the composed task models, produced by RUNNING THE REFERENCE's own model code on the CPU.

Run only in the build container (needs /root/reference):   python tests/golden/make_task_golden.py

Imports, unmodified, `RWKV` (src/model.py:1100-1242), `RwkvForSequenceEmbedding` (src/model_ext.py:1690-1769) and
`RwkvForClassification` (src/model_ext.py:172-212) with WKV=fla (so that nothing is JIT-compiled at import), points
`src.model.RUN_CUDA_RWKV6` at the reference's own CPU recurrence `run_rwkv6_forward`
(src/model_encoder_run.py:31-62) -- the Triton path needs a GPU -- and stubs the packages that are not installed and
not on the path (pytorch_lightning, deepspeed, sentence_transformers, bitsandbytes).  A small random-weight model
(2 layers, D = 128 = 2 heads, FFN 448, vocab 512; weights exactly representable in fp16) is run on a right-padded
batch; weights, ids and the three heads' outputs are stored.  The cross-encoder of the reference IS
`RwkvForClassification` on "query [sep] document [cls]" token rows (src/model_ext.py:1890-1960,
peft_train/data_collators.py), so its fixture is the classification one on such rows.  Only numbers are stored."""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def stubs():
    pl = types.ModuleType("pytorch_lightning")
    pl.__version__ = "2.0"
    pl.LightningModule = type("LightningModule", (nn.Module,), {})
    pl.Callback = object
    ut = types.ModuleType("pytorch_lightning.utilities")
    ut.rank_zero_info, ut.rank_zero_only = print, (lambda f: f)
    st = types.ModuleType("pytorch_lightning.strategies")
    st.DeepSpeedStrategy = type("DeepSpeedStrategy", (), {})
    import importlib.machinery
    ds = types.ModuleType("deepspeed")
    ds.__spec__ = importlib.machinery.ModuleSpec("deepspeed", None)
    ds_ops = types.ModuleType("deepspeed.ops")
    ds_adam = types.ModuleType("deepspeed.ops.adam")
    ds_adam.DeepSpeedCPUAdam = ds_adam.FusedAdam = object
    stf = types.ModuleType("sentence_transformers")
    stf.SentenceTransformer = object
    stf_u = types.ModuleType("sentence_transformers.util")
    stf_u.pairwise_cos_sim = stf_u.cos_sim = None
    sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.utilities": ut, "pytorch_lightning.strategies": st,
                        "deepspeed": ds, "deepspeed.ops": ds_ops, "deepspeed.ops.adam": ds_adam,
                        "sentence_transformers": stf, "sentence_transformers.util": stf_u,
                        "bitsandbytes": types.ModuleType("bitsandbytes")})


def main():
    os.environ.update(WKV="fla", RWKV_TRAIN_TYPE="", RWKV_MY_TESTING="x060", RWKV_JIT_ON="0", RWKV_HEAD_SIZE_A="64",
                      RWKV_CTXLEN="64", RWKV_FLOAT_MODE="fp32", RWKV_T_MAX="64", NO_CUDA="1")
    stubs()
    sys.path.insert(0, REF)
    import src.model as rm
    from src.model_encoder_run import run_rwkv6_forward
    rm.RUN_CUDA_RWKV6 = lambda B, T, C, H, r, k, v, w, u: run_rwkv6_forward(r, k, v, w, u)[0] if isinstance(
        run_rwkv6_forward(r, k, v, w, u), tuple) else run_rwkv6_forward(r, k, v, w, u)
    from src.model_ext import RwkvForClassification, RwkvForSequenceEmbedding
    args = types.SimpleNamespace(n_layer=2, n_embd=128, vocab_size=512, ctx_len=64, head_size_a=64, head_size_divisor=8,
                                 dim_att=128, dim_ffn=448, my_pos_emb=0, pre_ffn=0, head_qk=0, dropout=0.0, tiny_att_dim=0,
                                 tiny_att_layer=0, grad_cp=0, my_testing="x060", lora=False, state_tune=False, train_type="",
                                 chunk_ctx=64, my_qa_mask=0, lr_init=1e-4, weight_decay=0, layerwise_lr=0, my_pile_stage=0)
    torch.manual_seed(4321)
    base = rm.RWKV(args).eval().float()
    g = torch.Generator().manual_seed(77)
    with torch.no_grad():
        for name, p in base.named_parameters():
            if name.endswith(("time_maa_w1", "time_maa_w2", "time_decay_w1", "time_decay_w2")):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif "ln" in name and name.endswith("weight"):
                p.copy_(0.7 + 0.6 * torch.rand(p.shape, generator=g))
            elif "ln" in name and name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("emb.weight"):
                p.copy_(torch.randn(p.shape, generator=g))
            elif p.dim() == 2 and p.abs().max() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 / p.shape[1] ** 0.5))    # zero-initialised Linears
            p.copy_(p.half().float())
    B, T = 4, 48
    idx = torch.randint(4, args.vocab_size, (B, T), generator=g)
    lens = [47, 30, 12, 5]
    for b, n in enumerate(lens):
        idx[b, n] = 1                       # embedding / class id
        idx[b, n + 1:] = 0                  # padding
    idx[:, 7] = 2                           # a separator inside: the cross-encoder row layout "query [sep] document [cls]"
    out = {"idx": idx.numpy(), "n_layer": np.int64(2), "n_embd": np.int64(128), "dim_ffn": np.int64(448)}
    with torch.no_grad():
        for pool in ("weightedmean", "lasttoken", "avg"):
            emb = RwkvForSequenceEmbedding(base, embedding_id=1, pad_id=0, should_delete_head=False, pooling_type=pool)
            out["emb_" + pool] = emb(idx).float().numpy()
        mlp = RwkvForSequenceEmbedding(base, should_delete_head=False, pooling_type="lasttoken", add_mlp=True, output_dim=32)
        mlp.dense.weight.copy_((torch.randn(mlp.dense.weight.shape, generator=g) * 0.1).half().float())
        mlp.dense.bias.copy_((torch.randn(mlp.dense.bias.shape, generator=g) * 0.1).half().float())
        out["emb_mlp"] = mlp(idx).float().numpy()
        out["w:dense.weight"], out["w:dense.bias"] = mlp.dense.weight.half().numpy(), mlp.dense.bias.half().numpy()
        cls = RwkvForClassification(base, num_labels=3, class_id=1, pad_id=0, should_delete_head=False)
        cls.score.weight.copy_((torch.randn(cls.score.weight.shape, generator=g) * 0.1).half().float())
        out["logits"] = cls(idx).float().numpy()
        out["w:score.weight"] = cls.score.weight.half().numpy()
        x = base.emb(idx)
        for blk in base.blocks:
            x = blk(x)
        out["hidden"] = base.ln_out(x).float().numpy()
    for name, p in base.state_dict().items():
        out["w:rwkvModel." + name] = p.detach().half().numpy()
    path = os.path.join(OUT, "task_models_2x128.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if not k.startswith("w:")})


if __name__ == "__main__":
    main()
