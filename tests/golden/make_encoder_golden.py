"""Golden fixture of a whole bi-directional encoder, produced by RUNNING THE REFERENCE's own model code.

Run only in the build container (needs /root/reference):

    python tests/golden/make_encoder_golden.py

It imports, unmodified, `RwkvEncoder` from /root/reference/src/model_encoder_run.py with NO_CUDA=1 (its
CPU path: BiBlock / BiRWKV_Tmix_x060 / BiRWKV_CMix_x060 around run_rwkv6_forward, :31-75), builds a small
random-weight model (2 layers, D = 128 = 2 heads, FFN 448, vocab 512), runs `encode_sentence` on a padded
batch and stores the weights (fp16 to keep the file small; the model is run ON those rounded weights), the
token ids and the embeddings the reference produced.  Nothing of the reference is copied: only numbers.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
os.environ.update(NO_CUDA="1", RWKV_HEAD_SIZE_A="64", RWKV_JIT_ON="0", RWKV_MY_TESTING="x060",
                  RWKV_CTXLEN="64", RWKV_FLOAT_MODE="fp32")
sys.path.insert(0, REF)


def main():
    from src.model_encoder_run import RwkvEncoder
    args = types.SimpleNamespace(n_layer=2, n_embd=128, vocab_size=512, ctx_len=64, head_size_a=64, head_size_divisor=8,
                                 dim_att=128, dim_ffn=448, my_pos_emb=0, pre_ffn=0, head_qk=0, dropout=0.0,
                                 tiny_att_dim=0, tiny_att_layer=0, emb_id=1, pad_id=0)
    torch.manual_seed(1234)
    model = RwkvEncoder(args).eval()
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(("time_maa_w1", "time_maa_w2", "time_decay_w1", "time_decay_w2")):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)          # the reference inits these to ~1e-4
            elif "ln" in name and name.endswith("weight"):
                p.copy_(0.7 + 0.6 * torch.rand(p.shape, generator=g))
            elif "ln" in name and name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("emb.weight"):
                p.copy_(torch.randn(p.shape, generator=g))
            p.copy_(p.half().float())                                       # weights exactly representable in fp16
    B, T = 4, 48
    idx = torch.randint(4, args.vocab_size, (B, T), generator=g)
    lens = [47, 30, 12, 1]
    for b, n in enumerate(lens):
        idx[b, n] = args.emb_id
        idx[b, n + 1:] = args.pad_id
    with torch.no_grad():
        emb = model.encode_sentence(idx)
        _, hidden = model.forward(idx, True)
    out = {"idx": idx.numpy(), "emb": emb.numpy(), "hidden": hidden.numpy(),
           "n_layer": np.int64(args.n_layer), "n_embd": np.int64(args.n_embd), "dim_ffn": np.int64(args.dim_ffn),
           "emb_id": np.int64(args.emb_id), "pad_id": np.int64(args.pad_id)}
    for name, p in model.state_dict().items():
        out["w:" + name] = p.detach().half().numpy()
    path = os.path.join(OUT, "encoder_2x128.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", "emb", tuple(emb.shape), "hidden", tuple(hidden.shape))


if __name__ == "__main__":
    main()
