"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run only in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports, unmodified, from /root/reference:
  * src/model_encoder_run.py : run_rwkv6_forward (:31-62), create_mask (:7-11), reverse_x_idx (:19-26)
  * tests/test_cpu.py        : pytorch_forward (:190-231)  (function definitions only -- the tail of
                               that script JIT-builds CUDA and cannot run without a GPU)
  * fla/ops/rwkv6/recurrent_naive.py : naive_recurrent_rwkv6 (:8-36) for fp32 autograd gradients and
                               the initial-state form
  * src/model_ext.py / src/model_run.py : the two `pooling` methods (train :1708-1738, infer :777-797),
                               taken by compiling just those method bodies out of the source text
and writes inputs + reference outputs to small .npz files.  Nothing from the reference is copied
into the repo: only numbers it produced.
"""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

os.environ.update(NO_CUDA="1", RWKV_HEAD_SIZE_A="64", RWKV_JIT_ON="0", RWKV_MY_TESTING="x060",
                  RWKV_CTXLEN="4096", RWKV_FLOAT_MODE="bf16", RWKV_T_MAX="4096", WKV="fla",
                  RWKV_TRAIN_TYPE="")
sys.path.insert(0, REF)


def _functions_from(path, names):
    """exec only the named top-level function / method definitions of a reference file."""
    src = open(path).read()
    tree = ast.parse(src)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            found.setdefault(node.name, []).append(node)
    ns = {"torch": torch, "HEAD_SIZE": 64, "F": torch.nn.functional}
    out = {}
    for name, nodes in found.items():
        out[name] = []
        for node in nodes:
            seg = textwrap.dedent(ast.get_source_segment(src, node))
            loc = dict(ns)
            exec(compile(seg, path, "exec"), loc)
            out[name].append(loc[name])
    return out


def bf16_exact(*shape, gen, scale=1.0, shift=0.0):
    return (torch.randn(*shape, generator=gen) * scale + shift).bfloat16().float()


def main():
    from src.model_encoder_run import run_rwkv6_forward, create_mask, reverse_x_idx
    pytorch_forward = _functions_from(f"{REF}/tests/test_cpu.py", {"pytorch_forward"})["pytorch_forward"][0]
    from fla.ops.rwkv6.recurrent_naive import naive_recurrent_rwkv6

    def fla_grads(r, k, v, w, u, gy, s0_kv=None):
        """fp32 autograd through the reference's vendored naive recurrence ([B,H,T,K] layout,
        w = log decay = -exp(w_raw), state [K,V]); src/model.py:64-71 shows the same mapping."""
        B, T, C = r.shape
        H = u.shape[0]
        to = lambda x: x.view(B, T, H, 64).transpose(1, 2).contiguous()
        leaves = [x.clone().requires_grad_(True) for x in (r, k, v, w, u)]
        r_, k_, v_, w_, u_ = leaves
        s_ = None if s0_kv is None else s0_kv.clone().requires_grad_(True)
        o = naive_recurrent_rwkv6(to(r_), to(k_), to(v_), to(-torch.exp(w_)), u_, initial_state=s_)
        y = o.transpose(1, 2).reshape(B, T, C)
        y.backward(gy)
        res = dict(y_fla=y.detach(), gr=r_.grad, gk=k_.grad, gv=v_.grad, gw=w_.grad, gu=u_.grad)
        if s_ is not None:
            res["gs_kv"] = s_.grad
        return res

    cases = {}
    # ---- case 1: the reference's own op test shape and distribution (tests/test_cpu.py:260-267)
    g = torch.Generator().manual_seed(20241018)
    B, T, C, H = 2, 10, 256, 4
    r, k, v, w = (bf16_exact(B, T, C, gen=g) for _ in range(4))
    u = bf16_exact(H, 64, gen=g)
    gy = bf16_exact(B, T, C, gen=g)
    c = dict(r=r, k=k, v=v, w=w, u=u, gy=gy)
    c["y_run_rwkv6_forward"] = run_rwkv6_forward(r.clone(), k.clone(), v.clone(), w.clone(), u.clone())
    c["y_pytorch_forward"] = pytorch_forward(B, T, C, H, r.clone(), k.clone(), v.clone(), w.clone(), u.clone())
    c.update(fla_grads(r, k, v, w, u, gy))
    cases["wkv6_2x10x256_randn"] = c

    # ---- case 2: realistic decay range (src/model.py:407-411), longer T, crosses chunk borders
    g = torch.Generator().manual_seed(7)
    B, T, C, H = 2, 150, 128, 2
    r, k, v = (bf16_exact(B, T, C, gen=g) for _ in range(3))
    w = (torch.rand(B, T, C, generator=g) * 5 - 6 + 0.3 * torch.randn(B, T, C, generator=g)).bfloat16().float()
    u = bf16_exact(H, 64, gen=g, scale=0.3)
    gy = bf16_exact(B, T, C, gen=g)
    c = dict(r=r, k=k, v=v, w=w, u=u, gy=gy)
    c["y_run_rwkv6_forward"] = run_rwkv6_forward(r.clone(), k.clone(), v.clone(), w.clone(), u.clone())
    c.update(fla_grads(r, k, v, w, u, gy))
    cases["wkv6_2x150x128_decay"] = c

    # ---- case 3: non-zero initial state through the fla naive recurrence ([K,V] layout)
    g = torch.Generator().manual_seed(11)
    B, T, C, H = 2, 70, 128, 2
    r, k, v = (bf16_exact(B, T, C, gen=g) for _ in range(3))
    w = bf16_exact(B, T, C, gen=g, scale=0.7, shift=-1.0)
    u = bf16_exact(H, 64, gen=g, scale=0.3)
    gy = bf16_exact(B, T, C, gen=g)
    s0 = bf16_exact(B, H, 64, 64, gen=g, scale=0.5)          # [B,H,key,value]
    c = dict(r=r, k=k, v=v, w=w, u=u, gy=gy, s0_kv=s0)
    c.update(fla_grads(r, k, v, w, u, gy, s0_kv=s0))
    cases["wkv6state_2x70x128"] = c

    for name, c in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **{kk: vv.numpy() for kk, vv in c.items()})
        print(name, {kk: tuple(vv.shape) for kk, vv in c.items()})

    # ---- integer pieces: create_mask / reverse_x_idx
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(2, 1000, (6, 24), generator=g)
    idx[0, 23] = 1
    idx[1, 10] = 1; idx[1, 11:] = 0
    idx[2, 0] = 1; idx[2, 1:] = 0
    idx[3, 5] = 1; idx[3, 17] = 1; idx[3, 18:] = 0
    idx[4, :] = 0
    mask = create_mask(idx)
    rev = reverse_x_idx(mask, idx.size(1))
    np.savez_compressed(os.path.join(OUT, "mask_rev_idx.npz"), idx=idx.numpy(), mask=mask.numpy(), rev_idx=rev.numpy(),
                        eos_pos=torch.eq(idx, 1).int().argmax(-1).numpy())
    print("mask_rev_idx", tuple(idx.shape))

    # ---- pooling methods (class methods that only read self.pooling_type)
    pool_train = _functions_from(f"{REF}/src/model_ext.py", {"pooling"})["pooling"]
    pool_infer = _functions_from(f"{REF}/src/model_run.py", {"pooling"})["pooling"]
    g = torch.Generator().manual_seed(5)
    x = bf16_exact(4, 24, 96, gen=g).bfloat16()
    L = torch.tensor([23, 10, 5, 1])
    res = dict(x=x.float().numpy(), actual_len=L.numpy())
    # src/model_ext.py has several identical copies of `pooling`; the last one (:1708) has 'avg'
    for kind in ("weightedmean", "lasttoken", "avg"):
        me = types.SimpleNamespace(pooling_type=kind)
        res[f"train_{kind}"] = pool_train[-1](me, x, L).float().numpy()
    for kind in ("weightedmean", "lasttoken"):
        me = types.SimpleNamespace(pooling_type=kind)
        res[f"infer_{kind}"] = pool_infer[-1](me, x, L).float().numpy()
    np.savez_compressed(os.path.join(OUT, "pooling.npz"), **res)
    print("pooling", {kk: vv.shape for kk, vv in res.items()})


if __name__ == "__main__":
    main()
