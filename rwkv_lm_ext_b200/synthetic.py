"""Synthetic inputs of the benchmark / test shapes (there are no datasets or checkpoints offline)."""
import torch


def make_inputs(B, T, H, seed, decay="randn", device="cpu", u_scale=0.3):
    """bf16 r,k,v,w [B,T,H*64], u [H,64], gy [B,T,H*64].  decay: "randn" (the reference tests'
    w ~ N(0,1), tests/test_cpu.py:266) or "model" (time_decay-like U(-6,-1)+0.3N(0,1),
    src/model.py:407-411)."""
    g = torch.Generator().manual_seed(seed)
    C = H * 64
    r, k, v, gy = (torch.randn(B, T, C, generator=g).bfloat16() for _ in range(4))
    if decay == "randn":
        w = torch.randn(B, T, C, generator=g).bfloat16()
    else:
        w = (torch.rand(B, T, C, generator=g) * 5 - 6 + 0.3 * torch.randn(B, T, C, generator=g)).bfloat16()
    u = (torch.randn(H, 64, generator=g) * u_scale).bfloat16()
    return tuple(t.to(device) for t in (r, k, v, w, u, gy))


def make_bi_encoder(layers=24, D=2048, H=32, ffn=7168, vocab=65536, seed=0, device="cuda", dtype=torch.bfloat16):
    """A random-init module tree with the reference's attribute names for `RwkvEncoder`
    (src/model_encoder_run.py:296-350: emb, blocks[i].{ln0, ln1, ln2, att, ffn}, ln_out, emb_id / pad_id) at the
    RWKV-6 World shape given (defaults: 1B6 = L24, D2048, H32, FFN 7168), with
    parameter ranges like the reference's initialisation (src/model.py:375-432): there are no checkpoints offline."""
    from .tmix import Tmix_x060

    class CMix(torch.nn.Module):          # parameter names of RWKV_CMix_x060 (src/model.py:616-644)
        def __init__(self):
            super().__init__()
            self.time_maa_k = torch.nn.Parameter(torch.rand(1, 1, D))
            self.time_maa_r = torch.nn.Parameter(torch.rand(1, 1, D))
            self.key = torch.nn.Linear(D, ffn, bias=False)
            self.receptance = torch.nn.Linear(D, D, bias=False)
            self.value = torch.nn.Linear(ffn, D, bias=False)

    torch.manual_seed(seed)
    with torch.device(device):
        model = torch.nn.Module()
        model.emb = torch.nn.Embedding(vocab, D)
        model.blocks = torch.nn.ModuleList()
        for i in range(layers):
            b = torch.nn.Module()
            if i == 0:
                b.ln0 = torch.nn.LayerNorm(D)
            b.ln1, b.ln2 = torch.nn.LayerNorm(D), torch.nn.LayerNorm(D)
            b.att = Tmix_x060(D, H)
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
    return model.to(dtype).eval()


def make_passages(n, T=512, vocab=65536, seed=0, min_len=128, emb_id=1, pad_id=0):
    """[n, T] int64 token ids (CPU): random content of length U(min_len, T-1), then the embedding token, then
    padding -- the padded variant of SURVEY.md 8(d) config 3."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(2, vocab, (n, T), generator=g)
    lens = torch.randint(min_len, T, (n,), generator=g)
    for b_, ln in enumerate(lens.tolist()):
        idx[b_, ln] = emb_id
        idx[b_, ln + 1:] = pad_id
    return idx
