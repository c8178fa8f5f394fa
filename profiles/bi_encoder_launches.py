"""Two layers of the 1B6 bi-encoder forward (64 x 512 tokens) for an ncu launch list (per-kernel times of one layer).
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/bi_encoder_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_bi_encoder, make_passages

M.load()
model = make_bi_encoder(2, 2048, 32, 7168, 65536, seed=0, device="cuda")
idx = make_passages(64, 512, 65536, seed=100).cuda()
with torch.no_grad():
    for _ in range(3):
        e = M.bi_encoder_encode(model, idx)
torch.cuda.synchronize()
print("ok", float(e.float().abs().mean()))
