"""One SFT step (forward, loss, backward, AdamW) of a 2-layer 1B6-shape LoRA model on a 4 x 512 bucket, for an ncu launch list.
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/sft_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rwkv_lm_ext_b200 import sft

dev = torch.device("cuda")
with torch.device("meta"):
    model = sft.RwkvSft(2, 2048, 32, 7168, 65536, lora_r=8, lora_alpha=32)
model = model.to_empty(device=dev).bfloat16()
sft.init_like_reference(model, seed=0)
tr = sft.SftTrainer(model, graphs=False)
g = torch.Generator().manual_seed(0)
idx = torch.randint(2, 65536, (4, 512), generator=g).to(dev)
tgt = torch.randint(2, 65536, (4, 512), generator=g).to(dev)
for _ in range(3):
    loss = tr.step(idx, tgt)
torch.cuda.synchronize()
print("ok", float(loss))
