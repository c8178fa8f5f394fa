import os, sys, torch
sys.path.insert(0, "/root/repo")
import rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_inputs
M.load()
B, T, H = 4, 64, 2
C = H * 64
for masks in ([None]*4, [60, 40, 0, None]):
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=21, decay="model", device="cuda")
    mask = torch.ones(B, T, dtype=torch.int32, device="cuda")
    for b, p in enumerate(masks):
        if p is not None: mask[b, p:] = 0
    try:
        y = M.RUN_CUDA_RWKV6_BI(B, T, C, H, mask, r, k, v, w, u)
        torch.cuda.synchronize()
        M.set_impl("simt"); y2 = M.RUN_CUDA_RWKV6_BI(B, T, C, H, mask, r, k, v, w, u); M.set_impl("auto")
        print("ok", masks, ((y.float()-y2.float()).norm()/y2.float().norm()).item())
    except Exception as e:
        print("FAIL", masks, str(e)[-300:]); break
