import sys; sys.path.insert(0,'/root/repo')
import torch, rwkv_lm_ext_b200 as M
from rwkv_lm_ext_b200.synthetic import make_inputs
M.load()
T,H=1024,32; C=H*64
r,k,v,w,u,_=make_inputs(1,T,H,seed=0,decay="model",device="cuda")
r,k,v,w=(t[0].contiguous() for t in (r,k,v,w))
for impl in ("auto","simt"):
    M.set_impl(impl)
    st=torch.zeros(H,64,64,device="cuda")
    for _ in range(3): M.RUN_RWKV_6(1,T,C,H,st,r,k,v,w,u)
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): M.RUN_RWKV_6(1,T,C,H,st,r,k,v,w,u)
    b.record(); torch.cuda.synchronize()
    print(impl, a.elapsed_time(b)/20, "ms per 1024-token chunk (B=1, H=32)")
