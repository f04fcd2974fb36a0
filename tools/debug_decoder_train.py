import sys, os, copy
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from oracle import cases, synth, mhada_oracle as O
import mhada_style_transfer_b200 as M
DEV='cuda:0'
case=dict(cases.DECODER_CASES[0], B=2, hw=(16,16))
x, sd = cases.decoder_inputs(case)
m=M.Decoder(); m.load_state_dict({k[len("decoder."):]: v for k, v in synth.to_torch(sd, torch.float32).items()}, strict=True); m=m.to(DEV).train()
G_=torch.randn(2,3,128,128,device=DEV)
xt=torch.from_numpy(x).float().to(DEV)
def run(mod, impl, dt):
    mod.train_impl=impl; mod.zero_grad(set_to_none=True)
    xin=xt.detach().to(dt).clone().requires_grad_(True)
    out=mod(xin); (out.float()*G_).sum().backward()
    return out.detach().float(), {"x": xin.grad.float(), **{k:p.grad.float().clone() for k,p in mod.named_parameters()}}
a=run(m,"kernels",torch.float32)
b=run(m,"torch",torch.float32)
m16=copy.deepcopy(m).to(torch.bfloat16)
c=run(m16,"torch",torch.bfloat16)
f=lambda u,v: float((u-v).norm()/v.norm())
print("out  k-vs-f32 %.4f  bf16torch-vs-f32 %.4f  k-vs-bf16torch %.4f"%(f(a[0],b[0]), f(c[0],b[0]), f(a[0],c[0])))
for k in ["x","conv1.0.conv.conv.weight","conv1.2.conv.conv.weight","conv2.1.conv.conv.weight","conv3.1.conv.conv.weight","conv3.1.conv.conv.bias","conv1.0.conv.conv.bias"]:
    print("%-28s k-vs-f32 %.4f  bf16torch-vs-f32 %.4f  k-vs-bf16torch %.4f   |g| %.3g"%(k, f(a[1][k],b[1][k]), f(c[1][k],b[1][k]), f(a[1][k],c[1][k]), float(b[1][k].abs().max())))
