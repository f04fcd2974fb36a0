"""Development aid: kernel-time breakdown of the cfg5 training step (torch.profiler, one GPU)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

wl = bench.WORKLOADS["cfg5"]
device = torch.device("cuda:0")
vit_c, vit_s, model = bench.build_models(wl, device)
for m in (vit_c, vit_s, model):
    m.train()
opts = [torch.optim.Adam(m.parameters(), lr=1e-4) for m in (vit_c, vit_s, model)]
c_h, s_h = bench.make_images(wl, seed=0)
c, s = c_h.to(device), s_h.to(device)

def step():
    for o in opts:
        o.zero_grad(set_to_none=True)
    fc, fs = vit_c(c), vit_s(s)
    fcs, cs = model(fc, fs)
    loss = (cs.float() - c).pow(2).mean() * 1e-4 + fcs.float().pow(2).mean() * 1e-3
    loss.backward()
    for o in opts:
        o.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if e.device_type.name == "CUDA" and t > 0:
        rows.append((t / 3.0, e.count // 3, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time per step {tot/1e3:.2f} ms over {sum(r[1] for r in rows)} launches")
for t, n, k in rows[:45]:
    print(f"{t/1e3:8.3f} ms  {100*t/tot:5.1f}%  x{n:<4d} {k[:110]}")
