#!/usr/bin/env python
"""Development aid: the cfg5 training step (ViT x2 + MHAda x6 + decoder, forward + backward + Adam) captured as ONE CUDA
graph and replayed, against the eager launch loop (1150 launches per step).  One GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

wl = bench.WORKLOADS["cfg5"]
dev = torch.device("cuda", 0)

def make():
    torch.manual_seed(1234)
    vit_c, vit_s, model = bench.build_models(wl, dev)
    for m in (vit_c, vit_s, model):
        m.train()
    opts = [torch.optim.Adam(m.parameters(), lr=1e-4, capturable=True) for m in (vit_c, vit_s, model)]
    return vit_c, vit_s, model, opts

c_h, s_h = bench.make_images(wl, seed=0)
c, s = c_h.to(dev), s_h.to(dev)

def step(nets):
    vit_c, vit_s, model, opts = nets
    for o in opts:
        o.zero_grad(set_to_none=True)
    fcs, cs = model(vit_c(c), vit_s(s))
    loss = (cs.float() - c).pow(2).mean() * 1e-4 + fcs.float().pow(2).mean() * 1e-3
    loss.backward()
    for o in opts:
        o.step()
    return loss

def timed(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

nets = make()
eager_losses = [float(step(nets)) for _ in range(6)]
eager_ms = timed(lambda: step(nets))

nets = make()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    warm = [float(step(nets)) for _ in range(3)]
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    static_loss = step(nets)
graph_losses = list(warm)
for _ in range(3):
    g.replay()
    graph_losses.append(float(static_loss))
graph_ms = timed(g.replay)
print(json.dumps({"eager_ms": round(eager_ms, 3), "graph_ms": round(graph_ms, 3), "eager_losses": [round(x, 6) for x in eager_losses],
                  "graph_losses": [round(x, 6) for x in graph_losses]}))
