#!/usr/bin/env python
"""One cfg5 training step per iteration (ViT x2 + MHAda x6 + decoder, forward + backward + Adam).  Development aid for
ncu captures of the backward kernels (profiles/)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
wl = bench.WORKLOADS["cfg5"]
dev = torch.device("cuda", 0)
vit_c, vit_s, model = bench.build_models(wl, dev)
for m in (vit_c, vit_s, model):
    m.train()
opts = [torch.optim.Adam(m.parameters(), lr=1e-4) for m in (vit_c, vit_s, model)]
c, s = (t.to(dev) for t in bench.make_images(wl, 0))
for i in range(a.steps):
    for o in opts:
        o.zero_grad(set_to_none=True)
    fcs, cs = model(vit_c(c), vit_s(s))
    loss = (cs.float() - c).pow(2).mean() * 1e-4 + fcs.float().pow(2).mean() * 1e-3
    loss.backward()
    for o in opts:
        o.step()
    torch.cuda.synchronize()
    print("step", i, "loss", float(loss))
