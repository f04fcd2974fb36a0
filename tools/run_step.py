#!/usr/bin/env python
"""One e2e step of the cfg2 (or --workload) pipeline per iteration: images resident on the device -> vit_c, vit_s ->
MHAda x6 -> decoder.  Development aid for ncu captures (profiles/): `--steps 2` = one warm-up step + the captured one.
Prints the number of this library's launches per step (for ncu's -s / -c)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mhada_style_transfer_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
vit_c, vit_s, ada = bench.build_models(wl, dev)
c, s = (t.to(dev) for t in bench.make_images(wl, 0))
L = _lib.lib()
with torch.no_grad():
    for i in range(a.steps):
        n0 = L.mhada_total_launch_count()
        fcs, cs = ada(vit_c(c), vit_s(s))
        torch.cuda.synchronize()
        print("step", i, "launches", L.mhada_total_launch_count() - n0, "cs mean", float(cs.float().mean()))
