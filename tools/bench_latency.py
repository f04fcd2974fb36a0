#!/usr/bin/env python
"""Single-image latency of the whole pipeline (infer_image.py:82-86 with one 512 x 512 content / style pair, what the
reference's inference scripts run): eager launch loop vs one CUDA-graph replay (mhada_style_transfer_b200.graphs)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mhada_style_transfer_b200.graphs import GraphedStyleTransfer

ap = argparse.ArgumentParser()
ap.add_argument("--img", type=int, default=512)
ap.add_argument("--B", type=int, default=1)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
wl = dict(bench.WORKLOADS["cfg2"], B=a.B, hw=(a.img // 8, a.img // 8), hsws=(a.img // 8, a.img // 8))
dev = torch.device("cuda", 0)
vit_c, vit_s, ada = bench.build_models(wl, dev)
c, s = (t.to(dev) for t in bench.make_images(wl, 0))

def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters, (time.perf_counter() - t0) * 1e3 / a.iters

with torch.no_grad():
    eager = timed(lambda: ada(vit_c(c), vit_s(s)))
    want = ada(vit_c(c), vit_s(s))[1].float().clone()
g = GraphedStyleTransfer(vit_c, vit_s, ada, c, s)
graph = timed(lambda: g(c, s))
same = bool(torch.equal(g(c, s)[1].float(), want))
gv = GraphedStyleTransfer(vit_c, vit_s, ada, c, s, style="cached")
video = timed(lambda: gv(c))
print(json.dumps({"img": a.img, "batch": a.B, "eager_ms": round(eager[0], 4), "eager_wall_ms": round(eager[1], 4),
                  "graph_ms": round(graph[0], 4), "graph_wall_ms": round(graph[1], 4), "graph_bit_identical": same,
                  "graph_cached_style_ms": round(video[0], 4)}))
