#!/usr/bin/env python
"""ONE process driving TWO GPUs: per-device kernel attributes (DeviceOnce), SM counts, packed-weight
caches keyed per device, cluster (CTA-pair) launches on the second device.  Runs the image -> ViT x2 -> MHAda x6 -> decoder
pipeline, a training step and AdaAttnForLoss on cuda:1 FIRST (so nothing was initialised by device 0), then on cuda:0,
and compares the results bit for bit.  Needs a box with >= 2 GPUs (gpurun --gpus 2)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mhada_style_transfer_b200 as M
from mhada_style_transfer_b200.network import set_precision

assert torch.cuda.device_count() >= 2, "needs two GPUs"
torch.manual_seed(5)
vit_c, vit_s = M.VisionTransformer(pos_embedding=True), M.VisionTransformer(pos_embedding=False)
ada = set_precision(M.AdaAttnTransformerMultiHead(), "bf16")
loss_mod = M.AdaAttnForLoss(256, 448)
loss_mod.precision = "bf16"
c = (torch.rand(2, 3, 512, 512) * 255).floor()          # 2 x 4096 tokens: the CTA-pair GEMM / convolution paths
s = (torch.rand(2, 3, 512, 512) * 255).floor()
vgg = [torch.randn(1, ch, 32, 32) for ch in (256, 256, 448, 448)]
out = {}
for dev in ("cuda:1", "cuda:0", "cuda:1"):
    d = torch.device(dev)
    for m in (vit_c, vit_s, ada, loss_mod):
        m.to(d).eval()
    with torch.no_grad():
        fcs, cs = ada(vit_c(c.to(d)), vit_s(s.to(d)))
        fl = loss_mod(*[t.to(d) for t in vgg])
    for m in (vit_c, vit_s, ada):
        m.train()
        m.zero_grad(set_to_none=True)
    f2, c2 = ada(vit_c(c[:1, :, :256, :256].to(d)), vit_s(s[:1, :, :256, :256].to(d)))
    (c2.float().mean() + 1e-3 * f2.float().pow(2).mean()).backward()
    g = ada.adaAttnHead[0].out_conv.weight.grad
    torch.cuda.synchronize(d)
    res = (fcs.float().cpu(), cs.float().cpu(), fl.float().cpu(), g.float().cpu())
    if dev in out:
        same = all(torch.equal(a, b) for a, b in zip(out[dev], res))
        print(json.dumps({"device": dev, "second_visit_bit_identical": same}))
        assert same
    out[dev] = res
same = all(torch.equal(a, b) for a, b in zip(out["cuda:0"], out["cuda:1"]))
print(json.dumps({"devices": [torch.cuda.get_device_name(0), torch.cuda.get_device_name(1)], "cuda0_equals_cuda1": same,
                  "cs_mean": float(out["cuda:0"][1].mean()), "finite": bool(all(torch.isfinite(t).all() for t in out["cuda:0"]))}))
assert same
