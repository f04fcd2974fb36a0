"""Development aid: error table of the kernel backward (mhada_layer_backward) against the fp32 PyTorch recompute
backward and, for the golden case, against the reference's float64 autograd."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cases, synth, mhada_oracle as O
import mhada_style_transfer_b200 as M
DEV = "cuda:0"
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(DEV)

def build(case, sd):
    m = M.AdaAttnMultiHead(case["C"], case["H"])
    m.load_state_dict(synth.to_torch(sd, torch.float32), strict=True)
    return m.to(DEV).eval()

def grads(case, impl, alias=False, precision="bf16"):
    fc, fs, fcs, sd = cases.layer_inputs(case)
    G_ = dev(synth.bellish(997, fc.shape, 0.0, 1.0))
    m = build(case, sd); m.precision = precision; m.backward_impl = impl
    a, b = dev(fc).requires_grad_(True), dev(fs).requires_grad_(True)
    c = a if alias else dev(fcs).requires_grad_(True)
    out = m(a, b, c)
    (out.float() * G_).sum().backward()
    r = {"fc": a.grad, "fs": b.grad, **({} if alias else {"fcs": c.grad})}
    r.update({k: p.grad for k, p in m.named_parameters()})
    return r

for case, alias in ((dict(B=2, C=128, H=2, hw=(10, 10), hsws=(8, 9), gain=1.0, seed=71), False),
                    (dict(B=2, C=512, H=8, hw=(32, 32), hsws=(32, 32), gain=1.0, seed=73), False),
                    (dict(B=1, C=512, H=8, hw=(20, 13), hsws=(17, 9), gain=1.0, seed=73), True)):
    k_, t_ = grads(case, "kernels", alias), grads(case, "torch", alias)
    f_ = grads(case, "torch", alias, "fp32")
    scale = max(v.abs().max().item() for v in t_.values())
    print("case", case, "scale", scale)
    worst = {}
    for k in t_:
        e = O.errors(k_[k].float().cpu().numpy(), t_[k].float().cpu().numpy())
        e2 = O.errors(t_[k].float().cpu().numpy(), f_[k].float().cpu().numpy())
        grp = k.split(".")[0] + "." + k.split(".")[-1] if "." in k else k
        w = worst.get(grp)
        if w is None or e["max_abs"] / (e["absmax"] + 2e-3 * scale) > w[0]:
            worst[grp] = (e["max_abs"] / (e["absmax"] + 2e-3 * scale), k, e, e2["max_abs_rel"])
    for grp, (r, k, e, e2) in worst.items():
        print(f"  {k:22s} max_abs {e['max_abs']:.3e} absmax {e['absmax']:.3e} rel {e['max_abs_rel']:.3e} fro {e['fro_rel']:.3e}   [torch bf16-fwd vs fp32-fwd rel {e2:.2e}]")
