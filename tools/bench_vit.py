#!/usr/bin/env python
"""Development aid: device times of the ViT encoder (mhada_vit_forward) and of its GEMMs, CUDA events, L2 flushed
between iterations.  python tools/bench_vit.py [--B 8 --img 512]"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mhada_style_transfer_b200 as M
from mhada_style_transfer_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--img", type=int, default=512)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = "cuda:0"
L = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=a.iters):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

torch.manual_seed(0)
vit = M.VisionTransformer().to(dev).eval()
x = (torch.rand(a.B, 3, a.img, a.img, device=dev) * 255).floor()
with torch.no_grad():
    ms = timeit(lambda: vit(x))
N = (a.img // 8) ** 2
Mtok = a.B * N
flops = 2.0 * Mtok * (192 * 512 + 3 * 512 * ((1536 if a.B > 1 else 512) + 512 + 2048 + 2048))
out = {"B": a.B, "img": a.img, "tokens": Mtok, "vit_ms": round(ms, 4), "vit_tflops": round(flops / ms / 1e9, 1), "gemms": []}
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
for name, N_, K, mode in (("patch", 512, 192, "f32"), ("in_proj", 1536, 512, "bf16"), ("out_proj", 512, 512, "resid"),
                          ("fc1", 2048, 512, "bf16"), ("fc2", 512, 2048, "both")):
    xa = torch.randn(Mtok, K, device=dev).bfloat16()
    w = (torch.randn(N_, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N_, device=dev)
    yb = torch.empty(Mtok, N_, device=dev, dtype=torch.bfloat16) if mode in ("bf16", "both") else None
    yf = torch.empty(Mtok, N_, device=dev) if mode in ("f32", "resid", "both") else None
    r = torch.randn(Mtok, N_, device=dev) if mode in ("resid", "both") else None
    fn = lambda: _lib.check("gemm", L.mhada_gemm_bf16(P(xa), K, P(w), K, P(b), Mtok, N_, K, P(yb), N_, P(yf), N_, P(r), N_, 0, 0, st))
    t = timeit(fn)
    byt = 2 * Mtok * K + 2 * N_ * K + Mtok * N_ * ((2 if yb is not None else 0) + (4 if yf is not None else 0) + (4 if r is not None else 0))
    out["gemms"].append({"name": name, "N": N_, "K": K, "ms": round(t, 4), "tflops": round(2.0 * Mtok * N_ * K / t / 1e9, 1),
                         "gbs": round(byt / t / 1e6, 1)})
# small kernels
xf = torch.randn(Mtok, 512, device=dev)
g = torch.ones(512, device=dev); be = torch.zeros(512, device=dev)
y = torch.empty(Mtok, 512, device=dev, dtype=torch.bfloat16)
t = timeit(lambda: L.mhada_layernorm(P(xf), Mtok, 512, P(g), P(be), 1e-6, P(y), st))
out["layernorm"] = {"ms": round(t, 4), "gbs": round(Mtok * 512 * 6 / t / 1e6, 1)}
if a.B > 1:
    qkv = torch.randn(a.B, N, 1536, device=dev).bfloat16()
    o = torch.empty(a.B, N, 512, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: L.mhada_batch_attn(P(qkv), a.B, N, 8, 64, P(o), st))
    out["batch_attn"] = {"ms": round(t, 4), "gbs": round(Mtok * 2048 * 2 / t / 1e6, 1)}
print(json.dumps(out))
