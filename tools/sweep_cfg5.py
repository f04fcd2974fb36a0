"""BASELINE configs[4]: head-count and feature-level sweep (1/4/8 heads x three levels, AdaAttnForLoss shapes) and a
train_image.py-shaped forward + backward step, timed on one B200 (CUDA events, best of N).  Development aid; prints
one JSON line per case.

    python tools/sweep_cfg5.py [--hw 64] [--train-hw 32]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mhada_style_transfer_b200 as M  # noqa: E402
from mhada_style_transfer_b200.network import set_precision  # noqa: E402


def best_ms(fn, it=5):
    for _ in range(2):
        fn()
    out = []
    for _ in range(it):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return min(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", type=int, default=64)
    ap.add_argument("--train-hw", type=int, default=32)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    N = a.hw * a.hw
    # (1) heads sweep, one MHAda layer, B = 1, C = 512
    for heads in (1, 2, 4, 8):
        for prec in ("fp32", "bf16"):
            if False:
                continue                                   # the tensor-core path implements head_dim 64 and 128
            m = set_precision(M.AdaAttnMultiHead(512, heads).to(dev).eval(), prec)
            dt = torch.bfloat16 if prec == "bf16" else torch.float32
            fc, fs = (torch.randn(1, 512, a.hw, a.hw, device=dev).mul_(85).to(dt).contiguous(memory_format=torch.channels_last)
                      for _ in range(2))
            with torch.no_grad():
                ms = best_ms(lambda: m(fc, fs, fc))
            flops = 6.0 * N * N * 512
            print(json.dumps({"case": "layer", "heads": heads, "head_dim": 512 // heads, "precision": prec, "tokens": N,
                              "ms": round(ms, 4), "attn_tflops_equiv": round(flops / ms / 1e9, 1)}), flush=True)
    # (2) AdaAttnForLoss shapes (train_image.py:52-58) at 256 x 256 images: (v_dim, qk_dim, N)
    for name, v, qk, n_side in (("relu3_1", 256, 448, 64), ("relu4_1", 512, 960, 32), ("relu5_1", 512, 1472, 16)):
        m = M.AdaAttnForLoss(v, qk).to(dev).eval()
        cx, sx = (torch.randn(8, v, n_side, n_side, device=dev) for _ in range(2))
        c1, s1 = (torch.randn(8, qk, n_side, n_side, device=dev) for _ in range(2))
        for prec in ("fp32", "bf16"):
            m.precision = prec
            with torch.no_grad():
                ms = best_ms(lambda: m(cx, sx, c1, s1))
            print(json.dumps({"case": "forloss", "level": name, "v_dim": v, "qk_dim": qk, "tokens": n_side * n_side, "batch": 8,
                              "precision": prec, "ms": round(ms, 4)}), flush=True)
    # (3) train step: batch 8 of 256 x 256 images -> 32 x 32 tokens; kernels forward, recompute backward, Adam
    m = set_precision(M.AdaAttnTransformerMultiHead().to(dev).train(), "bf16")
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    th = a.train_hw
    fc = [torch.randn(8, 512, th, th, device=dev).mul_(85) for _ in range(3)]
    fs = [torch.randn(8, 512, th, th, device=dev).mul_(85) for _ in range(3)]

    def step():
        opt.zero_grad(set_to_none=True)
        fcs, cs = m(fc, fs)
        (cs.float().mean() + 1e-3 * fcs.float().pow(2).mean()).backward()
        opt.step()

    ms = best_ms(step, it=3)
    print(json.dumps({"case": "train_step", "batch": 8, "tokens": th * th, "forward": "bf16 kernels",
                      "backward": "fp32 recompute (PyTorch)", "ms": round(ms, 3), "images_per_s": round(8 / ms * 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
