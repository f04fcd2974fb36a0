"""Per-stage timing of one MHAda layer on the B200 (CUDA events on the launching stream, L2 flushed
between iterations).  Development aid; bench.py is the contract harness.

    python tools/bench_stages.py [--B 8] [--hw 64] [--dtype bf16]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhada_style_transfer_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--hw", type=int, default=64)
    ap.add_argument("--hws", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    L = _lib.lib()
    dev = "cuda:0"
    B, H, d, C = a.B, 8, 64, 512
    Nc = a.hw * a.hw
    Ns = (a.hws or a.hw) ** 2
    code = _lib.BF16 if a.dtype == "bf16" else _lib.F32
    dt = torch.bfloat16 if code == _lib.BF16 else torch.float32
    g = torch.Generator(device=dev).manual_seed(0)
    fc = (torch.randn(B, Nc, C, device=dev, generator=g) * 85).to(dt)
    fs = (torch.randn(B, Ns, C, device=dev, generator=g) * 85).to(dt)
    fcs = (torch.randn(B, Nc, C, device=dev, generator=g) * 30).to(dt)
    w = (torch.rand(3, H, d, d, device=dev, generator=g) - 0.5) / 4
    b = (torch.rand(3, H, d, device=dev, generator=g) - 0.5) / 4
    wo = (torch.rand(C, C, device=dev, generator=g) - 0.5) / 11
    bo = (torch.rand(C, device=dev, generator=g) - 0.5) / 11
    out = torch.empty(B, Nc, C, dtype=dt, device=dev)
    ws = torch.empty(L.mhada_layer_workspace(code, B, Nc, Ns, C, H), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def layer():
        _lib.check("layer", L.mhada_layer_forward(code, P(fc), P(fs), P(fcs), P(w), P(b), P(wo), P(bo), B, Nc, Ns, C, H, 0,
                                                  P(out), P(ws), ws.numel(), st))

    # stage pieces (re-using the layer's own buffers through the public stage entry points)
    mean = torch.empty(6, B, C, device=dev)
    sws = torch.empty(L.mhada_in_stats_workspace(B, max(Nc, Ns), C) + 1024, dtype=torch.uint8, device=dev)
    q = torch.empty(B, Nc, C, dtype=dt, device=dev)
    k = torch.empty(B, Ns, C, dtype=dt, device=dev)
    v = torch.empty(B, Ns, C * (2 if code == _lib.BF16 else 1), dtype=dt, device=dev)
    muv = torch.empty(B, C, device=dev)
    pws = torch.empty(L.mhada_proj_workspace(B, H, d) + 1024, dtype=torch.uint8, device=dev)
    heads = torch.empty(B, Nc, C, dtype=dt, device=dev)
    lws = torch.empty(L.mhada_linear_workspace(code, C, C) + 1024, dtype=torch.uint8, device=dev)

    def stats():
        _lib.check("stats", L.mhada_in_stats(P(fc), code, B, Nc, C, C, P(mean[0]), P(mean[1]), P(sws), sws.numel(), st))
        _lib.check("stats", L.mhada_in_stats(P(fs), code, B, Ns, C, C, P(mean[2]), P(mean[3]), P(sws), sws.numel(), st))
        _lib.check("stats", L.mhada_in_stats(P(fcs), code, B, Nc, C, C, P(mean[4]), P(mean[5]), P(sws), sws.numel(), st))

    def proj():
        _lib.check("proj", L.mhada_proj(code, 3, P(fc), P(fs), P(mean[0]), P(mean[1]), P(mean[2]), P(mean[3]), P(w), P(b),
                                        B, B, Nc, Ns, H, d, P(q), P(k), P(v), P(muv), P(pws), pws.numel(), st))

    args = _lib.AttnArgs()
    args.dtype = code
    args.B, args.H, args.Nc, args.Ns, args.dqk, args.dv = B, H, Nc, Ns, d, d
    args.q, args.k, args.v, args.x, args.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), fcs.data_ptr(), heads.data_ptr()
    args.ldq, args.ldk, args.ldv, args.ldx, args.ldo = C, C, v.shape[2], C, C
    args.x_mean, args.x_rstd, args.mu_v = mean[4].data_ptr(), mean[5].data_ptr(), muv.data_ptr()

    def attn():
        _lib.check("attn", L.mhada_attn(ctypes.byref(args), st))

    def linear():
        _lib.check("linear", L.mhada_linear(code, P(heads), C, P(wo), P(bo), B * Nc, C, C, P(out), C, P(lws), lws.numel(), st))

    def timeit(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    res = {"B": B, "Nc": Nc, "Ns": Ns, "dtype": a.dtype}
    stats(); proj()
    for name, fn in (("stats", stats), ("proj", proj), ("attn", attn), ("linear", linear), ("layer", layer)):
        med, best = timeit(fn, a.iters)
        res[name + "_ms"] = round(med, 4)
        res[name + "_best_ms"] = round(best, 4)
    flops = 6.0 * B * Nc * Ns * C
    res["attn_tflops"] = round(flops / (res["attn_ms"] * 1e-3) / 1e12, 1)
    res["attn_tflops_best"] = round(flops / (res["attn_best_ms"] * 1e-3) / 1e12, 1)
    e = 2 if code == _lib.BF16 else 4
    res["stats_gbs"] = round(e * B * C * (2 * Nc + Ns) / (res["stats_ms"] * 1e-3) / 1e9, 1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
