"""The whole inference step as ONE CUDA graph.

Every entry point of the C ABI launches on the caller's stream, allocates nothing and never synchronises, so a sequence
of calls is capturable (include/mhada_b200.h, "Conventions").  For the batch sizes the reference's scripts actually use
at inference time -- ONE image (infer_image.py:82-86) or one video frame at a time (infer_video.py:88-92) -- the 89
launches of a step take less GPU time than the Python / ctypes launch loop that issues them, so the step is CPU-bound;
replaying a captured graph removes that loop (tools/bench_latency.py: 512 x 512, one image).

    g = GraphedStyleTransfer(vit_c, vit_s, adaFormer, c_example, s_example)
    fcs, cs = g(c, s)            # same shapes / dtypes as the examples; results are views of graph-owned buffers
"""
from __future__ import annotations

import torch

from . import network

__all__ = ["GraphedStyleTransfer"]


class GraphedStyleTransfer:
    """`fc = vit_c(c); fs = vit_s(s); fcs, cs = adaFormer(fc, fs)` captured once for fixed input shapes.

    style: "per_call" re-encodes the style image every call (infer_image.py); "cached" encodes it once at construction
    through `adaFormer.precompute_style` and captures the content side only (infer_video.py: one style, many frames;
    `set_style(s)` replaces it without re-capturing)."""

    def __init__(self, vit_c, vit_s, ada, c_example: torch.Tensor, s_example: torch.Tensor, style: str = "per_call",
                 warmup: int = 2):
        if style not in ("per_call", "cached"):
            raise ValueError(f"Unknown style mode: {style}")
        if not (c_example.is_cuda and s_example.is_cuda):
            raise RuntimeError("GraphedStyleTransfer needs CUDA tensors (there is no CPU path)")
        self.vit_c, self.vit_s, self.ada, self.style = vit_c, vit_s, ada, style
        self.c = c_example.clone()
        self.s = s_example.clone()
        self._cache = None
        dev = c_example.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            if style == "cached":
                self._cache = ada.precompute_style(vit_s(self.s))
            for _ in range(max(1, warmup)):          # weight caches, kernel attributes, workspaces: all outside the graph
                self._step()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        before = set(network._WORKSPACES)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.fcs, self.cs = self._step()
        # scratch buffers taken during capture belong to the graph's memory pool: no eager call may pick them up
        for k in set(network._WORKSPACES) - before:
            del network._WORKSPACES[k]

    def _step(self):
        fc = self.vit_c(self.c)
        fs = self._cache if self._cache is not None else self.vit_s(self.s)
        return self.ada(fc, fs)

    def set_style(self, s: torch.Tensor):
        """style == "cached": new style image, same shape; the cache buffers are rewritten in place (no re-capture)."""
        if self._cache is None:
            raise RuntimeError('set_style() needs style="cached"')
        with torch.no_grad():
            new = self.ada.precompute_style(self.vit_s(s))
            for old, buf in zip(self._cache.buffers, new.buffers):
                old.copy_(buf)

    def __call__(self, c: torch.Tensor, s: torch.Tensor = None):
        if c.shape != self.c.shape:
            raise RuntimeError(f"captured for content {tuple(self.c.shape)}, got {tuple(c.shape)}")
        self.c.copy_(c, non_blocking=True)
        if s is not None and self._cache is None:
            if s.shape != self.s.shape:
                raise RuntimeError(f"captured for style {tuple(self.s.shape)}, got {tuple(s.shape)}")
            self.s.copy_(s, non_blocking=True)
        self.graph.replay()
        return self.fcs, self.cs
