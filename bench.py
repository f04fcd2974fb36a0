#!/usr/bin/env python
"""bench.py -- images/sec of the MHAda forward hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg3|cfg1]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (N=1 default) = BASELINE.json configs[1]: batch 8 of 512x512 content/style, bf16, MHAda x6 + decoder.
One step = one forward of AdaAttnTransformerMultiHead over one batch per GPU (weak scaling: 8 images per GPU,
sharded by image, no data-path collective; outputs gathered to rank 0 over NCCL at the end of every step).

One JSON line on rank 0:
  value     images/s of the hot path BASELINE.json names (MHAda x6 + decoder), the ViT feature maps resident in
            HBM (produced once, outside the timed region, by the B200 ViT from the synthetic images), device-timed
            (CUDA events), max over ranks
  e2e       images/s through the public API a user calls (infer_image.py:83-85): HOST (pinned) content / style
            IMAGES -> H2D -> vit_c, vit_s -> adaFormer -> D2H of the decoded images, all inside the timed region
            (r1 copied 201 MB of feature maps per step; with the ViT on the device the boundary is the image)
  roofline  the attention kernel (tcgen05): algorithmic FLOPs 6*B*Nc*Ns*C per launch / its device time
            measured with CUDA events on the launching stream INSIDE the timed region; frac is against the
            measured BURST peak (the timed region is ~0.1 s)
  also      the same measurements for BASELINE configs[2] (1024^2, the size the north-star target is stated on)
  cpu_baseline  the PyTorch-CPU port of the reference path (oracle/torch_port.py) on this box's cores
--impl reference times that CPU port alone (the reference is Python and cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (batch per GPU, content tokens h,w, style tokens h,w, dtype)
    "cfg2": dict(B=8, hw=(64, 64), hsws=(64, 64), dtype="bf16",
                 desc="BASELINE configs[1]: batch 8 of 512x512 content/style, bf16, MHAda x6 + decoder"),
    "cfg3": dict(B=1, hw=(128, 128), hsws=(128, 128), dtype="bf16",
                 desc="BASELINE configs[2]: 1024x1024 content/style (16384 tokens), bf16, MHAda x6 + decoder"),
    "cfg4": dict(B=2, hw=(135, 240), hsws=(64, 64), dtype="bf16", style_batch=1, cached_style=True,
                 desc="BASELINE configs[3]: 1080p frames (135x240 tokens) x one 512x512 style, bf16, style side cached, "
                      "2 frames per step per GPU"),
    "cfg5": dict(B=8, hw=(32, 32), hsws=(32, 32), dtype="bf16", train=True,
                 desc="BASELINE configs[4]: train_image.py-shaped step, batch 8 of 256x256 per GPU: ViT x2 + MHAda x6 + decoder "
                      "forward, backward, gradient all-reduce over NVLink, Adam"),
    "cfg1": dict(B=1, hw=(64, 64), hsws=(64, 64), dtype="fp32",
                 desc="BASELINE configs[0]: single 512x512 content/style, fp32, MHAda x6 + decoder"),
}
C, H, LAYERS = 512, 8, 3


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe), through
    NVML in a background thread (5 ms period; `nvidia-smi -lms` cannot sample a 100 ms region)."""

    def __init__(self, index: int):
        self.index, self.rows, self.stop_flag, self.thread, self.nvml = index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self.nvml = None
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def _run(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, pw, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self):
        if not self.nvml:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        n = self.nvml
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({k for r in self.rows for k, bit in names.items() if r[2] & bit})
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_min_mhz": (sm[0] if sm else None),
                "sm_max_mhz": self.max_sm, "power_w_max": max((r[1] for r in self.rows), default=None),
                "samples": len(sm), "reasons": reasons}


def bind_to_gpu_numa_node(local: int):
    """Pin this process (and so the pinned host buffers it allocates next) to the CPUs that are local to the GPU's
    PCIe root: the e2e number moves 201 MB per step over that link.  Best effort; returns what it did."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = open(f"{base}/numa_node").read().strip()
        cpus = open(f"{base}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        ids &= allowed
        if ids and ids != allowed:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": cpus, "bound": bool(ids)}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": repr(e)[:80]}


def h2d_bandwidth(device, nbytes=256 << 20):
    """Pinned host -> device copy bandwidth of this box (GB/s, best of 3)."""
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=device)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def make_images(wl, seed):
    """Synthetic content / style images, float32 U[0,255) like toTensor255 yields (utilities.py:11-16), pinned host
    memory.  Integer-valued (8-bit images): floor()."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    B = wl["B"]
    c = (torch.rand(B, 3, 8 * wl["hw"][0], 8 * wl["hw"][1], generator=g) * 255.0).floor()
    s = (torch.rand(wl.get("style_batch", B), 3, 8 * wl["hsws"][0], 8 * wl["hsws"][1], generator=g) * 255.0).floor()
    return c.pin_memory(), s.pin_memory()


def build_models(wl, device):
    """vit_c, vit_s, adaFormer of infer_image.py:51-53, random init in that construction order from one seed
    (SURVEY.md 8d).  Product modules only: nothing from oracle/."""
    import mhada_style_transfer_b200 as M
    from mhada_style_transfer_b200.network import set_precision
    torch.manual_seed(1234)
    vit_c = M.VisionTransformer(num_layers=LAYERS, num_heads=H, hidden_dim=C, pos_embedding=True)
    vit_s = M.VisionTransformer(num_layers=LAYERS, num_heads=H, hidden_dim=C, pos_embedding=False)
    ada = M.AdaAttnTransformerMultiHead(num_layers=LAYERS, qkv_dim=C, num_heads=H)
    vit_c, vit_s, ada = (m.to(device).eval() for m in (vit_c, vit_s, ada))
    bf16 = wl["dtype"] == "bf16"
    for v in (vit_c, vit_s):
        v.out_dtype = "bf16" if bf16 else "fp32"
    return vit_c, vit_s, set_precision(ada, "bf16" if bf16 else "fp32")


def cpu_port_run(wl, images: int, steps: int, warmup: int, budget_s: float):
    """Times oracle/torch_port.py (the reference's CPU path, re-stated) on `images` images per step."""
    from oracle import synth, torch_port
    torch.set_num_threads(os.cpu_count() or 1)
    sd = torch_port.prepare(synth.transformer_state(1234))
    g = torch.Generator().manual_seed(0)
    hw, hs = wl["hw"], wl["hsws"]
    fc = [torch.randn(images, C, hw[0], hw[1], generator=g) * 85 + 1.3 for _ in range(LAYERS)]
    fs = [torch.randn(images, C, hs[0], hs[1], generator=g) * 85 + 1.3 for _ in range(LAYERS)]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            torch_port.transformer(fc, fs, sd)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if sum(times) > budget_s and len(times) >= 1:
                break
    return times


def gpu_eager_port_run(wl, device, steps: int = 3, warmup: int = 1):
    """The reference's eager PyTorch path on THIS GPU (oracle/torch_port.py = the reference's ATen op sequence, fp32,
    PyTorch's default TF32 settings, Nc x Ns map materialised per head): the like-for-like bar SURVEY.md 8(d) asks
    for next to the CPU baseline.  Part of the baseline leg; device-resident inputs, CUDA events."""
    from oracle import synth, torch_port
    sd = {k: v.to(device) for k, v in torch_port.prepare(synth.transformer_state(1234)).items()}
    g = torch.Generator().manual_seed(0)
    hw, hs, B = wl["hw"], wl["hsws"], wl["B"]
    fc = [(torch.randn(B, C, hw[0], hw[1], generator=g) * 85 + 1.3).to(device) for _ in range(LAYERS)]
    fs = [(torch.randn(B, C, hs[0], hs[1], generator=g) * 85 + 1.3).to(device) for _ in range(LAYERS)]
    with torch.no_grad():
        for _ in range(warmup):
            torch_port.transformer(fc, fs, sd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            torch_port.transformer(fc, fs, sd)
        e1.record()
        torch.cuda.synchronize()
    return B * steps / (e0.elapsed_time(e1) * 1e-3)


def run_reference(args, wl, rank):
    if rank != 0:
        return
    images = 1
    times = cpu_port_run(wl, images, args.steps, args.warmup, budget_s=240.0)
    sec = sum(times) / len(times)
    val = images / sec
    # the ViT-inclusive pipeline (what the B200 arm's e2e runs) as an extra: vit_c + vit_s on one image pair
    from oracle import synth, torch_port
    g = torch.Generator().manual_seed(0)
    c = (torch.rand(1, 3, 8 * wl["hw"][0], 8 * wl["hw"][1], generator=g) * 255).floor()
    st = (torch.rand(1, 3, 8 * wl["hsws"][0], 8 * wl["hsws"][1], generator=g) * 255).floor()
    sd_c, sd_s = torch_port.prepare(synth.vit_state(1234, True)), torch_port.prepare(synth.vit_state(1734, False))
    vt = []
    with torch.no_grad():
        for i in range(3):
            t0 = time.perf_counter()
            torch_port.vit(c, sd_c); torch_port.vit(st, sd_s)
            vt.append(time.perf_counter() - t0)
    vit_sec = min(vt[1:])
    line = {
        "impl": "reference", "metric": "images_per_sec", "value": round(val, 4), "unit": "images/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "hot_path": "AdaAttnTransformerMultiHead forward (MHAda x6 + decoder)"},
        "cpu_baseline": {"value": round(val, 4), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{images} image(s) of the workload per step (the reference processes a batch "
                                   "head by head; its img/s does not grow with batch, BASELINE.md §2)"},
        "e2e": {"value": round(val, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "full_pipeline": {"value": round(images / (sec + vit_sec), 4), "unit": "images/s", "vit_c_plus_vit_s_s": round(vit_sec, 3),
                          "note": "ViT x2 + MHAda x6 + decoder per image on these cores = what the B200 arm's e2e runs; "
                                  "`value` keeps the hot path BASELINE.json names (MHAda x6 + decoder), so the driver's "
                                  "e2e ratio UNDERSTATES the B200 arm"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure(args, wl_name, ctx, steps, warmup, with_cpu_baseline):
    """All numbers of one workload on this rank; rank 0 gets the JSON-ready dict."""
    import torch.distributed as dist
    from mhada_style_transfer_b200 import _lib
    wl = WORKLOADS[wl_name]
    rank, world, device, L = ctx["rank"], ctx["world"], ctx["device"], ctx["L"]
    B = wl["B"]
    Nc, Ns = wl["hw"][0] * wl["hw"][1], wl["hsws"][0] * wl["hsws"][1]
    vit_c, vit_s, model = build_models(wl, device)
    c_h, s_h = make_images(wl, seed=rank)
    per_frame_vit = bool(wl.get("cached_style"))     # infer_video.py:91 runs the content ViT frame by frame (b = 1)

    def run_vit_c(c):
        if not per_frame_vit or c.shape[0] == 1:
            return vit_c(c)
        per = [vit_c(c[i:i + 1]) for i in range(c.shape[0])]
        return [torch.cat([p[l] for p in per], dim=0) for l in range(LAYERS)]

    # device-resident ViT feature maps for `value` (made once, outside every timed region)
    with torch.no_grad():
        fc_d = run_vit_c(c_h.to(device))
        fs_d = vit_s(s_h.to(device))
    torch.cuda.synchronize()
    out_dt = fc_d[0].dtype
    gather_buf = None
    if world > 1:
        gather_buf = [torch.empty((B, 3, 8 * wl["hw"][0], 8 * wl["hw"][1]), dtype=out_dt, device=device)
                      for _ in range(world)] if rank == 0 else None

    style = None
    if wl.get("cached_style"):
        with torch.no_grad():
            style = model.precompute_style(fs_d)        # once per style, outside the per-frame steps

    # The only collective on the path is the gather of the decoded images to rank 0.  It runs on a side stream so
    # that step i's gather (NCCL over NVLink) overlaps step i+1's kernels; the timed region ends only after the
    # last gather has completed (the compute stream waits for the side stream before the closing event).
    comm_stream = torch.cuda.Stream(device) if world > 1 else None

    def gather_async(cs):
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ready)
            c = cs.contiguous()
            cs.record_stream(comm_stream)
            dist.gather(c, gather_buf, dst=0)

    def step_device():
        fcs, cs = model(fc_d, style if style is not None else fs_d)
        if world > 1:
            gather_async(cs)
        return cs

    out_shape = (B, 3, 8 * wl["hw"][0], 8 * wl["hw"][1])
    cs_host = [torch.empty(out_shape, dtype=out_dt).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device)
    d2h_stream = torch.cuda.Stream(device)
    # Two preallocated sets of device image buffers (no allocation of inputs inside the timed loop).  Set k is
    # refilled on the copy stream as soon as the step that read it has been consumed.
    host_in = [c_h] if style is not None else [c_h, s_h]
    dev_in = [[torch.empty(t.shape, dtype=t.dtype, device=device) for t in host_in] for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [None, None]
    staged = set()

    def stage(i):
        """H2D of step i's images on the copy stream (pinned host memory -> device set i & 1)."""
        k = i & 1
        with torch.cuda.stream(copy_stream):
            if consumed[k] is not None:
                copy_stream.wait_event(consumed[k])
            for dst, src in zip(dev_in[k], host_in):
                dst.copy_(src, non_blocking=True)
            copied[k].record(copy_stream)
        staged.add(i)

    e2e_state = {"i": 0}

    def step_e2e():
        """One step through the public API with HOST images (infer_image.py:83-86): H2D, vit_c, vit_s, adaFormer,
        D2H.  The copies of step i+1 are issued on a second stream before step i computes (a two-deep pipeline,
        as a frame-streaming caller would run it); every step still moves its own inputs and its own result."""
        i = e2e_state["i"]
        k = i & 1
        if i not in staged:
            stage(i)
        stage(i + 1)
        staged.discard(i)
        cur = torch.cuda.current_stream()
        cur.wait_event(copied[k])
        fc = run_vit_c(dev_in[k][0])
        fs = style if style is not None else vit_s(dev_in[k][1])
        fcs, cs = model(fc, fs)
        consumed[k] = torch.cuda.Event()
        consumed[k].record(cur)
        if world > 1:
            gather_async(cs)
        # D2H of the decoded images on its own stream: on the compute stream the copy (~0.3 ms of PCIe) would
        # hold back the next step's kernels
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(consumed[k])
            cs_host[i & 1].copy_(cs, non_blocking=True)
            cs.record_stream(d2h_stream)
        e2e_state["i"] = i + 1
        return cs

    stage_ms = {}

    def timed(fn, steps, warmup, profile=False):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if profile:
                L.mhada_profile_begin()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            if comm_stream is not None:
                torch.cuda.current_stream().wait_stream(comm_stream)
            torch.cuda.current_stream().wait_stream(d2h_stream)     # the last result has reached the host buffer
            e1.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            attn_ms, attn_n = ctypes.c_float(0), ctypes.c_int(0)
            if profile:
                _lib.check("mhada_profile_end", L.mhada_profile_end(ctypes.byref(attn_ms), ctypes.byref(attn_n)))
                for name, code in (("stats", 0), ("proj", 1), ("linear", 3), ("vit", 4)):
                    sm, sn = ctypes.c_float(0), ctypes.c_int(0)
                    _lib.check("mhada_profile_stage", L.mhada_profile_stage(code, ctypes.byref(sm), ctypes.byref(sn)))
                    stage_ms[name] = (sm.value, sn.value)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, attn_ms.value, attn_n.value

    # count this library's kernel launches in one step of each kind
    with torch.no_grad():
        n0 = L.mhada_total_launch_count()
        step_device()
        launches_per_step = int(L.mhada_total_launch_count() - n0)
    torch.cuda.synchronize()

    sampler = ClockSampler(ctx["local"]) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, attn_ms, attn_n = timed(step_device, steps, warmup, profile=True)
    clocks = sampler.stop() if sampler else None
    stage_dev = dict(stage_ms)
    with torch.no_grad():
        n0 = L.mhada_total_launch_count()
        step_e2e()
        launches_per_e2e_step = int(L.mhada_total_launch_count() - n0)
    ms_e2e, _, _ = timed(step_e2e, steps, max(3, warmup // 2), profile=True)
    vit_ms, vit_n = stage_ms.get("vit", (0.0, 0))

    images = B * world * steps
    value = images / (ms_total * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    esz = 2 if wl["dtype"] == "bf16" else 4
    h2d = sum(t.numel() * t.element_size() for t in host_in)
    d2h = cs_host[0].numel() * cs_host[0].element_size()

    pk = peaks()
    traffic = None            # DRAM bytes per launch of the attention kernel from the committed ncu capture (cfg2 only)
    try:
        prof = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_attn_tc_ncu.json"))
        if prof and wl_name == "cfg2":
            traffic = json.load(open(os.path.join(ROOT, "profiles", prof[-1]))).get("dram_bytes_per_launch")
    except OSError:
        pass
    flops_per_launch = 6.0 * B * Nc * Ns * C
    attn_avg_ms = attn_ms / max(attn_n, 1)
    achieved = flops_per_launch / (attn_avg_ms * 1e-3) / 1e12
    if wl["dtype"] == "bf16":
        # The timed region is steps x a few ms (~0.1 s): the clocks sit at the boost clock (see "clocks"), so the
        # honest denominator is the measured BURST peak, not the sustained one.
        Bs_ = wl.get("style_batch", B)
        roof = {"bound": "tensor", "kernel": "attn_tc_kernel", "achieved": round(achieved, 1),
                "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": round(achieved / pk["bf16_tflops"], 4),
                "frac_of_sustained_peak": round(achieved / pk["bf16_tflops_sustained"], 4) if pk["bf16_tflops_sustained"] else None,
                "peak_sustained": pk["bf16_tflops_sustained"],
                "peak_source": pk["source"] + ", burst (bf16 matmul best of 10)",
                "algorithmic_flops_per_launch": flops_per_launch, "avg_launch_ms": round(attn_avg_ms, 4),
                "launches_timed": attn_n, "share_of_step": round(attn_ms / ms_total, 4), "traffic": traffic,
                "traffic_unit": "bytes/launch (dram read+write, ncu)",
                # read Q, fcs (2 Nc), K, V' (3 Ns) and write the head outputs (Nc), bf16
                "algorithmic_bytes_per_launch": 2.0 * C * (3 * B * Nc + 3 * Bs_ * Ns)}
    else:
        roof = {"bound": "fp32-simt", "kernel": "attn_f32_kernel", "achieved": round(achieved, 2), "peak": None,
                "unit": "TFLOP/s", "frac": None, "avg_launch_ms": round(attn_avg_ms, 4), "launches_timed": attn_n,
                "share_of_step": round(attn_ms / ms_total, 4), "traffic": None,
                "note": "fp32 parity path (FFMA); the roofline claim is made on the bf16 workload"}

    # HBM-bound stages of the six layers, timed by events inside the same steps: algorithmic bytes (every distinct
    # input of a stage read once, every output written once) over the measured device time
    tc_, ts_ = esz * B * C * Nc, esz * (wl.get("style_batch", B)) * C * Ns
    vmul = 2 if wl["dtype"] == "bf16" else 1                      # V' = [V~ | V~^2] on the bf16 path
    # statistics: layer 2i scans fc[i], fs[i] and (i > 0) fcs; layer 2i+1 scans its fc (= fcs), fs statistics are reused
    stats_bytes = (3 * LAYERS - 1) * tc_ + (LAYERS * ts_ if style is None else 0)
    # projections: read fc (+ fs), write Q (+ K and V')
    proj_bytes = 2 * LAYERS * (2 * tc_ + ((2 + vmul) * ts_ if style is None else 0))
    linear_bytes = 2 * LAYERS * 2 * tc_
    kernels = []
    for name, nbytes in (("stats", stats_bytes), ("proj", proj_bytes), ("linear", linear_bytes)):
        ms_s, n_s = stage_dev.get(name, (0.0, 0))
        if n_s:
            per_step_ms = ms_s / steps
            gbs = nbytes / (per_step_ms * 1e-3) / 1e9
            kernels.append({"stage": name, "bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": round(gbs / pk["hbm_gbs"], 4), "ms_per_step": round(per_step_ms, 4),
                            "algorithmic_bytes_per_step": nbytes, "brackets_per_step": n_s // steps})
    if vit_n:
        # the two encoders of an e2e step: GEMM FLOPs 2*M*(K0*D + layers*(D*(3D|D) + D*D + 2*D*F)) per encoder
        def vit_flops(b, n):
            return 2.0 * b * n * (192 * C + LAYERS * (C * (3 * C if b > 1 else C) + C * C + 2 * C * 4 * C))
        fl = (B * vit_flops(1, Nc) if per_frame_vit else vit_flops(B, Nc)) + (0 if style is not None else vit_flops(wl.get("style_batch", B), Ns))
        per_step_ms = vit_ms / steps
        kernels.append({"stage": "vit (e2e steps only: both encoders)", "bound": "tensor",
                        "achieved": round(fl / (per_step_ms * 1e-3) / 1e12, 1), "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": round(fl / (per_step_ms * 1e-3) / 1e12 / pk["bf16_tflops"], 4), "ms_per_step": round(per_step_ms, 4),
                        "algorithmic_flops_per_step": fl, "brackets_per_step": vit_n // steps})

    if rank != 0:
        return None
    line = {
        "metric": "images_per_sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms_total / steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": wl["desc"], "hot_path": "AdaAttnTransformerMultiHead forward (MHAda x6 + decoder)",
                   "images_per_gpu_per_step": B, "tokens": [Nc, Ns], "heads": H,
                   "channels": C, "layers": 2 * LAYERS, "sharding": "by image, one process per GPU, gather of the decoded images to rank 0 (side stream, overlapped)",
                   "value_inputs": "ViT feature maps resident in HBM (3 + 3 maps, made by the B200 ViT from the synthetic images before the timed region)",
                   "e2e_pipeline": "host images -> H2D -> vit_c, vit_s -> MHAda x6 -> decoder -> D2H (infer_image.py:83-86)",
                   "l2": f"features {sum(t.numel() * t.element_size() for t in fc_d + fs_d) / 1e6:.0f} MB/step + {L.mhada_layer_workspace(_lib.BF16 if esz == 2 else _lib.F32, B, Nc, Ns, C, H) / 1e6:.0f} MB workspace exceed the 126 MB L2"},
        "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e / steps, 4),
                "vit_ms_per_step": round(vit_ms / steps, 4), "gpu_launches_per_step": launches_per_e2e_step,
                "inputs": "float32 images 0..255 in pinned host memory", "host_numa": ctx["numa"]},
        "gpu_launches": launches_per_step * steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roof, "kernels": kernels, "clocks": clocks,
    }
    if with_cpu_baseline:
        line["e2e"]["h2d_gbs_this_box"] = round(h2d_bandwidth(device), 1)
        times = cpu_port_run(wl, 1, steps=3, warmup=1, budget_s=25.0)
        sec = sum(times) / len(times)
        line["cpu_baseline"] = {"value": round(1.0 / sec, 4), "unit": "images/s", "cores": torch.get_num_threads(),
                                "kind": "port", "sample": f"{len(times)} x 1 image of the workload (fp32, MHAda x6 + decoder, "
                                "oracle/torch_port.py = the reference's ATen op sequence) after 1 warm-up"}
        try:
            line["cpu_baseline"]["gpu_eager_port"] = {
                "value": round(gpu_eager_port_run(wl, device), 2), "unit": "images/s",
                "note": "same port run eagerly on this GPU in fp32 (PyTorch defaults), 3 steps after 1 warm-up: "
                        "the reference's own GPU path restated, not a bench value"}
        except RuntimeError as e:          # e.g. out of memory for the materialised maps at large sizes
            line["cpu_baseline"]["gpu_eager_port"] = {"unavailable": str(e).splitlines()[0][:120]}
    del fc_d, fs_d, dev_in, cs_host, model, vit_c, vit_s
    torch.cuda.empty_cache()
    return line


def measure_train(args, ctx, steps, warmup):
    """BASELINE configs[4]: one training step per GPU on its own batch (data parallel), gradients averaged with the
    bucketed all-reduce that is launched from autograd hooks DURING backward (sharding.OverlappedGradientAllReduce).
    Forward: the CUDA kernels (bf16) for the six MHAda layers; the ViTs run their Linear layers and batch-axis attention
    on own kernels, forward and backward (MHADA_VIT_TRAIN_IMPL=torch: plain PyTorch); the decoder runs its differentiable
    PyTorch op sequence; backward of the layers = mhada_layer_backward (own kernels; MHADA_BACKWARD_IMPL=torch
    switches to the fp32 PyTorch recompute for an A/B).
    The VGG loss network needs downloaded weights (absent offline): the loss is a synthetic stand-in with the same
    graph shape (pixel loss on cs against the content image + a feature term on fcs)."""
    import torch.distributed as dist
    from mhada_style_transfer_b200.sharding import OverlappedGradientAllReduce
    wl = WORKLOADS["cfg5"]
    rank, world, device = ctx["rank"], ctx["world"], ctx["device"]
    B = wl["B"]
    vit_c, vit_s, model = build_models(wl, device)
    for m in (vit_c, vit_s, model):
        m.train()
    # The whole step (forward, backward, gradient all-reduce, three Adam updates: ~1150 launches) is captured ONCE as a
    # CUDA graph and replayed (MHADA_TRAIN_GRAPH=0: eager launch loop).  Every launch of libmhada_b200.so goes on the
    # caller's stream without allocating or synchronising, so the autograd Functions capture like any PyTorch op.
    use_graph = os.environ.get("MHADA_TRAIN_GRAPH", "1") != "0"
    fused = os.environ.get("MHADA_TRAIN_FUSED_ADAM", "1") != "0"      # same update rule, one multi-tensor kernel per optimiser
    # (fused=None leaves PyTorch's default, the foreach implementation; an explicit False would select the per-tensor loop)
    opts = [torch.optim.Adam(m.parameters(), lr=1e-4, capturable=use_graph, fused=True if fused else None)
            for m in (vit_c, vit_s, model)]                                                                # train_image.py:70-72
    c_h, s_h = make_images(wl, seed=rank)
    c_d, s_d = c_h.to(device), s_h.to(device)
    sync = OverlappedGradientAllReduce([vit_c, vit_s, model]) if world > 1 else None
    ev = {"bwd_end": [], "sync_end": []}

    def step(c, s, record=False):
        for o in opts:
            o.zero_grad(set_to_none=True)
        fc, fs = vit_c(c), vit_s(s)
        fcs, cs = model(fc, fs)
        loss = (cs.float() - c).pow(2).mean() * 1e-4 + fcs.float().pow(2).mean() * 1e-3
        loss.backward()
        if record:
            e = torch.cuda.Event(enable_timing=True); e.record(); ev["bwd_end"].append(e)
        if sync is not None:
            sync.finish()
        if record:
            e = torch.cuda.Event(enable_timing=True); e.record(); ev["sync_end"].append(e)
        for o in opts:
            o.step()
        return loss

    def timed(fn, n, w):
        for _ in range(w):
            fn(False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn(True)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    graph, static_loss, graph_note = None, None, None
    if use_graph:
        # eager steps first: warm-up of every cache / kernel attribute, and the exposed part of the all-reduce (events
        # cannot be timed inside a graph)
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(max(3, warmup)):
                step(c_d, s_d, False)
            for _ in range(5):
                step(c_d, s_d, True)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize()
        tail_ms = sum(a.elapsed_time(b) for a, b in zip(ev["bwd_end"], ev["sync_end"])) / max(len(ev["bwd_end"]), 1)
        ev["bwd_end"].clear(); ev["sync_end"].clear()
        if world > 1:
            dist.barrier()
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = step(c_d, s_d, False)
        except Exception as e:      # noqa: BLE001 -- a benchmark falls back to the eager loop and says so
            graph, graph_note = None, f"capture failed, eager loop timed instead: {str(e).splitlines()[0][:160]}"
            torch.cuda.synchronize()

    def run_step(rec):
        if graph is not None:
            graph.replay()
            return static_loss
        return step(c_d, s_d, rec)

    sampler = ClockSampler(ctx["local"]) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = timed(run_step, steps, warmup)
    clocks = sampler.stop() if sampler else None
    if graph is None:
        tail_ms = sum(a.elapsed_time(b) for a, b in zip(ev["bwd_end"], ev["sync_end"])) / max(len(ev["bwd_end"]), 1)
    ev["bwd_end"].clear(); ev["sync_end"].clear()

    def step_host(rec):
        if graph is not None:
            c_d.copy_(c_h, non_blocking=True)       # the graph reads its inputs from these buffers
            s_d.copy_(s_h, non_blocking=True)
            graph.replay()
            return float(static_loss.detach())
        c = c_h.to(device, non_blocking=True)
        s = s_h.to(device, non_blocking=True)
        loss = step(c, s, rec)
        return float(loss)                      # D2H of the step's result (the loss), synchronises like a logging trainer

    ms_e2e = timed(step_host, steps, max(3, warmup // 2))
    if graph is not None and sync is not None:
        # what the all-reduce costs inside the graph: the same step captured WITHOUT its collectives, timed the same way
        # (after every reported measurement: the ranks' weights drift apart from here on)
        try:
            sync.enabled = False
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                step(c_d, s_d, False)
            ms_noar = timed(lambda rec: g2.replay(), steps, 3)
            tail_ms = (ms_dev - ms_noar) / steps
        except Exception:      # noqa: BLE001
            tail_ms = None
        finally:
            sync.enabled = True
    if rank != 0:
        return None
    images = B * world * steps
    n_params = sum(p.numel() for m in (vit_c, vit_s, model) for p in m.parameters())
    return {
        "metric": "images_per_sec", "value": round(images / (ms_dev * 1e-3), 2), "unit": "images/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms_dev / steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl["desc"], "images_per_gpu_per_step": B, "tokens": [wl["hw"][0] * wl["hw"][1]] * 2,
                   "parameters": n_params, "optimizer": "3 x Adam(lr=1e-4) (train_image.py:70-72)" + (", fused multi-tensor implementation" if fused else ""),
                   "loss": "synthetic (VGG19 weights cannot be downloaded offline): pixel term on cs + feature term on fcs",
                   "forward": ("MHAda layers on the bf16 CUDA kernels; ViT Linear layers + batch attention on own kernels "
                               "(tcgen05 token GEMM); decoder blocks on the inference kernels; LayerNorm / residuals = PyTorch ops"),
                   "backward": ("fp32 recompute of each layer with PyTorch ops (MHADA_BACKWARD_IMPL=torch)"
                                if os.environ.get("MHADA_BACKWARD_IMPL") == "torch" else
                                "mhada_layer_backward: flash-style attention backward kernels (V' = [V | V^2]), the other "
                                "contractions on the tcgen05 token GEMM; ViT Linear / attention backward on own kernels; "
                                "decoder backward = aten convolution / pad / bilinear backward ops (cuDNN) on bf16 channels_last"),
                   "gradient_sync": "bucketed (32 MB) all-reduce launched from autograd hooks during backward, NCCL",
                   "launch": ("the whole step replayed as one CUDA graph" if graph is not None else
                              ("eager launch loop" + (f" ({graph_note})" if graph_note else "")))},
        "e2e": {"value": round(images / (ms_e2e * 1e-3), 2), "unit": "images/s",
                "h2d_bytes_per_step": c_h.numel() * 4 + s_h.numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / steps, 4)},
        "gradient_allreduce": {"exposed_ms_per_step": round(tail_ms, 4) if tail_ms is not None else None, "bytes": 4 * n_params,
                               "note": "device time between the end of backward and the end of the last bucket's "
                                       "all-reduce + copy-back: what the overlap does NOT hide (0 at one GPU); with the step "
                                       "replayed as a CUDA graph: step time minus the time of the same graph captured "
                                       "without its collectives"},
        "gpu_launches": None, "roofline": None, "clocks": clocks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the 1024^2 (cfg3) measurement of the default run")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the MHAda path has no CPU fallback (use --impl reference for the CPU port)")
    import torch.distributed as dist
    from mhada_style_transfer_b200 import _lib
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    L = _lib.lib()
    _lib.check("mhada_device_check", L.mhada_device_check())
    ctx = {"rank": rank, "world": world, "local": local, "device": device, "L": L, "numa": numa}

    if wl.get("train"):
        line = measure_train(args, ctx, args.steps, args.warmup)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    line = measure(args, args.workload, ctx, args.steps, args.warmup, with_cpu_baseline=(world == 1 and not args.no_cpu_baseline))
    if args.workload == "cfg2" and not args.no_also:
        # BASELINE configs[2]: the 1024^2 size the north-star attention target is stated on, in the same run
        also = measure(args, "cfg3", ctx, args.steps, args.warmup, with_cpu_baseline=False)
        if rank == 0:
            line["also"] = {"cfg3": {k: also[k] for k in ("value", "unit", "ms_per_step", "dtype", "e2e", "roofline", "kernels",
                                                           "gpu_launches_per_step")}}
            line["also"]["cfg3"]["workload"] = WORKLOADS["cfg3"]["desc"]
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
