"""Host-side mirror of the reference's MHAda modules (MHAdaSTr/network/adaDecoder.py, conv.py).

Same class names, constructor signatures, attribute names, state_dict keys (318 for the default
AdaAttnTransformerMultiHead), return values and ValueErrors as the reference, so a checkpoint
written by the reference loads with strict=True and `adaFormer(fc, fs)` in infer_image.py /
infer_video.py / infer_time.py keeps working.  The forward does not run PyTorch ops for the hot
path: it hands device pointers to libmhada_b200.so (C ABI, include/mhada_b200.h) on the current
CUDA stream.  PyTorch is used for memory, streams and the decoder convolutions only.

Precision (`module.precision`, default "auto"):
  "fp32"  true-fp32 SIMT kernels (the reference's arithmetic; any head_dim)
  "bf16"  tcgen05 tensor-core kernels, bf16 storage / fp32 accumulate (head_dim 64, 128 or a multiple of 128)
  "auto"  bf16 path when the inputs are bf16/fp16 and the head width allows it, else fp32 path
There is no CPU path: CPU tensors raise.

Training (train_image.py:105-144): the FORWARD always runs the CUDA kernels.  When gradients are required the
layers become autograd Functions.  On the bf16 path at head_dim 64 (the configuration every reference script uses) the
BACKWARD runs own kernels too (mhada_layer_backward, SURVEY N4: flash-style attention backward with V' = [V~ | V~^2],
the other contractions on the tcgen05 token GEMM).  Elsewhere (fp32 path, head_dim != 64, cosine, AdaAttN,
AdaAttnForLoss) it recomputes the layer with PyTorch ops in fp32 and differentiates that.  The decoder, when the model
runs in bf16, runs its inference kernels forward with aten.convolution_backward + an own pad / up-sample backward kernel
behind them (Decoder._forward_train); otherwise the plain differentiable PyTorch path.  Gradients are checked against the
reference's float64 autograd (tests/golden/grad_*).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Sequence

import torch
import torch.nn as nn
from torch.nn import functional as F

from . import _lib

__all__ = ["Softmax", "CosineSimilarity", "AdaAttnForLoss", "AdaAttN", "AdaAttnMultiHead", "AdaAttnTransformer",
           "AdaAttnTransformerMultiHead", "Decoder", "StyleCache", "set_precision"]


# ------------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------------

def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_WORKSPACES: dict = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """One growing scratch buffer per (device, stream): launches on a stream are ordered, so layers
    that run back to back on it can share the buffer."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


def _require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("mhada_style_transfer_b200 runs on a B200 GPU only: got a CPU tensor "
                               "(there is no CPU fallback; use the reference package on CPU)")


def _check_layer_shapes(C: int, fc, fs, fcs):
    """What the reference's view / bmm calls would reject (adaDecoder.py:105-131, :168-198), checked before any
    pointer reaches the C ABI."""
    if fc.dim() != 4 or fs.dim() != 4 or fcs.dim() != 4:
        raise RuntimeError("fc, fs and fcs must be (b, qkv_dim, h, w)")
    if fc.shape[1] != C or fs.shape[1] != C or fcs.shape[1] != C:
        raise RuntimeError(f"expected {C} channels, got {fc.shape[1]}, {fs.shape[1]}, {fcs.shape[1]}")
    if fs.shape[0] != fc.shape[0]:
        # the reference reshapes K/V with the content batch (adaDecoder.py:177-183) and fails the same way
        raise RuntimeError(f"style batch {fs.shape[0]} must equal content batch {fc.shape[0]}")
    if fcs.shape != fc.shape:
        raise RuntimeError("fcs must have the shape of fc")
    if fc.device != fs.device or fc.device != fcs.device:
        raise RuntimeError("fc, fs and fcs must be on the same device")


def _needs_grad(module: nn.Module, *tensors) -> bool:
    return torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or
                                        any(p.requires_grad for p in module.parameters()))


def _no_autograd(module: nn.Module, *tensors):
    if _needs_grad(module, *tensors):
        raise RuntimeError("this entry point of the B200 MHAda path is forward-only: call it under torch.no_grad() "
                           "(the reference inference scripts do, e.g. infer_image.py:82)")


def _instance_norm_torch(x):                       # x [..., N]: per-row statistics over the last axis
    mu = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    return (x - mu) * torch.rsqrt(var + 1e-5)


def _layer_math_torch(fc, fs, fcs, wf, bf, wg, bg, wh, bh, wo, bo, num_heads: int, cosine: bool = False):
    """Differentiable fp32 PyTorch restatement of one layer (adaDecoder.py:162-206), used ONLY by the backward of
    _MhadaLayerFn (recompute-and-differentiate).  matmul / einsum only: no TF32 on the default settings."""
    B, C, h, w = fc.shape
    d = C // num_heads
    xc = fc.reshape(B, num_heads, d, -1)
    xs = fs.reshape(B, num_heads, d, -1)
    xx = fcs.reshape(B, num_heads, d, -1)
    heads = []
    for i in range(num_heads):                     # head by head: the Nc x Ns map of one head at a time
        q = torch.matmul(wf[i], _instance_norm_torch(xc[:, i])) + bf[i][None, :, None]        # [B,d,Nc]
        k = torch.matmul(wg[i], _instance_norm_torch(xs[:, i])) + bg[i][None, :, None]        # [B,d,Ns]
        v = (torch.matmul(wh[i], xs[:, i]) + bh[i][None, :, None]).transpose(1, 2)            # [B,Ns,d]
        logits = torch.matmul(q.transpose(1, 2), k)                                           # [B,Nc,Ns]
        if cosine:                                                                            # adaDecoder.py:29-33
            sim = logits / (q.norm(dim=1).unsqueeze(2) * k.norm(dim=1).unsqueeze(1)) + 1
            a = sim / sim.sum(dim=-1, keepdim=True)
        else:
            a = torch.softmax(logits, dim=-1)
        m = torch.matmul(a, v)
        var = torch.matmul(a, v * v) - m * m
        sd = torch.sqrt(var.clamp(min=1e-6))
        heads.append((sd * _instance_norm_torch(xx[:, i]).transpose(1, 2) + m).transpose(1, 2))   # [B,d,Nc]
    cat = torch.cat(heads, dim=1)                                                             # [B,C,Nc]
    if wo is not None:
        cat = torch.matmul(wo, cat) + bo[None, :, None]
    return cat.reshape(B, C, h, w)


def _layer_backward_kernels(num_heads, fc, fs, fcs, wf, bf, wg, bg, wh, bh, wo, bo, grad_out):
    """Backward of one layer on the B200 kernels (mhada_layer_backward: recomputed forward intermediates, flash-style
    attention backward, every other contraction on the tcgen05 token GEMM).  Returns the eleven gradients."""
    L = _lib.lib()
    dt = torch.bfloat16
    tfc, tfs = _token_major(fc, dt), _token_major(fs, dt)
    tfcs = tfc if _same_tensor(fc, fcs) else _token_major(fcs, dt)
    dout = _token_major(grad_out, dt)
    B, h, w, C = tfc.shape
    Nc, Ns = h * w, tfs.shape[1] * tfs.shape[2]
    dev = tfc.device
    with torch.no_grad():
        w_fgh = torch.stack([wf, wg, wh]).float().contiguous()
        b_fgh = torch.stack([bf, bg, bh]).float().contiguous()
        wo_f, bo_f = wo.float().contiguous(), bo.float().contiguous()
    f32 = dict(dtype=torch.float32, device=dev)
    d_fc, d_fcs = torch.empty((B, h, w, C), **f32), torch.empty((B, h, w, C), **f32)
    d_fs = torch.empty(tfs.shape, **f32)
    d_w_fgh, d_b_fgh = torch.empty_like(w_fgh), torch.empty_like(b_fgh)
    d_wo, d_bo = torch.empty_like(wo_f), torch.empty_like(bo_f)
    ws = _workspace(dev, L.mhada_layer_backward_workspace(B, Nc, Ns, C, num_heads))
    a = _lib.LayerBwdArgs()
    a.B, a.Nc, a.Ns, a.C, a.H = B, Nc, Ns, C, num_heads
    a.fc, a.fs, a.fcs = tfc.data_ptr(), tfs.data_ptr(), tfcs.data_ptr()
    a.w_fgh, a.b_fgh, a.w_out, a.b_out = w_fgh.data_ptr(), b_fgh.data_ptr(), wo_f.data_ptr(), bo_f.data_ptr()
    a.d_out = dout.data_ptr()
    a.d_fc, a.d_fs, a.d_fcs = d_fc.data_ptr(), d_fs.data_ptr(), d_fcs.data_ptr()
    a.d_w_fgh, a.d_b_fgh, a.d_w_out, a.d_b_out = d_w_fgh.data_ptr(), d_b_fgh.data_ptr(), d_wo.data_ptr(), d_bo.data_ptr()
    a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(dev):
        _lib.check("mhada_layer_backward", L.mhada_layer_backward(ctypes.byref(a), _stream()))
    nchw = lambda t, ref: t.permute(0, 3, 1, 2).to(ref.dtype)
    return (nchw(d_fc, fc), nchw(d_fs, fs), nchw(d_fcs, fcs),
            d_w_fgh[0].to(wf.dtype), d_b_fgh[0].to(bf.dtype), d_w_fgh[1].to(wg.dtype), d_b_fgh[1].to(bg.dtype),
            d_w_fgh[2].to(wh.dtype), d_b_fgh[2].to(bh.dtype), d_wo.to(wo.dtype), d_bo.to(bo.dtype))


class _MhadaLayerFn(torch.autograd.Function):
    """Forward: the CUDA kernels.  Backward: mhada_layer_backward (own kernels) when `kernel_bwd`, else recompute with
    _layer_math_torch in fp32 and differentiate that."""

    @staticmethod
    def forward(ctx, run_forward, num_heads, has_out, cosine, kernel_bwd, fc, fs, fcs, wf, bf, wg, bg, wh, bh, wo, bo):
        ctx.num_heads, ctx.has_out, ctx.cosine, ctx.kernel_bwd = num_heads, has_out, cosine, kernel_bwd
        ctx.save_for_backward(fc, fs, fcs, wf, bf, wg, bg, wh, bh, wo, bo)
        with torch.no_grad():
            return run_forward(fc, fs, fcs)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        if ctx.kernel_bwd:
            grads = _layer_backward_kernels(ctx.num_heads, *saved, grad_out)
            need = ctx.needs_input_grad[5:]
            return (None, None, None, None, None, *[g if n else None for g, n in zip(grads, need)])
        with torch.enable_grad():
            leaves = [t.detach().float().requires_grad_(True) for t in saved]
            fc, fs, fcs, wf, bf, wg, bg, wh, bh, wo, bo = leaves
            out = _layer_math_torch(fc, fs, fcs, wf, bf, wg, bg, wh, bh, wo if ctx.has_out else None,
                                    bo if ctx.has_out else None, ctx.num_heads, ctx.cosine)
            need = [i for i, ng in enumerate(ctx.needs_input_grad[5:]) if ng and (ctx.has_out or i < 9)]
            grads = torch.autograd.grad(out, [leaves[i] for i in need], grad_out.float(), allow_unused=True)
        full = [None] * len(leaves)
        for i, g in zip(need, grads):
            full[i] = None if g is None else g.to(saved[i].dtype)
        return (None, None, None, None, None, *full)


def _forloss_math_torch(c_x, s_x, c_1x, s_1x, cosine: bool):
    """Differentiable fp32 PyTorch restatement of AdaAttnForLoss.forward (adaDecoder.py:53-81), used ONLY by the
    backward of _ForLossFn."""
    B, dv, h, w = c_x.shape
    q = _instance_norm_torch(c_1x.reshape(B, c_1x.shape[1], -1)).transpose(1, 2)          # [B,Nc,dqk]
    k = _instance_norm_torch(s_1x.reshape(B, s_1x.shape[1], -1))                          # [B,dqk,Ns]
    v = s_x.reshape(B, dv, -1).transpose(1, 2)                                            # [B,Ns,dv]
    logits = torch.matmul(q, k)
    if cosine:
        sim = logits / (q.norm(dim=2, keepdim=True) * k.norm(dim=1, keepdim=True)) + 1
        a = sim / sim.sum(dim=-1, keepdim=True)
    else:
        a = torch.softmax(logits, dim=-1)
    m = torch.matmul(a, v)
    sd = torch.sqrt((torch.matmul(a, v * v) - m * m).clamp(min=1e-6))
    out = sd * _instance_norm_torch(c_x.reshape(B, dv, -1)).transpose(1, 2) + m
    return out.transpose(1, 2).reshape(B, dv, h, w)


class _ForLossFn(torch.autograd.Function):
    """Forward: the CUDA kernels.  Backward: recompute with _forloss_math_torch in fp32 and differentiate."""

    @staticmethod
    def forward(ctx, run_forward, cosine, c_x, s_x, c_1x, s_1x):
        ctx.cosine = cosine
        ctx.save_for_backward(c_x, s_x, c_1x, s_1x)
        with torch.no_grad():
            return run_forward(c_x, s_x, c_1x, s_1x)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        with torch.enable_grad():
            leaves = [t.detach().float().requires_grad_(True) for t in saved]
            out = _forloss_math_torch(*leaves, ctx.cosine)
            need = [i for i, ng in enumerate(ctx.needs_input_grad[2:]) if ng]
            grads = torch.autograd.grad(out, [leaves[i] for i in need], grad_out.float(), allow_unused=True)
        full = [None] * 4
        for i, g in zip(need, grads):
            full[i] = None if g is None else g.to(saved[i].dtype)
        return (None, None, *full)


def _token_major(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """(B,C,h,w) with any strides -> contiguous (B,h,w,C) of `dtype` (one fused copy, or none when the
    tensor already is channels_last in that dtype, which is what the reference ViT emits)."""
    t = x.permute(0, 2, 3, 1)
    if t.dtype == dtype and t.is_contiguous():
        return t
    out = torch.empty(t.shape, dtype=dtype, device=x.device)
    out.copy_(t)
    return out


def _same_tensor(a: torch.Tensor, b: torch.Tensor) -> bool:
    return a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride()
                      and a.dtype == b.dtype)


def _code(dtype: torch.dtype) -> int:
    return _lib.BF16 if dtype == torch.bfloat16 else _lib.F32


_TC_HEAD_DIMS = (64, 128)      # head dims the STREAMING tcgen05 kernel implements (128: value columns in two slices of 64)


def _tc_head_ok(head_dim: int) -> bool:
    """Streaming kernel for 64 / 128; wider heads (multiples of 128: 1- and 2-head layers, AdaAttN) run per-head
    projections on the token GEMM and the materialised tensor-core attention (forloss_tc.cu)."""
    return head_dim in _TC_HEAD_DIMS or (head_dim > 128 and head_dim % 128 == 0)


def _is_cosine(activation) -> bool:
    return isinstance(activation, CosineSimilarity)


def _resolve_precision(precision: str, head_dim: int, *inputs, activation=None) -> torch.dtype:
    if activation is not None and _is_cosine(activation):
        if precision == "bf16":
            raise NotImplementedError("activation='cosine' runs on the fp32 kernels only; use precision 'auto' or 'fp32'")
        return torch.float32
    return _resolve_precision_softmax(precision, head_dim, *inputs)


def _resolve_precision_softmax(precision: str, head_dim: int, *inputs) -> torch.dtype:
    if precision not in ("auto", "fp32", "bf16"):
        raise ValueError(f"Unknown precision: {precision}")
    if precision == "fp32":
        return torch.float32
    if precision == "bf16":
        if not _tc_head_ok(head_dim):
            raise NotImplementedError(f"the bf16 tensor-core path implements head_dim 64, 128 and multiples of 128, got "
                                      f"{head_dim}; use precision='fp32'")
        return torch.bfloat16
    low = all(t.dtype in (torch.bfloat16, torch.float16) for t in inputs)
    return torch.bfloat16 if (low and _tc_head_ok(head_dim)) else torch.float32


def _layer_forward(dt: torch.dtype, tfc, tfs, tfcs, w_fgh, b_fgh, w_out, b_out, num_heads: int, out=None,
                   flags: int = 0):
    """Token-major tensors in, token-major tensor out: AdaAttnMultiHead.forward, adaDecoder.py:162-206."""
    L = _lib.lib()
    B, h, w, C = tfc.shape
    Nc, Ns = h * w, tfs.shape[1] * tfs.shape[2]
    if out is None:
        out = torch.empty((B, h, w, C), dtype=dt, device=tfc.device)
    code = _code(dt)
    nbytes = L.mhada_layer_workspace(code, B, Nc, Ns, C, num_heads)
    ws = _workspace(tfc.device, nbytes)
    with torch.cuda.device(tfc.device):
        rc = L.mhada_layer_forward(code, _ptr(tfc), _ptr(tfs), _ptr(tfcs), _ptr(w_fgh), _ptr(b_fgh), _ptr(w_out),
                                   _ptr(b_out), B, Nc, Ns, C, num_heads, flags, _ptr(out), _ptr(ws), ws.numel(),
                                   _stream())
    _lib.check("mhada_layer_forward", rc)
    return out


class StyleCache:
    """Style side of the MHAda layers, computed once per style: per layer the K, V (V') and mu_v
    that `mhada_style_precompute` wrote.  They depend on the style features and the layer's g / h weights only,
    so one cache serves every content image / video frame; a style batch of 1 is broadcast over any content batch
    (the reference needs equal batches, adaDecoder.py:177-183, and recomputes all of this per frame,
    infer_video.py:91-92)."""

    def __init__(self, dtype: torch.dtype, style_batch: int, tokens: int, channels: int, buffers, weight_keys):
        self.dtype, self.style_batch, self.tokens, self.channels = dtype, style_batch, tokens, channels
        self.buffers = buffers            # one uint8 device tensor per layer
        self.weight_keys = weight_keys    # parameter versions the buffers were computed from

    def __len__(self):
        return len(self.buffers)


def _style_precompute(dt: torch.dtype, tfs: torch.Tensor, w_fgh, b_fgh, num_heads: int) -> torch.Tensor:
    L = _lib.lib()
    Bs, hs, ws_, C = tfs.shape
    Ns = hs * ws_
    code = _code(dt)
    cache = torch.empty(L.mhada_style_cache_bytes(code, Bs, Ns, C, num_heads), dtype=torch.uint8, device=tfs.device)
    ws = _workspace(tfs.device, L.mhada_layer_workspace(code, Bs, Ns, Ns, C, num_heads))
    with torch.cuda.device(tfs.device):
        rc = L.mhada_style_precompute(code, _ptr(tfs), _ptr(w_fgh), _ptr(b_fgh), Bs, Ns, C, num_heads, _ptr(cache),
                                      cache.numel(), _ptr(ws), ws.numel(), _stream())
    _lib.check("mhada_style_precompute", rc)
    return cache


def _layer_forward_cached(dt, tfc, tfcs, cache: torch.Tensor, Bs: int, Ns: int, w_fgh, b_fgh, w_out, b_out,
                          num_heads: int, flags: int = 0):
    L = _lib.lib()
    B, h, w, C = tfc.shape
    Nc = h * w
    out = torch.empty((B, h, w, C), dtype=dt, device=tfc.device)
    code = _code(dt)
    ws = _workspace(tfc.device, L.mhada_layer_workspace(code, B, Nc, Ns, C, num_heads))
    with torch.cuda.device(tfc.device):
        rc = L.mhada_layer_forward_cached(code, _ptr(tfc), _ptr(tfcs), _ptr(cache), Bs, _ptr(w_fgh), _ptr(b_fgh),
                                          _ptr(w_out), _ptr(b_out), B, Nc, Ns, C, num_heads, flags, _ptr(out),
                                          _ptr(ws), ws.numel(), _stream())
    _lib.check("mhada_layer_forward_cached", rc)
    return out


class _PackedWeights:
    """Packs per-head conv weights into the [3][H][d][d] / [3][H][d] fp32 buffers the ABI takes; re-packs
    only when a parameter was modified (version counter) or moved."""

    def __init__(self):
        self.cache = {}          # device -> (key, tensors): nn.DataParallel replicas share this object (shallow
                                 # __dict__ copy) but each has its own parameters on its own device

    def get(self, groups: Sequence[Sequence[nn.Conv2d]], out_conv):
        params = [p for g in groups for m in g for p in (m.weight, m.bias)]
        if out_conv is not None:
            params += [out_conv.weight, out_conv.bias]
        dev = params[0].device
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in params)
        hit = self.cache.get(dev)
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            w = torch.stack([torch.stack([m.weight.reshape(m.weight.shape[0], -1) for m in g]) for g in groups])
            b = torch.stack([torch.stack([m.bias for m in g]) for g in groups])
            w = w.float().contiguous()
            b = b.float().contiguous()
            wo = out_conv.weight.reshape(out_conv.weight.shape[0], -1).float().contiguous() if out_conv is not None else None
            bo = out_conv.bias.float().contiguous() if out_conv is not None else None
            wo16 = wo.to(torch.bfloat16).contiguous() if wo is not None else None
        self.cache[dev] = (key, (w, b, wo, bo, wo16))
        return self.cache[dev][1]


# ------------------------------------------------------------------------------------------------
# activation markers (adaDecoder.py:11-34).  They carry no parameters; the kernels implement softmax.
# ------------------------------------------------------------------------------------------------

class Softmax(nn.Module):
    """Marker for softmax(bmm(q, k)) -- computed inside the streaming attention kernels."""

    def forward(self, q, k):  # pragma: no cover - the attention map is never materialised on this path
        raise RuntimeError("the N x N attention map is never materialised on the B200 path")


class CosineSimilarity(nn.Module):
    """Marker for the 'cosine' activation (adaDecoder.py:20-34): a = (cos(q, k) + 1) / sum_k (cos(q, k) + 1).
    Runs on the fp32 kernels (attn_f32_kernel<COSINE>); no reference script uses it, so there is no
    tensor-core variant."""

    def forward(self, q, k):  # pragma: no cover - the attention map is never materialised on this path
        raise RuntimeError("the N x N attention map is never materialised on the B200 path")


def _make_activation(activation: str) -> nn.Module:
    if activation == "softmax":
        return Softmax()
    if activation == "cosine":
        return CosineSimilarity()
    raise ValueError(f"Unknown activation function: {activation}")


def _check_activation(m: nn.Module):
    if not isinstance(m, (Softmax, CosineSimilarity)):
        raise ValueError(f"Unknown activation module: {type(m).__name__}")


def _act_flags(m: nn.Module) -> int:
    return _lib.LAYER_COSINE if _is_cosine(m) else 0


# ------------------------------------------------------------------------------------------------
# AdaAttnForLoss (adaDecoder.py:38-81): parameter free, Q/K width != V width
# ------------------------------------------------------------------------------------------------

class AdaAttnForLoss(nn.Module):
    def __init__(self, v_dim, qk_dim, activation="softmax"):
        super().__init__()
        self.norm_q = nn.InstanceNorm2d(qk_dim, affine=False)
        self.norm_k = nn.InstanceNorm2d(qk_dim, affine=False)
        self.norm_v = nn.InstanceNorm2d(v_dim, affine=False)
        self.activation = _make_activation(activation)
        self.precision = "auto"        # "fp32": SIMT kernels (the reference's arithmetic, any widths); "bf16": tensor cores
                                       # (mhada_forloss_forward; softmax, widths multiples of 64); "auto": bf16 for
                                       # bf16 / fp16 inputs when the shape allows it, else fp32

    def forward(self, c_x, s_x, c_1x, s_1x):
        _require_cuda(c_x, s_x, c_1x, s_1x)
        _check_activation(self.activation)
        if _needs_grad(self, c_x, s_x, c_1x, s_1x):
            # train_image.py / lossfn.py:26-34 call this on VGG features that may carry gradients: kernels forward,
            # PyTorch recompute backward (like the layers)
            return _ForLossFn.apply(self._forward_nograd, _is_cosine(self.activation), c_x, s_x, c_1x, s_1x)
        return self._forward_nograd(c_x, s_x, c_1x, s_1x)

    def _tc_ok(self, c_x, c_1x) -> bool:
        return (not _is_cosine(self.activation)) and c_x.shape[1] % 64 == 0 and c_1x.shape[1] % 64 == 0

    def _forward_tc(self, c_x, s_x, c_1x, s_1x):
        """Tensor-core path (forloss_tc.cu): logits of one image materialised, contractions on the token GEMM.  f32
        inputs are handed over as f32 (the kernels normalise first and round / split afterwards)."""
        L = _lib.lib()
        dt = torch.float32 if all(t.dtype == torch.float32 for t in (c_x, s_x, c_1x, s_1x)) else torch.bfloat16
        tq, tk, tv, tx = (_token_major(t, dt) for t in (c_1x, s_1x, s_x, c_x))
        B, h, w, dv = tx.shape
        dqk = tq.shape[3]
        Nc, Ns = tq.shape[1] * tq.shape[2], tk.shape[1] * tk.shape[2]
        if tv.shape[1] * tv.shape[2] != Ns or h * w != Nc or tk.shape[3] != dqk or tv.shape[3] != dv:
            raise RuntimeError("AdaAttnForLoss: inconsistent shapes")
        if not (tq.shape[0] == tk.shape[0] == tv.shape[0] == B):
            raise RuntimeError(f"AdaAttnForLoss: batch sizes differ: c_x {B}, s_x {tv.shape[0]}, c_1x {tq.shape[0]}, "
                               f"s_1x {tk.shape[0]}")
        out = torch.empty((B, h, w, dv), dtype=dt, device=tx.device)
        ws = _workspace(tx.device, L.mhada_forloss_workspace(B, Nc, Ns, dqk, dv))
        a = _lib.ForlossArgs()
        a.dtype = _code(dt)
        a.B, a.Nc, a.Ns, a.dqk, a.dv = B, Nc, Ns, dqk, dv
        a.c_x, a.s_x, a.c_1x, a.s_1x, a.out = tx.data_ptr(), tv.data_ptr(), tq.data_ptr(), tk.data_ptr(), out.data_ptr()
        a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(tx.device):
            _lib.check("mhada_forloss_forward", L.mhada_forloss_forward(ctypes.byref(a), _stream()))
        res = out.permute(0, 3, 1, 2)
        return res if res.dtype == c_x.dtype else res.to(c_x.dtype)

    def _forward_nograd(self, c_x, s_x, c_1x, s_1x):
        if any(t.dim() != 4 for t in (c_x, s_x, c_1x, s_1x)):
            raise RuntimeError("AdaAttnForLoss: inputs must be (b, c, h, w)")
        if self.precision not in ("auto", "fp32", "bf16"):
            raise ValueError(f"Unknown precision: {self.precision}")
        if self.precision == "bf16":
            if not self._tc_ok(c_x, c_1x):
                raise NotImplementedError("AdaAttnForLoss: the tensor-core path needs the softmax activation and channel "
                                          "counts that are multiples of 64; use precision='fp32'")
            return self._forward_tc(c_x, s_x, c_1x, s_1x)
        if self.precision == "auto" and self._tc_ok(c_x, c_1x) and \
                all(t.dtype in (torch.bfloat16, torch.float16) for t in (c_x, s_x, c_1x, s_1x)):
            return self._forward_tc(c_x, s_x, c_1x, s_1x)
        L = _lib.lib()
        dt = torch.float32
        tq, tk, tv, tx = (_token_major(t, dt) for t in (c_1x, s_1x, s_x, c_x))
        B, h, w, dv = tx.shape
        dqk = tq.shape[3]
        Nc, Ns = tq.shape[1] * tq.shape[2], tk.shape[1] * tk.shape[2]
        if tv.shape[1] * tv.shape[2] != Ns or h * w != Nc or tk.shape[3] != dqk or tv.shape[3] != dv:
            raise RuntimeError("AdaAttnForLoss: inconsistent shapes")
        if not (tq.shape[0] == tk.shape[0] == tv.shape[0] == B):
            # the reference's bmm / view calls (adaDecoder.py:55-79) need one batch size for all four inputs
            raise RuntimeError(f"AdaAttnForLoss: batch sizes differ: c_x {B}, s_x {tv.shape[0]}, c_1x {tq.shape[0]}, "
                               f"s_1x {tk.shape[0]}")
        dev = tx.device
        stats = torch.empty((6, B, max(dqk, dv)), dtype=torch.float32, device=dev)
        st = _stream()
        with torch.cuda.device(dev):
            for i, (t, n, c) in enumerate(((tq, Nc, dqk), (tk, Ns, dqk), (tx, Nc, dv))):
                nb = L.mhada_in_stats_workspace(B, n, c)
                ws = _workspace(dev, nb)
                mean = stats[2 * i].view(-1)[: B * c]
                rstd = stats[2 * i + 1].view(-1)[: B * c]
                _lib.check("mhada_in_stats", L.mhada_in_stats(_ptr(t), _lib.F32, B, n, c, c, _ptr(mean), _ptr(rstd),
                                                              _ptr(ws), ws.numel(), st))
            out = torch.empty((B, h, w, dv), dtype=dt, device=dev)
            a = _lib.AttnArgs()
            a.dtype = _lib.F32
            a.B, a.H, a.Nc, a.Ns, a.dqk, a.dv = B, 1, Nc, Ns, dqk, dv
            a.q, a.k, a.v, a.x, a.out = tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), tx.data_ptr(), out.data_ptr()
            a.ldq, a.ldk, a.ldv, a.ldx, a.ldo = dqk, dqk, dv, dv, dv
            a.q_mean, a.q_rstd = stats[0].data_ptr(), stats[1].data_ptr()
            a.k_mean, a.k_rstd = stats[2].data_ptr(), stats[3].data_ptr()
            a.x_mean, a.x_rstd = stats[4].data_ptr(), stats[5].data_ptr()
            a.mu_v = None
            a.activation = _lib.ACT_COSINE if _is_cosine(self.activation) else _lib.ACT_SOFTMAX
            _lib.check("mhada_attn", L.mhada_attn(ctypes.byref(a), st))
        res = out.permute(0, 3, 1, 2)
        return res if res.dtype == c_x.dtype else res.to(c_x.dtype)


# ------------------------------------------------------------------------------------------------
# AdaAttN (adaDecoder.py:85-131): single head, learnable f/g/h, no out_conv
# ------------------------------------------------------------------------------------------------

class AdaAttN(nn.Module):
    def __init__(self, qkv_dim, activation="softmax"):
        super().__init__()
        self.f = nn.Conv2d(qkv_dim, qkv_dim, 1)
        self.g = nn.Conv2d(qkv_dim, qkv_dim, 1)
        self.h = nn.Conv2d(qkv_dim, qkv_dim, 1)
        self.norm_q = nn.InstanceNorm2d(qkv_dim, affine=False)
        self.norm_k = nn.InstanceNorm2d(qkv_dim, affine=False)
        self.norm_v = nn.InstanceNorm2d(qkv_dim, affine=False)
        self.activation = _make_activation(activation)
        self.precision = "auto"
        self._packed = _PackedWeights()

    def forward(self, fc: torch.Tensor, fs: torch.Tensor, fcs: torch.Tensor):
        _check_layer_shapes(self.f.in_channels, fc, fs, fcs)
        _require_cuda(fc, fs, fcs)
        _check_activation(self.activation)
        if _needs_grad(self, fc, fs, fcs):
            d = fc.shape[1]
            params = [t.reshape(1, *shape) for m in (self.f, self.g, self.h)
                      for t, shape in ((m.weight, (d, d)), (m.bias, (d,)))]
            none = fc.new_zeros(0)
            return _MhadaLayerFn.apply(self._forward_nograd, 1, False, _is_cosine(self.activation), False, fc, fs, fcs,
                                       *params, none, none)
        return self._forward_nograd(fc, fs, fcs)

    def _forward_nograd(self, fc, fs, fcs):
        dt = _resolve_precision(self.precision, fc.shape[1], fc, fs, fcs, activation=self.activation)
        w, b = self._packed.get([[self.f], [self.g], [self.h]], None)[:2]
        tfc, tfs = _token_major(fc, dt), _token_major(fs, dt)
        tfcs = tfc if _same_tensor(fc, fcs) else _token_major(fcs, dt)
        out = _layer_forward(dt, tfc, tfs, tfcs, w, b, None, None, 1, flags=_act_flags(self.activation)).permute(0, 3, 1, 2)
        return out if out.dtype == fc.dtype else out.to(fc.dtype)


# ------------------------------------------------------------------------------------------------
# AdaAttnMultiHead (adaDecoder.py:134-206): the hot path
# ------------------------------------------------------------------------------------------------

class AdaAttnMultiHead(nn.Module):
    def __init__(self, qkv_dim, num_heads, activation="softmax"):
        super().__init__()
        if qkv_dim % num_heads != 0:
            raise ValueError("qkv_dim 必須能被 num_heads 整除")   # same message as adaDecoder.py:138
        self.num_heads = num_heads
        self.head_dim = qkv_dim // num_heads
        d = self.head_dim
        self.f_list = nn.ModuleList([nn.Conv2d(d, d, kernel_size=1) for _ in range(num_heads)])
        self.g_list = nn.ModuleList([nn.Conv2d(d, d, kernel_size=1) for _ in range(num_heads)])
        self.h_list = nn.ModuleList([nn.Conv2d(d, d, kernel_size=1) for _ in range(num_heads)])
        # parameter-free; kept so the attribute surface matches the reference
        self.norm_q_list = nn.ModuleList([nn.InstanceNorm2d(d, affine=False) for _ in range(num_heads)])
        self.norm_k_list = nn.ModuleList([nn.InstanceNorm2d(d, affine=False) for _ in range(num_heads)])
        self.norm_v_out_list = nn.ModuleList([nn.InstanceNorm2d(d, affine=False) for _ in range(num_heads)])
        self.out_conv = nn.Conv2d(qkv_dim, qkv_dim, kernel_size=1)
        self.activation = _make_activation(activation)
        self.precision = "auto"
        # "auto": own backward kernels where implemented, else the PyTorch recompute; "kernels" / "torch" force one
        # (tests, A/B timing; MHADA_BACKWARD_IMPL sets the default of new modules)
        self.backward_impl = os.environ.get("MHADA_BACKWARD_IMPL", "auto")
        self._packed = _PackedWeights()

    def packed_weights(self, dt=None):
        """(w_fgh, b_fgh, w_out, b_out) as the ABI takes them; on the bf16 path w_out is the cached bf16 copy
        (MHADA_WOUT_BF16) when the channel count allows it."""
        w, b, wo, bo, wo16 = self._packed.get([self.f_list, self.g_list, self.h_list], self.out_conv)
        if dt == torch.bfloat16 and wo16 is not None and wo16.shape[0] % 128 == 0:
            return w, b, wo16, bo, _lib.WOUT_BF16
        return w, b, wo, bo, 0

    def _check_shapes(self, fc, fs, fcs):
        _check_layer_shapes(self.num_heads * self.head_dim, fc, fs, fcs)

    def forward_tokens(self, dt, tfc, tfs, tfcs, out=None, reuse_fs_stats: bool = False):
        """Token-major entry used by the transformer to chain layers without layout round trips.
        reuse_fs_stats: `tfs` is the tensor the previous layer call on this stream used (same workspace)."""
        w, b, wo, bo, wflag = self.packed_weights(dt)
        return _layer_forward(dt, tfc, tfs, tfcs, w, b, wo, bo, self.num_heads, out,
                              (_lib.REUSE_FS_STATS if reuse_fs_stats else 0) | _act_flags(self.activation) | wflag)

    def precompute_style_tokens(self, dt, tfs) -> torch.Tensor:
        """K, V', mu_v of this layer for token-major style features (one uint8 cache buffer)."""
        w, b = self.packed_weights()[:2]
        return _style_precompute(dt, tfs, w, b, self.num_heads)

    def forward_tokens_cached(self, dt, tfc, tfcs, cache: torch.Tensor, style_batch: int, style_tokens: int):
        w, b, wo, bo, wflag = self.packed_weights(dt)
        return _layer_forward_cached(dt, tfc, tfcs, cache, style_batch, style_tokens, w, b, wo, bo, self.num_heads,
                                     _act_flags(self.activation) | wflag)

    def _stacked_params(self):
        d = self.head_dim
        st = lambda lst, attr, shape: torch.stack([getattr(m, attr).reshape(shape) for m in lst])
        return (st(self.f_list, "weight", (d, d)), st(self.f_list, "bias", (d,)),
                st(self.g_list, "weight", (d, d)), st(self.g_list, "bias", (d,)),
                st(self.h_list, "weight", (d, d)), st(self.h_list, "bias", (d,)),
                self.out_conv.weight.reshape(self.out_conv.weight.shape[0], -1), self.out_conv.bias)

    def forward(self, fc: torch.Tensor, fs: torch.Tensor, fcs: torch.Tensor):
        self._check_shapes(fc, fs, fcs)
        _require_cuda(fc, fs, fcs)
        _check_activation(self.activation)
        if _needs_grad(self, fc, fs, fcs):
            # kernels forward AND backward on the bf16 path at head_dim 64 (mhada_layer_backward); otherwise the
            # PyTorch recompute backward (module docstring); torch.stack keeps the graph to the per-head Conv2d parameters
            return _MhadaLayerFn.apply(self._forward_nograd, self.num_heads, True, _is_cosine(self.activation),
                                       self._kernel_backward(fc, fs, fcs), fc, fs, fcs, *self._stacked_params())
        return self._forward_nograd(fc, fs, fcs)

    def _kernel_backward(self, fc, fs, fcs) -> bool:
        """mhada_layer_backward covers the bf16 path with head_dim 64, softmax, C a multiple of 128."""
        if self.backward_impl not in ("auto", "kernels", "torch"):
            raise ValueError(f"Unknown backward_impl: {self.backward_impl}")
        ok = (not _is_cosine(self.activation) and self.head_dim == 64 and (self.num_heads * 64) % 128 == 0 and
              _resolve_precision(self.precision, self.head_dim, fc, fs, fcs, activation=self.activation) == torch.bfloat16)
        if self.backward_impl == "kernels" and not ok:
            raise NotImplementedError("backward_impl='kernels' needs the bf16 path, softmax, head_dim 64 and C % 128 == 0")
        return ok and self.backward_impl != "torch"

    def _forward_nograd(self, fc, fs, fcs):
        dt = _resolve_precision(self.precision, self.head_dim, fc, fs, fcs, activation=self.activation)
        tfc, tfs = _token_major(fc, dt), _token_major(fs, dt)
        tfcs = tfc if _same_tensor(fc, fcs) else _token_major(fcs, dt)
        out = self.forward_tokens(dt, tfc, tfs, tfcs).permute(0, 3, 1, 2)
        return out if out.dtype == fc.dtype else out.to(fc.dtype)


# ------------------------------------------------------------------------------------------------
# Decoder (conv.py:23-100): boundary neighbour, runs on cuDNN through PyTorch (channels_last)
# ------------------------------------------------------------------------------------------------

class _PadConv(nn.Module):
    """ReflectionPad2d(k//2) + Conv2d (the reference's `Conv`, conv.py:23-33): keys `conv.weight/bias`."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int):
        super().__init__()
        self.pad = nn.ReflectionPad2d(kernel_size // 2)
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride)
        self._shadow = {}

    def _weights(self, dtype):
        if dtype == self.conv.weight.dtype:
            return self.conv.weight, self.conv.bias
        key = (dtype, self.conv.weight._version, self.conv.bias._version, self.conv.weight.data_ptr())
        if self._shadow.get("key") != key:
            self._shadow = {"key": key,
                            "w": self.conv.weight.detach().to(dtype).contiguous(memory_format=torch.channels_last),
                            "b": self.conv.bias.detach().to(dtype)}
        return self._shadow["w"], self._shadow["b"]

    def _packed_tc(self):
        """bf16 [Cout][3][3][Cin] weights + f32 bias for mhada_conv3x3 (K ordered (ky, kx, ci)); cached per device."""
        w, b = self.conv.weight, self.conv.bias
        key = ("tc", w._version, b._version, w.data_ptr(), b.data_ptr(), w.dtype, str(w.device))
        if self._shadow.get("tc_key") != key:
            with torch.no_grad():
                self._shadow["tc_w"] = w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()
                self._shadow["tc_b"] = b.detach().float().contiguous()
            self._shadow["tc_key"] = key
        return self._shadow["tc_w"], self._shadow["tc_b"]

    def forward(self, x):
        w, b = self._weights(x.dtype)
        return F.conv2d(self.pad(x), w, b, self.conv.stride)


class _ConvBlock(nn.Module):
    """ConvReLU / ConvReluInterpolate (conv.py:36-45, :61-72): key `conv.conv.*`."""

    def __init__(self, in_channels: int, out_channels: int, scale_factor: float = 0.0):
        super().__init__()
        self.conv = _PadConv(in_channels, out_channels, 3, 1)
        self.relu = nn.ReLU()
        self.scale_factor = scale_factor

    def forward(self, x):
        x = self.relu(self.conv(x))
        if self.scale_factor:
            x = F.interpolate(x, scale_factor=self.scale_factor, mode="bilinear", align_corners=False)
        return x


def _pad_reflect(x_tok: torch.Tensor, upsample: bool) -> torch.Tensor:
    """[B,H,W,C] contiguous -> [B,Ho+2,Wo+2,C]: ReflectionPad2d(1) (+ x2 bilinear first) in one kernel."""
    L = _lib.lib()
    B, H, W, C = x_tok.shape
    Ho, Wo = (2 * H, 2 * W) if upsample else (H, W)
    y = torch.empty((B, Ho + 2, Wo + 2, C), dtype=x_tok.dtype, device=x_tok.device)
    with torch.cuda.device(x_tok.device):
        rc = L.mhada_pad_reflect(_code(x_tok.dtype), _ptr(x_tok), B, H, W, C, 1 if upsample else 0, _ptr(y), _stream())
    _lib.check("mhada_pad_reflect", rc)
    return y


def _conv3x3_small_relu(x_tok: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """[B,H,W,64] bf16 -> [B,Cout,H,W] bf16: ReflectionPad2d(1) + 3x3 conv + bias + ReLU (mhada_conv3x3_small)."""
    L = _lib.lib()
    B, H, W, C = x_tok.shape
    cout = w.shape[0]
    wf, bf = w.detach().float().contiguous(), b.detach().float().contiguous()
    y = torch.empty((B, cout, H, W), dtype=torch.bfloat16, device=x_tok.device)
    with torch.cuda.device(x_tok.device):
        rc = L.mhada_conv3x3_small(_lib.BF16, _ptr(x_tok), _ptr(wf), _ptr(bf), B, H, W, C, cout, 1, _ptr(y), _stream())
    _lib.check("mhada_conv3x3_small", rc)
    return y


def _conv3x3_tc_relu(xp_tok: torch.Tensor, w_packed: torch.Tensor, b: torch.Tensor, out_padded: bool) -> torch.Tensor:
    """[B,H+2,W+2,Cin] bf16 reflect-padded -> conv3x3 + bias + ReLU on tcgen05 (mhada_conv3x3): [B,H,W,Cout], or
    [B,H+2,W+2,Cout] with the reflection ring already written when `out_padded` (= the next block's input)."""
    L = _lib.lib()
    B, Hp, Wp, Cin = xp_tok.shape
    H, W, Cout = Hp - 2, Wp - 2, w_packed.shape[0]
    shape = (B, Hp, Wp, Cout) if out_padded else (B, H, W, Cout)
    y = torch.empty(shape, dtype=torch.bfloat16, device=xp_tok.device)
    with torch.cuda.device(xp_tok.device):
        rc = L.mhada_conv3x3(_lib.BF16, _ptr(xp_tok), _ptr(w_packed), _ptr(b), B, H, W, Cin, Cout, 1, 1 if out_padded else 0,
                             _ptr(y), _stream())
    _lib.check("mhada_conv3x3", rc)
    return y


def _conv3x3_relu(xp_nchw: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """3x3 conv (input already padded) + bias + ReLU as ONE cuDNN fused conv-bias-relu call (fp32 path of the decoder;
    the bf16 path runs mhada_conv3x3).  No fallback: if cuDNN rejects the shape the error propagates."""
    return torch.cudnn_convolution_relu(xp_nchw, w, b, (1, 1), (0, 0), (1, 1), 1)


# ---- training path of the decoder: own kernels forward, library backward -----------------------------------------------
# train_image.py:105-144 back-propagates through the decoder.  The forward of every block runs the same kernels as
# inference (bf16, channels_last); the backward of the convolutions is aten.convolution_backward on those bf16
# channels_last tensors (cuDNN tensor-core dgrad / wgrad, no NCHW <-> NHWC conversion: the r2 profile of the plain PyTorch
# path had 1.5 ms of layout kernels and TF32 convolutions per cfg5 step), reflect-pad (+ bilinear) backward is
# pad_reflect_bwd_kernel.  Own convolution-backward kernels are not built.  bf16 activations through nine ReLU blocks
# move the gradients of the EARLY blocks by ~10 % against fp32 arithmetic -- exactly as much as PyTorch's own bf16
# decoder does (tools/debug_decoder_train.py) -- so the path is taken only when the model runs in bf16.

class _PadUpFn(torch.autograd.Function):
    """[B,H,W,C] bf16 -> [B,Ho+2,Wo+2,C]: (x2 bilinear +) ReflectionPad2d(1) by pad_reflect_kernel."""

    @staticmethod
    def forward(ctx, x_tok, upsample):
        ctx.upsample, ctx.in_shape = upsample, x_tok.shape
        return _pad_reflect(x_tok.contiguous(), upsample)

    @staticmethod
    def backward(ctx, g):
        B, H, W, C = ctx.in_shape
        g = g.contiguous()
        dx = torch.empty((B, H, W, C), dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            rc = _lib.lib().mhada_pad_reflect_bwd(_code(g.dtype), _ptr(g), B, H, W, C, 1 if ctx.upsample else 0, _ptr(dx), _stream())
        _lib.check("mhada_pad_reflect_bwd", rc)
        return dx, None


class _ConvReluFn(torch.autograd.Function):
    """Reflect-padded [B,H+2,W+2,Cin] bf16 -> conv3x3 + bias + ReLU on tcgen05 (mhada_conv3x3) -> [B,H,W,Cout] bf16."""

    @staticmethod
    def forward(ctx, xp_tok, weight, bias, block):
        w_packed, b32 = block._packed_tc()
        y = _conv3x3_tc_relu(xp_tok, w_packed, b32, False)
        ctx.save_for_backward(xp_tok, y, weight)
        return y

    @staticmethod
    def backward(ctx, g):
        xp_tok, y, weight = ctx.saved_tensors
        gz = (g * (y > 0)).permute(0, 3, 1, 2)                            # ReLU backward; NCHW view, channels_last memory
        w16 = weight.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        gxp, gw, gb = torch.ops.aten.convolution_backward(gz, xp_tok.permute(0, 3, 1, 2), w16, [weight.shape[0]], [1, 1], [0, 0],
                                                          [1, 1], False, [0, 0], 1, [True, True, True])
        return gxp.permute(0, 2, 3, 1).contiguous(), gw.to(weight.dtype), gb.to(weight.dtype), None


class _ConvSmallReluFn(torch.autograd.Function):
    """Last block (64 -> 3): UNPADDED [B,H,W,64] bf16 -> [B,3,H,W] bf16 by conv3x3_c64_small_kernel (pad resolved inside)."""

    @staticmethod
    def forward(ctx, x_tok, weight, bias):
        y = _conv3x3_small_relu(x_tok.contiguous(), weight, bias)
        ctx.save_for_backward(x_tok, y, weight)
        return y

    @staticmethod
    def backward(ctx, g):
        x_tok, y, weight = ctx.saved_tensors
        gz = (g * (y > 0)).contiguous(memory_format=torch.channels_last)
        xp = _pad_reflect(x_tok.contiguous(), False).permute(0, 3, 1, 2)
        w16 = weight.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        gxp, gw, gb = torch.ops.aten.convolution_backward(gz, xp, w16, [weight.shape[0]], [1, 1], [0, 0], [1, 1], False, [0, 0], 1,
                                                          [True, True, True])
        B, H, W, C = x_tok.shape
        gxp = gxp.permute(0, 2, 3, 1).contiguous()
        gx = torch.empty((B, H, W, C), dtype=gxp.dtype, device=gxp.device)
        with torch.cuda.device(gxp.device):
            rc = _lib.lib().mhada_pad_reflect_bwd(_code(gxp.dtype), _ptr(gxp), B, H, W, C, 0, _ptr(gx), _stream())
        _lib.check("mhada_pad_reflect_bwd", rc)
        return gx, gw.to(weight.dtype), gb.to(weight.dtype)


class Decoder(nn.Module):
    """Decoder.forward (conv.py:75-100).  bf16 path: own kernels only -- tcgen05 implicit-GEMM convolutions whose
    epilogue writes the next block's reflect-padded input, the fused x2-bilinear + pad kernel after the three
    up-sampling blocks, and the small 64 -> 3 kernel; activations stay channels_last.  fp32 path (the reference's
    arithmetic): the pad kernel + cuDNN fp32 convolutions.
    CPU tensors raise (no CPU path).  Under autograd: the same kernels forward with a library convolution backward and an
    own pad / up-sample backward kernel when the model runs in bf16 (_forward_train), else the plain differentiable
    PyTorch ops on the GPU."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Sequential(_ConvBlock(512, 256, 2), _ConvBlock(256, 256), _ConvBlock(256, 256),
                                   _ConvBlock(256, 256), _ConvBlock(256, 128, 2))
        self.conv2 = nn.Sequential(_ConvBlock(128, 128), _ConvBlock(128, 64, 2))
        self.conv3 = nn.Sequential(_ConvBlock(64, 64), _ConvBlock(64, 3))
        # training: "kernels" = own kernels forward + aten backward ops (bf16 activations); "torch" = the plain PyTorch op
        # sequence in the input's dtype; "auto" = kernels when the model runs in bf16 (precision "bf16", as
        # set_precision(model, "bf16") sets it, or bf16 / fp16 features), else torch (MHADA_DECODER_TRAIN_IMPL sets the default)
        self.train_impl = os.environ.get("MHADA_DECODER_TRAIN_IMPL", "auto")
        self.precision = "auto"

    def _blocks(self):
        return [*self.conv1, *self.conv2, *self.conv3]

    def _forward_train(self, fcs: torch.Tensor):
        """Training (train_image.py:105-144): own kernels forward (bf16, channels_last), aten backward ops."""
        x = fcs.permute(0, 2, 3, 1).to(torch.bfloat16)      # differentiable; no copy when fcs is bf16 channels_last already
        blocks = self._blocks()
        up = False
        for i, blk in enumerate(blocks):
            conv = blk.conv.conv
            if conv.in_channels == 64 and conv.out_channels <= 8 and not blk.scale_factor and i == len(blocks) - 1 and not up:
                y = _ConvSmallReluFn.apply(x, conv.weight, conv.bias)               # [B, 3, H, W]
                return y if y.dtype == fcs.dtype else y.to(fcs.dtype)
            xp = _PadUpFn.apply(x, up)
            x = _ConvReluFn.apply(xp, conv.weight, conv.bias, blk.conv)
            up = bool(blk.scale_factor)
        y = x.permute(0, 3, 1, 2)
        return y if y.dtype == fcs.dtype else y.to(fcs.dtype)

    def forward(self, fcs: torch.Tensor):
        _require_cuda(fcs)                                   # no CPU path (use the reference package on CPU)
        if _needs_grad(self, fcs):
            if self.train_impl not in ("auto", "kernels", "torch"):
                raise ValueError(f"Unknown train_impl: {self.train_impl}")
            tc_ok = fcs.shape[2] >= 2 and fcs.shape[3] >= 2 and all(
                b.conv.conv.in_channels % 64 == 0 and (b.conv.conv.out_channels in (64, 128, 256) or i == 8)
                for i, b in enumerate(self._blocks()))
            low = self.precision == "bf16" or fcs.dtype in (torch.bfloat16, torch.float16)
            if tc_ok and (self.train_impl == "kernels" or (self.train_impl == "auto" and low)):
                return self._forward_train(fcs)
            # the plain differentiable PyTorch ops of conv.py:96-100 on the GPU
            return self.conv3(self.conv2(self.conv1(fcs)))
        in_dtype = fcs.dtype
        if fcs.dtype not in (torch.float32, torch.bfloat16):
            fcs = fcs.to(torch.bfloat16)
        if fcs.shape[2] < 2 or fcs.shape[3] < 2:
            raise RuntimeError("Decoder needs at least 2x2 feature maps (ReflectionPad2d(1))")
        x = _token_major(fcs, fcs.dtype)                    # [B,h,w,512]
        blocks = self._blocks()
        if x.dtype == torch.bfloat16:
            # bf16 path, own kernels only: blocks 0..7 = tcgen05 implicit GEMM (mhada_conv3x3) on reflect-padded input;
            # a block whose successor is a plain convolution writes its output already padded (no pad pass between
            # them), a block followed by the x2 up-sample hands it to the fused up-sample + pad kernel; the last block
            # (64 -> 3) is the HBM-bound small kernel on unpadded input.
            xp = _pad_reflect(x, False)                      # conv.py:26-27
            for i, blk in enumerate(blocks):
                cin, cout = blk.conv.conv.in_channels, blk.conv.conv.out_channels
                if cin == 64 and cout <= 8 and not blk.scale_factor:
                    if xp is not None:
                        raise RuntimeError("decoder: the 64 -> %d block expects an unpadded input" % cout)
                    y = _conv3x3_small_relu(x, blk.conv.conv.weight, blk.conv.conv.bias)     # conv.py:90-93
                    if i != len(blocks) - 1:
                        raise RuntimeError("decoder: the small block must be the last one")
                    return y if y.dtype == in_dtype else y.to(in_dtype)
                nxt = blocks[i + 1] if i + 1 < len(blocks) else None
                nxt_small = nxt is not None and nxt.conv.conv.in_channels == 64 and nxt.conv.conv.out_channels <= 8
                out_padded = nxt is not None and not blk.scale_factor and not nxt_small
                w, b = blk.conv._packed_tc()
                y = _conv3x3_tc_relu(xp, w, b, out_padded)
                if out_padded:
                    xp = y
                elif blk.scale_factor:
                    xp = _pad_reflect(y, True)               # conv.py:71 + the next block's pad, one pass
                else:
                    x, xp = y, None
            y = x.permute(0, 3, 1, 2)
            return y if y.dtype == in_dtype else y.to(in_dtype)
        up = False
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):     # fp32 path: the reference's fp32 arithmetic
            for blk in blocks:                               # (cuDNN convolutions, never TF32), pads by pad_reflect_kernel
                xp = _pad_reflect(x, up)                     # conv.py:26-27 (+ :71 of the previous block)
                w, b = blk.conv._weights(x.dtype)
                y = _conv3x3_relu(xp.permute(0, 3, 1, 2), w, b)
                x = _token_major(y, y.dtype)                 # no copy when cuDNN answered in channels_last
                up = bool(blk.scale_factor)
        y = x.permute(0, 3, 1, 2)
        return y if y.dtype == in_dtype else y.to(in_dtype)       # fp16 in -> fp16 out like the reference


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ------------------------------------------------------------------------------------------------
# Transformers (adaDecoder.py:209-268)
# ------------------------------------------------------------------------------------------------

class AdaAttnTransformer(nn.Module):
    def __init__(self, num_layers: int = 3, qkv_dim: int = 512, activation: str = "softmax"):
        super().__init__()
        self.num_layers = num_layers
        self.adaAttNs = nn.ModuleList([AdaAttN(qkv_dim, activation) for _ in range(num_layers)])
        self.decoder = Decoder()

    def forward(self, fc: List[torch.Tensor], fs: List[torch.Tensor]) -> torch.Tensor:
        fcs = fc[0]
        for i in range(self.num_layers):
            fcs = self.adaAttNs[i](fc[i], fs[i], fcs)
        return self.decoder(fcs)


class AdaAttnTransformerMultiHead(nn.Module):
    def __init__(self, num_layers: int = 3, qkv_dim: int = 512, num_heads: int = 8, activation: str = "softmax"):
        super().__init__()
        self.num_layers = num_layers
        self.adaAttnHead = nn.ModuleList(
            [AdaAttnMultiHead(qkv_dim=qkv_dim, num_heads=num_heads, activation=activation)
             for _ in range(num_layers * 2)])
        self.decoder = Decoder()
        self.precision = "auto"

    def _weight_keys(self):
        # version counter alone misses module.to() / .half() / load_state_dict(assign=True): those rebind storage
        return tuple((p._version, p.data_ptr(), p.dtype, str(p.device)) for m in self.adaAttnHead for p in m.parameters())

    def precompute_style(self, fs, precision: str = None) -> StyleCache:
        """Extension: run the style side of all 2*num_layers layers once.  `fs` is the list of
        style feature maps `vit_s(s)` (batch 1 or the content batch).  Pass the result instead of `fs`:
            cache = adaFormer.precompute_style(fs);  fcs, cs = adaFormer(fc, cache)      # every frame"""
        _require_cuda(*fs)
        _no_autograd(self, *fs)
        L0 = self.adaAttnHead[0]
        _check_activation(L0.activation)
        dt = _resolve_precision(precision or self.precision, L0.head_dim, *fs[: self.num_layers], activation=L0.activation)
        if L0.head_dim not in _TC_HEAD_DIMS:
            dt = torch.float32           # the style cache holds the streaming kernels' K / V' layout: wide heads cache in fp32
        bufs = []
        Bs, C, hs, ws_ = fs[0].shape
        for i in range(self.num_layers):
            if fs[i].shape != fs[0].shape:
                raise RuntimeError("all style feature maps must have the same shape")
            tfs = _token_major(fs[i], dt)
            bufs.append(self.adaAttnHead[2 * i].precompute_style_tokens(dt, tfs))
            bufs.append(self.adaAttnHead[2 * i + 1].precompute_style_tokens(dt, tfs))
        return StyleCache(dt, Bs, hs * ws_, C, bufs, self._weight_keys())

    def _forward_cached(self, fc, cache: StyleCache):
        L0 = self.adaAttnHead[0]
        if len(cache) != 2 * self.num_layers or cache.channels != L0.num_heads * L0.head_dim:
            raise RuntimeError("StyleCache was built for a different model")
        if cache.weight_keys != self._weight_keys():
            raise RuntimeError("StyleCache is stale: the model parameters changed since precompute_style()")
        _require_cuda(*fc)
        _no_autograd(self, *fc)
        B = fc[0].shape[0]
        if cache.style_batch not in (1, B):
            raise RuntimeError(f"style batch {cache.style_batch} must be 1 or the content batch {B}")
        for i in range(self.num_layers):
            if fc[i].dim() != 4 or fc[i].shape != fc[0].shape or fc[i].shape[1] != cache.channels:
                raise RuntimeError(f"fc[{i}] must be (b, {cache.channels}, h, w) with the shape of fc[0], got "
                                   f"{tuple(fc[i].shape)}")
            if fc[i].device != cache.buffers[0].device:
                raise RuntimeError("content features and StyleCache are on different devices")
        in_dtype, dt = fc[0].dtype, cache.dtype
        tfc = [_token_major(t, dt) for t in fc[: self.num_layers]]
        fcs = tfc[0]
        for i in range(self.num_layers):
            fcs = self.adaAttnHead[2 * i].forward_tokens_cached(dt, tfc[i], fcs, cache.buffers[2 * i], cache.style_batch,
                                                                cache.tokens)
            fcs = self.adaAttnHead[2 * i + 1].forward_tokens_cached(dt, fcs, fcs, cache.buffers[2 * i + 1],
                                                                    cache.style_batch, cache.tokens)
        fcs = fcs.permute(0, 3, 1, 2)
        cs = self.decoder(fcs)
        if dt != in_dtype:
            fcs, cs = fcs.to(in_dtype), cs.to(in_dtype)
        return fcs, cs

    def forward(self, *args):
        # model(fc_list, fs_list) or model((fc_list, fs_list)) -- adaDecoder.py:253-260
        if len(args) == 1:
            fc, fs = args[0]
        else:
            fc, fs = args
        if isinstance(fs, StyleCache):
            return self._forward_cached(fc, fs)
        L0 = self.adaAttnHead[0]
        for i in range(self.num_layers):
            L0._check_shapes(fc[i], fs[i], fc[0])
        _require_cuda(*fc, *fs)
        _check_activation(L0.activation)
        if _needs_grad(self, *fc, *fs):
            fcs = fc[0]                                                          # adaDecoder.py:262-267, layer by layer
            for i in range(self.num_layers):
                fcs = self.adaAttnHead[2 * i](fc[i], fs[i], fcs)
                fcs = self.adaAttnHead[2 * i + 1](fcs, fs[i], fcs)
            return fcs, self.decoder(fcs)
        in_dtype = fc[0].dtype
        dt = _resolve_precision(self.precision, L0.head_dim, *fc[: self.num_layers], *fs[: self.num_layers],
                                activation=L0.activation)
        tfc = [_token_major(t, dt) for t in fc[: self.num_layers]]
        tfs = [_token_major(t, dt) for t in fs[: self.num_layers]]
        fcs = tfc[0]                                                         # :262
        for i in range(self.num_layers):                                     # :263-265
            fcs = self.adaAttnHead[2 * i].forward_tokens(dt, tfc[i], tfs[i], fcs)
            fcs = self.adaAttnHead[2 * i + 1].forward_tokens(dt, fcs, tfs[i], fcs, reuse_fs_stats=True)
        fcs = fcs.permute(0, 3, 1, 2)            # (B,C,h,w) view over channels_last memory
        cs = self.decoder(fcs)                   # :267
        if dt != in_dtype:
            fcs, cs = fcs.to(in_dtype), cs.to(in_dtype)
        return fcs, cs


def set_precision(module: nn.Module, precision: str) -> nn.Module:
    """Set `.precision` on a module tree ("auto" | "fp32" | "bf16")."""
    if precision not in ("auto", "fp32", "bf16"):
        raise ValueError(f"Unknown precision: {precision}")
    for m in module.modules():
        if hasattr(m, "precision"):
            m.precision = precision
    return module
