"""PyTorch-CPU port of the reference's MHAda forward, for TIMING the reference's CPU path on the GPU
box's host cores (bench.py: cpu_baseline and --impl reference), where /root/reference does not exist.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.

It issues the same ATen operator sequence per head as the reference (instance_norm -> 1x1 conv ->
bmm -> softmax -> bmm x2 -> elementwise -> cat -> 1x1 conv; adaDecoder.py:162-206, decoder
conv.py:96-100), in fp32, materialising the Nc x Ns map like the reference does, so its wall time
on N threads is a fair stand-in for the unmodified module (kind = "port").  Numerically it is
checked against the same golden vectors as the numpy oracle (tests/test_oracle.py).
"""
from __future__ import annotations

import torch
from torch.nn import functional as F


def _t(sd, key, dtype):
    v = sd[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(v).to(dtype)


def prepare(sd: dict, dtype=torch.float32) -> dict:
    """numpy / torch state dict -> torch tensors of `dtype` (done once, outside any timed region)."""
    return {k: (v.to(dtype) if isinstance(v, torch.Tensor) else torch.from_numpy(v).to(dtype)) for k, v in sd.items()}


def mhada_layer(fc, fs, fcs, sd: dict, prefix: str, num_heads: int):
    """AdaAttnMultiHead.forward (adaDecoder.py:162-206) with functional ops."""
    b, c, h, w = fc.shape
    d = c // num_heads
    hs, ws = fs.shape[2:]
    outs = []
    for i in range(num_heads):
        sl = slice(i * d, (i + 1) * d)
        q = F.conv2d(F.instance_norm(fc[:, sl]), sd[f"{prefix}f_list.{i}.weight"], sd[f"{prefix}f_list.{i}.bias"])
        k = F.conv2d(F.instance_norm(fs[:, sl]), sd[f"{prefix}g_list.{i}.weight"], sd[f"{prefix}g_list.{i}.bias"])
        v = F.conv2d(fs[:, sl], sd[f"{prefix}h_list.{i}.weight"], sd[f"{prefix}h_list.{i}.bias"])
        q = q.reshape(b, d, h * w).permute(0, 2, 1)
        k = k.reshape(b, d, hs * ws)
        v = v.reshape(b, d, hs * ws).permute(0, 2, 1)
        a = torch.softmax(torch.bmm(q, k), dim=-1)
        m = torch.bmm(a, v)
        var = torch.bmm(a, v ** 2) - m ** 2
        s = torch.sqrt(var.clamp(min=1e-6))
        m = m.reshape(b, h, w, d).permute(0, 3, 1, 2)
        s = s.reshape(b, h, w, d).permute(0, 3, 1, 2)
        outs.append(s * F.instance_norm(fcs[:, sl]) + m)
    cat = torch.cat(outs, dim=1)
    return F.conv2d(cat, sd[f"{prefix}out_conv.weight"], sd[f"{prefix}out_conv.bias"])


_PLAN = (("conv1.0", True), ("conv1.1", False), ("conv1.2", False), ("conv1.3", False), ("conv1.4", True),
         ("conv2.0", False), ("conv2.1", True), ("conv3.0", False), ("conv3.1", False))


def decoder(x, sd: dict, prefix: str = "decoder."):
    """Decoder.forward (conv.py:96-100)."""
    for stem, up in _PLAN:
        x = F.relu(F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[f"{prefix}{stem}.conv.conv.weight"],
                            sd[f"{prefix}{stem}.conv.conv.bias"]))
        if up:
            x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return x


def transformer(fc, fs, sd: dict, num_layers: int = 3, num_heads: int = 8, decode: bool = True):
    """AdaAttnTransformerMultiHead.forward (adaDecoder.py:253-268)."""
    fcs = fc[0]
    for i in range(num_layers):
        fcs = mhada_layer(fc[i], fs[i], fcs, sd, f"adaAttnHead.{2 * i}.", num_heads)
        fcs = mhada_layer(fcs, fs[i], fcs, sd, f"adaAttnHead.{2 * i + 1}.", num_heads)
    return fcs, (decoder(fcs, sd) if decode else None)


def vit(img, sd: dict, num_layers: int = 3, num_heads: int = 8, patch: int = 8):
    """VisionTransformer.forward (vit.py:148-169) with functional ops: strided conv patch embedding, (resized)
    positional table, per layer LayerNorm -> nn.MultiheadAttention semantics WITHOUT batch_first on a (B, N, D)
    tensor (sequence axis = batch, vit.py:48,59) -> residual -> LayerNorm -> Linear / ReLU / Linear -> residual."""
    b, _, hh, ww = img.shape
    h, w = hh // patch, ww // patch
    x = F.conv2d(img, sd["patch_embedding.conv_proj.weight"], sd["patch_embedding.conv_proj.bias"], stride=patch)
    d = x.shape[1]
    x = x.reshape(b, d, h * w).permute(0, 2, 1)
    if "pos_embedding.pos_embed" in sd:
        pe = sd["pos_embedding.pos_embed"]
        if pe.shape[2] != h or pe.shape[3] != w:
            pe = F.interpolate(pe, size=(h, w), mode="bilinear", align_corners=False)
        x = x + pe.expand(b, -1, -1, -1).reshape(b, d, h * w).permute(0, 2, 1)
    z = []
    for l in range(num_layers):
        p = f"encoder.{l}."
        y = F.layer_norm(x, (d,), sd[p + "ln1.weight"], sd[p + "ln1.bias"], 1e-6)
        y, _ = F.multi_head_attention_forward(
            y, y, y, d, num_heads, sd[p + "attention.in_proj_weight"], sd[p + "attention.in_proj_bias"], None, None, False, 0.0,
            sd[p + "attention.out_proj.weight"], sd[p + "attention.out_proj.bias"], training=False, need_weights=False)
        x = y + x
        y = F.layer_norm(x, (d,), sd[p + "ln2.weight"], sd[p + "ln2.bias"], 1e-6)
        y = F.linear(F.relu(F.linear(y, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])), sd[p + "mlp.2.weight"], sd[p + "mlp.2.bias"])
        x = x + y
        z.append(x.permute(0, 2, 1).reshape(-1, d, h, w))
    return z
