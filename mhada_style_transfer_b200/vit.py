"""Host-side mirror of the reference's ViT encoder (MHAdaSTr/network/vit.py:45-169), SURVEY.md N3.

Same class names, constructor signatures, attribute names and state_dict keys as the reference
(`patch_embedding.conv_proj.*`, `pos_embedding.pos_embed`, `encoder.{i}.attention.in_proj_weight / in_proj_bias /
out_proj.*`, `encoder.{i}.mlp.{0,2}.*`, `encoder.{i}.ln{1,2}.*`), so `vit_c.load_state_dict(torch.load(VITC_PATH),
strict=True)` (infer_image.py:55) keeps working and, with the same torch seed, the random init is the reference's.

The inference forward hands the image to `mhada_vit_forward` (include/mhada_b200.h): patch gather, tcgen05 GEMMs
with fused bias / ReLU / residual / positional table, LayerNorm and the batch-axis attention as hand-written
sm_100a kernels.  It returns the three feature maps as (B, hidden, h, w) views over TOKEN-MAJOR bf16 memory --
exactly what `AdaAttnTransformerMultiHead` consumes without a copy -- so the pipeline's host boundary is the image
(infer_image.py:83-85) and the features never leave HBM.

Reference quirk kept on purpose: `nn.MultiheadAttention` is built without batch_first and fed
(B, N, D), so it attends ACROSS THE BATCH for every token position.  The kernels reproduce that, so an image's
features depend on the other images of its batch exactly as in the reference.

There is no CPU path.  Under autograd (train_image.py:103-108) the encoder runs the reference's op sequence with every
Linear layer (in_proj, out_proj, the MLP: > 95 % of its FLOPs) on the tcgen05 token GEMM, forward AND backward
(`_LinearTC`: dx = dy W and dW = dy^T x are the same GEMM kernel on transposed operands); LayerNorm, ReLU, the residual
adds, the 8-long batch-axis attention and the patch convolution stay differentiable PyTorch ops.  `train_impl = "torch"`
(or MHADA_VIT_TRAIN_IMPL=torch) keeps the plain fp32 PyTorch sequence (r1 / early r2: fp32 SIMT GEMMs, 20 ms per step).
"""
from __future__ import annotations

import ctypes
import os
from typing import List

import torch
import torch.nn as nn
from torch.nn import functional as F

from . import _lib
from .network import _needs_grad, _require_cuda, _stream, _workspace

__all__ = ["VisionTransformer", "EncoderBlock", "PatchEmbedding", "PosEmbedding"]


def _gemm_bf16(a16: torch.Tensor, w16: torch.Tensor, bias, M: int, N: int, K: int, out_dtype=torch.float32) -> torch.Tensor:
    """[M, N] (f32 or bf16) = a16 [M, >=K] . w16 [N, >=K]^T (+ bias) on the tcgen05 token GEMM (mhada_gemm_bf16)."""
    L = _lib.lib()
    out = torch.empty((M, N), dtype=out_dtype, device=a16.device)
    o16, o32 = (out.data_ptr(), None) if out_dtype == torch.bfloat16 else (None, out.data_ptr())
    with torch.cuda.device(a16.device):
        rc = L.mhada_gemm_bf16(a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0),
                               bias.data_ptr() if bias is not None else None, M, N, K, o16, N, o32, N, None, 0, 0, 0,
                               _stream())
    _lib.check("mhada_gemm_bf16", rc)
    return out


def _gemm_bf16_splitk(a16: torch.Tensor, w16: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    """f32 [M, N] = a16 . w16^T with the contraction split over extra work items (weight gradients: K = tokens)."""
    L = _lib.lib()
    out = torch.empty((M, N), dtype=torch.float32, device=a16.device)
    nb = L.mhada_gemm_splitk_workspace(M, N, K)
    ws = _workspace(a16.device, nb) if nb else None
    with torch.cuda.device(a16.device):
        rc = L.mhada_gemm_bf16_splitk(a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0), M, N, K, out.data_ptr(), N,
                                      ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0, _stream())
    _lib.check("mhada_gemm_bf16_splitk", rc)
    return out


def _transpose_bf16(x: torch.Tensor, M: int, C: int) -> torch.Tensor:
    """bf16 [C, Mpad] = x[M, C]^T, token axis padded with zeros to a multiple of 64 (mhada_transpose_bf16)."""
    L = _lib.lib()
    Mpad = (M + 63) // 64 * 64
    out = torch.empty((C, Mpad), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.mhada_transpose_bf16(x.data_ptr(), _lib.BF16 if x.dtype == torch.bfloat16 else _lib.F32, x.stride(0), M, C, Mpad,
                                    out.data_ptr(), _stream())
    _lib.check("mhada_transpose_bf16", rc)
    return out


def _colsum(x: torch.Tensor, M: int, C: int) -> torch.Tensor:
    L = _lib.lib()
    out = torch.empty((C,), dtype=torch.float32, device=x.device)
    ws = _workspace(x.device, L.mhada_colsum_workspace(M, C))
    with torch.cuda.device(x.device):
        rc = L.mhada_colsum(x.data_ptr(), _lib.BF16 if x.dtype == torch.bfloat16 else _lib.F32, M, C, ws.data_ptr(), ws.numel(),
                            out.data_ptr(), _stream())
    _lib.check("mhada_colsum", rc)
    return out


class _LinearTC(torch.autograd.Function):
    """y = x W^T + b with bf16 operands / f32 accumulation on the tcgen05 token GEMM, forward and backward:
    dx = dy W (the weight transposed), dW = dy^T x (both operands transposed so that the tokens are the contraction axis),
    db = column sums of dy.  Needs in_features and out_features to be multiples of 128."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_dtype=torch.float32):
        K, N = x.shape[-1], weight.shape[0]
        x2 = x.reshape(-1, K)
        M = x2.shape[0]
        x16 = x2.to(torch.bfloat16).contiguous()
        w16 = weight.detach().to(torch.bfloat16).contiguous()
        y = _gemm_bf16(x16, w16, bias.detach().float().contiguous() if bias is not None else None, M, N, K, out_dtype)
        ctx.save_for_backward(x16, w16)
        ctx.in_shape, ctx.has_bias, ctx.dtypes = x.shape, bias is not None, (x.dtype, weight.dtype)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x16, w16 = ctx.saved_tensors
        M, K = x16.shape
        N = w16.shape[0]
        dy2 = dy.reshape(M, N)
        dy2 = dy2 if dy2.is_contiguous() else dy2.contiguous()
        dy16 = dy2.to(torch.bfloat16)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _gemm_bf16(dy16, _transpose_bf16(w16, N, K), None, M, K, N).view(ctx.in_shape).to(ctx.dtypes[0])
        if ctx.needs_input_grad[1]:
            dyT, xT = _transpose_bf16(dy16, M, N), _transpose_bf16(x16, M, K)
            dw = _gemm_bf16_splitk(dyT, xT, N, K, dyT.shape[1]).to(ctx.dtypes[1])
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _colsum(dy2, M, N)
        return dx, dw, db, None


class _BatchAttnFn(torch.autograd.Function):
    """The batch_first=False attention of the encoder block (vit.py:48,59) on own kernels, forward
    (mhada_batch_attn) and backward (mhada_batch_attn_bwd): qkv bf16 [B, N, 3 D] -> bf16 [B, N, D], B <= 8."""

    @staticmethod
    def forward(ctx, qkv, heads):
        B, N, D3 = qkv.shape
        D = D3 // 3
        qkv = qkv.contiguous()
        out = torch.empty((B, N, D), dtype=torch.bfloat16, device=qkv.device)
        with torch.cuda.device(qkv.device):
            rc = _lib.lib().mhada_batch_attn(qkv.data_ptr(), B, N, heads, D // heads, out.data_ptr(), _stream())
        _lib.check("mhada_batch_attn", rc)
        ctx.save_for_backward(qkv)
        ctx.heads = heads
        return out

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        B, N, D3 = qkv.shape
        D = D3 // 3
        dout = dout.to(torch.bfloat16).contiguous()
        dqkv = torch.empty_like(qkv)
        with torch.cuda.device(qkv.device):
            rc = _lib.lib().mhada_batch_attn_bwd(qkv.data_ptr(), dout.data_ptr(), B, N, ctx.heads, D // ctx.heads, dqkv.data_ptr(),
                                                 _stream())
        _lib.check("mhada_batch_attn_bwd", rc)
        return dqkv, None


class EncoderBlock(nn.Module):
    """vit.py:45-64.  Parameters only; `forward` is the reference op sequence (used under autograd)."""

    def __init__(self, num_heads: int, hidden_dim: int, mlp_dim: int):
        super().__init__()
        self.attention = nn.MultiheadAttention(embed_dim=hidden_dim, num_heads=num_heads)    # batch_first=False, :48
        self.mlp = nn.Sequential(nn.Linear(hidden_dim, mlp_dim), nn.ReLU(), nn.Linear(mlp_dim, hidden_dim))
        self.ln1 = nn.LayerNorm(hidden_dim, eps=1e-6)
        self.ln2 = nn.LayerNorm(hidden_dim, eps=1e-6)

    def forward(self, input: torch.Tensor):
        x = self.ln1(input)
        x, _ = self.attention(x, x, x, need_weights=False)
        x = x + input
        return x + self.mlp(self.ln2(x))

    def forward_train_tc(self, input: torch.Tensor):
        """The same block (vit.py:54-64) with its four Linear layers on the tcgen05 GEMM (forward and backward).
        nn.MultiheadAttention without batch_first reads (B, N, D) as (sequence = B, batch = N): attention across the
        images of the batch per token position, scale 1 / sqrt(head_dim), no dropout."""
        a = self.attention
        L, Nn, D = input.shape
        H = a.num_heads
        bf16 = torch.bfloat16        # in_proj and fc1 feed ops that round to bf16 anyway: the GEMM writes bf16 directly
        qkv = _LinearTC.apply(self.ln1(input), a.in_proj_weight, a.in_proj_bias, bf16)
        if L <= 8 and D // H == 64:
            o = _BatchAttnFn.apply(qkv, H)                                                            # own kernels
        else:
            q, k, v = (t.reshape(L, Nn, H, D // H).permute(1, 2, 0, 3) for t in qkv.float().chunk(3, dim=-1))   # (N, H, L, hd)
            o = F.scaled_dot_product_attention(q, k, v).permute(2, 0, 1, 3).reshape(L, Nn, D)
        x = _LinearTC.apply(o, a.out_proj.weight, a.out_proj.bias) + input
        hdn = torch.relu(_LinearTC.apply(self.ln2(x), self.mlp[0].weight, self.mlp[0].bias, bf16))
        return x + _LinearTC.apply(hdn, self.mlp[2].weight, self.mlp[2].bias)


class PosEmbedding(nn.Module):
    """vit.py:67-102: learned (1, D, 32, 32) table, bilinearly resized to the token grid."""

    def __init__(self, patch_size: int = 8, embed_dim: int = 512, base_embed_size: int = 32):
        super().__init__()
        self.patch_size, self.embed_dim, self.base_embed_size = patch_size, embed_dim, base_embed_size
        self.pos_embed = nn.Parameter(torch.empty(1, embed_dim, base_embed_size, base_embed_size).normal_(std=0.02))
        self._table = {}

    def grid(self, out_h: int, out_w: int) -> torch.Tensor:
        """(1, D, out_h, out_w) table (vit.py:91-94)."""
        if out_h != self.base_embed_size or out_w != self.base_embed_size:
            return F.interpolate(self.pos_embed, size=(out_h, out_w), mode="bilinear", align_corners=False)
        return self.pos_embed

    def token_table(self, out_h: int, out_w: int) -> torch.Tensor:
        """f32 [N, D] token-major table for the GEMM epilogue; cached per grid size and parameter version."""
        p = self.pos_embed
        key = (out_h, out_w, p._version, p.data_ptr(), p.dtype, str(p.device))
        hit = self._table.get(str(p.device))
        if hit is None or hit[0] != key:
            with torch.no_grad():
                t = self.grid(out_h, out_w).float().reshape(self.embed_dim, out_h * out_w).t().contiguous()
            self._table[str(p.device)] = hit = (key, t)
        return hit[1]

    def forward(self, x_shape: torch.Size) -> torch.Tensor:
        b, _, h, w = x_shape
        out_h, out_w = h // self.patch_size, w // self.patch_size
        pe = self.grid(out_h, out_w).expand(b, -1, -1, -1)
        return pe.reshape(b, self.embed_dim, out_h * out_w).permute(0, 2, 1)


class PatchEmbedding(nn.Module):
    """vit.py:105-117."""

    def __init__(self, in_channels: int, patch_size: int, hidden_dim: int):
        super().__init__()
        self.conv_proj = nn.Conv2d(in_channels=in_channels, out_channels=hidden_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.conv_proj(x)
        b, c, h, w = x.shape
        return x.reshape(b, c, h * w).permute(0, 2, 1)


class _VitWeights:
    """bf16 copies of the GEMM weights and f32 copies of everything else, per device; rebuilt when a parameter was
    modified (version counter), moved or rebound."""

    def __init__(self):
        self.cache = {}

    def get(self, vit: "VisionTransformer"):
        params = list(vit.parameters())
        dev = params[0].device
        key = tuple((p._version, p.data_ptr(), p.dtype) for p in params)
        hit = self.cache.get(dev)
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
            f32 = lambda t: t.detach().float().contiguous()
            cp = vit.patch_embedding.conv_proj
            packed = {"w_patch": bf(cp.weight.reshape(cp.weight.shape[0], -1)), "b_patch": f32(cp.bias), "layers": []}
            for blk in vit.encoder:
                a = blk.attention
                packed["layers"].append(dict(
                    w_in=bf(a.in_proj_weight), b_in=f32(a.in_proj_bias), w_out=bf(a.out_proj.weight), b_out=f32(a.out_proj.bias),
                    w_fc1=bf(blk.mlp[0].weight), b_fc1=f32(blk.mlp[0].bias), w_fc2=bf(blk.mlp[2].weight), b_fc2=f32(blk.mlp[2].bias),
                    ln1_g=f32(blk.ln1.weight), ln1_b=f32(blk.ln1.bias), ln2_g=f32(blk.ln2.weight), ln2_b=f32(blk.ln2.bias)))
        self.cache[dev] = (key, packed)
        return packed


class VisionTransformer(nn.Module):
    """VisionTransformer (vit.py:120-169).  `forward(x)` -> list of `num_layers` feature maps (B, hidden, H/p, W/p).

    Extensions (not in the reference): `x` may be uint8 (0..255) as well as float; `out_dtype` selects what is
    returned -- "bf16" (default: token-major bf16, the MHAda tensor-core path's input) or "fp32" (the f32 residual
    stream itself)."""

    def __init__(self, patch_size: int = 8, num_layers: int = 3, num_heads: int = 8, hidden_dim: int = 512,
                 mlp_dim: int = 2048, pos_embedding: bool = True):
        super().__init__()
        self.patch_size, self.num_layers, self.hidden_dim = patch_size, num_layers, hidden_dim
        self.patch_embedding = PatchEmbedding(in_channels=3, patch_size=patch_size, hidden_dim=hidden_dim)
        self.pos_embedding = PosEmbedding(patch_size=patch_size, embed_dim=hidden_dim) if pos_embedding else None
        self.encoder = nn.ModuleList([EncoderBlock(num_heads=num_heads, hidden_dim=hidden_dim, mlp_dim=mlp_dim)
                                      for _ in range(num_layers)])
        self.num_heads, self.mlp_dim = num_heads, mlp_dim
        self.precision = "auto"          # "auto" | "bf16": tcgen05 path.  "fp32" is not built (NotImplementedError)
        self.out_dtype = "bf16"
        # training forward / backward: "auto" = Linear layers on the tcgen05 GEMM when the widths allow it, "kernels"
        # forces that, "torch" = the plain fp32 PyTorch op sequence
        self.train_impl = os.environ.get("MHADA_VIT_TRAIN_IMPL", "auto")
        self._weights = _VitWeights()

    # ---- the reference op sequence (vit.py:148-169), differentiable; used only when gradients are required
    def _forward_torch(self, x: torch.Tensor) -> List[torch.Tensor]:
        x_shape = x.shape
        out_h, out_w = x_shape[2] // self.patch_size, x_shape[3] // self.patch_size
        x = self.patch_embedding(x)
        if self.pos_embedding is not None:
            x = x + self.pos_embedding(x_shape)
        if self.train_impl not in ("auto", "kernels", "torch"):
            raise ValueError(f"Unknown train_impl: {self.train_impl}")
        tc_ok = self.hidden_dim % 128 == 0 and self.mlp_dim % 128 == 0
        if self.train_impl == "kernels" and not tc_ok:
            raise NotImplementedError("train_impl='kernels' needs hidden_dim and mlp_dim to be multiples of 128")
        use_tc = tc_ok and self.train_impl != "torch"
        z = []
        for layer in self.encoder:
            x = layer.forward_train_tc(x) if use_tc else layer(x)
            z.append(x.permute(0, 2, 1).reshape(-1, self.hidden_dim, out_h, out_w))
        return z

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected an image batch (B, 3, H, W), got {tuple(x.shape)}")
        _require_cuda(x)
        if _needs_grad(self, x):
            return self._forward_torch(x)
        if self.precision not in ("auto", "bf16", "fp32"):
            raise ValueError(f"Unknown precision: {self.precision}")
        if self.precision == "fp32":
            raise NotImplementedError("the B200 ViT runs on the bf16 tensor-core kernels (f32 residual stream); "
                                      "an fp32-arithmetic encoder is not built")
        if self.out_dtype not in ("bf16", "fp32"):
            raise ValueError(f"Unknown out_dtype: {self.out_dtype}")
        P, D = self.patch_size, self.hidden_dim
        B, _, Himg, Wimg = x.shape
        if Himg % P or Wimg % P:
            # the reference's strided convolution silently drops the remainder rows / columns (vit.py:109): crop alike
            x = x[:, :, : Himg - Himg % P, : Wimg - Wimg % P]
            Himg, Wimg = x.shape[2:]
        if Himg < P or Wimg < P:
            raise RuntimeError("image smaller than one patch")
        if x.dtype == torch.uint8:
            img, code = x.contiguous(), _lib.U8
        else:
            img, code = x.contiguous().float(), _lib.F32
        L = _lib.lib()
        h, w = Himg // P, Wimg // P
        N = h * w
        Wt = self._weights.get(self)
        dev = x.device
        a = _lib.VitArgs()
        a.img_dtype, a.img = code, img.data_ptr()
        a.B, a.Himg, a.Wimg, a.patch = B, Himg, Wimg, P
        a.D, a.F, a.heads, a.n_layers = D, self.mlp_dim, self.num_heads, self.num_layers
        a.w_patch, a.b_patch = Wt["w_patch"].data_ptr(), Wt["b_patch"].data_ptr()
        pos = self.pos_embedding.token_table(h, w) if self.pos_embedding is not None else None
        a.pos = pos.data_ptr() if pos is not None else None
        f32_out = [torch.empty((B, h, w, D), dtype=torch.float32, device=dev) for _ in range(self.num_layers)]
        bf_out = [torch.empty((B, h, w, D), dtype=torch.bfloat16, device=dev) for _ in range(self.num_layers)] \
            if self.out_dtype == "bf16" else [None] * self.num_layers
        if self.num_layers > _lib.VIT_MAX_LAYERS:
            raise NotImplementedError(f"at most {_lib.VIT_MAX_LAYERS} encoder layers")
        for l, lw in enumerate(Wt["layers"]):
            for k, t in lw.items():
                setattr(a.layers[l], k, t.data_ptr())
            a.feat_f32[l] = f32_out[l].data_ptr()
            a.feat_bf16[l] = bf_out[l].data_ptr() if bf_out[l] is not None else None
        nbytes = L.mhada_vit_workspace(B, N, D, self.mlp_dim, 3 * P * P)
        ws = _workspace(dev, nbytes)
        a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(dev):
            rc = L.mhada_vit_forward(ctypes.byref(a), _stream())
        _lib.check("mhada_vit_forward", rc)
        outs = bf_out if self.out_dtype == "bf16" else f32_out
        return [t.permute(0, 3, 1, 2) for t in outs]          # (B, D, h, w) views, channels_last memory (vit.py:163-166)
