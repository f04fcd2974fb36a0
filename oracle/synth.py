"""Deterministic synthetic inputs and weights for the MHAda hot path.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): imported by tests/, bench.py and
__graft_entry__.smoke() to make seeded inputs; never by the product package.

Everything here is a counter-based hash (splitmix64) evaluated with exact IEEE
integer / add / multiply arithmetic in numpy, so the same (seed, shape) gives the
same bits on any host.  No libm calls (log/cos) are used on purpose: the golden
vectors under tests/golden/ were produced from these inputs in the build
container and are re-derived on the GPU box.

Shapes and value ranges follow the reference:
  * images / features are floats, images in [0, 255]   (MHAdaSTr/utilities.py:11-16)
  * ViT feature maps at random init have mean ~1.3, std ~85 (SURVEY.md §8c probe)
  * nn.Conv2d default init is U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
    (the reference never re-initialises: MHAdaSTr/network/adaDecoder.py:143-152,
    MHAdaSTr/network/conv.py:23-33)
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _key(seed: int, stream: int = 0) -> np.uint64:
    k = _splitmix64(np.array([seed & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))
    k = _splitmix64(k ^ np.uint64(stream & 0xFFFFFFFFFFFFFFFF))
    return k[0]


def uniform01(seed: int, shape, stream: int = 0) -> np.ndarray:
    """float64 U[0,1) with 53 random bits, C-order over `shape`."""
    n = int(np.prod(shape)) if len(tuple(shape)) else 1
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        bits = _splitmix64((idx * np.uint64(0xD1342543DE82EF95)) ^ _key(seed, stream))
    u = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return u.reshape(shape)


def uniform(seed: int, shape, lo: float, hi: float, stream: int = 0) -> np.ndarray:
    return lo + (hi - lo) * uniform01(seed, shape, stream)


def bellish(seed: int, shape, mean: float = 0.0, std: float = 1.0) -> np.ndarray:
    """Irwin-Hall(4) variate, centred and scaled to (mean, std): bell shaped, bounded at
    +-3.46 sigma, and bit-reproducible (adds/multiplies only)."""
    s = np.zeros(shape, dtype=np.float64)
    for k in range(4):
        s = s + uniform01(seed, shape, stream=101 + k)
    # Irwin-Hall(4): mean 2, variance 4/12
    return mean + std * ((s - 2.0) * np.sqrt(3.0))


def features(seed: int, b: int, c: int, h: int, w: int, std: float = 85.0, mean: float = 1.3) -> np.ndarray:
    """A (B,C,h,w) feature map with the statistics the reference ViT emits at random init.
    Each channel gets its own offset and gain so instance-norm has real work to do."""
    x = bellish(seed, (b, c, h, w))
    gain = uniform(seed, (1, c, 1, 1), 0.5, 1.5, stream=7)
    off = uniform(seed, (b, c, 1, 1), -0.5, 0.5, stream=9)
    return mean + std * (gain * x + off)


def image(seed: int, b: int, h: int, w: int) -> np.ndarray:
    """(B,3,H,W) image in [0,255) (MHAdaSTr/utilities.py:11-16 value range)."""
    return uniform(seed, (b, 3, h, w), 0.0, 255.0)


# --------------------------------------------------------------------------------------
# state_dict builders: same keys / shapes as the reference modules
# --------------------------------------------------------------------------------------

def _conv_params(seed: int, out_c: int, in_c: int, k: int, gain: float = 1.0):
    bound = 1.0 / np.sqrt(in_c * k * k)
    w = uniform(seed, (out_c, in_c, k, k), -bound, bound, stream=1) * gain
    b = uniform(seed, (out_c,), -bound, bound, stream=2) * gain
    return w, b


def mhada_layer_state(seed: int, qkv_dim: int, num_heads: int, prefix: str = "", qk_gain: float = 1.0,
                      out_conv: bool = True) -> dict:
    """Keys of AdaAttnMultiHead (MHAdaSTr/network/adaDecoder.py:143-152):
    {f,g,h}_list.{i}.{weight,bias}, out_conv.{weight,bias}.  qk_gain>1 makes the logits
    larger (stress case for the streaming softmax)."""
    d = qkv_dim // num_heads
    sd = {}
    for li, name in enumerate(("f_list", "g_list", "h_list")):
        for i in range(num_heads):
            g = qk_gain if name != "h_list" else 1.0
            w, b = _conv_params(seed * 1000 + li * 100 + i, d, d, 1, g)
            sd[f"{prefix}{name}.{i}.weight"] = w
            sd[f"{prefix}{name}.{i}.bias"] = b
    if out_conv:
        w, b = _conv_params(seed * 1000 + 900, qkv_dim, qkv_dim, 1)
        sd[f"{prefix}out_conv.weight"] = w
        sd[f"{prefix}out_conv.bias"] = b
    return sd


def adaattn_state(seed: int, qkv_dim: int, prefix: str = "") -> dict:
    """Keys of the single-head AdaAttN (MHAdaSTr/network/adaDecoder.py:88-90): f, g, h."""
    sd = {}
    for li, name in enumerate(("f", "g", "h")):
        w, b = _conv_params(seed * 1000 + li * 100, qkv_dim, qkv_dim, 1)
        sd[f"{prefix}{name}.weight"] = w
        sd[f"{prefix}{name}.bias"] = b
    return sd


DECODER_CONVS = (  # (key stem, out, in)  -- MHAdaSTr/network/conv.py:78-94
    ("conv1.0", 256, 512), ("conv1.1", 256, 256), ("conv1.2", 256, 256), ("conv1.3", 256, 256),
    ("conv1.4", 128, 256), ("conv2.0", 128, 128), ("conv2.1", 64, 128), ("conv3.0", 64, 64),
    ("conv3.1", 3, 64),
)


def decoder_state(seed: int, prefix: str = "decoder.") -> dict:
    sd = {}
    for j, (stem, oc, ic) in enumerate(DECODER_CONVS):
        w, b = _conv_params(seed * 1000 + 500 + j, oc, ic, 3)
        sd[f"{prefix}{stem}.conv.conv.weight"] = w
        sd[f"{prefix}{stem}.conv.conv.bias"] = b
    return sd


def transformer_state(seed: int, num_layers: int = 3, qkv_dim: int = 512, num_heads: int = 8,
                      qk_gain: float = 1.0) -> dict:
    """All 318 keys (defaults) of AdaAttnTransformerMultiHead
    (MHAdaSTr/network/adaDecoder.py:235-251)."""
    sd = {}
    for l in range(2 * num_layers):
        sd.update(mhada_layer_state(seed + 17 * l, qkv_dim, num_heads, prefix=f"adaAttnHead.{l}.",
                                    qk_gain=qk_gain))
    sd.update(decoder_state(seed))
    return sd


def single_head_transformer_state(seed: int, num_layers: int = 3, qkv_dim: int = 512) -> dict:
    """Keys of AdaAttnTransformer (MHAdaSTr/network/adaDecoder.py:209-225): adaAttNs.{i}.{f,g,h} + decoder."""
    sd = {}
    for l in range(num_layers):
        sd.update(adaattn_state(seed + 29 * l, qkv_dim, prefix=f"adaAttNs.{l}."))
    sd.update(decoder_state(seed))
    return sd


def vit_state(seed: int, pos_embedding: bool = True, patch: int = 8, num_layers: int = 3, hidden: int = 512,
              mlp: int = 2048) -> dict:
    """Keys of VisionTransformer (MHAdaSTr/network/vit.py:120-146): patch_embedding.conv_proj, pos_embedding.pos_embed
    (vit_c only), encoder.{l}.{attention.in_proj_weight/bias, attention.out_proj, mlp.0, mlp.2, ln1, ln2}.
    Scales follow the PyTorch default initialisers; biases and LayerNorm affine parameters are perturbed away from
    their 0 / 1 defaults so that every term of the forward is exercised."""
    sd = {}
    w, b = _conv_params(seed * 1000 + 1, hidden, 3, patch)
    sd["patch_embedding.conv_proj.weight"], sd["patch_embedding.conv_proj.bias"] = w, b
    if pos_embedding:
        sd["pos_embedding.pos_embed"] = bellish(seed * 1000 + 2, (1, hidden, 32, 32), 0.0, 0.02)
    for l in range(num_layers):
        s0 = seed * 1000 + 10 + 20 * l
        pre = f"encoder.{l}."
        xav = np.sqrt(6.0 / (hidden + 3 * hidden))                    # xavier_uniform_ of in_proj_weight
        sd[pre + "attention.in_proj_weight"] = uniform(s0, (3 * hidden, hidden), -xav, xav, stream=1)
        sd[pre + "attention.in_proj_bias"] = uniform(s0, (3 * hidden,), -0.05, 0.05, stream=2)
        kb = 1.0 / np.sqrt(hidden)
        sd[pre + "attention.out_proj.weight"] = uniform(s0 + 1, (hidden, hidden), -kb, kb, stream=1)
        sd[pre + "attention.out_proj.bias"] = uniform(s0 + 1, (hidden,), -0.05, 0.05, stream=2)
        sd[pre + "mlp.0.weight"] = uniform(s0 + 2, (mlp, hidden), -kb, kb, stream=1)
        sd[pre + "mlp.0.bias"] = uniform(s0 + 2, (mlp,), -kb, kb, stream=2)
        km = 1.0 / np.sqrt(mlp)
        sd[pre + "mlp.2.weight"] = uniform(s0 + 3, (hidden, mlp), -km, km, stream=1)
        sd[pre + "mlp.2.bias"] = uniform(s0 + 3, (hidden,), -km, km, stream=2)
        sd[pre + "ln1.weight"] = uniform(s0 + 4, (hidden,), 0.9, 1.1, stream=1)
        sd[pre + "ln1.bias"] = uniform(s0 + 4, (hidden,), -0.1, 0.1, stream=2)
        sd[pre + "ln2.weight"] = uniform(s0 + 5, (hidden,), 0.9, 1.1, stream=1)
        sd[pre + "ln2.bias"] = uniform(s0 + 5, (hidden,), -0.1, 0.1, stream=2)
    return sd


def image_u8(seed: int, b: int, h: int, w: int) -> np.ndarray:
    """(B,3,H,W) image with INTEGER values 0..255 as float64 (what toTensor255 yields for an 8-bit image,
    MHAdaSTr/utilities.py:11-16); smooth + noisy so that patches differ."""
    u = uniform01(seed, (b, 3, h, w))
    yy = np.arange(h).reshape(1, 1, h, 1) / max(h - 1, 1)
    xx = np.arange(w).reshape(1, 1, 1, w) / max(w - 1, 1)
    ph = uniform01(seed, (b, 3, 1, 1), stream=5)
    base = 0.5 + 0.35 * (2.0 * ((yy * (1.0 + ph) + xx * (2.0 - ph)) % 1.0) - 1.0)       # saw-tooth ramps, no libm
    return np.floor(np.clip(0.6 * base + 0.4 * u, 0.0, 1.0) * 255.0)


def to_torch(sd: dict, dtype=None):
    import torch
    out = {}
    for k, v in sd.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        out[k] = t.to(dtype) if dtype is not None else t
    return out
