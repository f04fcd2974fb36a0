"""CPU oracle for the MHAda forward hot path -- a numpy restatement of the reference algorithm.

TEST INFRASTRUCTURE ONLY.  Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this file; the product package never does, and it fails
loudly when its CUDA library is missing instead of falling back to anything here.

Parity status: the reference ships no tests, golden vectors or fixtures for this path
 -- "parity unpinned" by the reference's own tests.  The pins are therefore made
by running the *unmodified reference module* (imported from /root/reference/MHAdaSTr in the build
container) on the seeded inputs of oracle/synth.py: see oracle/gen_golden.py and tests/golden/.
tests/test_oracle.py checks every function below against those vectors.

Each function cites the reference lines it follows.  Arithmetic is float64 by default (the
reference is float32; its float64 run is the tighter yardstick).

Layout: feature maps are (B, C, h, w) arrays exactly like the reference's tensors.
"""
from __future__ import annotations

import numpy as np

IN_EPS = 1e-5        # nn.InstanceNorm2d default eps         (adaDecoder.py:147-149)
VAR_FLOOR = 1e-6     # Var.clamp(min=1e-6)                    (adaDecoder.py:191)


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------

def instance_norm_stats(x: np.ndarray):
    """Per-(b, channel) mean and 1/sqrt(biased var + eps) over the spatial positions.
    nn.InstanceNorm2d(affine=False, track_running_stats=False): adaDecoder.py:147-149."""
    b, c = x.shape[:2]
    flat = x.reshape(b, c, -1)
    mean = flat.mean(axis=2)
    var = flat.var(axis=2)  # biased (ddof=0), like instance_norm
    return mean, 1.0 / np.sqrt(var + IN_EPS)


def instance_norm(x: np.ndarray) -> np.ndarray:
    mean, rstd = instance_norm_stats(x)
    return (x - mean[:, :, None, None]) * rstd[:, :, None, None]


def conv1x1(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """nn.Conv2d(kernel_size=1): x (B,I,h,w), w (O,I,1,1) or (O,I), b (O,)."""
    w2 = w.reshape(w.shape[0], -1)
    y = np.einsum("oi,bihw->bohw", w2, x, optimize=True)
    return y + b[None, :, None, None]


def softmax_attention(q: np.ndarray, k: np.ndarray) -> np.ndarray:
    """Softmax.forward: softmax(bmm(q, k), dim=-1), no 1/sqrt(d) (adaDecoder.py:11-17).
    q (B,Nc,d), k (B,d,Ns) -> (B,Nc,Ns)."""
    s = np.matmul(q, k)
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s)
    return e / e.sum(axis=-1, keepdims=True)


def cosine_attention(q: np.ndarray, k: np.ndarray) -> np.ndarray:
    """CosineSimilarity.forward (adaDecoder.py:20-34): a = (cos+1) / sum_j (cos+1)."""
    qn = np.linalg.norm(q, axis=-1, keepdims=True)
    kn = np.linalg.norm(k, axis=1, keepdims=True)
    s = np.matmul(q, k) / np.matmul(qn, kn) + 1.0
    return s / s.sum(axis=-1, keepdims=True)


def _activation(name: str):
    if name == "softmax":
        return softmax_attention
    if name == "cosine":
        return cosine_attention
    raise ValueError(f"Unknown activation function: {name}")   # adaDecoder.py:160


def _attend(q, k, v, activation):
    """A = act(Q,K); M = A V; S = sqrt(clamp(A V^2 - M^2, 1e-6))   (adaDecoder.py:186-191).
    q (B,Nc,dq) k (B,dq,Ns) v (B,Ns,dv) -> M,S (B,Nc,dv)."""
    a = activation(q, k)
    m = np.matmul(a, v)
    var = np.matmul(a, v * v) - m * m
    s = np.sqrt(np.maximum(var, VAR_FLOOR))
    return m, s


# --------------------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------------------

def ada_attn_multi_head(fc, fs, fcs, sd: dict, num_heads: int, prefix: str = "",
                        activation: str = "softmax", dtype=np.float64) -> np.ndarray:
    """AdaAttnMultiHead.forward (adaDecoder.py:162-206).
    fc, fcs (B,C,h,w); fs (B,C,hs,ws) with the same B (adaDecoder.py:177-183 reshapes K/V with
    the content batch).  sd holds f_list/g_list/h_list/out_conv weights under `prefix`."""
    fc = np.asarray(fc, dtype=dtype); fs = np.asarray(fs, dtype=dtype); fcs = np.asarray(fcs, dtype=dtype)
    b, c, h, w = fc.shape
    if c % num_heads != 0:
        raise ValueError("qkv_dim must be divisible by num_heads")      # adaDecoder.py:137-138
    if fs.shape[0] != b:
        raise RuntimeError("style batch must equal content batch")      # adaDecoder.py:179 reshape
    act = _activation(activation)
    d = c // num_heads
    hs, ws = fs.shape[2:]
    heads = []
    for i in range(num_heads):
        sl = slice(i * d, (i + 1) * d)                                    # :168-170
        wq = np.asarray(sd[f"{prefix}f_list.{i}.weight"], dtype=dtype); bq = np.asarray(sd[f"{prefix}f_list.{i}.bias"], dtype=dtype)
        wk = np.asarray(sd[f"{prefix}g_list.{i}.weight"], dtype=dtype); bk = np.asarray(sd[f"{prefix}g_list.{i}.bias"], dtype=dtype)
        wv = np.asarray(sd[f"{prefix}h_list.{i}.weight"], dtype=dtype); bv = np.asarray(sd[f"{prefix}h_list.{i}.bias"], dtype=dtype)
        q = conv1x1(instance_norm(fc[:, sl]), wq, bq).reshape(b, d, h * w).transpose(0, 2, 1)   # :173-174
        k = conv1x1(instance_norm(fs[:, sl]), wk, bk).reshape(b, d, hs * ws)                    # :177-179
        v = conv1x1(fs[:, sl], wv, bv).reshape(b, d, hs * ws).transpose(0, 2, 1)                # :182-183
        m, s = _attend(q, k, v, act)                                                              # :186-191
        m = m.reshape(b, h, w, d).transpose(0, 3, 1, 2)                                           # :194-195
        s = s.reshape(b, h, w, d).transpose(0, 3, 1, 2)
        heads.append(s * instance_norm(fcs[:, sl]) + m)                                           # :198
    cat = np.concatenate(heads, axis=1)                                                           # :202
    wo = np.asarray(sd[f"{prefix}out_conv.weight"], dtype=dtype); bo = np.asarray(sd[f"{prefix}out_conv.bias"], dtype=dtype)
    return conv1x1(cat, wo, bo)                                                                   # :205


def ada_attn(fc, fs, fcs, sd: dict, prefix: str = "", activation: str = "softmax", dtype=np.float64):
    """AdaAttN.forward (adaDecoder.py:102-131): single head, learnable f/g/h, no out_conv."""
    fc = np.asarray(fc, dtype=dtype); fs = np.asarray(fs, dtype=dtype); fcs = np.asarray(fcs, dtype=dtype)
    b, c, h, w = fc.shape
    hs, ws = fs.shape[2:]
    act = _activation(activation)
    g = lambda n: np.asarray(sd[f"{prefix}{n}"], dtype=dtype)
    q = conv1x1(instance_norm(fc), g("f.weight"), g("f.bias")).reshape(b, c, h * w).transpose(0, 2, 1)
    k = conv1x1(instance_norm(fs), g("g.weight"), g("g.bias")).reshape(b, c, hs * ws)
    v = conv1x1(fs, g("h.weight"), g("h.bias")).reshape(b, c, hs * ws).transpose(0, 2, 1)
    m, s = _attend(q, k, v, act)
    m = m.reshape(b, h, w, c).transpose(0, 3, 1, 2)
    s = s.reshape(b, h, w, c).transpose(0, 3, 1, 2)
    return s * instance_norm(fcs) + m


def ada_attn_for_loss(c_x, s_x, c_1x, s_1x, activation: str = "softmax", dtype=np.float64):
    """AdaAttnForLoss.forward (adaDecoder.py:53-81): parameter-free; Q/K width != V width."""
    c_x = np.asarray(c_x, dtype=dtype); s_x = np.asarray(s_x, dtype=dtype)
    c_1x = np.asarray(c_1x, dtype=dtype); s_1x = np.asarray(s_1x, dtype=dtype)
    act = _activation(activation)
    b, cq, h, w = c_1x.shape
    q = instance_norm(c_1x).reshape(b, cq, h * w).transpose(0, 2, 1)     # :55-57
    bs, _, hs, ws = s_1x.shape
    k = instance_norm(s_1x).reshape(bs, cq, hs * ws)                     # :60-62
    cv = s_x.shape[1]
    v = s_x.reshape(s_x.shape[0], cv, -1).transpose(0, 2, 1)             # :65-67
    m, s = _attend(q, k, v, act)                                         # :70-75
    b, _, h, w = c_x.shape
    m = m.reshape(b, h, w, cv).transpose(0, 3, 1, 2)                     # :78-79
    s = s.reshape(b, h, w, cv).transpose(0, 3, 1, 2)
    return s * instance_norm(c_x) + m                                    # :81


# --------------------------------------------------------------------------------------
# decoder (MHAdaSTr/network/conv.py)
# --------------------------------------------------------------------------------------

def reflection_pad1(x: np.ndarray) -> np.ndarray:
    """nn.ReflectionPad2d(1) (conv.py:26-27 with kernel_size 3)."""
    return np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="reflect")


def conv3x3_reflect(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Conv.forward: ReflectionPad2d(1) then Conv2d(k=3, stride=1) (conv.py:23-33)."""
    bsz, ic, h, wd = x.shape
    oc = w.shape[0]
    xp = reflection_pad1(x)
    out = np.zeros((bsz, oc, h, wd), dtype=x.dtype)
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, :, ky:ky + h, kx:kx + wd].reshape(bsz, ic, h * wd)
            out += np.matmul(w[:, :, ky, kx][None], patch).reshape(bsz, oc, h, wd)
    return out + b[None, :, None, None]


def _up2_axis(x: np.ndarray, axis: int) -> np.ndarray:
    """x2 bilinear, align_corners=False, along one axis: src = (i + 0.5)/2 - 0.5 clamped at 0."""
    n = x.shape[axis]
    i = np.arange(2 * n)
    src = np.maximum((i + 0.5) / 2.0 - 0.5, 0.0)
    i0 = np.floor(src).astype(np.int64)
    i1 = np.minimum(i0 + 1, n - 1)
    lam = (src - i0).astype(x.dtype)
    shp = [1] * x.ndim
    shp[axis] = 2 * n
    lam = lam.reshape(shp)
    return np.take(x, i0, axis=axis) * (1 - lam) + np.take(x, i1, axis=axis) * lam


def bilinear_up2(x: np.ndarray) -> np.ndarray:
    """F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) (conv.py:71)."""
    return _up2_axis(_up2_axis(x, 2), 3)


_DECODER_PLAN = (  # (key stem, upsample after relu?)  conv.py:78-94
    ("conv1.0", True), ("conv1.1", False), ("conv1.2", False), ("conv1.3", False), ("conv1.4", True),
    ("conv2.0", False), ("conv2.1", True), ("conv3.0", False), ("conv3.1", False),
)


def decoder(fcs, sd: dict, prefix: str = "decoder.", dtype=np.float64) -> np.ndarray:
    """Decoder.forward (conv.py:96-100): nine reflect-pad 3x3 conv + ReLU blocks, three of them
    followed by a x2 bilinear up-sample; the last block is ConvReLU(64,3) so the output is >= 0."""
    x = np.asarray(fcs, dtype=dtype)
    for stem, up in _DECODER_PLAN:
        w = np.asarray(sd[f"{prefix}{stem}.conv.conv.weight"], dtype=dtype)
        b = np.asarray(sd[f"{prefix}{stem}.conv.conv.bias"], dtype=dtype)
        x = np.maximum(conv3x3_reflect(x, w, b), 0.0)
        if up:
            x = bilinear_up2(x)
    return x


# --------------------------------------------------------------------------------------
# transformers
# --------------------------------------------------------------------------------------

def transformer_multi_head(fc_list, fs_list, sd: dict, num_layers: int = 3, num_heads: int = 8,
                           activation: str = "softmax", dtype=np.float64, decode: bool = True):
    """AdaAttnTransformerMultiHead.forward (adaDecoder.py:253-268): returns (fcs, cs)."""
    fcs = np.asarray(fc_list[0], dtype=dtype)
    for i in range(num_layers):
        fcs = ada_attn_multi_head(fc_list[i], fs_list[i], fcs, sd, num_heads,
                                  prefix=f"adaAttnHead.{2 * i}.", activation=activation, dtype=dtype)
        fcs = ada_attn_multi_head(fcs, fs_list[i], fcs, sd, num_heads,
                                  prefix=f"adaAttnHead.{2 * i + 1}.", activation=activation, dtype=dtype)
    cs = decoder(fcs, sd, dtype=dtype) if decode else None
    return fcs, cs


def transformer_single_head(fc_list, fs_list, sd: dict, num_layers: int = 3, activation: str = "softmax",
                            dtype=np.float64):
    """AdaAttnTransformer.forward (adaDecoder.py:227-232): returns cs only."""
    fcs = np.asarray(fc_list[0], dtype=dtype)
    for i in range(num_layers):
        fcs = ada_attn(fc_list[i], fs_list[i], fcs, sd, prefix=f"adaAttNs.{i}.", activation=activation, dtype=dtype)
    return decoder(fcs, sd, dtype=dtype)


# --------------------------------------------------------------------------------------
# error metrics used by every parity test (SURVEY.md §8c / BASELINE.md §5.6)
# --------------------------------------------------------------------------------------

def errors(got, want) -> dict:
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    diff = np.abs(got - want)
    absmax = float(np.abs(want).max())
    return {
        "max_abs": float(diff.max()),
        "absmax": absmax,
        "max_abs_rel": float(diff.max() / max(absmax, 1e-30)),
        "fro_rel": float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)),
    }
