"""Golden-case catalogue shared by oracle/gen_golden.py (which runs the unmodified reference)
and tests/ (which re-derive the same seeded inputs and compare).  TEST INFRASTRUCTURE ONLY.

A case is a dict; inputs and weights are never stored, only re-generated from seeds through
oracle/synth.py, so the fixtures under tests/golden/ hold reference OUTPUTS only.

Case kinds (reference symbol each one pins):
  layer       AdaAttnMultiHead.forward                      adaDecoder.py:162-206
  adaattn     AdaAttN.forward                               adaDecoder.py:102-131
  forloss     AdaAttnForLoss.forward                        adaDecoder.py:53-81
  transformer AdaAttnTransformerMultiHead.forward (+decoder) adaDecoder.py:253-268, conv.py:96-100
  decoder     Decoder.forward                               conv.py:96-100
"""
from __future__ import annotations

import numpy as np

from . import synth

LAYER_CASES = [
    # name, B, C, heads, (h,w) content, (hs,ws) style, qk_gain, fcs_is_fc
    dict(name="layer_c512_h8_16x16", kind="layer", B=2, C=512, H=8, hw=(16, 16), hsws=(16, 16), gain=1.0, seed=11),
    dict(name="layer_c512_h8_ragged", kind="layer", B=1, C=512, H=8, hw=(9, 15), hsws=(11, 13), gain=1.0, seed=12),
    dict(name="layer_c512_h8_cross", kind="layer", B=1, C=512, H=8, hw=(16, 24), hsws=(8, 8), gain=1.0, seed=13),
    dict(name="layer_c512_h8_stress", kind="layer", B=1, C=512, H=8, hw=(16, 16), hsws=(12, 20), gain=4.0, seed=14),
    dict(name="layer_c512_h4_12x12", kind="layer", B=1, C=512, H=4, hw=(12, 12), hsws=(12, 12), gain=1.0, seed=15),
    dict(name="layer_c512_h1_8x8", kind="layer", B=1, C=512, H=1, hw=(8, 8), hsws=(8, 8), gain=1.0, seed=16),
    dict(name="layer_c128_h2_20x20", kind="layer", B=2, C=128, H=2, hw=(20, 20), hsws=(20, 12), gain=1.0, seed=17),
    dict(name="layer_c64_h1_2keys", kind="layer", B=1, C=64, H=1, hw=(6, 6), hsws=(1, 2), gain=1.0, seed=18),
    dict(name="layer_c512_h8_selffcs", kind="layer", B=1, C=512, H=8, hw=(12, 12), hsws=(12, 12), gain=1.0, seed=19,
         fcs_is_fc=True),
    # activation="cosine" (CosineSimilarity, adaDecoder.py:20-34): a = (cos + 1) / sum(cos + 1)
    dict(name="layer_c512_h8_cosine", kind="layer", B=1, C=512, H=8, hw=(12, 12), hsws=(10, 14), gain=1.0, seed=20,
         activation="cosine"),
    dict(name="layer_c128_h2_cosine", kind="layer", B=2, C=128, H=2, hw=(9, 7), hsws=(20, 12), gain=1.0, seed=21,
         activation="cosine"),
]

ADAATTN_CASES = [
    dict(name="adaattn_c64_10x10", kind="adaattn", B=2, C=64, hw=(10, 10), hsws=(7, 9), seed=31),
    dict(name="adaattn_c64_cosine", kind="adaattn", B=1, C=64, hw=(8, 8), hsws=(9, 9), seed=32, activation="cosine"),
]

FORLOSS_CASES = [
    # AdaAttnForLoss shapes from train_image.py:52-58 at reduced spatial size: (v_dim, qk_dim)
    dict(name="forloss_relu3_1", kind="forloss", B=1, v=256, qk=448, hw=(12, 12), hsws=(12, 12), seed=41),
    dict(name="forloss_relu4_1", kind="forloss", B=2, v=512, qk=960, hw=(6, 6), hsws=(6, 6), seed=42),
    dict(name="forloss_relu5_1", kind="forloss", B=1, v=512, qk=1472, hw=(3, 3), hsws=(3, 3), seed=43),
    dict(name="forloss_relu3_1_cosine", kind="forloss", B=1, v=256, qk=448, hw=(8, 8), hsws=(8, 10), seed=44,
         activation="cosine"),
]

TRANSFORMER_CASES = [
    dict(name="transformer_8x8", kind="transformer", B=1, hw=(8, 8), hsws=(8, 8), seed=51, sub=1, img_sub=1),
    dict(name="transformer_b2_12x10", kind="transformer", B=2, hw=(12, 10), hsws=(6, 9), seed=52, sub=1, img_sub=2),
    # cfg1-sized (512x512 image -> 64x64 tokens): outputs stored on a token / pixel sub-lattice
    dict(name="transformer_64x64_sub", kind="transformer", B=1, hw=(64, 64), hsws=(64, 64), seed=53, sub=37,
         img_sub=8),
    # the sizes BASELINE.json states its targets on (r2): cfg2 batch (8 x 512^2), cfg3 (1024^2 -> 16384 tokens),
    # cfg4 (1080p frame x 512^2 style: 32400 x 4096 tokens), and the 4-head six-layer chain of configs[4]
    dict(name="transformer_b8_64x64_sub", kind="transformer", B=8, hw=(64, 64), hsws=(64, 64), seed=56, sub=149,
         img_sub=16),
    dict(name="transformer_128x128_sub", kind="transformer", B=1, hw=(128, 128), hsws=(128, 128), seed=57, sub=149,
         img_sub=16),
    dict(name="transformer_135x240_x_64x64_sub", kind="transformer", B=1, hw=(135, 240), hsws=(64, 64), seed=58,
         sub=293, img_sub=24),
    dict(name="transformer_h4_32x32_sub", kind="transformer", B=1, hw=(32, 32), hsws=(32, 32), seed=59, sub=7,
         img_sub=4, heads=4),
    dict(name="transformer_h4_64x64_sub", kind="transformer", B=2, hw=(64, 64), hsws=(48, 56), seed=63, sub=61,
         img_sub=8, heads=4),
]

SINGLE_HEAD_TRANSFORMER_CASES = [
    # AdaAttnTransformer (adaDecoder.py:209-232): three single-head AdaAttN layers + decoder, returns cs only
    dict(name="single_head_transformer_6x8", kind="single_head_transformer", B=1, hw=(6, 8), hsws=(7, 5), seed=55, img_sub=1),
]

DECODER_CASES = [
    dict(name="decoder_5x7", kind="decoder", B=1, hw=(5, 7), seed=61),
]

GRAD_CASES = [
    # d(loss)/d(inputs, weights) of one AdaAttnMultiHead layer, loss = sum(out * G); reference autograd in float64
    dict(name="grad_layer_c128_h2", kind="grad", B=2, C=128, H=2, hw=(10, 10), hsws=(8, 9), gain=1.0, seed=71),
    # the shape of the training step (train_image.py: 256 x 256 images -> 32 x 32 tokens, 8 heads of 64); input
    # gradients stored on every 13th token
    dict(name="grad_layer_c512_h8_32x32_sub", kind="grad", B=2, C=512, H=8, hw=(32, 32), hsws=(32, 32), gain=1.0, seed=72, sub=13),
]

VIT_CASES = [
    # VisionTransformer (vit.py:120-169).  B > 1 pins the batch-axis attention of the reference.
    dict(name="vit_b1_48x64", kind="vit", B=1, img=(48, 64), pos=True, seed=81, sub=1),
    dict(name="vit_b3_40x64_nopos", kind="vit", B=3, img=(40, 64), pos=False, seed=82, sub=1),
    dict(name="vit_b8_32x32", kind="vit", B=8, img=(32, 32), pos=True, seed=83, sub=1),
    dict(name="vit_b2_256x256_sub", kind="vit", B=2, img=(256, 256), pos=True, seed=84, sub=13),       # 32 x 32 grid: table as is
    dict(name="vit_b1_512x512_sub", kind="vit", B=1, img=(512, 512), pos=True, seed=85, sub=37),       # BASELINE configs[0] image
    dict(name="vit_b1_512x512_nopos_sub", kind="vit", B=1, img=(512, 512), pos=False, seed=86, sub=37),
]

PIPELINE_CASES = [
    # images -> vit_c, vit_s -> AdaAttnTransformerMultiHead -> decoded image (infer_image.py:82-86), end to end
    dict(name="pipeline_b1_64x64", kind="pipeline", B=1, img=(64, 64), simg=(64, 64), seed=91, sub=1, img_sub=2),
    dict(name="pipeline_b2_128x96", kind="pipeline", B=2, img=(128, 96), simg=(64, 80), seed=97, sub=3, img_sub=4),
    dict(name="pipeline_b1_512x512_sub", kind="pipeline", B=1, img=(512, 512), simg=(512, 512), seed=93, sub=37, img_sub=8),
]

ALL_CASES = (LAYER_CASES + ADAATTN_CASES + FORLOSS_CASES + TRANSFORMER_CASES + SINGLE_HEAD_TRANSFORMER_CASES +
             DECODER_CASES + GRAD_CASES + VIT_CASES + PIPELINE_CASES)

GRAD_KEYS = ("fc", "fs", "fcs", "f_list.0.weight", "f_list.1.bias", "g_list.1.weight", "g_list.0.bias",
             "h_list.0.weight", "h_list.1.bias", "out_conv.weight", "out_conv.bias")


def grad_inputs(case: dict):
    fc, fs, fcs, sd = layer_inputs(case)
    B, C = case["B"], case["C"]
    h, w = case["hw"]
    G = synth.bellish(case["seed"] * 10 + 9, (B, C, h, w), 0.0, 1.0)
    return fc, fs, fcs, sd, G


def by_name(name: str) -> dict:
    for c in ALL_CASES:
        if c["name"] == name:
            return c
    raise KeyError(name)


# --------------------------------------------------------------------------------------
# seeded inputs (float64 numpy); the same arrays feed the reference, the oracle and the GPU path
# --------------------------------------------------------------------------------------

def layer_inputs(case: dict):
    B, C = case["B"], case["C"]
    h, w = case["hw"]
    hs, ws = case["hsws"]
    s = case["seed"]
    fc = synth.features(s * 10 + 1, B, C, h, w)
    fs = synth.features(s * 10 + 2, B, C, hs, ws)
    fcs = fc if case.get("fcs_is_fc") else synth.features(s * 10 + 3, B, C, h, w, std=30.0, mean=-1.6)
    sd = synth.mhada_layer_state(s, C, case["H"], qk_gain=case.get("gain", 1.0))
    return fc, fs, fcs, sd


def adaattn_inputs(case: dict):
    B, C = case["B"], case["C"]
    h, w = case["hw"]
    hs, ws = case["hsws"]
    s = case["seed"]
    fc = synth.features(s * 10 + 1, B, C, h, w)
    fs = synth.features(s * 10 + 2, B, C, hs, ws)
    fcs = synth.features(s * 10 + 3, B, C, h, w, std=30.0)
    return fc, fs, fcs, synth.adaattn_state(s, C)


def forloss_inputs(case: dict):
    B = case["B"]
    h, w = case["hw"]
    hs, ws = case["hsws"]
    s = case["seed"]
    # VGG relu features are non-negative; shift so most values are > 0 (lossfn.py:26-34 feeds relu maps)
    c_x = np.abs(synth.features(s * 10 + 1, B, case["v"], h, w, std=2.0, mean=1.0))
    s_x = np.abs(synth.features(s * 10 + 2, B, case["v"], hs, ws, std=2.0, mean=1.0))
    c_1x = np.abs(synth.features(s * 10 + 3, B, case["qk"], h, w, std=2.0, mean=1.0))
    s_1x = np.abs(synth.features(s * 10 + 4, B, case["qk"], hs, ws, std=2.0, mean=1.0))
    return c_x, s_x, c_1x, s_1x


def transformer_inputs(case: dict, num_layers: int = 3):
    B = case["B"]
    h, w = case["hw"]
    hs, ws = case["hsws"]
    s = case["seed"]
    fc = [synth.features(s * 100 + i, B, 512, h, w) for i in range(num_layers)]
    fs = [synth.features(s * 100 + 10 + i, B, 512, hs, ws) for i in range(num_layers)]
    sd = synth.transformer_state(s, num_heads=case.get("heads", 8))
    return fc, fs, sd


def single_head_transformer_inputs(case: dict, num_layers: int = 3):
    fc, fs, _ = transformer_inputs(case, num_layers)
    return fc, fs, synth.single_head_transformer_state(case["seed"], num_layers)


def decoder_inputs(case: dict):
    h, w = case["hw"]
    x = synth.features(case["seed"], case["B"], 512, h, w, std=29.0, mean=-1.6)
    return x, synth.decoder_state(case["seed"])


def vit_inputs(case: dict):
    H, W = case["img"]
    return synth.image_u8(case["seed"] * 10 + 1, case["B"], H, W), synth.vit_state(case["seed"], pos_embedding=case["pos"])


def pipeline_inputs(case: dict):
    """content / style images and the three state dicts of infer_image.py:51-57."""
    s = case["seed"]
    c = synth.image_u8(s * 10 + 1, case["B"], *case["img"])
    st = synth.image_u8(s * 10 + 2, case["B"], *case["simg"])
    return c, st, synth.vit_state(s, pos_embedding=True), synth.vit_state(s + 500, pos_embedding=False), synth.transformer_state(s)


def token_sublattice(x: np.ndarray, sub: int) -> np.ndarray:
    """(B,C,h,w) -> (B,C,ceil(h*w/sub)): every `sub`-th token in row-major order."""
    b, c = x.shape[:2]
    return x.reshape(b, c, -1)[:, :, ::sub]


def pixel_sublattice(x: np.ndarray, sub: int) -> np.ndarray:
    return x[:, :, ::sub, ::sub]
