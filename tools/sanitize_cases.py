#!/usr/bin/env python
"""Small, edge-heavy invocations of every hand-written kernel for compute-sanitizer (memcheck / racecheck / synccheck):
ragged tiles, a single key, the single-tile tail of the persistent attention kernel, head_dim 128, the GEMM's M < tile
and slab-reuse paths, the ViT small kernels and one whole image -> image pipeline.  Each case also checks its result
(the tests' own assertions), so a sanitizer run that passes is a parity run too.

    compute-sanitizer --tool racecheck python tools/sanitize_cases.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: F401,E402
import test_gpu_stages as S  # noqa: E402
import test_gpu_vit as V  # noqa: E402
from oracle import cases  # noqa: E402

idx = conftest.json.load(open(os.path.join(conftest.GOLDEN, "index.json")))
S._attn_bf16_case(1, 8, 135, 143, 0.6, 0, 64)        # ragged query and key tails
S._attn_bf16_case(1, 1, 70, 1, 0.6, 0, 64)           # one key
S._attn_bf16_case(20, 8, 200, 130, 0.6, 1.0, 64)     # persistent CTAs + single-tile tail items, reference-update path
S._attn_bf16_case(1, 2, 135, 143, 0.45, 2.0, 128)    # head_dim 128 (two chunks / two value slices)
S.test_linear(S.BF16, 130, 128, 64)
S.test_linear(S.BF16, 300, 512, 512)
S.test_in_stats(S.BF16, 1, 512, (9, 15))
S.test_conv3x3_tc(1, 3, 3, 64, 64, 1)
S.test_conv3x3_tc(2, 10, 14, 256, 256, 1)
S.test_conv3x3_tc(1, 20, 28, 256, 128, 0)
for args in ((300, 512, 192, "pos"), (70, 384, 64, "bf16"), (129, 512, 2048, "both"), (260, 2048, 512, "relu")):
    V.test_gemm(*args)
V.test_gemm(20000, 512, 512, "resid")                 # CTA pairs (cluster of two, tcgen05.mma.cta_group::2), ragged last pair tile
S.test_conv3x3_tc(3, 64, 64, 512, 256, 0)             # CTA-pair convolution
S.test_attn_bwd(2, 2, 100, 72, 1.0)                   # flash-style backward kernels: ragged tiles
S.test_attn_bwd(2, 2, 130, 1, 1.0)                    # one key
V.test_batch_attn_bwd(3, 37, 2)
V.test_layernorm(7, 128)
V.test_batch_attn(5, 7)
V.test_batch_attn(8, 64)
V.test_patch_im2col("u8")
V.test_vit_vs_reference_golden(cases.by_name("vit_b3_40x64_nopos"), idx)
V.test_pipeline_vs_reference_golden(cases.by_name("pipeline_b1_64x64"), idx)
print("sanitize cases ok")
