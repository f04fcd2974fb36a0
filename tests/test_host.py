"""Host-side mirror of the reference module surface + the C-ABI library (CPU only, no compute)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from oracle import synth

import mhada_style_transfer_b200 as M
from mhada_style_transfer_b200 import _lib, build as build_mod


@pytest.fixture(scope="module")
def lib():
    build_mod.build()
    return _lib.lib()


def test_state_dict_is_reference_compatible():
    # 318 keys, same names / shapes / order as the reference (SURVEY.md §5 checkpoint row)
    m = M.AdaAttnTransformerMultiHead()
    sd = synth.to_torch(synth.transformer_state(5), torch.float32)
    assert len(sd) == 318
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert m.load_state_dict(sd, strict=True).missing_keys == []
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    # single layer / single head variants
    l = M.AdaAttnMultiHead(512, 8)
    l.load_state_dict(synth.to_torch(synth.mhada_layer_state(3, 512, 8), torch.float32), strict=True)
    a = M.AdaAttN(64)
    a.load_state_dict(synth.to_torch(synth.adaattn_state(3, 64), torch.float32), strict=True)
    assert M.AdaAttnForLoss(256, 448).state_dict() == {}


def test_vit_state_dict_is_reference_compatible():
    """vit.py:120-146 key surface (what infer_image.py:55-56 loads with strict=True)."""
    for pos in (True, False):
        m = M.VisionTransformer(pos_embedding=pos)
        sd = synth.to_torch(synth.vit_state(3, pos_embedding=pos), torch.float32)
        assert sorted(m.state_dict().keys()) == sorted(sd.keys())
        m.load_state_dict(sd, strict=True)
        for k, v in m.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), k
    m = M.VisionTransformer()
    assert m.patch_size == 8 and m.num_layers == 3 and m.hidden_dim == 512 and len(m.encoder) == 3
    assert m.pos_embedding.pos_embed.shape == (1, 512, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"), torch.no_grad():
        m(torch.zeros(1, 3, 16, 16))


def test_attribute_surface():
    l = M.AdaAttnMultiHead(512, 8)
    for name in ("num_heads", "head_dim", "f_list", "g_list", "h_list", "norm_q_list", "norm_k_list",
                 "norm_v_out_list", "out_conv", "activation"):
        assert hasattr(l, name), name
    assert l.num_heads == 8 and l.head_dim == 64 and len(l.f_list) == 8
    assert l.f_list[0].weight.shape == (64, 64, 1, 1) and l.out_conv.weight.shape == (512, 512, 1, 1)
    t = M.AdaAttnTransformerMultiHead(num_layers=2, qkv_dim=256, num_heads=4)
    assert t.num_layers == 2 and len(t.adaAttnHead) == 4 and hasattr(t, "decoder")
    assert M.AdaAttnTransformer().num_layers == 3


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError):
        M.AdaAttnMultiHead(512, 7)                      # adaDecoder.py:137-138
    with pytest.raises(ValueError, match="Unknown activation"):
        M.AdaAttnMultiHead(512, 8, activation="relu")   # adaDecoder.py:160
    with pytest.raises(ValueError, match="Unknown activation"):
        M.AdaAttnForLoss(64, 64, "tanh")
    M.AdaAttnMultiHead(512, 8, activation="cosine")     # accepted at construction like the reference


def test_no_cpu_fallback():
    m = M.AdaAttnMultiHead(128, 2)
    x = torch.zeros(1, 128, 4, 4)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, x, x)
    with pytest.raises(RuntimeError):
        m(x, torch.zeros(2, 128, 4, 4), x)              # batch mismatch (adaDecoder.py:177-183)


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "mhada_b200.h")).read()
    declared = set(re.findall(r"MHADA_API\s+[\w\s\*]+?\b(mhada_\w+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mhada_abi_version() == _lib.ABI_VERSION == 13


def test_abi_rejects_without_gpu_or_bad_args(lib):
    # argument validation happens before any CUDA call, so it is testable on a CPU box
    rc = lib.mhada_in_stats(None, 0, 1, 1, 4, 4, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.mhada_last_error()
    assert lib.mhada_layer_workspace(_lib.BF16, 8, 4096, 4096, 512, 8) > 0
    assert lib.mhada_layer_workspace(_lib.BF16, 8, 4096, 4096, 512, 7) == 0
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.mhada_layer_forward(_lib.BF16, p, p, p, p, p, None, None, 1, 4, 4, 512, 4, 0, p, p, 1 << 30, None)
    assert rc == -1   # out aliases an input
    if not torch.cuda.is_available():
        assert lib.mhada_device_check() == -3


def test_workspace_formula_matches_design(lib):
    # bf16, cfg2: Q,K,heads [8,4096,512] bf16 + V' [8,4096,1024] bf16 dominate
    n = lib.mhada_layer_workspace(_lib.BF16, 8, 4096, 4096, 512, 8)
    tensors = 8 * 4096 * 512 * 2 * 5
    assert tensors <= n <= tensors + (16 << 20)          # + statistics partial sums (12.6 MB), folded weights, stats


def test_abi_argument_checks_of_the_newer_entry_points(lib):
    """Validation that happens before any CUDA call (so it runs on a CPU box): cosine / head-dim rules, the decoder
    block kernel, the stage profiler."""
    buf = (ctypes.c_float * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    aligned = ctypes.c_void_p((p.value + 63) & ~63)
    other = ctypes.c_void_p(aligned.value + 2048)
    # cosine is an fp32-path activation
    rc = lib.mhada_layer_forward(_lib.BF16, aligned, aligned, aligned, aligned, aligned, None, None, 1, 4, 4, 512, 8,
                                 _lib.LAYER_COSINE, other, other, 1 << 30, None)
    assert rc == -2 and b"cosine" in lib.mhada_last_error()
    # tensor-core path: head_dim 64, 128 and multiples of 128 (192 is neither)
    rc = lib.mhada_layer_forward(_lib.BF16, aligned, aligned, aligned, aligned, aligned, None, None, 1, 4, 4, 384, 2, 0,
                                 other, other, 1 << 30, None)
    assert rc == -2 and b"head_dim" in lib.mhada_last_error()
    # wide heads (256) pass the shape rules and need more workspace than the streaming path's layout
    assert lib.mhada_layer_workspace(_lib.BF16, 1, 64, 64, 512, 2) > lib.mhada_layer_workspace(_lib.BF16, 1, 64, 64, 512, 8)
    a = _lib.AttnArgs()
    a.dtype = _lib.F32
    a.B, a.H, a.Nc, a.Ns, a.dqk, a.dv = 1, 1, 4, 4, 64, 64
    for f in ("q", "k", "v", "x", "out", "x_mean", "x_rstd"):
        setattr(a, f, aligned.value)
    a.ldq = a.ldk = a.ldv = a.ldx = a.ldo = 64
    a.activation = 7
    assert lib.mhada_attn(ctypes.byref(a), None) == -1 and b"activation" in lib.mhada_last_error()
    a.dtype, a.activation = _lib.BF16, _lib.ACT_COSINE
    assert lib.mhada_attn(ctypes.byref(a), None) == -2
    # decoder block kernel
    assert lib.mhada_conv3x3_small(_lib.BF16, None, aligned, aligned, 1, 8, 8, 64, 3, 1, other, None) == -1
    assert lib.mhada_conv3x3_small(_lib.F32, aligned, aligned, aligned, 1, 8, 8, 64, 3, 1, other, None) == -2
    assert lib.mhada_conv3x3_small(_lib.BF16, aligned, aligned, aligned, 1, 1, 8, 64, 3, 1, other, None) == -1
    # stage profiler
    ms, n = ctypes.c_float(0), ctypes.c_int(0)
    assert lib.mhada_profile_stage(9, ctypes.byref(ms), ctypes.byref(n)) == -1
    assert lib.mhada_profile_stage(0, ctypes.byref(ms), ctypes.byref(n)) == 0


def test_backward_path_selection_is_host_logic():
    """which backward a layer takes (own kernels / fp32 recompute) is decided on the host from the precision,
    the head width and the activation; bad switches raise where they are read."""
    x16, x32 = torch.zeros(1, 512, 4, 4, dtype=torch.bfloat16), torch.zeros(1, 512, 4, 4)
    m = M.AdaAttnMultiHead(512, 8)
    assert m.backward_impl in ("auto", "kernels", "torch")
    m.backward_impl = "auto"
    assert m._kernel_backward(x16, x16, x16) is True                 # bf16 inputs, head_dim 64 -> mhada_layer_backward
    assert m._kernel_backward(x32, x32, x32) is False                # fp32 inputs -> fp32 path -> recompute backward
    m.precision = "bf16"
    assert m._kernel_backward(x32, x32, x32) is True
    m.backward_impl = "torch"
    assert m._kernel_backward(x16, x16, x16) is False
    m.backward_impl = "nope"
    with pytest.raises(ValueError, match="backward_impl"):
        m._kernel_backward(x16, x16, x16)
    m4 = M.AdaAttnMultiHead(512, 4)                                   # head_dim 128: tensor-core forward, recompute backward
    assert m4._kernel_backward(x16, x16, x16) is False
    m4.backward_impl = "kernels"
    with pytest.raises(NotImplementedError):
        m4._kernel_backward(x16, x16, x16)
    mc = M.AdaAttnMultiHead(512, 8, activation="cosine")
    assert mc._kernel_backward(x32, x32, x32) is False
    v = M.VisionTransformer()
    assert v.train_impl in ("auto", "kernels", "torch")


def test_backward_abi_argument_checks(lib):
    a = _lib.LayerBwdArgs()
    assert lib.mhada_layer_backward(None, None) == -1
    assert lib.mhada_layer_backward(ctypes.byref(a), None) == -1 and b"null" in lib.mhada_last_error()
    assert lib.mhada_layer_backward_workspace(8, 1024, 1024, 512, 8) > lib.mhada_layer_workspace(_lib.BF16, 8, 1024, 1024, 512, 8)
    assert lib.mhada_layer_backward_workspace(8, 1024, 1024, 512, 7) == 0
    assert lib.mhada_attn_bwd(1, 1, 1, 1, *([None] * 15)) == -1
    assert lib.mhada_batch_attn_bwd(None, None, 1, 1, 1, 64, None, None) == -1
    assert lib.mhada_transpose_bf16(None, 0, 4, 4, 4, 64, None, None) == -1
    assert lib.mhada_colsum(None, 0, 4, 4, None, 0, None, None) == -1
    assert lib.mhada_gemm_splitk_workspace(512, 512, 8192) > 0 and lib.mhada_gemm_splitk_workspace(512, 500, 8192) == 0
    assert lib.mhada_gemm_bf16_splitk(None, 64, None, 64, 1, 128, 64, None, 128, None, 0, None) == -1


def test_graph_wrapper_and_training_switches_reject_bad_arguments_on_the_host():
    from mhada_style_transfer_b200.graphs import GraphedStyleTransfer
    x = torch.zeros(1, 3, 16, 16)
    with pytest.raises(ValueError, match="style mode"):
        GraphedStyleTransfer(None, None, None, x, x, style="sometimes")
    with pytest.raises(RuntimeError, match="CUDA"):
        GraphedStyleTransfer(None, None, None, x, x)                      # no CPU path
    d = M.Decoder()
    assert d.train_impl in ("auto", "kernels", "torch") and d.precision == "auto"
    from mhada_style_transfer_b200.network import set_precision
    set_precision(M.AdaAttnTransformerMultiHead(), "bf16")               # reaches the decoder too
    t = set_precision(M.AdaAttnTransformerMultiHead(), "bf16")
    assert t.decoder.precision == "bf16" and all(l.precision == "bf16" for l in t.adaAttnHead)
    with pytest.raises(ValueError):
        set_precision(t, "fp16")
