"""Module-level parity on the B200 against the golden vectors made by the unmodified reference
(tests/golden, oracle/gen_golden.py) and, at BASELINE.json sizes, against the oracle and through
size-independent properties.  Tolerances are BASELINE.json's: fp32 path max-abs <= 1e-3 (scaled with
the output range where the reference's own fp32-vs-fp64 deviation is already larger), bf16 path
<= 2e-2 relative (max-abs / absmax)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import cases, synth
from oracle import mhada_oracle as O

pytestmark = pytest.mark.gpu

import mhada_style_transfer_b200 as M  # noqa: E402
from mhada_style_transfer_b200.network import set_precision  # noqa: E402

DEV = "cuda:0"
FP32_MAX_ABS = 1e-3
BF16_REL = 2e-2


def dev(x, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype).to(DEV)


def fp32_tol(meta_key):
    # 1e-3 max-abs, or 3x the reference's own fp32-vs-fp64 deviation on this case when that is larger
    return max(FP32_MAX_ABS, 3.0 * meta_key["max_abs"])


def bf16_tol(case):
    """2e-2 (BASELINE.json) from ~100 tokens up.  Below that every output averages over too few keys /
    tokens for the independent bf16 roundings to cancel: an exact-arithmetic emulation of the same
    roundings (tools/error_budget.py) gives 2.0e-2 at 64 tokens and 1.4e-2 at 4096, so the tiny fixtures
    get 4e-2.  The stress case multiplies the logits by 16 (bf16 Q/K rounding then moves probability mass)."""
    if case.get("gain", 1.0) != 1.0:
        return 8e-2
    if case.get("heads", 8) == 4 and case.get("kind") == "transformer":
        # SIX chained 4-head layers (head_dim 128): the logits are sqrt(2) wider than with 8 heads at the default
        # initialisation, the attention is sharp, and the chain amplifies every rounding.  An exact-arithmetic emulation
        # with ONLY the input feature maps rounded to bf16 already deviates 2.4e-2 from the reference's float64 result on
        # transformer_h4_32x32_sub, and with every storage rounding of the pipeline 5.0e-2 / 2.55e-2 Frobenius
        # (tools/error_budget.py transformer_h4_32x32_sub) -- the kernels measure 5.03e-2 / 2.57e-2: they add nothing to
        # what bf16 storage costs.  North-star 2e-2 holds per layer (layer_c512_h4_12x12) and for the 8-head chain.
        return 6e-2
    ntok = min(case["hw"][0] * case["hw"][1], case["hsws"][0] * case["hsws"][1])
    return BF16_REL if ntok >= 100 else 4e-2


def build_layer(case, sd):
    m = M.AdaAttnMultiHead(case["C"], case["H"], case.get("activation", "softmax"))
    m.load_state_dict(synth.to_torch(sd, torch.float32), strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize("case", cases.LAYER_CASES, ids=lambda c: c["name"])
def test_layer_fp32_vs_reference_golden(case, golden_index):
    fc, fs, fcs, sd = cases.layer_inputs(case)
    m = build_layer(case, sd)
    tfc, tfs = dev(fc), dev(fs)
    tfcs = tfc if case.get("fcs_is_fc") else dev(fcs)
    with torch.no_grad():
        out = m(tfc, tfs, tfcs)
    assert out.shape == tfc.shape and out.dtype == torch.float32
    e = O.errors(out.cpu().numpy(), load_golden(case["name"])["out"])
    assert e["max_abs"] <= fp32_tol(golden_index[case["name"]]["ref32_vs_ref64"]), e


@pytest.mark.parametrize("case", [c for c in cases.LAYER_CASES if c["C"] // c["H"] in (64, 128, 512) and "activation" not in c],
                         ids=lambda c: c["name"])
def test_layer_bf16_vs_reference_golden(case, golden_index):
    fc, fs, fcs, sd = cases.layer_inputs(case)
    m = build_layer(case, sd)
    m.precision = "bf16"
    tfc, tfs = dev(fc), dev(fs)
    tfcs = tfc if case.get("fcs_is_fc") else dev(fcs)
    with torch.no_grad():
        out = m(tfc, tfs, tfcs)
    assert out.dtype == torch.float32
    e = O.errors(out.cpu().numpy(), load_golden(case["name"])["out"])
    tol = bf16_tol(case)
    assert e["max_abs_rel"] <= tol, e
    # bf16 tensors in -> bf16 out on the "auto" path
    m.precision = "auto"
    with torch.no_grad():
        ob = m(tfc.bfloat16(), tfs.bfloat16(), tfcs.bfloat16())
    assert ob.dtype == torch.bfloat16
    eb = O.errors(ob.float().cpu().numpy(), load_golden(case["name"])["out"])
    assert eb["max_abs_rel"] <= 2 * tol, eb


def test_layer_accepts_nchw_and_channels_last():
    case = cases.by_name("layer_c512_h8_ragged")
    fc, fs, fcs, sd = cases.layer_inputs(case)
    m = build_layer(case, sd)
    a = [dev(x) for x in (fc, fs, fcs)]
    b = [t.contiguous(memory_format=torch.channels_last) for t in a]
    with torch.no_grad():
        o1, o2 = m(*a), m(*b)
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("case", cases.ADAATTN_CASES, ids=lambda c: c["name"])
def test_adaattn_fp32(case, golden_index):
    fc, fs, fcs, sd = cases.adaattn_inputs(case)
    m = M.AdaAttN(case["C"], case.get("activation", "softmax"))
    m.load_state_dict(synth.to_torch(sd, torch.float32), strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        out = m(dev(fc), dev(fs), dev(fcs))
    e = O.errors(out.cpu().numpy(), load_golden(case["name"])["out"])
    assert e["max_abs"] <= fp32_tol(golden_index[case["name"]]["ref32_vs_ref64"]), e


@pytest.mark.parametrize("case", cases.SINGLE_HEAD_TRANSFORMER_CASES, ids=lambda c: c["name"])
def test_single_head_transformer_fp32(case, golden_index):
    """AdaAttnTransformer (adaDecoder.py:209-232): head_dim 512 runs on the fp32 kernels; decoded image vs reference."""
    fc, fs, sd = cases.single_head_transformer_inputs(case)
    m = M.AdaAttnTransformer()
    m.load_state_dict(synth.to_torch(sd, torch.float32), strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        cs = m([dev(x) for x in fc], [dev(x) for x in fs])
    meta = golden_index[case["name"]]
    e = O.errors(cases.pixel_sublattice(cs.cpu().numpy(), case["img_sub"]), load_golden(case["name"])["cs"])
    assert e["max_abs_rel"] <= max(1e-4, 5 * meta["cs_ref32_vs_ref64"]["max_abs_rel"]), e


@pytest.mark.parametrize("case", cases.FORLOSS_CASES, ids=lambda c: c["name"])
def test_forloss_fp32(case, golden_index):
    args = [dev(a) for a in cases.forloss_inputs(case)]
    m = M.AdaAttnForLoss(case["v"], case["qk"], case.get("activation", "softmax")).to(DEV).eval()
    with torch.no_grad():
        out = m(*args)
    e = O.errors(out.cpu().numpy(), load_golden(case["name"])["out"])
    assert e["max_abs"] <= fp32_tol(golden_index[case["name"]]["ref32_vs_ref64"]), e


@pytest.mark.parametrize("case", [c for c in cases.FORLOSS_CASES if "activation" not in c], ids=lambda c: c["name"])
def test_forloss_bf16_tensor_core_path(case, golden_index):
    """AdaAttnForLoss on the tensor cores (mhada_forloss_forward: logits materialised per image, every contraction on the
    tcgen05 token GEMM, Q / K split in two bf16 terms, exact V^2) against the reference goldens of the three VGG shapes
    (d_qk 448 / 960 / 1472).  bf16 INPUTS are judged against the float64 oracle evaluated on those rounded inputs: the
    rounding of a caller's tensors is not the module's error (these logits amplify it to several per cent)."""
    args = [dev(a) for a in cases.forloss_inputs(case)]
    m = M.AdaAttnForLoss(case["v"], case["qk"]).to(DEV).eval()
    m.precision = "bf16"
    args16 = [a.bfloat16() for a in args]
    with torch.no_grad():
        out = m(*args)
        auto = M.AdaAttnForLoss(case["v"], case["qk"]).to(DEV).eval()(*args16)      # bf16 in -> same path
    assert out.dtype == torch.float32 and auto.dtype == torch.bfloat16
    e = O.errors(out.cpu().numpy(), load_golden(case["name"])["out"])
    assert e["max_abs_rel"] <= BF16_REL and e["fro_rel"] <= 5e-3, e
    want16 = O.ada_attn_for_loss(*[a.float().cpu().numpy().astype(np.float64) for a in args16])
    e16 = O.errors(auto.float().cpu().numpy(), want16)
    assert e16["max_abs_rel"] <= BF16_REL and e16["fro_rel"] <= 5e-3, e16


def test_forloss_bf16_vs_oracle_at_training_size():
    """relu3_1 shape at 1024 x 900 tokens (ragged key count -> padded key tile), batch 2, against the float64 oracle."""
    case = dict(B=2, v=256, qk=448, hw=(32, 32), hsws=(30, 30), seed=45)
    args = cases.forloss_inputs(case)
    want = O.ada_attn_for_loss(*args)
    m = M.AdaAttnForLoss(256, 448).to(DEV).eval()
    m.precision = "bf16"
    with torch.no_grad():
        got = m(*[dev(a) for a in args])
        m.precision = "fp32"
        ref32 = m(*[dev(a) for a in args])
    e = O.errors(got.cpu().numpy(), want)
    assert e["max_abs_rel"] <= BF16_REL and e["fro_rel"] <= 5e-3, e
    # fp32 kernels on the same case: sharp rows make Var = E - M^2 cancel, so fp32 arithmetic (the reference's included,
    # SURVEY D8) is good to ~2^-12 |v| here, not to 1e-3 absolute: absmax 16 -> 3e-3
    e32 = O.errors(ref32.cpu().numpy(), want)
    assert e32["max_abs"] <= max(FP32_MAX_ABS, 3e-4 * e32["absmax"]), e32
    with torch.no_grad(), pytest.raises(NotImplementedError):
        mc = M.AdaAttnForLoss(256, 448, "cosine").to(DEV)
        mc.precision = "bf16"
        mc(*[dev(a) for a in args])


@pytest.mark.parametrize("case", cases.DECODER_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_decoder_vs_reference_golden(case, dtype, tol, golden_index):
    """Decoder.forward (conv.py:96-100): fused reflect-pad / up-sample kernels + cuDNN convolutions."""
    x, sd = cases.decoder_inputs(case)
    m = M.Decoder()
    m.load_state_dict({k[len("decoder."):]: v for k, v in synth.to_torch(sd, torch.float32).items()}, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        out = m(dev(x, dtype))
    assert out.shape == (case["B"], 3, 8 * case["hw"][0], 8 * case["hw"][1])
    e = O.errors(out.float().cpu().numpy(), load_golden(case["name"])["out"])
    assert e["max_abs_rel"] <= tol, e


def build_transformer(sd, precision="auto", heads=8):
    m = M.AdaAttnTransformerMultiHead(num_heads=heads)
    m.load_state_dict(synth.to_torch(sd, torch.float32), strict=True)
    m = m.to(DEV).eval()
    return set_precision(m, precision)


@pytest.mark.parametrize("case", cases.TRANSFORMER_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_transformer_vs_reference_golden(case, precision, golden_index):
    """6 MHAda layers + decoder (adaDecoder.py:253-268) against the UNMODIFIED reference's float64 output, at every
    size BASELINE.json states a target on: transformer_64x64_sub = configs[0] (512x512 image -> 4096 tokens),
    _b8_64x64 = configs[1] (batch 8), _128x128 = configs[2] (1024^2 -> 16384 tokens, streaming softmax),
    _135x240_x_64x64 = configs[3] (1080p frame x 512^2 style), _h4_* = the 4-head chain of configs[4]."""
    fc, fs, sd = cases.transformer_inputs(case)
    m = build_transformer(sd, precision, case.get("heads", 8))
    with torch.no_grad():
        fcs, cs = m([dev(x) for x in fc], [dev(x) for x in fs])          # call form 1
        fcs2, cs2 = m(([dev(x) for x in fc], [dev(x) for x in fs]))      # call form 2 (ptflops style)
    assert torch.equal(fcs, fcs2) and torch.equal(cs, cs2)
    B = case["B"]
    h, w = case["hw"]
    assert fcs.shape == (B, 512, h, w) and cs.shape == (B, 3, 8 * h, 8 * w)
    g = load_golden(case["name"])
    meta = golden_index[case["name"]]
    ef = O.errors(cases.token_sublattice(fcs.cpu().numpy(), case["sub"]), g["fcs"])
    ec = O.errors(cases.pixel_sublattice(cs.cpu().numpy(), case["img_sub"]), g["cs"])
    print(case["name"], precision, "fcs", ef, "cs", ec)
    if meta["cs"]["absmax"] == 0.0:
        # the reference's decoded image is identically zero for this seed (the final ReLU clips everything):
        # relative error is undefined, the image must be (near) zero too
        assert float(cs.float().abs().max()) <= 1e-3, float(cs.float().abs().max())
        ec = {"max_abs_rel": 0.0}
    if precision == "fp32":
        assert ef["max_abs"] <= fp32_tol(meta["fcs_ref32_vs_ref64"]), ef
        assert ec["max_abs_rel"] <= 1e-4, ec
    else:
        assert ef["max_abs_rel"] <= bf16_tol(case), ef
        assert ec["max_abs_rel"] <= bf16_tol(case), ec
        assert ef["fro_rel"] <= (3e-2 if case.get("heads", 8) == 4 else bf16_tol(case)), ef


# ---------------------------------------------------------------------------------------------------
# BASELINE.json sizes: oracle where it finishes in seconds, properties above that
# ---------------------------------------------------------------------------------------------------

def test_layer_512px_vs_oracle():
    """One layer at configs[0]/[1] size (64x64 tokens) against the float64 oracle, both paths."""
    case = dict(B=1, C=512, H=8, hw=(64, 64), hsws=(64, 64), seed=71, gain=1.0)
    fc, fs, fcs, sd = cases.layer_inputs(case)
    want = O.ada_attn_multi_head(fc, fs, fcs, sd, 8)
    m = build_layer(case, sd)
    with torch.no_grad():
        o32 = m(dev(fc), dev(fs), dev(fcs))
        m.precision = "bf16"
        o16 = m(dev(fc), dev(fs), dev(fcs))
    e32, e16 = O.errors(o32.cpu().numpy(), want), O.errors(o16.cpu().numpy(), want)
    assert e32["max_abs"] <= FP32_MAX_ABS, e32
    assert e16["max_abs_rel"] <= BF16_REL, e16


@pytest.mark.parametrize("B,hw,hsws", [(8, (64, 64), (64, 64)), (1, (128, 128), (128, 128)), (1, (135, 240), (64, 64))])
def test_full_size_properties(B, hw, hsws):
    """configs[1] (batch 8 @512^2), configs[2] (1024^2 -> 16384 tokens), configs[3] (1080p frame x 512^2
    style): (a) tensor-core path vs fp32 SIMT path -- two independent implementations -- within the bf16
    tolerance on a token sub-lattice; (b) images are independent: element 0 of the batch equals the B=1 run
    bit for bit; (c) constant style => S = sqrt(1e-6), M = V: out = 1e-3 * IN(fcs) + const, closed form."""
    case = dict(B=B, C=512, H=8, hw=hw, hsws=hsws, seed=81, gain=1.0)
    fc, fs, fcs, sd = cases.layer_inputs(case)
    m = build_layer(case, sd)
    tfc, tfs, tfcs = dev(fc, torch.bfloat16), dev(fs, torch.bfloat16), dev(fcs, torch.bfloat16)
    with torch.no_grad():
        o16 = m(tfc, tfs, tfcs)                                   # auto -> tensor cores
        m.precision = "fp32"
        # fp32 path on a query subset keeps the SIMT run short at 16k / 32k tokens
        rows = min(hw[0], 8)
        o32 = m(tfc[:1, :, :rows].float(), tfs[:1].float(), tfcs[:1, :, :rows].float())
    # (a) NOTE: instance-norm statistics of fc / fcs are taken over the tokens present, so compare on a
    # run where both paths see the same content tokens
    with torch.no_grad():
        m.precision = "bf16"
        o16_sub = m(tfc[:1, :, :rows], tfs[:1], tfcs[:1, :, :rows])
    e = O.errors(o16_sub.float().cpu().numpy(), o32.cpu().numpy())
    assert e["max_abs_rel"] <= BF16_REL, e
    # (b)
    with torch.no_grad():
        o1 = m(tfc[:1], tfs[:1], tfcs[:1])
    assert torch.equal(o1, o16[:1])
    # (c)
    const_fs = tfs[:1, :, :1, :1].expand(-1, -1, hsws[0], hsws[1]).contiguous()
    with torch.no_grad():
        oc = m(tfc[:1], const_fs, tfcs[:1]).float()
        m.precision = "fp32"
        oc32 = m(tfc[:1, :, :rows].float(), const_fs.float(), tfcs[:1, :, :rows].float())
    with torch.no_grad():
        _check_constant_style(m, tfcs, const_fs, rows, oc, oc32)


def _check_constant_style(m, tfcs, const_fs, rows, oc, oc32):
    """Expected value in float64 on the CPU (torch's GPU convs would use TF32)."""
    f64 = lambda t: t.detach().double().cpu()
    wh = torch.stack([f64(m.h_list[i].weight).reshape(64, 64) for i in range(8)])           # [8,64,64]
    bh = torch.stack([f64(m.h_list[i].bias) for i in range(8)])
    fs0 = f64(const_fs[0, :, 0, 0]).reshape(8, 64)
    v = (torch.einsum("hoi,hi->ho", wh, fs0) + bh).reshape(1, 512, 1, 1)                    # V, constant per channel
    wo, bo = f64(m.out_conv.weight).reshape(512, 512), f64(m.out_conv.bias)

    def closed_form(x):
        x = f64(x)
        xin = (x - x.mean((2, 3), keepdim=True)) / torch.sqrt(x.var((2, 3), unbiased=False, keepdim=True) + 1e-5)
        heads = 1e-3 * xin + v                     # S = sqrt(1e-6), M = V
        return torch.einsum("oc,bchw->bohw", wo, heads) + bo.reshape(1, -1, 1, 1)

    ec = O.errors(oc.cpu().numpy(), closed_form(tfcs[:1]).numpy())
    assert ec["max_abs_rel"] <= BF16_REL, ec
    ec32 = O.errors(oc32.cpu().numpy(), closed_form(tfcs[:1, :, :rows]).numpy())
    assert ec32["max_abs"] <= FP32_MAX_ABS, ec32


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_style_cache_matches_uncached(precision):
    """the style side computed once (precompute_style) gives bit-identical results to the plain call,
    and one style (batch 1) serves a batch of frames -- the infer_video.py situation (one style, many frames)."""
    case = dict(B=3, hw=(12, 20), hsws=(8, 8), seed=91)           # cross sizes like infer_video.py (Nc != Ns)
    fc, fs, sd = cases.transformer_inputs(case)
    m = build_transformer(sd, precision)
    dfc = [dev(x) for x in fc]
    dfs_all = [dev(x) for x in fs]
    with torch.no_grad():
        ref_fcs, ref_cs = m(dfc, dfs_all)
        cache = m.precompute_style(dfs_all)
        got_fcs, got_cs = m(dfc, cache)
    assert torch.equal(ref_fcs, got_fcs) and torch.equal(ref_cs, got_cs)
    # one style for all frames: equals the plain call with that style repeated per frame
    dfs_one = [t[:1] for t in dfs_all]
    with torch.no_grad():
        cache1 = m.precompute_style(dfs_one)
        b_fcs, b_cs = m(dfc, cache1)
        rep_fcs, rep_cs = m(dfc, [t.expand(3, -1, -1, -1).contiguous() for t in dfs_one])
    assert cache1.style_batch == 1
    assert torch.equal(b_fcs, rep_fcs) and torch.equal(b_cs, rep_cs)
    # a stale cache (weights changed) is refused
    with torch.no_grad():
        m.adaAttnHead[0].out_conv.bias.add_(1.0)
    with torch.no_grad(), pytest.raises(RuntimeError, match="stale"):
        m(dfc, cache1)


def test_four_heads_tensor_core_path():
    """num_heads=4 (head_dim 128) on the tcgen05 kernels: every layer of the chain against the fp32 kernels of the same
    module ON THE SAME INPUTS (the fp32 kernels are golden-tested for 4 heads; with PyTorch's default init the
    128-wide logits are sharp and six chained bf16 layers drift apart by ~8 %, so the chain is not compared end to
    end), and the style cache bit-identical to the plain call."""
    torch.manual_seed(5)
    m = M.AdaAttnTransformerMultiHead(num_heads=4).to(DEV).eval()
    case = dict(B=2, hw=(12, 20), hsws=(16, 9), seed=92)
    fc, fs, _ = cases.transformer_inputs(case)
    dfc, dfs = [dev(x) for x in fc], [dev(x) for x in fs]
    with torch.no_grad():
        fcs = dfc[0]
        for i in range(3):
            for k in range(2):
                L = m.adaAttnHead[2 * i + k]
                a_fc = dfc[i] if k == 0 else fcs
                L.precision = "fp32"
                ref = L(a_fc, dfs[i], fcs)
                L.precision = "bf16"
                got = L(a_fc, dfs[i], fcs)
                e = O.errors(got.cpu().numpy(), ref.cpu().numpy())
                assert e["max_abs_rel"] <= 4e-2 and e["fro_rel"] <= 1.5e-2, (2 * i + k, e)
                fcs = ref
        set_precision(m, "bf16")
        got_fcs, got_cs = m(dfc, dfs)
        cache = m.precompute_style(dfs)
        c_fcs, c_cs = m(dfc, cache)
    assert torch.isfinite(got_cs).all()
    assert torch.equal(got_fcs, c_fcs) and torch.equal(got_cs, c_cs)


def test_errors_on_device():
    m = M.AdaAttnMultiHead(512, 8, activation="cosine").to(DEV)
    m.precision = "bf16"                # the cosine activation exists on the fp32 kernels only
    x = torch.zeros(1, 512, 4, 4, device=DEV)
    with torch.no_grad(), pytest.raises(NotImplementedError):
        m(x, x, x)
    m2 = M.AdaAttnMultiHead(384, 2).to(DEV)       # head_dim 192: neither the streaming kernel nor a multiple of 128
    m2.precision = "bf16"
    x2 = torch.zeros(1, 384, 4, 4, device=DEV)
    with torch.no_grad(), pytest.raises(NotImplementedError):
        m2(x2, x2, x2)


@pytest.mark.parametrize("H,B,hw,hsws", [(1, 2, (32, 32), (30, 30)), (2, 1, (24, 40), (16, 16)), (1, 1, (64, 64), (64, 64))])
def test_wide_head_layers_on_the_tensor_cores(H, B, hw, hsws):
    """head_dim 512 / 256 (1- and 2-head AdaAttnMultiHead: BASELINE configs[4] sweeps 1 / 4 / 8 heads): per-head
    projections on the token GEMM + materialised tensor-core attention with split operands, against the float64 oracle;
    ragged key counts, cross sizes, 4096 x 4096 tokens.  The fp32 SIMT kernels stay the reference-arithmetic path."""
    case = dict(B=B, C=512, H=H, hw=hw, hsws=hsws, gain=1.0, seed=31 + H)
    fc, fs, fcs, sd = cases.layer_inputs(case)
    want = O.ada_attn_multi_head(fc, fs, fcs, sd, H)
    m = build_layer(case, sd)
    m.precision = "bf16"
    t16 = [dev(x).bfloat16() for x in (fc, fs, fcs)]
    with torch.no_grad():
        got = m(dev(fc), dev(fs), dev(fcs))
        auto = m(*t16)
    assert got.dtype == torch.float32 and auto.dtype == torch.bfloat16
    # (a) against the oracle on the inputs the kernels see (bf16 storage): what the PATH adds -- 2e-2 / 5e-3
    want16 = O.ada_attn_multi_head(*[t.float().cpu().numpy().astype(np.float64) for t in t16], sd, H)
    e16 = O.errors(got.cpu().numpy(), want16)
    assert e16["max_abs_rel"] <= BF16_REL and e16["fro_rel"] <= 5e-3, e16
    # (b) against the oracle on the unrounded inputs: the bf16 rounding of the feature maps in front of 256- / 512-wide
    # logits (std ~ sqrt(d): sharper rows than at 64) costs up to 3e-2 by itself
    e = O.errors(got.cpu().numpy(), want)
    assert e["max_abs_rel"] <= 3.5e-2 and e["fro_rel"] <= 1.5e-2, e
    ea = O.errors(auto.float().cpu().numpy(), want16)
    assert ea["max_abs_rel"] <= BF16_REL, ea


@pytest.mark.parametrize("heads", [2, 1])
def test_wide_head_transformer_chain(heads):
    """AdaAttnTransformerMultiHead with 2 / 1 heads (head_dim 256 / 512), six chained layers + decoder, precision "bf16":
    the second layer of every level reuses the style statistics of the first (MHADA_REUSE_FS_STATS on the wide-head path).
    Against the float64 oracle on the bf16-rounded inputs (what the kernels see): the chain amplifies every rounding
    (3e-2 Frobenius as for the 4-head chain); and the cached-style call (fp32 kernels for these widths) agrees."""
    case = dict(B=1, hw=(24, 20), hsws=(16, 18), seed=57, heads=heads)
    fc, fs, sd = cases.transformer_inputs(case)
    m = build_transformer(sd, "bf16", heads=heads)
    t16 = lambda xs: [dev(x).bfloat16() for x in xs]
    r64 = lambda xs: [t.float().cpu().numpy().astype(np.float64) for t in t16(xs)]
    with torch.no_grad():
        fcs, cs = m(t16(fc), t16(fs))
        cache = m.precompute_style(t16(fs))
        fcs_c, cs_c = m(t16(fc), cache)
    want_fcs, want_cs = O.transformer_multi_head(r64(fc), r64(fs), sd, num_heads=heads)
    e = O.errors(fcs.float().cpu().numpy(), want_fcs)
    print("wide chain heads", heads, e)
    # Frobenius is the criterion; the maximum is dominated by single tokens whose (512-wide, sharp) attention row picks a
    # different key after the bf16 rounding of the previous layer's output
    assert e["fro_rel"] <= 3e-2 and e["max_abs_rel"] <= (0.25 if heads == 1 else 0.1), e
    ec = O.errors(cs.float().cpu().numpy(), want_cs)
    assert ec["max_abs_rel"] <= 0.1 and ec["fro_rel"] <= 3e-2, ec        # measured 6.4e-2 / 2.4e-2 (bf16 decoder on top of the chain)
    assert cache.dtype == torch.float32
    e2 = O.errors(fcs_c.float().cpu().numpy(), want_fcs)
    assert e2["max_abs_rel"] <= 2e-2, e2


def test_single_head_modules_on_the_tensor_cores(golden_index):
    """AdaAttN / AdaAttnTransformer (head_dim 512, adaDecoder.py:85-131, :209-232) with precision="bf16" against the
    reference goldens (48 tokens: the small-fixture tolerance)."""
    case = cases.SINGLE_HEAD_TRANSFORMER_CASES[0]
    fc, fs, sd = cases.single_head_transformer_inputs(case)
    m = M.AdaAttnTransformer()
    m.load_state_dict(synth.to_torch(sd, torch.float32), strict=True)
    m = set_precision(m.to(DEV).eval(), "bf16")
    with torch.no_grad():
        cs = m([dev(x) for x in fc], [dev(x) for x in fs])
    e = O.errors(cases.pixel_sublattice(cs.float().cpu().numpy(), case["img_sub"]), load_golden(case["name"])["cs"])
    assert e["max_abs_rel"] <= 4e-2, e


# ---------------------------------------------------------------------------------------------------
# training step (BASELINE configs[4]): kernels forward, gradients against the reference's float64 autograd
# ---------------------------------------------------------------------------------------------------

def _as_golden(case, key, t):
    """Gradient tensor -> numpy in the form the golden stores it (input gradients on every sub-th token)."""
    a = t.float().cpu().numpy()
    if "sub" in case and key in ("fc", "fs", "fcs"):
        a = cases.token_sublattice(a, case["sub"])
    return a


@pytest.mark.parametrize("case", cases.GRAD_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4)])
def test_layer_gradients_vs_reference_autograd(case, precision, tol, golden_index):
    fc, fs, fcs, sd, G = cases.grad_inputs(case)
    m = build_layer(case, sd)
    m.precision = precision
    tin = [dev(a).requires_grad_(True) for a in (fc, fs, fcs)]
    out = m(*tin)                                   # grad mode: CUDA forward + autograd node
    assert out.requires_grad
    e = O.errors(out.detach().cpu().numpy(), O.ada_attn_multi_head(fc, fs, fcs, sd, case["H"]))
    assert e["max_abs"] <= FP32_MAX_ABS, e
    (out * dev(G)).sum().backward()
    g = load_golden(case["name"])
    named = dict(m.named_parameters())
    got = {"fc": tin[0].grad, "fs": tin[1].grad, "fcs": tin[2].grad}
    for k in cases.GRAD_KEYS[3:]:
        got[k] = named[k].grad
    # d/d(g bias) is exactly zero (a constant added to every logit of a row cancels in the softmax): judge every
    # gradient against its own range plus 1e-6 of the largest gradient in the layer
    scale = max(float(np.abs(g[k.replace(".", "__")]).max()) for k in got)
    for k, t in got.items():
        eg = O.errors(_as_golden(case, k, t), g[k.replace(".", "__")])
        assert eg["max_abs"] <= tol * eg["absmax"] + 1e-6 * scale, (k, eg)


def test_layer_gradients_bf16_forward():
    """bf16 tensor-core forward under autograd: gradients are those of the fp32 function at the same point."""
    case = cases.GRAD_CASES[0]
    fc, fs, fcs, sd, G = cases.grad_inputs(case)
    m = build_layer(case, sd)
    m.precision = "bf16"
    m.backward_impl = "torch"                       # the fp32 recompute backward behind the bf16 kernels forward
    tin = [dev(a).requires_grad_(True) for a in (fc, fs, fcs)]
    out = m(*tin)
    (out * dev(G)).sum().backward()
    g = load_golden(case["name"])
    assert O.errors(tin[0].grad.cpu().numpy(), g["fc"])["max_abs_rel"] <= 2e-4
    assert O.errors(m.out_conv.weight.grad.cpu().numpy(), g["out_conv__weight"])["max_abs_rel"] <= 2e-4


def _grads_of(m, tin, G_):
    out = m(*tin)
    (out.float() * G_).sum().backward()
    got = {"fc": tin[0].grad, "fs": tin[1].grad, "fcs": tin[2].grad}
    got.update({k: p.grad for k, p in m.named_parameters()})
    return out.detach(), got


def _check_kernel_grads(got, want, max_rel, fro_rel):
    """bf16 gradients: each within max_rel of its own range plus 5e-3 of the largest gradient of the layer (d/d(g bias)
    is exactly zero in exact arithmetic: a constant added to every logit of a row cancels in the softmax), and within
    fro_rel in Frobenius norm unless it is one of those vanishing gradients."""
    scale = max(float(np.abs(v).max()) for v in want.values())
    for k, w in want.items():
        eg = O.errors(got[k], w)
        assert eg["max_abs"] <= max_rel * eg["absmax"] + 5e-3 * scale, (k, eg)
        if eg["absmax"] > 1e-2 * scale:
            assert eg["fro_rel"] <= fro_rel, (k, eg)


@pytest.mark.parametrize("case", cases.GRAD_CASES, ids=lambda c: c["name"])
def test_layer_backward_kernels_vs_reference_autograd(case, golden_index):
    """forward AND backward on own kernels (mhada_layer_backward: flash-style attention backward, the other
    contractions on the tcgen05 GEMM) against the reference's float64 autograd gradients, at the shape of the training
    step (2 x 1024 tokens, 8 heads: 3e-2 / 2e-2) and on a small ragged fixture.  100 x 72 tokens: the bf16
    rounding of the INPUT feature maps alone moves the instance-norm statistics by a per cent at this token count
    (forward parity on these fixtures is 2-4e-2 too); measured 7e-2 max / 2e-2 Frobenius."""
    fc, fs, fcs, sd, G = cases.grad_inputs(case)
    m = build_layer(case, sd)
    m.precision = "bf16"
    m.backward_impl = "kernels"
    tin = [dev(a).requires_grad_(True) for a in (fc, fs, fcs)]
    _, got = _grads_of(m, tin, dev(G))
    g = load_golden(case["name"])
    big = case["hw"][0] * case["hw"][1] >= 1024            # the training-step shape: 3e-2 / 2e-2
    _check_kernel_grads({k: _as_golden(case, k, got[k]) for k in cases.GRAD_KEYS},
                        {k: g[k.replace(".", "__")] for k in cases.GRAD_KEYS}, 3e-2 if big else 1.2e-1, 2e-2 if big else 3.5e-2)


@pytest.mark.parametrize("B,H,hw,hsws,alias,max_rel,fro_rel", [(2, 8, (32, 32), (32, 32), False, 3e-2, 2e-2),
                                                               (1, 8, (20, 13), (17, 9), True, 1.5e-1, 4e-2)])
def test_layer_backward_kernels_vs_recompute(B, H, hw, hsws, alias, max_rel, fro_rel):
    """The kernel backward against the fp32 PyTorch recompute backward of the same module at the training resolution
    (1024 tokens, 8 heads: 3e-2 / 2e-2), and on small ragged sizes with fcs aliasing fc (layer 0 of the transformer,
    adaDecoder.py:262; few tokens -> looser, see the test above)."""
    C = 64 * H
    case = dict(B=B, C=C, H=H, hw=hw, hsws=hsws, gain=1.0, seed=73)
    fc, fs, fcs, sd = cases.layer_inputs(case)
    G_ = dev(synth.bellish(997, fc.shape, 0.0, 1.0))
    res = {}
    for impl in ("kernels", "torch"):
        m = build_layer(case, sd)
        m.precision = "bf16"
        m.backward_impl = impl
        a, b_ = dev(fc).requires_grad_(True), dev(fs).requires_grad_(True)
        c = a if alias else dev(fcs).requires_grad_(True)
        out = m(a, b_, c)
        (out.float() * G_).sum().backward()
        res[impl] = {"fc": a.grad, "fs": b_.grad, **({} if alias else {"fcs": c.grad}),
                     **{k: p.grad for k, p in m.named_parameters()}}
    _check_kernel_grads({k: v.float().cpu().numpy() for k, v in res["kernels"].items()},
                        {k: v.float().cpu().numpy() for k, v in res["torch"].items()}, max_rel, fro_rel)


def test_cosine_layer_gradient_matches_oracle_finite_difference():
    """activation="cosine" under autograd: CUDA forward (fp32 kernels), recompute backward; d(loss)/d(fc) and
    d(loss)/d(fs) against directional finite differences of the float64 oracle."""
    case = cases.by_name("layer_c128_h2_cosine")
    fc, fs, fcs, sd = cases.layer_inputs(case)
    m = build_layer(case, sd)
    G = synth.bellish(991, fc.shape, 0.0, 1.0)
    tin = [dev(a).requires_grad_(True) for a in (fc, fs, fcs)]
    out = m(*tin)
    assert O.errors(out.detach().cpu().numpy(), load_golden(case["name"])["out"])["max_abs"] <= FP32_MAX_ABS
    (out * dev(G)).sum().backward()
    rng = np.random.default_rng(1)
    loss = lambda a, b: float((O.ada_attn_multi_head(a, b, fcs, sd, case["H"], activation="cosine") * G).sum())
    eps = 1e-5
    for idx, base in ((0, fc), (1, fs)):
        v = rng.standard_normal(base.shape)
        plus = [fc, fs]; minus = [fc, fs]
        plus[idx] = base + eps * v
        minus[idx] = base - eps * v
        fd = (loss(*plus) - loss(*minus)) / (2 * eps)
        an = float((tin[idx].grad.double().cpu().numpy() * v).sum())
        assert an == pytest.approx(fd, rel=2e-3, abs=1e-4 * abs(fd) + 1e-5), (idx, an, fd)


def test_forloss_gradient_matches_oracle_finite_difference():
    """AdaAttnForLoss under autograd (lossfn.py:26-34 feeds it VGG features): CUDA forward, recompute backward;
    gradients w.r.t. all four inputs against directional finite differences of the float64 oracle."""
    case = cases.by_name("forloss_relu5_1")
    args = cases.forloss_inputs(case)
    m = M.AdaAttnForLoss(case["v"], case["qk"]).to(DEV)
    tin = [dev(a).requires_grad_(True) for a in args]
    out = m(*tin)
    G = synth.bellish(993, out.shape, 0.0, 1.0)
    (out * dev(G)).sum().backward()
    rng = np.random.default_rng(2)
    loss = lambda *a: float((O.ada_attn_for_loss(*a) * G).sum())
    eps = 1e-5
    for idx in range(4):
        v = rng.standard_normal(args[idx].shape)
        plus, minus = list(args), list(args)
        plus[idx] = args[idx] + eps * v
        minus[idx] = args[idx] - eps * v
        fd = (loss(*plus) - loss(*minus)) / (2 * eps)
        an = float((tin[idx].grad.double().cpu().numpy() * v).sum())
        assert an == pytest.approx(fd, rel=2e-3, abs=1e-4 * abs(fd) + 1e-5), (idx, an, fd)


@pytest.mark.parametrize("B,hw", [(2, (16, 16)), (1, (9, 14))])
def test_decoder_training_path_vs_torch(B, hw):
    """Decoder under autograd (train_image.py:105-144): own kernels forward (bf16, the inference kernels) + aten backward
    ops, against the plain fp32 PyTorch op sequence of the same module.  Decoded image: 2e-2.  Gradients: bf16 activations
    through nine ReLU blocks flip masks, so the gradients of the EARLY blocks sit ~0.10-0.12 (Frobenius) from the fp32
    ones -- PyTorch's own bf16 decoder deviates by the same amount and as much from this path (measured,
    tools/debug_decoder_train.py) -- while the last block's agree to < 1e-2: bounds 0.2 everywhere, 2e-2 on the last block."""
    case = dict(cases.DECODER_CASES[0], B=B, hw=hw)
    x, sd = cases.decoder_inputs(case)
    m = M.Decoder()
    m.load_state_dict({k[len("decoder."):]: v for k, v in synth.to_torch(sd, torch.float32).items()}, strict=True)
    m = m.to(DEV).train()
    G_ = torch.randn(B, 3, 8 * hw[0], 8 * hw[1], device=DEV)
    res = {}
    for impl in ("kernels", "torch"):
        m.train_impl = impl
        m.zero_grad(set_to_none=True)
        xin = dev(x).requires_grad_(True)
        out = m(xin)
        assert out.dtype == torch.float32
        (out.float() * G_).sum().backward()
        res[impl] = (out.detach().float(), {"x": xin.grad, **{k: p.grad.clone() for k, p in m.named_parameters()}})
    e = O.errors(res["kernels"][0].cpu().numpy(), res["torch"][0].cpu().numpy())
    assert e["max_abs_rel"] <= 2e-2, e
    for k, want in res["torch"][1].items():
        got = res["kernels"][1][k]
        assert torch.isfinite(got).all(), k
        fro = float((got.float() - want.float()).norm() / want.float().norm())
        assert fro <= (2e-2 if k.startswith("conv3.1.") else 0.2), (k, fro)
    m.train_impl, m.precision = "auto", "auto"                     # fp32 features, precision not set: the fp32 torch path
    m.zero_grad(set_to_none=True)
    out = m(dev(x).requires_grad_(True))
    assert torch.equal(out.detach().float(), res["torch"][0])


def test_single_head_transformer_trains():
    """AdaAttnTransformer under autograd: every parameter receives a finite gradient."""
    case = cases.SINGLE_HEAD_TRANSFORMER_CASES[0]
    fc, fs, sd = cases.single_head_transformer_inputs(case)
    m = M.AdaAttnTransformer()
    m.load_state_dict(synth.to_torch(sd, torch.float32), strict=True)
    m = m.to(DEV).train()
    cs = m([dev(x) for x in fc], [dev(x) for x in fs])
    cs.float().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all().item() for p in m.parameters())


def test_transformer_train_step_runs():
    """forward + backward through all six layers and the decoder (train_image.py:105-144 shape of use)."""
    case = dict(B=2, hw=(8, 8), hsws=(8, 8), seed=95)
    fc, fs, sd = cases.transformer_inputs(case)
    m = build_transformer(sd, "bf16")
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    fcs, cs = m([dev(x) for x in fc], [dev(x) for x in fs])
    loss = cs.float().mean() + 1e-3 * fcs.float().pow(2).mean()
    loss.backward()
    n_grad = sum(p.grad is not None and torch.isfinite(p.grad).all().item() for p in m.parameters())
    assert n_grad == len(list(m.parameters()))
    before = m.adaAttnHead[0].out_conv.weight.detach().clone()
    opt.step()
    assert not torch.equal(before, m.adaAttnHead[0].out_conv.weight)


def _ddp_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    from mhada_style_transfer_b200.sharding import allreduce_gradients, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)      # two ranks share the one test GPU
    torch.backends.cudnn.allow_tf32 = False       # training-mode decoder = stock cuDNN convs: keep them fp32 so the
    torch.backends.cuda.matmul.allow_tf32 = False # shard / full-batch difference is summation order only
    case = dict(B=4, hw=(8, 8), hsws=(8, 8), seed=96)
    fc, fs, sd = cases.transformer_inputs(case)
    m = build_transformer(sd, "bf16")
    s, e = shard_range(4, rank, world)
    fcs, cs = m([dev(x[s:e]) for x in fc], [dev(x[s:e]) for x in fs])
    (cs.float().sum() / 4).backward()                                  # global-batch mean
    for p in m.parameters():                                           # sum of shard gradients = full-batch gradient
        p.grad.mul_(world)
    allreduce_gradients(m)
    if rank == 0:
        full = build_transformer(sd, "bf16")
        fcs2, cs2 = full([dev(x) for x in fc], [dev(x) for x in fs])
        (cs2.float().sum() / 4).backward()
        scale = max(b.grad.abs().max().item() for b in full.parameters())     # some gradients are exactly zero
        worst = 0.0
        for a, b in zip(m.parameters(), full.parameters()):
            d = (a.grad - b.grad).abs().max().item()
            worst = max(worst, d / (b.grad.abs().max().item() + 1e-2 * scale))   # vanishing gradients (g biases): f32
                                                                               # summation order of O(scale) terms
        q.put(worst)
    dist.barrier()
    dist.destroy_process_group()


def test_train_step_gradient_allreduce():
    """Data-parallel training step: each rank runs forward + backward on its shard of the batch, gradients are
    averaged with an all-reduce (gloo here: two ranks on the single test GPU; NCCL on a multi-GPU box)."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    # images are independent through the path (instance norm is per image), so the shard gradients add up to the
    # full-batch gradient; what is left is cuDNN picking batch-size-dependent algorithms in the decoder backward
    assert q.get(timeout=10) <= 5e-3
