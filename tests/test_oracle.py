"""The numpy oracle against the golden vectors produced by the unmodified reference
(oracle/gen_golden.py).  CPU only.  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import cases, synth
from oracle import mhada_oracle as O

# golden arrays are the reference's float64 outputs rounded to float32 -> 2^-24 relative storage error
STORE_RTOL = 2e-7


def _check(got64, gold32, meta, key="out"):
    e = O.errors(got64, gold32)
    assert e["max_abs"] <= STORE_RTOL * max(meta[key]["absmax"], 1.0) + 1e-9, e
    return e


def _check_sums(got64, meta, key="out"):
    # float64 sums recorded from the reference's float64 run: pins the oracle to ~1e-10
    assert list(got64.shape) == meta[key]["shape"]
    assert got64.sum() == pytest.approx(meta[key]["sum"], rel=1e-9, abs=1e-7 * meta[key]["absmax"])
    assert (got64 * got64).sum() == pytest.approx(meta[key]["sumsq"], rel=1e-9)


@pytest.mark.parametrize("case", cases.LAYER_CASES, ids=lambda c: c["name"])
def test_layer_matches_reference(case, golden_index):
    fc, fs, fcs, sd = cases.layer_inputs(case)
    got = O.ada_attn_multi_head(fc, fs, fcs, sd, case["H"], activation=case.get("activation", "softmax"))
    meta = golden_index[case["name"]]
    _check(got, load_golden(case["name"])["out"], meta)
    _check_sums(got, meta)


@pytest.mark.parametrize("case", cases.ADAATTN_CASES, ids=lambda c: c["name"])
def test_adaattn_matches_reference(case, golden_index):
    fc, fs, fcs, sd = cases.adaattn_inputs(case)
    got = O.ada_attn(fc, fs, fcs, sd, activation=case.get("activation", "softmax"))
    meta = golden_index[case["name"]]
    _check(got, load_golden(case["name"])["out"], meta)
    _check_sums(got, meta)


@pytest.mark.parametrize("case", cases.FORLOSS_CASES, ids=lambda c: c["name"])
def test_forloss_matches_reference(case, golden_index):
    got = O.ada_attn_for_loss(*cases.forloss_inputs(case), activation=case.get("activation", "softmax"))
    meta = golden_index[case["name"]]
    _check(got, load_golden(case["name"])["out"], meta)
    _check_sums(got, meta)


@pytest.mark.parametrize("case", cases.SINGLE_HEAD_TRANSFORMER_CASES, ids=lambda c: c["name"])
def test_single_head_transformer_matches_reference(case, golden_index):
    fc, fs, sd = cases.single_head_transformer_inputs(case)
    got = O.transformer_single_head(fc, fs, sd)
    meta = golden_index[case["name"]]
    g = load_golden(case["name"])
    e = O.errors(cases.pixel_sublattice(got, case["img_sub"]), g["cs"])
    assert e["max_abs_rel"] <= 2e-6, e
    _check_sums(got, meta, key="cs")


@pytest.mark.parametrize("case", cases.DECODER_CASES, ids=lambda c: c["name"])
def test_decoder_matches_reference(case, golden_index):
    x, sd = cases.decoder_inputs(case)
    got = O.decoder(x, sd)
    meta = golden_index[case["name"]]
    _check(got, load_golden(case["name"])["out"], meta)
    _check_sums(got, meta)


@pytest.mark.parametrize("case", [c for c in cases.TRANSFORMER_CASES if c["hw"][0] <= 16], ids=lambda c: c["name"])
def test_transformer_matches_reference(case, golden_index):
    fc, fs, sd = cases.transformer_inputs(case)
    fcs, cs = O.transformer_multi_head(fc, fs, sd)
    meta = golden_index[case["name"]]
    g = load_golden(case["name"])
    _check(cases.token_sublattice(fcs, case["sub"]), g["fcs"], meta, "fcs")
    _check(cases.pixel_sublattice(cs, case["img_sub"]), g["cs"], meta, "cs")
    _check_sums(fcs, meta, "fcs")
    _check_sums(cs, meta, "cs")


@pytest.mark.parametrize("case", cases.VIT_CASES, ids=lambda c: c["name"])
def test_vit_matches_reference(case, golden_index):
    """oracle/vit_oracle.py (numpy float64 restatement of vit.py:45-169) against the unmodified reference's
    VisionTransformer, including the batch-axis attention (B > 1) and the resized positional table."""
    from oracle import vit_oracle as V
    img, sd = cases.vit_inputs(case)
    z = V.vision_transformer(img, sd)
    meta = golden_index[case["name"]]
    g = load_golden(case["name"])
    for l in range(3):
        _check(cases.token_sublattice(z[l], case["sub"]), g[f"z{l}"], meta, f"z{l}")
        _check_sums(z[l], meta, f"z{l}")


@pytest.mark.parametrize("case", [c for c in cases.PIPELINE_CASES if c["img"][0] <= 128], ids=lambda c: c["name"])
def test_pipeline_matches_reference(case, golden_index):
    """images -> ViT x2 -> MHAda x6 -> decoder (infer_image.py:82-86) through both oracles."""
    from oracle import vit_oracle as V
    c, st, sd_c, sd_s, sd_a = cases.pipeline_inputs(case)
    fcs, cs = O.transformer_multi_head(V.vision_transformer(c, sd_c), V.vision_transformer(st, sd_s), sd_a)
    meta = golden_index[case["name"]]
    g = load_golden(case["name"])
    e = O.errors(cases.token_sublattice(fcs, case["sub"]), g["fcs"])
    assert e["max_abs"] <= 1e-6 * meta["fcs"]["absmax"], e          # six chained layers amplify the storage rounding
    ec = O.errors(cases.pixel_sublattice(cs, case["img_sub"]), g["cs"])
    assert ec["max_abs_rel"] <= 1e-5, ec


def test_fp32_oracle_is_inside_reference_noise(golden_index):
    """The oracle evaluated in float32 must sit within a few x of the reference's own
    float32-vs-float64 deviation: it is the 'port' timed as cpu_baseline."""
    case = cases.by_name("layer_c512_h8_16x16")
    fc, fs, fcs, sd = cases.layer_inputs(case)
    got = O.ada_attn_multi_head(fc, fs, fcs, sd, case["H"], dtype=np.float32)
    e = O.errors(got, load_golden(case["name"])["out"])
    floor = golden_index[case["name"]]["ref32_vs_ref64"]["max_abs"]
    assert e["max_abs"] < max(5 * floor, 2e-3), (e, floor)


def test_errors_and_validation():
    with pytest.raises(ValueError):
        O.ada_attn_multi_head(np.zeros((1, 6, 2, 2)), np.zeros((1, 6, 2, 2)), np.zeros((1, 6, 2, 2)), {}, 4)
    with pytest.raises(ValueError):
        O._activation("relu")
    with pytest.raises(RuntimeError):
        case = cases.by_name("layer_c64_h1_2keys")
        fc, fs, fcs, sd = cases.layer_inputs(case)
        O.ada_attn_multi_head(fc, np.concatenate([fs, fs]), fcs, sd, 1)


def test_synth_is_bit_stable():
    # literal bits: guards the "same seeds -> same inputs on any host" contract of the fixtures
    u = synth.uniform01(7, (4,))
    assert u.tolist() == synth.uniform01(7, (4,)).tolist()
    assert [float.hex(v) for v in u] == EXPECTED_U7
    f = synth.features(3, 1, 2, 2, 2)
    assert f.shape == (1, 2, 2, 2) and np.isfinite(f).all()


EXPECTED_U7 = ['0x1.38028f22c378bp-1', '0x1.02def267c57d3p-1', '0x1.7211ab3566398p-4', '0x1.698cfc3157b8ap-2']


def test_transformer_cfg1_size_matches_reference(golden_index):
    """BASELINE.json configs[0] size (512x512 image -> 64x64 tokens, 6 layers + decoder); the
    golden holds the reference's output on a token / pixel sub-lattice."""
    case = cases.by_name("transformer_64x64_sub")
    fc, fs, sd = cases.transformer_inputs(case)
    fcs, cs = O.transformer_multi_head(fc, fs, sd)
    meta = golden_index[case["name"]]
    g = load_golden(case["name"])
    _check(cases.token_sublattice(fcs, case["sub"]), g["fcs"], meta, "fcs")
    _check(cases.pixel_sublattice(cs, case["img_sub"]), g["cs"], meta, "cs")
    _check_sums(fcs, meta, "fcs")
    _check_sums(cs, meta, "cs")


def test_torch_port_matches_reference(golden_index):
    """oracle/torch_port.py (the CPU port bench.py times as cpu_baseline / --impl reference) against the
    same golden vectors, at the reference's native float32."""
    import torch
    from oracle import torch_port
    case = cases.by_name("transformer_b2_12x10")
    fc, fs, sd = cases.transformer_inputs(case)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()
    with torch.no_grad():
        fcs, cs = torch_port.transformer([t(x) for x in fc], [t(x) for x in fs], torch_port.prepare(sd))
    g = load_golden(case["name"])
    meta = golden_index[case["name"]]
    ef = O.errors(cases.token_sublattice(fcs.numpy(), case["sub"]), g["fcs"])
    ec = O.errors(cases.pixel_sublattice(cs.numpy(), case["img_sub"]), g["cs"])
    assert ef["max_abs"] <= 5 * meta["fcs_ref32_vs_ref64"]["max_abs"] + 1e-4, ef
    assert ec["max_abs_rel"] <= 1e-4, ec


def test_torch_port_vit_matches_reference(golden_index):
    """oracle/torch_port.vit (timed as part of the full-pipeline CPU number) against the reference ViT goldens,
    batch-axis attention included."""
    import torch
    from oracle import torch_port
    for name in ("vit_b3_40x64_nopos", "vit_b1_48x64"):
        case = cases.by_name(name)
        img, sd = cases.vit_inputs(case)
        with torch.no_grad():
            z = torch_port.vit(torch.from_numpy(img).float(), torch_port.prepare(sd))
        g = load_golden(name)
        for l in range(3):
            e = O.errors(cases.token_sublattice(z[l].numpy(), case["sub"]), g[f"z{l}"])
            assert e["max_abs_rel"] <= 5e-6, (name, l, e)


@pytest.mark.parametrize("case", cases.GRAD_CASES, ids=lambda c: c["name"])
def test_reference_gradients_agree_with_oracle_finite_differences(case, golden_index):
    """The golden gradients (reference autograd, float64) against directional finite differences of the numpy
    oracle: pins the gradient fixtures independently of any autograd."""
    fc, fs, fcs, sd, G = cases.grad_inputs(case)
    g = load_golden(case["name"])
    meta = golden_index[case["name"]]

    def loss(fc_, fs_, fcs_, sd_):
        return float((O.ada_attn_multi_head(fc_, fs_, fcs_, sd_, case["H"]) * G).sum())

    assert loss(fc, fs, fcs, sd) == pytest.approx(meta["loss"], rel=1e-10)
    rng = np.random.default_rng(0)
    eps = 1e-5
    for key, arr in (("fc", fc), ("fs", fs), ("fcs", fcs), ("g_list.1.weight", sd["g_list.1.weight"]),
                     ("out_conv.bias", sd["out_conv.bias"])):
        args = dict(fc=fc, fs=fs, fcs=fcs)
        gk = g[key.replace(".", "__")].astype(np.float64)
        if "sub" in case and key in args:
            # input gradients are stored on every sub-th token: probe along a direction supported on those tokens
            vs = rng.standard_normal(gk.shape)
            v = np.zeros((arr.shape[0], arr.shape[1], arr.shape[2] * arr.shape[3]))
            v[:, :, ::case["sub"]] = vs
            v = v.reshape(arr.shape)
        else:
            vs = v = rng.standard_normal(arr.shape)

        def at(sign):
            if key in args:
                a = dict(args)
                a[key] = arr + sign * eps * v
                return loss(a["fc"], a["fs"], a["fcs"], sd)
            sd2 = dict(sd)
            sd2[key] = arr + sign * eps * v
            return loss(fc, fs, fcs, sd2)

        fd = (at(+1) - at(-1)) / (2 * eps)
        an = float((gk * vs).sum())
        assert fd == pytest.approx(an, rel=2e-4, abs=1e-6 * abs(an) + 1e-6), key


def test_backward_formulas_of_the_kernels_match_autograd():
    """SURVEY A.2 / DESIGN 4.7: the backward the B200 kernels implement -- written out in float64 with the CENTRED
    quantities they use (V~ = V - mu_V, O' = [M~ | E~], dO' = [g - 2 M~ dVar | dVar], delta = dO'.O', dS = A (dA - delta),
    dV = dV'_m + 2 V~ dV'_e, instance-norm backward) -- against torch.autograd of the reference's formulation
    (adaDecoder.py:162-206).  Pins the claim that the centring needs no correction terms."""
    import torch
    torch.manual_seed(0)
    B, H, d, Nc, Ns = 2, 2, 8, 12, 9
    C = H * d
    dd = torch.float64
    fc = torch.randn(B, Nc, C, dtype=dd) * 3 + 1
    fs = torch.randn(B, Ns, C, dtype=dd) * 2 + 5
    fcs = torch.randn(B, Nc, C, dtype=dd) + 2
    W = torch.randn(3, H, d, d, dtype=dd) * 0.3
    bias = torch.randn(3, H, d, dtype=dd)
    Wo = torch.randn(C, C, dtype=dd) * 0.1
    bo = torch.randn(C, dtype=dd)
    G = torch.randn(B, Nc, C, dtype=dd)

    def inorm(x):
        mu = x.mean(1, keepdim=True)
        var = x.var(1, unbiased=False, keepdim=True)
        return (x - mu) * torch.rsqrt(var + 1e-5)

    def fwd(fc, fs, fcs, W, bias, Wo, bo):
        xc, xs = inorm(fc).view(B, Nc, H, d), inorm(fs).view(B, Ns, H, d)
        xr, xx = fs.view(B, Ns, H, d), inorm(fcs).view(B, Nc, H, d)
        q = torch.einsum("bnhi,hoi->bhno", xc, W[0]) + bias[0][None, :, None, :]
        k = torch.einsum("bnhi,hoi->bhno", xs, W[1]) + bias[1][None, :, None, :]
        v = torch.einsum("bnhi,hoi->bhno", xr, W[2]) + bias[2][None, :, None, :]
        a = torch.softmax(q @ k.transpose(2, 3), -1)
        m, e = a @ v, a @ (v * v)
        sd = torch.sqrt((e - m * m).clamp(min=1e-6))
        y = (sd * xx.transpose(1, 2) + m).transpose(1, 2).reshape(B, Nc, C)
        return y @ Wo.T + bo

    leaves = [t.clone().requires_grad_(True) for t in (fc, fs, fcs, W, bias, Wo, bo)]
    (fwd(*leaves) * G).sum().backward()
    ref = [t.grad for t in leaves]

    def stats(x):
        return x.mean(1), torch.rsqrt(x.var(1, unbiased=False) + 1e-5)

    (mc, rc), (ms, rs), (mx, rx) = stats(fc), stats(fs), stats(fcs)
    xc, xs, xx = (fc - mc[:, None]) * rc[:, None], (fs - ms[:, None]) * rs[:, None], (fcs - mx[:, None]) * rx[:, None]
    hv = lambda t, N: t.view(B, N, H, d).transpose(1, 2)
    q = torch.einsum("bhni,hoi->bhno", hv(xc, Nc), W[0]) + bias[0][None, :, None, :]
    k = torch.einsum("bhni,hoi->bhno", hv(xs, Ns), W[1]) + bias[1][None, :, None, :]
    v = torch.einsum("bhni,hoi->bhno", hv(fs, Ns), W[2]) + bias[2][None, :, None, :]
    muv = v.mean(2, keepdim=True)
    vt = v - muv                                                        # the kernels only ever see the centred values
    a = torch.softmax(q @ k.transpose(2, 3), -1)
    Mt, Et = a @ vt, a @ (vt * vt)
    var = Et - Mt * Mt
    sd = torch.sqrt(var.clamp(min=1e-6))
    cat = (sd * hv(xx, Nc) + Mt + muv).transpose(1, 2).reshape(B, Nc, C)
    dcat = G @ Wo
    dWo, dbo = G.reshape(-1, C).T @ cat.reshape(-1, C), G.sum((0, 1))
    g = hv(dcat, Nc)
    dxhat = g * sd
    dvar = torch.where(var >= 1e-6, g * hv(xx, Nc) / (2 * sd), torch.zeros_like(g))
    dM, dE = g - 2 * Mt * dvar, dvar
    delta = (dM * Mt + dE * Et).sum(-1, keepdim=True)
    dA = dM @ vt.transpose(2, 3) + dE @ (vt * vt).transpose(2, 3)
    dS = a * (dA - delta)
    dq, dk = dS @ k, dS.transpose(2, 3) @ q
    dv = a.transpose(2, 3) @ dM + 2 * vt * (a.transpose(2, 3) @ dE)
    gq = torch.einsum("bhno,hoi->bhni", dq, W[0])
    gk = torch.einsum("bhno,hoi->bhni", dk, W[1])
    gv = torch.einsum("bhno,hoi->bhni", dv, W[2])
    dW = torch.stack([torch.einsum("bhno,bhni->hoi", dq, hv(xc, Nc)), torch.einsum("bhno,bhni->hoi", dk, hv(xs, Ns)),
                      torch.einsum("bhno,bhni->hoi", dv, hv(fs, Ns))])
    db = torch.stack([dq.sum((0, 2)), dk.sum((0, 2)), dv.sum((0, 2))])
    tok = lambda t, N: t.transpose(1, 2).reshape(B, N, C)
    inbwd = lambda g_, xh, r: r[:, None] * (g_ - g_.mean(1, keepdim=True) - xh * (g_ * xh).mean(1, keepdim=True))
    mine = [inbwd(tok(gq, Nc), xc, rc), inbwd(tok(gk, Ns), xs, rs) + tok(gv, Ns), inbwd(tok(dxhat, Nc), xx, rx), dW, db, dWo, dbo]
    for name, a_, b_ in zip("fc fs fcs W b Wo bo".split(), mine, ref):
        assert float((a_ - b_).abs().max()) <= 1e-10 * max(1.0, float(b_.abs().max())), name
