"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded inputs.

Runs only in the build container (needs /root/reference).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.gen_golden            # all cases
    python -m oracle.gen_golden layer_     # cases whose name starts with the prefix

For every case the reference module is built, loaded (strict) with oracle/synth.py weights, and
run twice: in float64 (module.double(); stored, rounded to float32, as the golden output) and in
float32 (the reference's native precision; only its deviation from the float64 run is recorded,
as the noise floor any fp32 implementation should be judged against -- SURVEY.md D8).
float64 sums of the float64 outputs are stored in index.json to pin the oracle tightly.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

REF = os.environ.get("MHADA_REFERENCE", "/root/reference/MHAdaSTr")
sys.path.insert(0, REF)
from network.adaDecoder import (AdaAttN, AdaAttnForLoss, AdaAttnMultiHead,  # noqa: E402
                                AdaAttnTransformer, AdaAttnTransformerMultiHead)
from network.conv import Decoder  # noqa: E402
from network.vit import VisionTransformer  # noqa: E402

from . import cases, synth  # noqa: E402
from .mhada_oracle import errors  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def T(x, dtype):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def run_both(build, sd, args):
    outs = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        m = build().to(dt)      # cast first so float64 weights are not rounded through float32
        if sd is not None:
            m.load_state_dict(synth.to_torch(sd, dt), strict=True)
        m = m.eval()
        with torch.no_grad():
            o = m(*[[T(a, dt) for a in x] if isinstance(x, list) else T(x, dt) for x in args])
        outs[name] = o
    return outs


def summarize(x64: np.ndarray) -> dict:
    return {"sum": float(x64.sum()), "sumsq": float((x64 * x64).sum()), "absmax": float(np.abs(x64).max()),
            "shape": list(x64.shape)}


def main(prefix: str = ""):
    os.makedirs(OUT, exist_ok=True)
    idx_path = os.path.join(OUT, "index.json")
    index = json.load(open(idx_path)) if os.path.exists(idx_path) else {}
    index["_meta"] = {"torch": torch.__version__, "reference": "Maboroshi0327/MHAda-Style-Transfer MHAdaSTr/network",
                      "generator": "oracle/gen_golden.py"}
    torch.set_num_threads(os.cpu_count())
    for case in cases.ALL_CASES:
        if not case["name"].startswith(prefix):
            continue
        t0 = time.time()
        kind = case["kind"]
        arrays, meta = {}, {"kind": kind}
        if kind == "layer":
            fc, fs, fcs, sd = cases.layer_inputs(case)
            o = run_both(lambda: AdaAttnMultiHead(case["C"], case["H"], case.get("activation", "softmax")), sd, (fc, fs, fcs))
            o64, o32 = o["f64"].numpy(), o["f32"].numpy()
            arrays["out"] = o64.astype(np.float32)
            meta["out"] = summarize(o64); meta["ref32_vs_ref64"] = errors(o32, o64)
        elif kind == "adaattn":
            fc, fs, fcs, sd = cases.adaattn_inputs(case)
            o = run_both(lambda: AdaAttN(case["C"], case.get("activation", "softmax")), sd, (fc, fs, fcs))
            o64, o32 = o["f64"].numpy(), o["f32"].numpy()
            arrays["out"] = o64.astype(np.float32)
            meta["out"] = summarize(o64); meta["ref32_vs_ref64"] = errors(o32, o64)
        elif kind == "forloss":
            args = cases.forloss_inputs(case)
            o = run_both(lambda: AdaAttnForLoss(case["v"], case["qk"], case.get("activation", "softmax")), None, args)
            o64, o32 = o["f64"].numpy(), o["f32"].numpy()
            arrays["out"] = o64.astype(np.float32)
            meta["out"] = summarize(o64); meta["ref32_vs_ref64"] = errors(o32, o64)
        elif kind == "transformer":
            fc, fs, sd = cases.transformer_inputs(case)
            o = run_both(lambda: AdaAttnTransformerMultiHead(num_heads=case.get("heads", 8)), sd, (fc, fs))
            fcs64, cs64 = (t.numpy() for t in o["f64"])
            fcs32, cs32 = (t.numpy() for t in o["f32"])
            arrays["fcs"] = cases.token_sublattice(fcs64, case["sub"]).astype(np.float32)
            arrays["cs"] = cases.pixel_sublattice(cs64, case["img_sub"]).astype(np.float32)
            meta["fcs"] = summarize(fcs64); meta["cs"] = summarize(cs64)
            meta["fcs_ref32_vs_ref64"] = errors(fcs32, fcs64); meta["cs_ref32_vs_ref64"] = errors(cs32, cs64)
        elif kind == "single_head_transformer":
            fc, fs, sd = cases.single_head_transformer_inputs(case)
            o = run_both(lambda: AdaAttnTransformer(), sd, (fc, fs))
            cs64, cs32 = o["f64"].numpy(), o["f32"].numpy()
            arrays["cs"] = cases.pixel_sublattice(cs64, case["img_sub"]).astype(np.float32)
            meta["cs"] = summarize(cs64); meta["cs_ref32_vs_ref64"] = errors(cs32, cs64)
        elif kind == "grad":
            fc, fs, fcs, sd, G = cases.grad_inputs(case)
            m = AdaAttnMultiHead(case["C"], case["H"]).to(torch.float64)
            m.load_state_dict(synth.to_torch(sd, torch.float64), strict=True)
            tin = [T(a, torch.float64).requires_grad_(True) for a in (fc, fs, fcs)]
            out = m(*tin)
            loss = (out * T(G, torch.float64)).sum()
            loss.backward()
            named = dict(m.named_parameters())
            grads = {"fc": tin[0].grad, "fs": tin[1].grad, "fcs": tin[2].grad}
            for k in cases.GRAD_KEYS[3:]:
                grads[k] = named[k].grad
            for k, g in grads.items():
                a = g.numpy()
                if "sub" in case and k in ("fc", "fs", "fcs"):
                    a = cases.token_sublattice(a, case["sub"])
                arrays[k.replace(".", "__")] = a.astype(np.float32)
                meta[k] = summarize(g.numpy())
            meta["loss"] = float(loss)
        elif kind == "decoder":
            x, sd = cases.decoder_inputs(case)
            sd_local = {k[len("decoder."):]: v for k, v in sd.items()}
            o = run_both(lambda: Decoder(), sd_local, (x,))
            o64, o32 = o["f64"].numpy(), o["f32"].numpy()
            arrays["out"] = o64.astype(np.float32)
            meta["out"] = summarize(o64); meta["ref32_vs_ref64"] = errors(o32, o64)
        elif kind == "vit":
            img, sd = cases.vit_inputs(case)
            o = run_both(lambda: VisionTransformer(pos_embedding=case["pos"]), sd, (img,))
            for l in range(3):
                z64, z32 = o["f64"][l].numpy(), o["f32"][l].numpy()
                arrays[f"z{l}"] = cases.token_sublattice(z64, case["sub"]).astype(np.float32)
                meta[f"z{l}"] = summarize(z64); meta[f"z{l}_ref32_vs_ref64"] = errors(z32, z64)
        elif kind == "pipeline":
            c, st, sd_c, sd_s, sd_a = cases.pipeline_inputs(case)
            res = {}
            for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
                vc = VisionTransformer(pos_embedding=True).to(dt); vc.load_state_dict(synth.to_torch(sd_c, dt), strict=True)
                vs = VisionTransformer(pos_embedding=False).to(dt); vs.load_state_dict(synth.to_torch(sd_s, dt), strict=True)
                ada = AdaAttnTransformerMultiHead().to(dt); ada.load_state_dict(synth.to_torch(sd_a, dt), strict=True)
                with torch.no_grad():                                           # infer_image.py:82-86
                    fc = vc.eval()(T(c, dt)); fs = vs.eval()(T(st, dt))
                    fcs, cs = ada.eval()(fc, fs)
                res[name] = (fcs.numpy(), cs.numpy())
            fcs64, cs64 = res["f64"]; fcs32, cs32 = res["f32"]
            arrays["fcs"] = cases.token_sublattice(fcs64, case["sub"]).astype(np.float32)
            arrays["cs"] = cases.pixel_sublattice(cs64, case["img_sub"]).astype(np.float32)
            meta["fcs"] = summarize(fcs64); meta["cs"] = summarize(cs64)
            meta["fcs_ref32_vs_ref64"] = errors(fcs32, fcs64); meta["cs_ref32_vs_ref64"] = errors(cs32, cs64)
        else:
            raise ValueError(kind)
        np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **arrays)
        index[case["name"]] = meta
        print(f"{case['name']:32s} {time.time() - t0:6.1f}s", {k: v for k, v in meta.items() if 'vs' in k})
    json.dump(index, open(idx_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "")
