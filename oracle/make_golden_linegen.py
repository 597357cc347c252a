"""Golden vectors of the line-level generator produced by the UNMODIFIED reference (line_generation/model/pure_gen.py, imported
in place from /root/reference; nothing is copied).  Container-only:  python -m oracle.make_golden_linegen   (TEST INFRASTRUCTURE)
Writes tests/golden/linegen.npz + linegen_spec.json and checks oracle.linegen_oracle against the reference on the same weights,
inputs and noise (torch.randn_like is intercepted for the duration of the reference's forward)."""
import importlib.util
import json
import os

import numpy as np
import torch

from oracle import linegen_oracle as LG
from oracle import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF = "/root/reference/line_generation/model/pure_gen.py"


def make_state(spec):
    """deterministic weights from the {key: shape} spec; the Blur buffers keep their fixed kernel (pure_gen.py:127-133)"""
    sd = W.make_state(spec)
    for k in spec:
        if k.endswith(".weight_flip") or (k.endswith(".weight") and k[:-len("weight")] + "weight_flip" in spec):
            sd[k] = LG.BLUR.view(1, 1, 3, 3).repeat(spec[k][0], 1, 1, 1).clone()
        if k.startswith("gen."):                       # `gen` aliases `conv` (pure_gen.py:40)
            sd[k] = sd["conv." + k[4:]]
    return sd


def main():
    sp = importlib.util.spec_from_file_location("ref_pure_gen", REF)
    ref = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(ref)
    gen = ref.SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True, small=False)
    spec = W.spec_of(gen)
    sd = make_state(spec)
    gen.load_state_dict(sd)
    gen.eval()
    out, report = {}, []
    for case, (batch, T) in {"b2_t24": (2, 24), "b3_t7": (3, 7)}.items():
        content, style, noises = LG.synthetic_inputs(batch, T)
        feed = list(noises)
        real = torch.randn_like

        def fake(t, *a, **k):
            n = feed.pop(0)
            assert n.shape == t.shape, (n.shape, t.shape)
            return n
        torch.randn_like = fake
        try:
            with torch.no_grad():
                y_ref = gen(content, style)
        finally:
            torch.randn_like = real
        assert not feed and y_ref.shape == (batch, 1, 64, 4 * T)
        with torch.no_grad():
            y = LG.spaced_generator(content, style, sd, noises)
        err = float((y - y_ref).abs().max())
        report.append({"name": f"linegen.{case}.image", "max_abs": err, "tol": 1e-5})
        print(f"{case}: oracle vs reference image max-abs {err:.2e}; range [{float(y_ref.min()):.3f}, {float(y_ref.max()):.3f}]")
        assert err <= 1e-5
        out[f"{case}.image"] = y_ref.numpy()
    np.savez_compressed(os.path.join(OUT, "linegen.npz"), **out)
    json.dump({"spec": spec, "report": report}, open(os.path.join(OUT, "linegen_spec.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
