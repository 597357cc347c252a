"""Line-level generator (BASELINE.json configs[4]): the CPU oracle against the image the UNMODIFIED reference
produced (tests/golden/linegen.npz, oracle/make_golden_linegen.py), the drop-in's state_dict layout, and - on the GPU - the
libaffgw forward against the same reference image with the same injected noise."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import linegen_oracle as LG
from oracle import weights as W

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"b2_t24": (2, 24), "b3_t7": (3, 7)}


def _state():
    spec = json.load(open(os.path.join(GOLDEN, "linegen_spec.json")))["spec"]
    sd = W.make_state(spec)
    for k in spec:
        if k.endswith(".weight_flip") or (k.endswith(".weight") and k[:-len("weight")] + "weight_flip" in spec):
            sd[k] = LG.BLUR.view(1, 1, 3, 3).repeat(spec[k][0], 1, 1, 1).clone()
        if k.startswith("gen."):
            sd[k] = sd["conv." + k[4:]]
    return spec, sd


def test_linegen_oracle_matches_reference_image():
    rep = json.load(open(os.path.join(GOLDEN, "linegen_spec.json")))["report"]
    assert rep and all(r["max_abs"] <= r["tol"] for r in rep)
    spec, sd = _state()
    gold = np.load(os.path.join(GOLDEN, "linegen.npz"))
    for case, (b, t) in CASES.items():
        content, style, noises = LG.synthetic_inputs(b, t)
        assert [tuple(n.shape) for n in noises] == LG.noise_shapes(b, t)
        with torch.no_grad():
            y = LG.spaced_generator(content, style, sd, noises)
        assert y.shape == (b, 1, 64, 4 * t)
        assert float((y - torch.from_numpy(gold[case + ".image"])).abs().max()) <= 1e-5


def test_linegen_dropin_has_the_reference_state_dict_layout():
    from affganwriting_b200.linegen import SpacedGenerator
    spec, sd = _state()
    g = SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True, small=False)
    own = g.state_dict()
    assert list(own.keys()) == list(spec.keys())
    assert all(list(own[k].shape) == spec[k] for k in spec)
    g.load_state_dict(sd, strict=True)
    assert torch.equal(own["conv.1.conv1.2.weight"], sd["conv.1.conv1.2.weight"])        # the Blur buffers are the binomial kernel
    assert own["gen.3.conv1.0.weight"].data_ptr() == own["conv.3.conv1.0.weight"].data_ptr()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "f16", "bf16"])
def test_linegen_forward_matches_reference_image(mode):
    import affganwriting_b200 as A
    from affganwriting_b200.linegen import SpacedGenerator
    spec, sd = _state()
    gold = np.load(os.path.join(GOLDEN, "linegen.npz"))
    A.set_precision(mode)
    try:
        g = SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True, small=False)
        g.load_state_dict(sd)
        g = g.cuda().eval()
        for case, (b, t) in CASES.items():
            content, style, noises = LG.synthetic_inputs(b, t)
            n0 = A.launch_count()
            y = g(content.cuda(), style.cuda(), noise=[n.cuda() for n in noises])
            ref = torch.from_numpy(gold[case + ".image"])
            assert y.shape == ref.shape
            err = float((y.float().cpu() - ref).abs().max() / ref.abs().max())
            print(f"\n[{mode}] line generator {case}: max error relative to max|image| {err:.2e}, {A.launch_count() - n0} launches")
            assert err <= {"fp32": 1e-5, "f16": 1e-4, "bf16": 1e-3}[mode]      # measured on B200: 4e-7 / 7e-7 / 1e-5
        with pytest.raises(RuntimeError):                    # generation only: no silent autograd through the transposed convolutions
            with torch.enable_grad():
                g.conv[3].conv1[0](torch.randn(1, 64, 4, 8, device="cuda", requires_grad=True))
        A.check_device_errors()
    finally:
        A.set_precision("fp32")


@pytest.mark.gpu
def test_linegen_full_size_line():
    """configs[4] shape: T = 256 spaced characters -> 64 x 1024 line images; device-drawn noise; deterministic under a seed."""
    import affganwriting_b200 as A
    from affganwriting_b200.linegen import SpacedGenerator
    _, sd = _state()
    A.set_precision("f16")
    try:
        g = SpacedGenerator(80, 128, 256, append_style=True)
        g.load_state_dict(sd)
        g = g.cuda().eval()
        content, style, _ = LG.synthetic_inputs(4, 256)
        torch.manual_seed(3)
        a = g(content.cuda(), style.cuda())
        torch.manual_seed(3)
        b = g(content.cuda(), style.cuda())
        assert a.shape == (4, 1, 64, 1024) and torch.isfinite(a).all() and float(a.abs().max()) <= 1.0
        assert float((a - b).abs().max()) <= 1e-3 * float(a.abs().max())
    finally:
        A.set_precision("fp32")


@pytest.mark.gpu
def test_linegen_graph_replay_matches_eager():
    """inference.GraphedForward around the line generator: with the noise tensors passed in (no random draw inside) a replay
    returns the eager image; with device-drawn noise two replays differ (fresh noise per replay) and stay in range."""
    import affganwriting_b200 as A
    from affganwriting_b200.linegen import SpacedGenerator
    from affganwriting_b200.inference import GraphedForward
    _, sd = _state()
    A.set_precision("f16")
    try:
        g = SpacedGenerator(80, 128, 256, append_style=True)
        g.load_state_dict(sd)
        g = g.cuda().eval()
        content, style, _ = LG.synthetic_inputs(4, 64)
        content, style = content.cuda(), style.cuda()
        fast = GraphedForward(g)
        with torch.no_grad():
            outs = [fast(content, style).clone() for _ in range(5)]
        assert fast._graph is not None
        a, b = outs[-2], outs[-1]
        assert a.shape[0] == 4 and torch.isfinite(b).all() and float(b.abs().max()) <= 1.0
        assert float((a - b).abs().max()) > 1e-4                     # every replay draws its own noise

        class Fixed(torch.nn.Module):                                # same generator, noise as an input
            def __init__(self, gen, noise):
                super().__init__()
                self.gen, self.noise = gen, noise

            def forward(self, c, s):
                return self.gen(c, s, noise=self.noise)
        with torch.no_grad():
            probe = []
            for blk in g.gen if hasattr(g, "gen") else []:
                probe.append(blk)
            ref0 = g(content, style, return_intermediate=False)
        # noise shapes: run once eagerly and record the shapes the injections see
        shapes = []
        hooks = [m.register_forward_pre_hook(lambda mod, inp: shapes.append(tuple(inp[1].shape)))
                 for m in g.modules() if type(m).__name__ == "NoiseInjection"]
        with torch.no_grad():
            g(content, style)
        for h in hooks:
            h.remove()
        gen_ = torch.Generator(device="cuda").manual_seed(11)
        noise = [torch.randn(s, device="cuda", generator=gen_) for s in shapes]
        fixed = Fixed(g, noise).eval()
        fast2 = GraphedForward(fixed)
        with torch.no_grad():
            ref = fixed(content, style).clone()
            for _ in range(4):
                out = fast2(content, style)
        assert fast2._graph is not None
        assert float((out - ref).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max()))
    finally:
        A.set_precision("fp32")
