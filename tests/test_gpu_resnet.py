"""GPU parity of the torchvision-ResNet style encoders: ImageEncoderResNet50 is the encoder that is
active in the reference's GenModel_FC (modules_tro.py:219,464-533); ImageEncoderResNet18 follows modules_tro2.py:447-516.
Fixtures: fp64 run of the reference classes (tests/golden/resnet_enc.npz, oracle/make_golden_resnet.py); because these
BatchNorm stacks at batch 2 amplify fp32 rounding (the reference's own fp32 run is 3 % off its fp64 run on the ResNet-50
input gradient), gradient tolerances are expressed relative to that recorded noise floor."""
import json
import os

import numpy as np
import pytest
import torch

import affganwriting_b200 as A
from affganwriting_b200 import modules_tro as M
from affganwriting_b200.resnet_encoder import ImageEncoderResNet18, ImageEncoderResNet50
from affgw_testutil import cosine, rel_err
from oracle import affgw_oracle as O
from oracle import weights as W

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLS = {"resnet18": ImageEncoderResNet18, "resnet50": ImageEncoderResNet50}


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _spec(arch):
    return json.load(open(os.path.join(GOLDEN, "resnet_spec.json")))[arch]


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_state_dict_keys_match_reference(arch):
    enc = CLS[arch](weight_path=None, in_channels=50)
    mine = {k: list(v.shape) for k, v in enc.state_dict().items()}
    assert mine == _spec(arch) and list(mine) == list(_spec(arch))


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_encoder_forward_backward(arch, mode, golden):
    g = golden("resnet_enc.npz")
    sd = W.alias_extractor(W.make_state(_spec(arch)))
    enc = CLS[arch](weight_path=None, in_channels=50)
    enc.load_state_dict(sd)
    enc = enc.cuda().train()
    x = O.synthetic_batch(2, 50)["tr_img"].cuda().requires_grad_()
    res = enc(x)
    loss = sum(r.float().square().mean() for r in res)
    loss.backward()
    noise_f, noise_g = float(g[f"{arch}.noise.fwd"]), float(g[f"{arch}.noise.dx"])
    # one bf16 rounding per operand is amplified chaotically by the 50-layer BatchNorm stack (same effect as DESIGN.md §3)
    ftol = {"fp32": 1e-4, "bf16": 5e-3, "bf16x1": 6e-2 if arch == "resnet18" else 1.0}[mode] + 3 * noise_f
    worst = 0.0
    for i, r in enumerate(res):
        assert list(r.shape) == g[f"{arch}.result{i}.shape"].tolist()
        worst = max(worst, rel_err(r[:, :8], _t(g[f"{arch}.result{i}.head"])))
    print(f"\n[{mode}] {arch}: worst map error vs reference {worst:.3e} (tolerance {ftol:.1e})")
    assert worst <= ftol
    assert abs(float(loss) - float(g[f"{arch}.loss"])) <= 10 * ftol * float(g[f"{arch}.loss"])
    if mode == "bf16x1":
        return
    gtol = {"fp32": 2e-3, "bf16": 2e-2}[mode] + 4 * noise_g
    dx = x.grad.cpu()
    assert abs(float(dx.norm()) / float(g[f"{arch}.dx.norm"]) - 1) <= gtol
    assert cosine(dx[:, :4, ::8, ::8], _t(g[f"{arch}.dx.head"])) >= 1 - gtol
    ref_norm = dict(zip(g[f"{arch}.grad.keys"].tolist(), g[f"{arch}.grad.norms"].tolist()))
    bad = []
    for k, p in enc.named_parameters():
        if ref_norm[k] < 0:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        if ref_norm[k] > 1e-6 and abs(float(p.grad.norm()) / ref_norm[k] - 1) > 5 * gtol:
            bad.append((k, float(p.grad.norm()) / ref_norm[k]))
    assert len(bad) <= len(ref_norm) // 20, bad[:5]
    post = enc.state_dict()
    assert int(post["model.layer3.1.bn1.num_batches_tracked"]) == 1
    for k in ("model.bn1.running_mean", "model.layer4.0.bn2.running_var", "model.layer2.0.downsample.1.running_mean"):
        assert rel_err(post[k], _t(g[f"{arch}.post.{k}"])) <= {"fp32": 1e-4, "bf16": 5e-3}[mode] + 3 * noise_f, k


def test_generator_with_resnet50_encoder_matches_oracle(specs):
    """GenModel_FC as the reference builds it (ResNet-50 style encoder, modules_tro.py:219): image against the CPU oracle
    (oracle.resnet_encoder + the decoder path pinned in tests/test_oracle_golden.py), fp32 and bf16 modes."""
    gsd = {k: v for k, v in W.make_state(specs["gen_c50"]).items() if not k.startswith("enc_image.")}
    esd = W.alias_extractor(W.make_state(_spec("resnet50")))
    full = dict(gsd)
    full.update({"enc_image." + k: v for k, v in esd.items()})
    cpu = O.synthetic_batch(4, 50)
    with torch.no_grad():
        ref = O.gen_forward(cpu["tr_img"], cpu["label_xt"], full,
                            results=O.resnet_encoder(cpu["tr_img"], full, "enc_image.", "resnet50", True, None))
    gen = M.GenModel_FC(12, encoder=ImageEncoderResNet50(weight_path=None, in_channels=50))
    # fp32 bar: two fp32 evaluations of this 50-layer train-mode BatchNorm stack differ by ~1e-4 at the encoder maps (the
    # reference's own fp32 run vs its fp64 run, resnet_enc.npz `noise.fwd`), which the decoder amplifies to a few 1e-4
    for mode, tol in (("fp32", 2e-3), ("bf16", 2e-2)):
        A.set_precision(mode)
        try:
            gen.load_state_dict(full)
            gen = gen.cuda().train()
            xg = gen(cpu["tr_img"].cuda(), cpu["label_xt"].cuda())
            err = float((xg.detach().cpu() - ref).abs().max())
            print(f"\n[{mode}] ResNet-50 generator image max-abs error vs CPU oracle: {err:.3e}")
            assert xg.shape == (4, 1, 64, 216) and err <= tol
            xg.float().square().mean().backward()
            assert gen.enc_image.model.conv1.weight.grad is not None
            gen.zero_grad()
        finally:
            A.set_precision("fp32")
