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


# ---------------------------------------------------------------------------------------------------------------------
# Resnet18.py:4-88 - the stand-alone ResNet18 / BasicBlock classes (row stride 2, column stride 1 in the stem and first pool)
S18_CASES = {"nb384_c50": (384, 50, 2), "nb512_c3": (512, 3, 3)}


def _s18_spec(case):
    return json.load(open(os.path.join(GOLDEN, "resnet18_standalone_spec.json")))[case]


@pytest.mark.parametrize("case", list(S18_CASES))
def test_resnet18_standalone_keys_match_reference(case):
    from affganwriting_b200.Resnet18 import ResNet18
    nb, cin, _ = S18_CASES[case]
    mine = {k: list(v.shape) for k, v in ResNet18(nb_feat=nb, in_channels=cin).state_dict().items()}
    assert mine == _s18_spec(case) and list(mine) == list(_s18_spec(case))


@pytest.mark.parametrize("case", list(S18_CASES))
def test_resnet18_standalone_forward_backward(case, mode, golden):
    from affganwriting_b200.Resnet18 import ResNet18
    g = golden("resnet18_standalone.npz")
    nb, cin, batch = S18_CASES[case]
    net = ResNet18(nb_feat=nb, in_channels=cin)
    net.load_state_dict(W.make_state(_s18_spec(case)))
    net = net.cuda().train()
    x = O.synthetic_batch(batch, cin)["tr_img"].cuda().requires_grad_()
    res = net(x)
    loss = sum(r.float().square().mean() for r in res)
    loss.backward()
    noise_f, noise_g = float(g[f"{case}.noise.fwd"]), float(g[f"{case}.noise.dx"])
    ftol = {"fp32": 1e-4, "bf16": 5e-3, "bf16x1": 6e-2}[mode] + 3 * noise_f
    worst = 0.0
    for i, r in enumerate(res):
        assert list(r.shape) == g[f"{case}.result{i}.shape"].tolist()
        worst = max(worst, rel_err(r[:, :8], _t(g[f"{case}.result{i}.head"])))
    print(f"\n[{mode}] Resnet18.py {case}: worst map error vs reference {worst:.3e} (tolerance {ftol:.1e})")
    assert worst <= ftol
    assert abs(float(loss) - float(g[f"{case}.loss"])) <= 10 * ftol * float(g[f"{case}.loss"])
    if mode == "bf16x1":
        return
    gtol = {"fp32": 2e-3, "bf16": 2e-2}[mode] + 4 * noise_g
    dx = x.grad.cpu()
    assert abs(float(dx.norm()) / float(g[f"{case}.dx.norm"]) - 1) <= gtol
    assert cosine(dx[:, :3, ::4, ::4], _t(g[f"{case}.dx.head"])) >= 1 - gtol
    assert cosine(net.conv1.weight.grad[:8].cpu(), _t(g[f"{case}.grad.conv1"])) >= 1 - gtol
    ref_norm = dict(zip(g[f"{case}.grad.keys"].tolist(), g[f"{case}.grad.norms"].tolist()))
    bad = [(k, float(p.grad.norm()) / ref_norm[k]) for k, p in net.named_parameters()
           if ref_norm[k] > 1e-6 and abs(float(p.grad.norm()) / ref_norm[k] - 1) > 5 * gtol]
    assert len(bad) <= len(ref_norm) // 20, bad[:5]
    post = net.state_dict()
    assert int(post["layer2.0.bn1.num_batches_tracked"]) == 1
    for k in ("bn1.running_mean", "layer1.0.downsample.1.running_var", "layer3.1.bn2.running_mean"):
        assert rel_err(post[k], _t(g[f"{case}.post.{k}"])) <= {"fp32": 1e-4, "bf16": 5e-3}[mode] + 3 * noise_f, k


@pytest.mark.parametrize("strides", [(2, 1), (1, 1), (2, 2), (1, 2)])
def test_max_pool3_strides_match_torch(strides):
    """nn.MaxPool2d(3, stride, 1) forward and the first-maximum gradient rule, including ties (Resnet18.py:45-46)."""
    import torch.nn.functional as F
    from affganwriting_b200 import ops
    torch.manual_seed(3)
    x = torch.randint(-3, 4, (2, 16, 9, 14)).float().cuda().requires_grad_()      # small integers -> many ties
    y = ops.max_pool3(ops.input_to_internal(x), strides)
    yr = F.max_pool2d(x, 3, strides, 1)
    assert y.shape == yr.shape and torch.equal(y.contiguous(), yr)
    w = torch.randn_like(yr)
    (gx,) = torch.autograd.grad((y * w).sum(), x)
    (gr,) = torch.autograd.grad((yr * w).sum(), x)
    assert torch.allclose(gx, gr, atol=1e-6)


@pytest.mark.parametrize("stride", [(2, 1), (1, 2), (3, 1)])
def test_conv_anisotropic_stride(stride, mode):
    """affgw_conv_desc.stride_w: nn.Conv2d(..., stride=(rows, columns)) forward, input gradient and weight gradient."""
    import torch.nn.functional as F
    from affganwriting_b200 import ops
    torch.manual_seed(5)
    x = torch.randn(3, 24, 17, 29).cuda().requires_grad_()
    w = (torch.randn(40, 24, 3, 3) * 0.1).cuda().requires_grad_()
    y = ops.conv2d(ops.input_to_internal(x), w, None, stride=stride, pad=1)
    yr = F.conv2d(x.double(), w.double(), None, stride=stride, padding=1)
    assert y.shape == yr.shape
    tol = {"fp32": 1e-5, "bf16": 1e-4, "bf16x1": 2e-2}[mode]
    assert rel_err(y, yr.float()) <= tol
    gy = torch.randn_like(yr)
    gx, gw = torch.autograd.grad((y.double() * gy).sum(), (x, w))
    gxr, gwr = torch.autograd.grad((yr * gy).sum(), (x, w))
    btol = 2e-2 if mode == "bf16" else tol          # 'bf16': backward GEMMs run one MMA per product
    assert rel_err(gx, gxr.float()) <= btol and rel_err(gw, gwr.float()) <= btol
