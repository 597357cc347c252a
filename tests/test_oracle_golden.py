"""CPU: the oracle (oracle/affgw_oracle.py) against vectors produced by the UNMODIFIED reference
(tests/golden/*.npz, written by oracle/make_golden.py in the build container).  This is the pin that lets the GPU
tests use the oracle as the stand-in for the reference on a box where /root/reference does not exist."""
import json
import os

import numpy as np
import pytest
import torch

from affgw_testutil import rel_err
from oracle import affgw_oracle as O
from oracle import weights as W

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_generation_report_is_green():
    rep = json.load(open(os.path.join(GOLDEN, "oracle_vs_reference.json")))
    assert len(rep) > 80
    assert all(r["max_abs"] <= r["tol"] for r in rep)


@pytest.mark.parametrize("name", ["conv_zero_in_relu", "conv_reflect_none_tanh_k7", "conv_reflect_in_relu_k5",
                                  "conv_actfirst_lrelu", "conv_head_k2s7", "conv_1x1_nobias", "conv_replicate"])
def test_conv2dblock_cases(name, specs, golden):
    g = golden("blocks.npz")
    sp = specs["blocks." + name]
    sd = {k: v.requires_grad_() for k, v in W.make_state(sp["spec"]).items()}
    kw = dict(sp["ctor"])
    kw.pop("in_dim"), kw.pop("out_dim"), kw.pop("use_bias", None)
    x = _t(g[name + ".x"]).requires_grad_()
    y = O.conv2d_block(x, sd, "", **kw)
    y.backward(_t(g[name + ".gy"]))
    assert rel_err(y, _t(g[name + ".y"])) <= 1e-5
    assert rel_err(x.grad, _t(g[name + ".dx"])) <= 1e-4
    assert rel_err(sd["conv.weight"].grad, _t(g[name + ".dw"])) <= 1e-4


@pytest.mark.parametrize("name,train", [("adain_plain", True), ("adain_iaff_train", True), ("adain_iaff_eval", False)])
def test_adain_iaff(name, train, specs, golden):
    g = golden("blocks.npz")
    sd = W.make_state(specs["blocks." + name]["spec"])
    x = _t(g[name + ".x"]).requires_grad_()
    style = _t(g[name + ".style"]) if name + ".style" in g.files else None
    stats = {}
    y = O.adaptive_instance_norm(x, sd, "", _t(g[name + ".weight"]), _t(g[name + ".bias"]), style, train, stats)
    y.backward(_t(g[name + ".gy"]))
    assert rel_err(y, _t(g[name + ".y"])) <= 5e-5
    assert rel_err(x.grad, _t(g[name + ".dx"])) <= 2e-4
    if name == "adain_iaff_train":
        assert int(stats["iAff.global_att.2.num_batches_tracked"]) == 2      # blocks.py:295 reuses global_att
        assert "iAff.global_att2.2.num_batches_tracked" not in stats
        for k in ("iAff.global_att.2.running_mean", "iAff.global_att.5.running_var"):
            assert rel_err(stats[k], _t(g[name + ".post." + k])) <= 1e-5


def test_blocks_misc(specs, golden):
    g = golden("blocks.npz")
    sd = W.make_state(specs["blocks.resblocks_in"]["spec"])
    y = _t(g["resblocks_in.x"])
    for i in range(2):
        y = O.res_block(y, sd, f"model.{i}.", "in", "relu", "reflect")
    assert rel_err(y, _t(g["resblocks_in.y"])) <= 1e-5
    for name, (fin, fout) in {"actfirst_same": (8, 8), "actfirst_grow": (8, 16)}.items():
        sd = W.make_state(specs["blocks." + name]["spec"])
        assert rel_err(O.act_first_res_block(_t(g[name + ".x"]), sd, "", fin, fout), _t(g[name + ".y"])) <= 1e-5
    sd = W.make_state(specs["blocks.mlp"]["spec"])
    assert rel_err(O.mlp(_t(g["mlp.x"]), sd, "", 3), _t(g["mlp.y"])) <= 1e-5
    x = torch.empty(2, 8, 8, 27)
    assert rel_err(O.get_key(x, _t(g["get_key.style"])), _t(g["get_key.y"])) <= 1e-6


def test_generator_forward_and_bn_updates(specs, golden):
    gold = golden("gen_fwd_c15_b4.npz")
    sd = W.make_state(specs["gen_c15"])
    batch = O.synthetic_batch(4, 15)
    stats = {}
    with torch.no_grad():
        res = O.image_encoder(batch["tr_img"], sd)
        for i in range(6):
            assert tuple(res[i].shape) == tuple(gold[f"result{i}.shape"])
            assert abs(float(res[i].abs().mean()) - float(gold[f"result{i}.abs_mean"])) <= 1e-4
        xg = O.gen_forward(None, batch["label_xt"], sd, stats=stats, results=res)
    assert float((xg - _t(gold["xg"])).abs().max()) <= 1e-4
    for k in gold.files:
        if k.startswith("post."):
            assert rel_err(stats[k[5:]].float(), _t(gold[k]).float()) <= 1e-4, k
    with torch.no_grad():
        xe = O.gen_forward(batch["tr_img"][:1], batch["label_xt"][:1], sd, training=False)
    assert float((xe - _t(gold["xg_eval_b1"])).abs().max()) <= 1e-4


def test_dis_cla_losses(specs, golden):
    gold, gg = golden("dis_cla_b4.npz"), golden("gen_fwd_c15_b4.npz")
    dsd, csd = W.make_state(specs["dis"]), W.make_state(specs["cla"])
    batch = O.synthetic_batch(4, 15)
    xg = _t(gg["xg"])
    with torch.no_grad():
        assert rel_err(O.dis_forward(xg, dsd), _t(gold["dis.out"])) <= 1e-4
        assert abs(float(O.dis_loss(batch["img_xt"], dsd, target=1.0)) - float(gold["dis.real_loss"])) <= 1e-5
        assert abs(float(O.dis_loss(xg, dsd, target=0.0)) - float(gold["dis.fake_loss"])) <= 1e-5
        assert abs(float(O.cla_loss(batch["img_xt"], batch["tr_wid"], csd)) - float(gold["cla.loss"])) <= 1e-4


def test_train_mode_batch_of_one_raises(specs):
    sd = W.make_state(specs["gen_c15"])
    batch = O.synthetic_batch(1, 15)
    with pytest.raises(ValueError):
        O.text_encoder(batch["label_xt"], (1, 512, 8, 27), sd, "enc_text.")


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_resnet_encoder_oracle_matches_reference_fixture(arch, golden):
    """the oracle's torchvision-ResNet encoder against the fp64 run of the reference class
    (tests/golden/resnet_enc.npz, written by oracle/make_golden_resnet.py)."""
    g = golden("resnet_enc.npz")
    spec = json.load(open(os.path.join(GOLDEN, "resnet_spec.json")))[arch]
    rep = json.load(open(os.path.join(GOLDEN, "oracle_vs_reference_resnet.json")))
    assert all(r["max_abs"] <= r["tol"] for r in rep) and len(rep) >= 14
    sd = {k: v.requires_grad_(v.is_floating_point()) for k, v in W.alias_extractor(W.make_state(spec)).items()}
    x = O.synthetic_batch(2, 50)["tr_img"].requires_grad_()
    stats = {}
    res = O.resnet_encoder(x, sd, "", arch, True, stats)
    sum(r.square().mean() for r in res).backward()
    noise_f, noise_g = float(g[f"{arch}.noise.fwd"]), float(g[f"{arch}.noise.dx"])
    for i, r in enumerate(res):
        assert list(r.shape) == g[f"{arch}.result{i}.shape"].tolist()
        assert rel_err(r[:, :8], _t(g[f"{arch}.result{i}.head"])) <= 1e-5 + 3 * noise_f
    assert [int(v) for v in res[-1].shape] == [2, 512, 8, 27]
    assert abs(float(x.grad.norm()) / float(g[f"{arch}.dx.norm"]) - 1) <= 1e-4 + 3 * noise_g
    assert int(stats["model.layer3.1.bn1.num_batches_tracked"]) == 1
    assert rel_err(stats["model.bn1.running_mean"], _t(g[f"{arch}.post.model.bn1.running_mean"])) <= 1e-5


@pytest.mark.parametrize("case", ["nb384_c50", "nb512_c3"])
def test_resnet18_standalone_oracle_matches_reference_fixture(case, golden):
    """SURVEY.md §8 row a9, `Resnet18.py:9-88` (row stride 2 / column stride 1 stem and pool): the oracle against the fp64 run
    of the reference module (tests/golden/resnet18_standalone.npz, oracle/make_golden_resnet18_standalone.py)."""
    g = golden("resnet18_standalone.npz")
    spec = json.load(open(os.path.join(GOLDEN, "resnet18_standalone_spec.json")))[case]
    rep = json.load(open(os.path.join(GOLDEN, "oracle_vs_reference_resnet18_standalone.json")))
    assert all(r["max_abs"] <= r["tol"] for r in rep) and len(rep) >= 14
    batch, cin = (2, 50) if case == "nb384_c50" else (3, 3)
    sd = {k: v.requires_grad_(v.is_floating_point()) for k, v in W.make_state(spec).items()}
    x = O.synthetic_batch(batch, cin)["tr_img"].requires_grad_()
    stats = {}
    res = O.resnet18_standalone(x, sd, "", True, stats)
    sum(r.square().mean() for r in res).backward()
    noise_f, noise_g = float(g[f"{case}.noise.fwd"]), float(g[f"{case}.noise.dx"])
    for i, r in enumerate(res):
        assert list(r.shape) == g[f"{case}.result{i}.shape"].tolist()
        assert rel_err(r[:, :8], _t(g[f"{case}.result{i}.head"])) <= 1e-5 + 3 * noise_f
    assert list(res[0].shape[2:]) == [16, 216] and list(res[4].shape[2:]) == [2, 27]      # rows / 4, columns kept
    assert abs(float(x.grad.norm()) / float(g[f"{case}.dx.norm"]) - 1) <= 1e-4 + 3 * noise_g
    assert rel_err(sd["conv1.weight"].grad[:8], _t(g[f"{case}.grad.conv1"])) <= 1e-4 + 3 * noise_g
    assert int(stats["layer2.0.bn1.num_batches_tracked"]) == 1
    assert rel_err(stats["bn1.running_mean"], _t(g[f"{case}.post.bn1.running_mean"])) <= 1e-5


def test_u8_wire_format_oracle_matches_reference_loader(golden):
    """the oracle's restatement of load_data.py:152-166 against float32 canvases returned by the
    reference's own `read_image_single` (tests/golden/u8_wire.npz, oracle/make_golden_u8.py) - bit for bit."""
    g = golden("u8_wire.npz")
    levels = set()
    for i in range(int(g["count"])):
        u8, ref = g[f"u8.{i}"], g[f"ref.{i}"]
        mine, width = O.normalize_resized_u8(u8)
        assert width == int(g[f"width.{i}"]) and mine.dtype == np.float32 and np.array_equal(mine, ref)
        wire = O.pad_resized_u8(u8)
        assert wire.dtype == np.uint8 and wire.shape == (64, 216) and np.array_equal(O.decode_u8(wire), ref)
        levels |= set(np.unique(u8).tolist())
    assert len(levels) == 256                                    # every grey level went through the reference
    assert O.decode_u8(np.array([255, 0], dtype=np.uint8)).tolist() == [-1.0, 1.0]


@pytest.mark.parametrize("case,batch", [("b3", 3), ("b2", 2)])
def test_recogniser_oracle_matches_reference_fixture(case, batch, golden):
    """SURVEY.md §8(f).1 (next row, oracle only so far): oracle.rec_oracle against logits the reference's RecModel produced
    under the same torch seed (tests/golden/rec.npz, oracle/make_golden_rec.py).  Dropout is active on this path, so the pin
    also checks that the restatement draws its random numbers in the reference's order."""
    from oracle import rec_oracle as R
    g = golden("rec.npz")
    rep = json.load(open(os.path.join(GOLDEN, "oracle_vs_reference_rec.json")))
    assert all(r["max_abs"] <= r["tol"] for r in rep) and len(rep) >= 6
    sd = W.make_state(json.load(open(os.path.join(GOLDEN, "rec_spec.json"))))
    b = O.synthetic_batch(batch, 15)
    ref = _t(g[f"{case}.logits"])
    stats = {}
    torch.manual_seed(int(g[f"{case}.seed"]))
    with torch.no_grad():
        out = R.rec_forward(b["img_xt"], b["label_xt"], sd, [216] * batch, True, stats)
    assert out.shape == ref.shape == (batch, 11, 55)
    ok = ~torch.isnan(ref)
    assert torch.equal(torch.isnan(out), ~ok)
    assert float((out - ref)[ok].abs().max()) <= 1e-4 * max(1.0, float(ref[ok].abs().max()))
    assert torch.equal(out.argmax(-1), _t(g[f"{case}.tokens"]))
    target = b["label_xt"][:, 1:]
    assert bool(torch.isnan(R.label_smoothing_loss(out, target))) == bool(g[f"{case}.loss_is_nan"])
    loss = float(R.label_smoothing_loss(torch.nan_to_num(out, nan=0.0), target))
    assert abs(loss - float(g[f"{case}.loss_nan_zeroed"])) <= 1e-4 * max(1.0, abs(loss))
    k = "seq2seq.encoder.layer.features.1."
    assert int(stats[k + "num_batches_tracked"]) == 1
    assert rel_err(stats[k + "running_mean"], _t(g[f"{case}.post.features1.running_mean"])) <= 1e-5
    # backward of rec_update (network_tro.py:39-48) against the reference's gradients: per-tensor norms and the image gradient
    sdo = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
    x = b["img_xt"].clone().requires_grad_()
    torch.manual_seed(int(g[f"{case}.seed"]))
    R.label_smoothing_loss(R.rec_forward(x, b["label_xt"], sdo, [216] * batch, True, None), target).backward()
    ref_norm = dict(zip(g[f"{case}.grad.keys"].tolist(), g[f"{case}.grad.norms"].tolist()))
    big = max(ref_norm.values())
    for key, n in ref_norm.items():
        mine = float(sdo[key].grad.norm())
        assert abs(mine - n) <= 1e-2 * n + 1e-4 * big, (key, mine, n)      # structurally zero gradients: absolute bound only
    dx_head = x.grad[:, :, ::8, ::8]
    assert abs(float(x.grad.norm()) / float(g[f"{case}.dx.norm"]) - 1) <= 1e-2
    assert float((dx_head.flatten() @ _t(g[f"{case}.dx.head"]).flatten()) /
                 (dx_head.norm() * _t(g[f"{case}.dx.head"]).norm())) >= 0.999
    # the path consumes random numbers (Dropout2d + GRU dropout under the forced train()): another seed, other logits
    torch.manual_seed(int(g[f"{case}.seed"]) + 100)
    with torch.no_grad():
        other = R.rec_forward(b["img_xt"], b["label_xt"], sd, [216] * batch, True, None)
    assert float((other - out)[~torch.isnan(other) & ok].abs().max()) > 1e-3


def test_recogniser_oracle_with_explicit_dropout_masks(golden):
    """The mask-injectable form of the recogniser oracle (the comparison target for a device implementation, which cannot
    reproduce torch's CPU generator): drawing the masks with torch's generator in the documented order reproduces the
    `_VF.gru` / F.dropout2d form - and so the reference - bit for bit up to rounding; feeding the recorded masks back in gives the
    same logits again."""
    from oracle import rec_oracle as R
    g = golden("rec.npz")
    sd = W.make_state(json.load(open(os.path.join(GOLDEN, "rec_spec.json"))))
    b = O.synthetic_batch(2, 15)
    ref = _t(g["b2.logits"])
    with torch.no_grad():
        rec = {}
        torch.manual_seed(int(g["b2.seed"]))
        drawn = R.rec_forward_explicit(b["img_xt"], b["label_xt"], sd, record=rec)
        assert float((drawn - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
        assert tuple(rec["enc.drop2d"].shape) == (2, 512, 1, 1) and tuple(rec["enc.gru"].shape) == (13, 2, 1024)
        assert len(rec) == 2 + 2 * (1 + 10 * 3) and set(rec["enc.gru"].unique().tolist()) == {0.0, 2.0}
        again = R.rec_forward_explicit(b["img_xt"], b["label_xt"], sd, masks=rec)
        assert torch.equal(again, drawn)
        # the device-shaped prototype: one batched decoder call per step over all live hypotheses, selection on the host
        batched = R.rec_forward_batched(b["img_xt"], b["label_xt"], sd, rec)
        assert float((batched - drawn).abs().max()) <= 1e-5 and torch.equal(batched.argmax(-1), drawn.argmax(-1))
        # other masks, other logits; no masks at all (eval-style call) is yet another function
        flipped = {k: (2.0 - v) for k, v in rec.items()}
        assert float((R.rec_forward_explicit(b["img_xt"], b["label_xt"], sd, masks=flipped) - drawn).abs().max()) > 1e-3
