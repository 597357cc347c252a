"""Native recogniser (affganwriting_b200.recognizer.RecModel) against the CPU oracle's mask-injectable
restatement of the reference (oracle.rec_oracle.rec_forward_explicit, itself pinned to the unmodified reference in
tests/test_oracle_golden.py): same weights, same images, the SAME dropout keep-masks (recorded from the oracle's draw and
injected), fp32 mode.  Checked: logits of the best hypotheses, arg-max tokens, BatchNorm running statistics, the
label-smoothed loss and the rec_update gradients of every parameter."""
import json
import os

import numpy as np
import pytest
import torch

import affganwriting_b200 as A
from affganwriting_b200 import ops
from affganwriting_b200.recognizer import RecModel
from affgw_testutil import cosine
from oracle import affgw_oracle as O
from oracle import rec_oracle as R
from oracle import weights as W

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _alias(sd):
    out = dict(sd)
    for k, v in sd.items():
        if k.startswith("seq2seq.encoder."):
            out["enc." + k[len("seq2seq.encoder."):]] = v
        elif k.startswith("seq2seq.decoder."):
            out["dec." + k[len("seq2seq.decoder."):]] = v
    return out


def _state():
    spec = json.load(open(os.path.join(GOLDEN, "rec_spec.json")))
    spec = spec.get("spec", spec)
    return W.make_state({k: v for k, v in spec.items() if k.startswith("seq2seq.")})


@pytest.mark.parametrize("batch,seed", [(3, 11), (2, 12)])
def test_native_recogniser_matches_oracle(batch, seed):
    sd = _state()
    leaves = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    b = O.synthetic_batch(batch, 15)
    img, lab = b["img_xt"], b["label_xt"]
    record, stats = {}, {}
    torch.manual_seed(seed)
    ref = R.rec_forward_explicit(img, lab, leaves, masks=None, record=record, stats=stats)
    target = lab[:, 1:]
    loss_ref = R.label_smoothing_loss(ref, target)
    loss_ref.backward()

    A.set_precision("fp32")
    rec = RecModel().cuda()
    assert list(rec.state_dict().keys())[:2] == ["enc.layer.features.0.weight", "enc.layer.features.0.bias"]
    rec.load_state_dict(_alias(sd))
    n0 = A.launch_count()
    out = rec(img.cuda(), lab.cuda(), img_width=torch.from_numpy(np.array([216] * batch)), masks=record)
    assert out.shape == ref.shape == (batch, 11, 55)
    got = out.detach().cpu()
    nan_r, nan_g = torch.isnan(ref.detach()), torch.isnan(got)
    assert torch.equal(nan_r, nan_g)
    err = float((ref.detach() - got)[~nan_r].abs().max() / max(1.0, float(ref.detach()[~nan_r].abs().max())))
    same_tokens = torch.equal(ref.detach().argmax(-1), got.argmax(-1))
    print(f"\n[fp32] native recogniser vs oracle (batch {batch}): logits rel max error {err:.2e}, arg-max tokens identical: "
          f"{same_tokens}, {A.launch_count() - n0} libaffgw launches")
    assert err <= 2e-4 and same_tokens
    # BatchNorm running statistics after the one training-mode call
    post = rec.state_dict()
    for leaf in ("running_mean", "running_var"):
        k = "seq2seq.encoder.layer.features.1." + leaf
        assert float((post[k].cpu() - stats[k]).abs().max()) <= 1e-5, k
    assert int(post["seq2seq.encoder.layer.features.1.num_batches_tracked"]) == 1
    # loss + rec_update gradients (network_tro.py:44-47)
    loss = ops.label_smoothing_kl(out.reshape(-1, 55), target.cuda().reshape(-1), 2, 0.4)
    if torch.isnan(loss_ref):
        assert torch.isnan(loss)
        return
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
    loss.backward()
    top = max(float(v.grad.norm()) for v in leaves.values() if v.grad is not None)
    worst, n_checked = 1.0, 0
    for k, p in rec.named_parameters():
        # named_parameters() lists the shared Parameter objects once, under their first names enc.* / dec.*
        k = ("seq2seq.encoder." + k[4:]) if k.startswith("enc.") else ("seq2seq.decoder." + k[4:]) if k.startswith("dec.") else k
        g = leaves[k].grad
        if g is None or float(g.norm()) < 1e-4 * top:   # unused (attention.proj) or exact-zero (biases in front of BatchNorm)
            continue
        assert p.grad is not None, k
        c = cosine(p.grad, g)
        worst = min(worst, c)
        n_checked += 1
        assert c >= 0.999, (k, c)
        assert abs(float(p.grad.norm()) / float(g.norm()) - 1.0) <= 1e-2, k
    print(f"  rec_update gradients: {n_checked} tensors, worst cosine {worst:.6f}")
    assert n_checked >= 40
    A.check_device_errors()


def test_native_recogniser_bf16_and_device_masks():
    """Shipping precision, masks drawn on the device: shapes, finiteness, determinism under a fixed seed, gradients to every
    live parameter and to the image (gen_update differentiates the recogniser with respect to the generated image)."""
    sd = _state()
    A.set_precision("bf16")
    try:
        rec = RecModel().cuda()
        rec.load_state_dict(_alias(sd))
        b = O.synthetic_batch(4, 15)
        img = ops.to_internal(b["img_xt"].cuda()).requires_grad_()
        lab = b["label_xt"].cuda()
        torch.manual_seed(5)
        a = rec(img, lab, img_width=torch.from_numpy(np.array([216] * 4)))
        rec.load_state_dict(_alias(sd))
        torch.manual_seed(5)
        c = rec(img, lab, img_width=torch.from_numpy(np.array([216] * 4)))
        assert a.shape == (4, 11, 55) and torch.isfinite(a).all()
        assert float((a - c).abs().max()) <= 5e-3 * float(a.abs().max())
        ops.label_smoothing_kl(c.reshape(-1, 55), lab[:, 1:].reshape(-1), 2, 0.4).backward()
        assert img.grad is not None and torch.isfinite(img.grad).all() and float(img.grad.abs().max()) > 0
        missing = [k for k, p in rec.named_parameters() if p.grad is None and "attention.proj." not in k]
        assert not missing, missing
        with pytest.raises(RuntimeError):
            rec(img, lab, img_width=torch.from_numpy(np.array([216, 216, 100, 216])))
    finally:
        A.set_precision("fp32")
