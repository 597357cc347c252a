"""Golden vectors for the torchvision-ResNet style encoders, produced by the UNMODIFIED reference:
  resnet50: modules_tro.ImageEncoderResNet50 imported in place (GAN_word/modules_tro.py:464-533)
  resnet18: the class body of GAN_word/modules_tro2.py:447-516 executed from the reference file (that module itself is not
            importable); nothing is copied into the repo.
Container-only:  python -m oracle.make_golden_resnet      (TEST INFRASTRUCTURE)
Writes tests/golden/resnet_enc.npz + resnet_spec.json and checks oracle.resnet_encoder against the reference."""
import json
import os
import re

import numpy as np
import torch

from oracle import affgw_oracle as O
from oracle import ref_bootstrap as rb
from oracle import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SLICE = 8        # channels of every map kept in the fixture


alias_extractor = W.alias_extractor


def ref_class(ns, arch):
    if arch == "resnet50":
        return ns.modules_tro.ImageEncoderResNet50
    src = open(os.path.join(rb.REF_WORD, "modules_tro2.py")).read()
    body = src[src.index("class ImageEncoderResNet50(nn.Module):"):]
    body = body[:re.search(r"\n(?=\S)", body[10:]).start() + 10]          # up to the next top-level statement
    from torch import nn
    import torch.nn.functional as F
    from torchvision.models import resnet18
    from torchvision.models.feature_extraction import create_feature_extractor
    scope = dict(nn=nn, torch=torch, F=F, resnet18=resnet18, create_feature_extractor=create_feature_extractor)
    exec(compile(body, "modules_tro2.py:447-516", "exec"), scope)
    return scope["ImageEncoderResNet50"]


def main():
    torch.manual_seed(0)
    ns = rb.load(50)
    out, spec_out, report = {}, {}, []
    x = O.synthetic_batch(2, 50)["tr_img"]
    for arch in ("resnet18", "resnet50"):
        enc = ref_class(ns, arch)(weight_path=None, in_channels=50).train()
        spec = W.spec_of(enc)
        sd = alias_extractor(W.make_state(spec))
        # the reference in fp32 (what it really computes) ...
        enc.load_state_dict(sd)
        x32 = x.clone().requires_grad_()
        res32 = enc(x32)
        sum(r.square().mean() for r in res32).backward()
        g32 = {k: p.grad.clone() for k, p in enc.named_parameters() if p.grad is not None}
        enc.zero_grad()
        # ... and in fp64: these deep BatchNorm stacks at batch 2 amplify fp32 rounding (ResNet-50: ~3 % on dx), so the
        # fixtures hold the fp64 run of the reference and the fp32-vs-fp64 gap of the reference itself as the noise floor
        enc.load_state_dict(sd)
        enc = enc.double()
        xr = x.double().clone().requires_grad_()
        res = enc(xr)
        loss = sum(r.square().mean() for r in res)
        loss.backward()
        post = {k: v.clone() for k, v in enc.state_dict().items()}
        sdo = {k: (v.double().clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        xo = x.double().clone().requires_grad_()
        stats = {}
        reso = O.resnet_encoder(xo, sdo, "", arch, True, stats)
        sum(r.square().mean() for r in reso).backward()
        for i, (a, b) in enumerate(zip(res, reso)):
            err = float((a - b).abs().max() / max(1.0, float(a.abs().max())))
            report.append({"name": f"{arch}.result{i}", "max_abs": err, "tol": 1e-9})
            assert err <= 1e-9, (arch, i, err)
        err = float((xr.grad - xo.grad).abs().max() / float(xr.grad.abs().max()))
        report.append({"name": f"{arch}.dx", "max_abs": err, "tol": 1e-8})
        assert err <= 1e-8, err
        for k, v in stats.items():
            e = float((post[k].double() - v.double()).abs().max())
            assert e <= 1e-9, (k, e)
        worst = 0.0
        for k, p in enc.named_parameters():
            if p.grad is None:
                assert sdo[k].grad is None or float(sdo[k].grad.abs().max()) == 0.0, k
                continue
            e = float((p.grad - sdo[k].grad).abs().max() / max(1e-30, float(p.grad.abs().max())))
            worst = max(worst, e)
        report.append({"name": f"{arch}.param_grads", "max_abs": worst, "tol": 1e-7})
        assert worst <= 1e-7, worst
        noise_dx = float((x32.grad.double() - xr.grad).abs().max() / xr.grad.abs().max())
        noise_fw = max(float((a.double() - b).abs().max() / max(1.0, float(b.abs().max()))) for a, b in zip(res32, res))
        out[f"{arch}.noise.dx"] = np.float32(noise_dx)
        out[f"{arch}.noise.fwd"] = np.float32(noise_fw)
        print(f"{arch}: reference fp32 vs its own fp64: forward {noise_fw:.2e}, dx {noise_dx:.2e}")
        res = [r.float() for r in res]
        xr_grad = xr.grad.float()
        # fixture
        spec_out[arch] = spec
        for i, r in enumerate(res):
            out[f"{arch}.result{i}.shape"] = np.array(r.shape)
            out[f"{arch}.result{i}.head"] = r[:, :SLICE].detach().numpy().copy()
            out[f"{arch}.result{i}.abs_mean"] = np.float32(r.abs().mean().item())
        out[f"{arch}.loss"] = np.float32(loss.item())
        out[f"{arch}.dx.head"] = xr_grad[:, :4, ::8, ::8].numpy().copy()
        out[f"{arch}.dx.norm"] = np.float32(xr_grad.norm().item())
        keys, norms = [], []
        for k, p in enc.named_parameters():
            keys.append(k)
            norms.append(-1.0 if p.grad is None else float(p.grad.norm()))
        out[f"{arch}.grad.keys"] = np.array(keys)
        out[f"{arch}.grad.norms"] = np.array(norms, dtype=np.float64)
        for k in ("model.bn1.running_mean", "model.layer4.0.bn2.running_var", "model.layer2.0.downsample.1.running_mean",
                  "model.layer3.1.bn1.num_batches_tracked"):
            out[f"{arch}.post.{k}"] = (post[k].float() if post[k].is_floating_point() else post[k]).numpy().copy()
        print(arch, "ok: loss", float(loss), "worst param-grad rel err", worst)
    np.savez_compressed(os.path.join(OUT, "resnet_enc.npz"), **out)
    json.dump(spec_out, open(os.path.join(OUT, "resnet_spec.json"), "w"))
    json.dump(report, open(os.path.join(OUT, "oracle_vs_reference_resnet.json"), "w"), indent=1)
    print("wrote resnet_enc.npz,", os.path.getsize(os.path.join(OUT, "resnet_enc.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
