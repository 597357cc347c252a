"""Golden vectors for the reference's stand-alone ResNet-18 (GAN_word/Resnet18.py:4-88), produced
by the UNMODIFIED reference module imported in place; nothing is copied into the repo.
Container-only:  python -m oracle.make_golden_resnet18_standalone      (TEST INFRASTRUCTURE)
Writes tests/golden/resnet18_standalone.npz + resnet18_standalone_spec.json and checks oracle.resnet18_standalone against
the reference in fp64 (fixtures hold the fp64 run; the reference's own fp32-vs-fp64 gap is stored as the noise floor)."""
import importlib
import json
import os
import sys

import numpy as np
import torch

from oracle import affgw_oracle as O
from oracle import ref_bootstrap as rb
from oracle import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SLICE = 8
CASES = {"nb384_c50": dict(nb_feat=384, in_channels=50, batch=2), "nb512_c3": dict(nb_feat=512, in_channels=3, batch=3)}


def main():
    sys.path.insert(0, rb.REF_WORD)
    R = importlib.import_module("Resnet18")
    assert os.path.abspath(R.__file__).startswith(rb.REF_WORD)
    out, spec_out, report = {}, {}, []
    for name, c in CASES.items():
        torch.manual_seed(0)
        net = R.ResNet18(nb_feat=c["nb_feat"], in_channels=c["in_channels"]).train()
        spec = W.spec_of(net)
        sd = W.make_state(spec)
        x = O.synthetic_batch(c["batch"], c["in_channels"])["tr_img"]
        net.load_state_dict(sd)
        x32 = x.clone().requires_grad_()
        res32 = net(x32)
        sum(r.square().mean() for r in res32).backward()
        net.zero_grad()
        net.load_state_dict(sd)
        net = net.double()
        xr = x.double().clone().requires_grad_()
        res = net(xr)
        loss = sum(r.square().mean() for r in res)
        loss.backward()
        post = {k: v.clone() for k, v in net.state_dict().items()}
        sdo = {k: (v.double().clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        xo = x.double().clone().requires_grad_()
        stats = {}
        reso = O.resnet18_standalone(xo, sdo, "", True, stats)
        sum(r.square().mean() for r in reso).backward()
        assert len(res) == len(reso) == 5
        for i, (a, b) in enumerate(zip(res, reso)):
            assert a.shape == b.shape
            err = float((a - b).abs().max() / max(1.0, float(a.abs().max())))
            report.append({"name": f"{name}.result{i}", "max_abs": err, "tol": 1e-9})
            assert err <= 1e-9, (name, i, err)
        err = float((xr.grad - xo.grad).abs().max() / float(xr.grad.abs().max()))
        report.append({"name": f"{name}.dx", "max_abs": err, "tol": 1e-8})
        assert err <= 1e-8, err
        for k, v in stats.items():
            e = float((post[k].double() - v.double()).abs().max())
            assert e <= 1e-9, (k, e)
        worst = 0.0
        for k, p in net.named_parameters():
            e = float((p.grad - sdo[k].grad).abs().max() / max(1e-30, float(p.grad.abs().max())))
            worst = max(worst, e)
        report.append({"name": f"{name}.param_grads", "max_abs": worst, "tol": 1e-7})
        assert worst <= 1e-7, worst
        noise_dx = float((x32.grad.double() - xr.grad).abs().max() / xr.grad.abs().max())
        noise_fw = max(float((a.double() - b).abs().max() / max(1.0, float(b.abs().max()))) for a, b in zip(res32, res))
        out[f"{name}.noise.dx"] = np.float32(noise_dx)
        out[f"{name}.noise.fwd"] = np.float32(noise_fw)
        print(f"{name}: reference fp32 vs its own fp64: forward {noise_fw:.2e}, dx {noise_dx:.2e}")
        spec_out[name] = spec
        for i, r in enumerate(res):
            r = r.float()
            out[f"{name}.result{i}.shape"] = np.array(r.shape)
            out[f"{name}.result{i}.head"] = r[:, :SLICE].detach().numpy().copy()
            out[f"{name}.result{i}.abs_mean"] = np.float32(r.abs().mean().item())
        g = xr.grad.float()
        out[f"{name}.loss"] = np.float32(loss.item())
        out[f"{name}.dx.head"] = g[:, :3, ::4, ::4].numpy().copy()
        out[f"{name}.dx.norm"] = np.float32(g.norm().item())
        keys, norms = [], []
        for k, p in net.named_parameters():
            keys.append(k)
            norms.append(float(p.grad.norm()))
        out[f"{name}.grad.keys"] = np.array(keys)
        out[f"{name}.grad.norms"] = np.array(norms, dtype=np.float64)
        out[f"{name}.grad.conv1"] = net.conv1.weight.grad.float()[:8].numpy().copy()
        for k in ("bn1.running_mean", "layer1.0.downsample.1.running_var", "layer3.1.bn2.running_mean",
                  "layer2.0.bn1.num_batches_tracked"):
            out[f"{name}.post.{k}"] = (post[k].float() if post[k].is_floating_point() else post[k]).numpy().copy()
        print(name, "ok: loss", float(loss), "shapes", [tuple(r.shape) for r in res], "worst param-grad rel err", worst)
    np.savez_compressed(os.path.join(OUT, "resnet18_standalone.npz"), **out)
    json.dump(spec_out, open(os.path.join(OUT, "resnet18_standalone_spec.json"), "w"))
    json.dump(report, open(os.path.join(OUT, "oracle_vs_reference_resnet18_standalone.json"), "w"), indent=1)
    print("wrote resnet18_standalone.npz,", os.path.getsize(os.path.join(OUT, "resnet18_standalone.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
