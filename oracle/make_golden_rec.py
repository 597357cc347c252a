"""Golden vectors for the recogniser as the GAN step calls it, produced by the UNMODIFIED reference:
`modules_tro.RecModel` (GAN_word/modules_tro.py:610-638 -> recognizer/models/{encoder_vgg,decoder,attention,seq2seqnew2}.py)
imported in place with the shims of oracle/ref_bootstrap.py (+ PRE_TRAIN_VGG = False, SURVEY.md §8c shim 6); nothing is copied.
Container-only:  python -m oracle.make_golden_rec      (TEST INFRASTRUCTURE)
Writes tests/golden/rec.npz + rec_spec.json and checks oracle.rec_oracle against the reference under the same torch seed
(dropout is active in this path: see the header of oracle/rec_oracle.py)."""
import json
import os
import warnings

import numpy as np
import torch

from oracle import affgw_oracle as O
from oracle import rec_oracle as R
from oracle import ref_bootstrap as rb
from oracle import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 11


def alias(sd):
    """`enc.*` / `dec.*` are the same modules as `seq2seq.encoder.*` / `seq2seq.decoder.*` (modules_tro.py:624-626)."""
    out = dict(sd)
    for k, v in sd.items():
        if k.startswith("seq2seq.encoder."):
            out["enc." + k[len("seq2seq.encoder."):]] = v
        elif k.startswith("seq2seq.decoder."):
            out["dec." + k[len("seq2seq.decoder."):]] = v
    return out


def main():
    warnings.filterwarnings("ignore")
    ns = rb.load(50)
    import recognizer.models.encoder_vgg as ev
    ev.PRE_TRAIN_VGG = False
    rec = ns.modules_tro.RecModel(pretrain=False)
    full = W.spec_of(rec)
    spec = {k: v for k, v in full.items() if k.startswith("seq2seq.")}
    assert set(full) == set(alias(spec)), "unexpected key families in RecModel.state_dict()"
    sd = W.make_state(spec)
    rec.load_state_dict(alias(sd))
    out, report = {}, []
    for case, (batch, seed) in {"b3": (3, SEED), "b2": (2, SEED + 1)}.items():
        b = O.synthetic_batch(batch, 15)
        img, lab = b["img_xt"], b["label_xt"]
        widths = torch.from_numpy(np.array([img.shape[-1]] * batch))
        rec.load_state_dict(alias(sd))                          # reset the BatchNorm running statistics
        torch.manual_seed(seed)
        ref = rec(img, lab, img_width=widths).detach()
        post = {k: v.clone() for k, v in rec.state_dict().items()}
        torch.manual_seed(seed)
        stats = {}
        with torch.no_grad():
            mine = R.rec_forward(img, lab, sd, widths.numpy(), True, stats)
        assert mine.shape == ref.shape == (batch, 11, 55)
        nan_ref, nan_mine = torch.isnan(ref), torch.isnan(mine)
        assert torch.equal(nan_ref, nan_mine)
        err = float((ref - mine)[~nan_ref].abs().max() / max(1.0, float(ref[~nan_ref].abs().max())))
        same_tokens = torch.equal(ref.argmax(-1), mine.argmax(-1))
        report.append({"name": f"rec.{case}.logits", "max_abs": err, "tol": 1e-4})
        print(f"{case}: oracle vs reference logits {err:.2e}, arg-max tokens identical: {same_tokens}")
        assert err <= 1e-4 and same_tokens, (case, err)
        bnp = "seq2seq.encoder.layer.features.1."
        for leaf in ("running_mean", "running_var"):
            e = float((post[bnp + leaf] - stats[bnp + leaf]).abs().max())
            report.append({"name": f"rec.{case}.bn.{leaf}", "max_abs": e, "tol": 1e-5})
            assert e <= 1e-5, (leaf, e)
        assert int(post[bnp + "num_batches_tracked"]) == int(stats[bnp + "num_batches_tracked"]) == 1
        # dropout really is active on this path: another seed gives other logits
        torch.manual_seed(seed + 100)
        other = rec(img, lab, img_width=widths).detach()
        moved = float((other - ref).abs().max())
        assert moved > 1e-3, "the recogniser call did not consume random numbers"
        # the loss the GAN step forms from these logits (network_tro.py:44-45, loss_tro.py:8-35); NaN logits give a NaN loss
        lt = ns.load_data  # noqa: F841
        import loss_tro
        target = lab[:, 1:]
        l_ref = loss_tro.crit(loss_tro.log_softmax(ref.reshape(-1, 55)), target.reshape(-1))
        l_mine = R.label_smoothing_loss(mine, target)
        if torch.isnan(l_ref):
            assert torch.isnan(l_mine)
        else:
            assert abs(float(l_ref) - float(l_mine)) <= 1e-4 * max(1.0, abs(float(l_ref))), (float(l_ref), float(l_mine))
        finite = torch.nan_to_num(ref, nan=0.0)
        l_ref_f = loss_tro.crit(loss_tro.log_softmax(finite.reshape(-1, 55)), target.reshape(-1))
        l_mine_f = R.label_smoothing_loss(torch.nan_to_num(mine, nan=0.0), target)
        e = abs(float(l_ref_f) - float(l_mine_f)) / max(1.0, abs(float(l_ref_f)))
        report.append({"name": f"rec.{case}.loss", "max_abs": e, "tol": 1e-4})
        assert e <= 1e-4, e
        print(f"{case}: label-smoothed KL loss {float(l_ref):.4f} (NaN logits zeroed: {float(l_ref_f):.4f}), oracle differs by {e:.1e}")
        # backward of rec_update (network_tro.py:39-48): gradients of the loss w.r.t. the image and the parameters.  The reference
        # cannot run in fp64 (attention.py:144 builds a FloatTensor), and this 16-layer train-mode BatchNorm stack at batch 2-3
        # amplifies fp32 rounding (single weight-gradient elements of two fp32 evaluations differ by up to several %), so the
        # pin is per-tensor direction + norm; conv biases in front of a BatchNorm and attention.out.bias have a structurally
        # zero gradient (1e-8 of noise) and are only required to stay that small.
        canon = lambda k: ("seq2seq.encoder." + k[4:]) if k.startswith("enc.") else ("seq2seq.decoder." + k[4:]) if \
            k.startswith("dec.") else k                        # named_parameters() lists shared modules under their first name
        rec.load_state_dict(alias(sd))
        rec.zero_grad()
        x_ref = img.clone().requires_grad_()
        torch.manual_seed(seed)
        p_ref = rec(x_ref, lab, img_width=widths)
        loss_tro.crit(loss_tro.log_softmax(p_ref.reshape(-1, 55)), target.reshape(-1)).backward()
        g_ref = {canon(k): p.grad.clone() for k, p in rec.named_parameters() if p.grad is not None}
        assert all(k in sd for k in g_ref)
        sdo = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        x_mine = img.clone().requires_grad_()
        torch.manual_seed(seed)
        R.label_smoothing_loss(R.rec_forward(x_mine, lab, sdo, widths.numpy(), True, None), target).backward()
        cos = lambda a, c: float((a.double().flatten() @ c.double().flatten()) / (a.double().norm() * c.double().norm() + 1e-300))  # noqa: E731
        gmax = max(float(g.abs().max()) for g in g_ref.values())
        worst_cos, worst_norm, zeros = cos(x_ref.grad, x_mine.grad), 0.0, 0
        for k, gr in g_ref.items():
            assert sdo[k].grad is not None, k
            if float(gr.abs().max()) < 1e-4 * gmax:
                zeros += 1
                assert float(sdo[k].grad.abs().max()) < 1e-4 * gmax, k
                continue
            worst_cos = min(worst_cos, cos(gr, sdo[k].grad))
            if cos(gr, sdo[k].grad) < 0.999 or abs(float(sdo[k].grad.norm() / gr.norm()) - 1.0) > 1e-2:
                print('   ', k, cos(gr, sdo[k].grad), float(sdo[k].grad.norm() / gr.norm()), float(gr.abs().max()) / gmax)
            worst_norm = max(worst_norm, abs(float(sdo[k].grad.norm() / gr.norm()) - 1.0))
        silent = [k for k, v in sdo.items() if v.is_floating_point() and v.grad is not None and k not in g_ref]
        assert not silent, silent[:3]
        report.append({"name": f"rec.{case}.grad_cosine_deficit", "max_abs": 1.0 - worst_cos, "tol": 1e-3})
        report.append({"name": f"rec.{case}.grad_norm", "max_abs": worst_norm, "tol": 1e-2})
        print(f"{case}: gradients of the rec_update loss (image + {len(g_ref)} parameters, {zeros} structurally zero): "
              f"worst cosine {worst_cos:.6f}, worst norm deviation {worst_norm:.1e}")
        assert worst_cos >= 1 - 1e-3 and worst_norm <= 1e-2
        out[f"{case}.dx.norm"] = np.float32(float(x_ref.grad.norm()))
        out[f"{case}.dx.head"] = x_ref.grad[:, :, ::8, ::8].numpy().copy()
        keys = sorted(g_ref)
        out[f"{case}.grad.keys"] = np.array(keys)
        out[f"{case}.grad.norms"] = np.array([float(g_ref[k].norm()) for k in keys], dtype=np.float64)
        out[f"{case}.loss_nan_zeroed"] = np.float32(float(l_ref_f))
        out[f"{case}.loss_is_nan"] = np.bool_(bool(torch.isnan(l_ref)))
        out[f"{case}.logits"] = ref.numpy().copy()
        out[f"{case}.seed"] = np.int64(seed)
        out[f"{case}.tokens"] = ref.argmax(-1).numpy().copy()
        out[f"{case}.other_seed_delta"] = np.float32(moved)
        out[f"{case}.post.features1.running_mean"] = post[bnp + "running_mean"].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "rec.npz"), **out)
    json.dump(spec, open(os.path.join(OUT, "rec_spec.json"), "w"))
    json.dump(report, open(os.path.join(OUT, "oracle_vs_reference_rec.json"), "w"), indent=1)
    print("wrote rec.npz,", os.path.getsize(os.path.join(OUT, "rec.npz")) // 1024, "KB;", len(spec), "tensors in the spec")


if __name__ == "__main__":
    main()
