"""Generate tests/golden/* by running the UNMODIFIED reference (imported in place from /root/reference)
and, in the same pass, check oracle/affgw_oracle.py against it.  Container-only; run as

    python -m oracle.make_golden

TEST INFRASTRUCTURE.  The generated files are small (outputs, gradient digests, key/shape specs) because
weights are regenerated deterministically by oracle/weights.py.
"""
import json
import os
import sys

import numpy as np
import torch

from oracle import affgw_oracle as O
from oracle import ref_bootstrap as rb
from oracle import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
report = []


def check(name, a, b, tol):
    err = (a.detach() - b.detach()).abs().max().item() / max(1.0, a.detach().abs().max().item())
    report.append((name, err, tol))
    print(f"  oracle-vs-reference {name}: max-abs/max(1,|ref|) {err:.3e} (tol {tol:g})")
    assert err <= tol, name
    return err


def load_into(module, seed=0):
    spec = W.spec_of(module)
    sd = W.make_state(spec, seed)
    module.load_state_dict(sd)
    return spec, sd


def grad_digest(named_grads):
    """per-parameter L2 norm + first 8 values; None grads are recorded as norm -1."""
    keys, norms, heads = [], [], []
    for k, g in named_grads:
        keys.append(k)
        if g is None:
            norms.append(-1.0)
            heads.append(np.zeros(8, np.float32))
        else:
            f = g.detach().reshape(-1)
            norms.append(float(f.double().norm()))
            h = np.zeros(8, np.float32)
            h[:min(8, f.numel())] = f[:8].numpy()
            heads.append(h)
    return keys, np.array(norms, np.float64), np.stack(heads)


def cosine(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    ns = rb.load(15)
    B, M = ns.blocks, ns.modules_tro
    specs = {}

    # ---------------------------------------------------------------- labels (bit-exact pins)
    print("labels")

    class _Fake:
        output_max_len = ns.load_data.OUTPUT_MAX_LEN
    words = ["three", "a", "Z", "abcdefghij", "Hello", "zZ", "qwertyuiop", "I", "of", "Writing"]
    kat = {w: [int(v) for v in ns.load_data.IAM_words.label_padding(_Fake(), w, ns.load_data.num_tokens)]
           for w in words}
    for w, ids in kat.items():
        assert ids == O.label_padding(w), w
    assert kat["three"] == [0, 22, 10, 20, 7, 7, 1, 2, 2, 2, 2, 2]
    te = M.TextEncoder_FC(12)
    with torch.no_grad():
        te.embed.weight.copy_(torch.arange(55).float().view(55, 1).expand(55, 64))
        te.linear.weight.zero_()
        te.linear.weight[:, 0] = 1.0
        te.linear.bias.zero_()
    te.eval()
    lab = torch.tensor([[0, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 1]])
    colmaps = {}
    for width in (27, 12, 24, 30, 7):
        _, fe = te(lab, (1, 512, 8, width))
        vals = fe[0, 0, 0].round().long().tolist()
        cmap = [lab[0].tolist().index(v) if v != 2 else -1 for v in vals]
        assert cmap == O.text_column_map(width), (width, cmap)
        colmaps[str(width)] = cmap
    json.dump({"label_padding": kat, "text_column_map": colmaps,
               "vocab_size": ns.load_data.vocab_size, "tokens": ns.load_data.tokens},
              open(os.path.join(OUT, "labels.json"), "w"), indent=1)

    # ---------------------------------------------------------------- building blocks (a1-a7)
    print("blocks")
    blk = {}
    g = torch.Generator().manual_seed(7)

    def rnd(*s):
        return torch.randn(*s, generator=g)

    cases = {
        "conv_zero_in_relu": dict(ctor=dict(in_dim=6, out_dim=10, ks=3, st=1, padding=1, norm="in", activation="relu",
                                            pad_type="zero"), x=(2, 6, 9, 11)),
        "conv_reflect_none_tanh_k7": dict(ctor=dict(in_dim=8, out_dim=1, ks=7, st=1, padding=3, norm="none",
                                                    activation="tanh", pad_type="reflect"), x=(2, 8, 10, 13)),
        "conv_reflect_in_relu_k5": dict(ctor=dict(in_dim=8, out_dim=4, ks=5, st=1, padding=2, norm="in",
                                                  activation="relu", pad_type="reflect"), x=(2, 8, 6, 10)),
        "conv_actfirst_lrelu": dict(ctor=dict(in_dim=5, out_dim=7, ks=3, st=1, padding=1, norm="none",
                                              activation="lrelu", pad_type="reflect", activation_first=True),
                                    x=(3, 5, 8, 9)),
        "conv_head_k2s7": dict(ctor=dict(in_dim=16, out_dim=12, ks=2, st=7, norm="none", activation="lrelu",
                                         activation_first=True), x=(3, 16, 2, 7)),
        "conv_1x1_nobias": dict(ctor=dict(in_dim=8, out_dim=16, ks=1, st=1, activation="none", use_bias=False),
                                x=(2, 8, 5, 7)),
        "conv_replicate": dict(ctor=dict(in_dim=4, out_dim=4, ks=3, st=1, padding=1, norm="none", activation="relu",
                                         pad_type="replicate"), x=(1, 4, 5, 6)),
    }
    for name, c in cases.items():
        m = B.Conv2dBlock(**c["ctor"])
        spec, sd = load_into(m)
        x = rnd(*c["x"]).requires_grad_()
        y = m(x)
        gy = rnd(*y.shape)
        y.backward(gy)
        kw = dict(c["ctor"])
        kw.pop("in_dim"), kw.pop("out_dim"), kw.pop("use_bias", None)
        xo = x.detach().clone().requires_grad_()
        sdo = {k: v.clone().requires_grad_() for k, v in sd.items()}
        yo = O.conv2d_block(xo, sdo, "", **kw)
        yo.backward(gy)
        check(name + ".y", y, yo, 1e-5)
        check(name + ".dx", x.grad, xo.grad, 1e-4)
        check(name + ".dw", m.conv.weight.grad, sdo["conv.weight"].grad, 1e-4)
        blk[name + ".x"], blk[name + ".gy"] = x.detach().numpy(), gy.numpy()
        blk[name + ".y"], blk[name + ".dx"] = y.detach().numpy(), x.grad.numpy()
        blk[name + ".dw"] = m.conv.weight.grad.numpy()
        if m.conv.bias is not None:
            blk[name + ".db"] = m.conv.bias.grad.numpy()
        specs["blocks." + name] = dict(ctor=c["ctor"], spec=spec)

    # ResBlock with plain instance norm, ActFirstResBlock (both shortcut kinds), LinearBlock, MLP
    m = B.ResBlocks(2, 8, "in", "relu", "reflect")
    spec, sd = load_into(m)
    x = rnd(2, 8, 6, 9).requires_grad_()
    y = m(x); gy = rnd(*y.shape); y.backward(gy)
    xo = x.detach().clone().requires_grad_()
    yo = xo
    for i in range(2):
        yo = O.res_block(yo, sd, f"model.{i}.", "in", "relu", "reflect")
    yo.backward(gy)
    check("resblocks_in.y", y, yo, 1e-5); check("resblocks_in.dx", x.grad, xo.grad, 1e-4)
    blk.update({"resblocks_in.x": x.detach().numpy(), "resblocks_in.gy": gy.numpy(),
                "resblocks_in.y": y.detach().numpy(), "resblocks_in.dx": x.grad.numpy()})
    specs["blocks.resblocks_in"] = dict(ctor=dict(num_blocks=2, dim=8, norm="in", activation="relu",
                                                  pad_type="reflect"), spec=spec)

    for name, (fin, fout) in {"actfirst_same": (8, 8), "actfirst_grow": (8, 16)}.items():
        m = B.ActFirstResBlock(fin, fout, None, "lrelu", "none")
        spec, sd = load_into(m)
        x = rnd(2, fin, 7, 10).requires_grad_()
        y = m(x); gy = rnd(*y.shape); y.backward(gy)
        xo = x.detach().clone().requires_grad_()
        yo = O.act_first_res_block(xo, sd, "", fin, fout); yo.backward(gy)
        check(name + ".y", y, yo, 1e-5); check(name + ".dx", x.grad, xo.grad, 1e-4)
        blk.update({name + ".x": x.detach().numpy(), name + ".gy": gy.numpy(), name + ".y": y.detach().numpy(),
                    name + ".dx": x.grad.numpy()})
        specs["blocks." + name] = dict(ctor=dict(fin=fin, fout=fout), spec=spec)

    for name, kw in {"linear_bn_relu": dict(norm="bn", activation="relu"),
                     "linear_none_lrelu": dict(norm="none", activation="lrelu"),
                     "linear_none_tanh": dict(norm="none", activation="tanh")}.items():
        m = B.LinearBlock(24, 16, **kw)
        spec, sd = load_into(m)
        x = rnd(5, 24).requires_grad_()
        y = m(x); gy = rnd(*y.shape); y.backward(gy)
        xo = x.detach().clone().requires_grad_()
        yo = O.linear_block(xo, sd, "", kw["norm"], kw["activation"], True, {}); yo.backward(gy)
        check(name + ".y", y, yo, 1e-5); check(name + ".dx", x.grad, xo.grad, 1e-4)
        blk.update({name + ".x": x.detach().numpy(), name + ".gy": gy.numpy(), name + ".y": y.detach().numpy(),
                    name + ".dx": x.grad.numpy()})
        specs["blocks." + name] = dict(ctor=dict(in_dim=24, out_dim=16, **kw), spec=spec)

    m = M.MLP(in_dim=32, out_dim=48, dim=40, n_blk=3, norm="none", activ="relu")
    spec, sd = load_into(m)
    x = rnd(4, 2, 16)
    y = m(x)
    check("mlp.y", y, O.mlp(x, sd, "", 3), 1e-5)
    blk.update({"mlp.x": x.numpy(), "mlp.y": y.detach().numpy()})
    specs["blocks.mlp"] = dict(ctor=dict(in_dim=32, out_dim=48, dim=40, n_blk=3, norm="none", activ="relu"), spec=spec)

    # AdaptiveInstanceNorm2d with and without the iAFF style path; get_key; iAFF train + eval
    for name, with_input, train in (("adain_plain", False, True), ("adain_iaff_train", True, True),
                                    ("adain_iaff_eval", True, False)):
        m = B.AdaptiveInstanceNorm2d(512)
        spec, sd = load_into(m)
        m.train(train)
        x = rnd(3, 512, 4, 6).requires_grad_()
        wgt, bia = rnd(3 * 512).requires_grad_(), rnd(3 * 512).requires_grad_()
        style = rnd(3, 512, 2, 3).requires_grad_() if with_input else None
        m.weight, m.bias, m.input = wgt, bia, style
        y = m(x); gy = rnd(*y.shape); y.backward(gy)
        post = {k: v.clone() for k, v in m.state_dict().items()}
        xo = x.detach().clone().requires_grad_()
        wo, bo = wgt.detach().clone().requires_grad_(), bia.detach().clone().requires_grad_()
        so = style.detach().clone().requires_grad_() if with_input else None
        stats = {}
        yo = O.adaptive_instance_norm(xo, sd, "", wo, bo, so, train, stats); yo.backward(gy)
        check(name + ".y", y, yo, 5e-5); check(name + ".dx", x.grad, xo.grad, 2e-4)
        check(name + ".dweight", wgt.grad, wo.grad, 2e-4)
        if with_input:
            check(name + ".dstyle", style.grad, so.grad, 2e-4)
            blk[name + ".style"], blk[name + ".dstyle"] = style.detach().numpy(), style.grad.numpy()
        for k, v in stats.items():
            check(name + ".stat." + k, post[k].float(), v.float(), 1e-5)
        if with_input and train:
            assert int(post["iAff.global_att.2.num_batches_tracked"]) == 2   # updated twice (blocks.py:295)
            assert int(post["iAff.global_att2.2.num_batches_tracked"]) == 0  # never used
            assert torch.equal(post["running_mean"], sd["running_mean"])     # .repeat(b) copies are updated
            for k in ("iAff.global_att.2.running_mean", "iAff.global_att.5.running_var",
                      "iAff.local_att.1.running_var", "iAff.local_att2.4.running_mean"):
                blk[name + ".post." + k] = post[k].numpy()
        blk.update({name + ".x": x.detach().numpy(), name + ".gy": gy.numpy(), name + ".weight": wgt.detach().numpy(),
                    name + ".bias": bia.detach().numpy(), name + ".y": y.detach().numpy(), name + ".dx": x.grad.numpy(),
                    name + ".dweight": wgt.grad.numpy(), name + ".dbias": bia.grad.numpy()})
        specs["blocks." + name] = dict(ctor=dict(num_features=512), spec=spec)
    x, s = rnd(2, 8, 8, 27), rnd(2, 8, 2, 7)
    k = B.get_key(x, s)
    check("get_key", k, O.get_key(x, s), 1e-6)
    blk.update({"get_key.style": s.numpy(), "get_key.y": k.numpy()})
    np.savez_compressed(os.path.join(OUT, "blocks.npz"), **blk)

    # ---------------------------------------------------------------- generator forward / dis / cla / grads
    print("generator (C_s=15, B=4; B=2 makes the train-mode BatchNorms ill-conditioned)")
    torch.manual_seed(0)
    gen = ns.Gen()
    dis = M.DisModel()
    cla = M.WriterClaModel(O.NUM_WRITERS)
    gspec, gsd = load_into(gen)
    dspec, dsd = load_into(dis)
    cspec, csd = load_into(cla)
    specs["gen_c15"], specs["dis"], specs["cla"] = gspec, dspec, cspec
    batch = O.synthetic_batch(4, 15)
    gen.train(); dis.train(); cla.train()
    res = gen.enc_image(batch["tr_img"])
    f_xt, f_embed = gen.enc_text(batch["label_xt"], res[-1].shape)
    f_mix = gen.mix(res, f_embed)
    xg = gen.decode(f_mix, res, f_embed, f_xt)
    post = {k: v.clone() for k, v in gen.state_dict().items()}
    stats = {}
    res_o = O.image_encoder(batch["tr_img"], gsd)
    for i in range(6):
        check(f"enc.result{i}", res[i], res_o[i], 2e-4)
    xg_o = O.gen_forward(None, batch["label_xt"], gsd, stats=stats, results=res_o)
    check("gen.xg", xg, xg_o, 1e-4)
    for k, v in stats.items():
        check("gen.stat." + k, post[k].float(), v.float(), 1e-4)
    gold = {"xg": xg.detach().numpy(), "f_xt": f_xt.detach().numpy(), "f_mix_mean_abs": f_mix.abs().mean().item()}
    for i in range(6):
        gold[f"result{i}.mean"] = res[i].mean().item()
        gold[f"result{i}.abs_mean"] = res[i].abs().mean().item()
        gold[f"result{i}.shape"] = np.array(res[i].shape)
    bn_keys = [k for k in stats if k.endswith("running_var")][:6] + [k for k in stats if "num_batches" in k][:4]
    for k in bn_keys:
        gold["post." + k] = post[k].numpy()
    # eval-mode forward (tt scripts: BN running stats, AdaIN/IN still instance stats), batch 1
    gen.load_state_dict(gsd); gen.eval()
    b1 = {k: (v[:1] if torch.is_tensor(v) else v) for k, v in batch.items()}
    xg_e = rb.gen_forward(gen, b1["tr_img"], b1["label_xt"])
    check("gen.xg_eval_b1", xg_e, O.gen_forward(b1["tr_img"], b1["label_xt"], gsd, training=False), 1e-4)
    gold["xg_eval_b1"] = xg_e.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "gen_fwd_c15_b4.npz"), **gold)

    print("dis / cla")
    dc = {}
    r = dis(xg.detach())
    check("dis.out", r, O.dis_forward(xg.detach(), dsd), 1e-4)
    dc["dis.out"] = r.detach().numpy()
    dc["dis.real_loss"] = dis.calc_dis_real_loss(batch["img_xt"]).item()
    dc["dis.fake_loss"] = dis.calc_dis_fake_loss(xg.detach()).item()
    dc["dis.gen_loss"] = dis.calc_gen_loss(xg.detach()).item()
    dc["cla.loss"] = cla(batch["img_xt"], batch["tr_wid"]).item()
    assert abs(dc["dis.real_loss"] - O.dis_loss(batch["img_xt"], dsd, target=1.0).item()) < 1e-5
    assert abs(dc["cla.loss"] - O.cla_loss(batch["img_xt"], batch["tr_wid"], csd).item()) < 1e-4

    print("gradients: gen_update without recogniser (network_tro.py:57-103), dis_update, cla_update")
    gen.load_state_dict(gsd); gen.train()
    for mod in (gen, dis, cla):
        mod.zero_grad()
    res = gen.enc_image(batch["tr_img"])
    outs = []
    for lab in (batch["label_xt"], batch["label_xt_swap"]):
        f_xt, f_embed = gen.enc_text(lab, res[-1].shape)
        outs.append(gen.decode(gen.mix(res, f_embed), res, f_embed, f_xt))
    l_dis = (dis.calc_gen_loss(outs[0]) + dis.calc_gen_loss(outs[1])) / 2
    l_cla = (cla(outs[0], batch["tr_wid"]) + cla(outs[1], batch["tr_wid"])) / 2
    (l_dis + l_cla).backward()
    full = {"gen." + k: v.clone().requires_grad_(v.is_floating_point()) for k, v in gsd.items()}
    full.update({"dis." + k: v.clone().requires_grad_(v.is_floating_point()) for k, v in dsd.items()})
    full.update({"cla." + k: v.clone().requires_grad_(v.is_floating_point()) for k, v in csd.items()})
    lt, ld, lc, xg_o, xgs_o = O.gen_update(batch, full)
    lt.backward()
    check("gen_update.l_dis", l_dis, ld, 1e-5); check("gen_update.l_cla", l_cla, lc, 1e-4)
    worst, dead, noise = 1.0, 0, []
    for k, p in gen.named_parameters():
        go = full["gen." + k].grad
        if p.grad is None:
            dead += 1
            assert go is None or float(go.abs().max()) == 0.0, k
            continue
        cs = cosine(p.grad, go)
        if cs < 0.999:
            # two fp32 CPU evaluations of the same graph disagree: the exact gradient is structurally zero (a bias
            # in front of an instance/batch norm) and what is left is rounding noise.  Recorded, excluded from cosine.
            assert k.endswith("bias"), (k, cs)
            noise.append(k)
            continue
        worst = min(worst, cs)
    print(f"  gen_update grad cosine (min over {len(list(gen.parameters())) - dead} live tensors): {worst:.6f}; "
          f"{dead} tensors never receive a gradient")
    assert worst > 0.9999
    keys, norms, heads = grad_digest([(k, p.grad) for k, p in gen.named_parameters()])
    print(f"  {len(noise)} bias tensors carry only rounding noise (structurally zero gradient)")
    gd = {"gen.keys": np.array(keys), "gen.norms": norms, "gen.heads": heads, "gen.noise_keys": np.array(noise),
          "gen.l_dis": l_dis.item(), "gen.l_cla": l_cla.item(), "xg": outs[0].detach().numpy(),
          "xg_swap": outs[1].detach().numpy()}
    # dis_update / cla_update
    for mod in (dis, cla):
        mod.zero_grad()
    im1 = batch["tr_img"][:, 0:1].clone().requires_grad_()
    l_real = (dis.calc_dis_real_loss(im1) + dis.calc_dis_real_loss(batch["tr_img"][:, 1:2])) / 2
    l_real.backward(retain_graph=True)
    l_fake = (dis.calc_dis_fake_loss(outs[0].detach()) + dis.calc_dis_fake_loss(outs[1].detach())) / 2
    l_fake.backward()
    keys, norms, heads = grad_digest([(k, p.grad) for k, p in dis.named_parameters()])
    gd.update({"dis.keys": np.array(keys), "dis.norms": norms, "dis.heads": heads, "dis.l_real": l_real.item(),
               "dis.l_fake": l_fake.item(), "dis.dimg_norm": float(im1.grad.norm())})
    lr_o, lf_o = O.dis_update(batch, full)
    check("dis_update.l_real", l_real, lr_o, 1e-5); check("dis_update.l_fake", l_fake, lf_o, 1e-4)
    l_c = cla(batch["tr_img"][:, 0:1], batch["tr_wid"]); l_c.backward()
    keys, norms, heads = grad_digest([(k, p.grad) for k, p in cla.named_parameters()])
    gd.update({"cla.keys": np.array(keys), "cla.norms": norms, "cla.heads": heads, "cla.loss": l_c.item()})
    np.savez_compressed(os.path.join(OUT, "grads_c15_b4.npz"), **gd)
    np.savez_compressed(os.path.join(OUT, "dis_cla_b4.npz"), **dc)

    # spec for the 50-style-image generator: only the first conv differs
    g50 = dict(gspec)
    g50["enc_image.model.features.0.weight"] = [64, 50, 3, 3]
    specs["gen_c50"] = g50
    json.dump(specs, open(os.path.join(OUT, "state_spec.json"), "w"))
    json.dump([dict(name=n, max_abs=e, tol=t) for n, e, t in report],
              open(os.path.join(OUT, "oracle_vs_reference.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
