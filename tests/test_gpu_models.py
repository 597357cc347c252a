"""Whole-network parity on the GPU through the drop-in classes.

fp32 mode   image <= 1e-4 max-abs against the image the REFERENCE produced (tests/golden/gen_fwd_c15_b4.npz);
            per-parameter gradient cosine >= 0.999 against autograd over the CPU oracle, norms against the reference's.
bf16 mode   (tcgen05 convolutions on split-bf16 operands, three MMAs per product) is held to BASELINE.json's bf16 bars
            against the same fp32 references: image <= 2e-2 max-abs, gradient cosine >= 0.999.
bf16x1 mode (one MMA per product) is checked against the oracle evaluated under the bf16 storage model
            (oracle.storage_model("bf16"): the same algorithm with convolution operands rounded to bfloat16).  Its
            distance to the fp32 reference is reported and bounded by what that model itself costs (measured on the CPU:
            a 1e-6 input perturbation moves the image by ~1e-4 on this random-init network, so a single 2^-9 operand
            rounding cannot stay within 2e-2 of fp32 for ANY implementation; see DESIGN.md "precision").
"""
import numpy as np
import pytest
import torch

import affganwriting_b200 as A
from affganwriting_b200 import modules_tro as M
from affgw_testutil import cosine, rel_err
from oracle import affgw_oracle as O
from oracle import weights as W

pytestmark = pytest.mark.gpu


def _encoder(num_channel):
    from affganwriting_b200 import load_data
    old = load_data.NUM_CHANNEL
    load_data.NUM_CHANNEL = num_channel
    try:
        return M.ImageEncoder()
    finally:
        load_data.NUM_CHANNEL = old


def _gen(specs, key="gen_c15"):
    g = M.GenModel_FC(12, encoder=_encoder(specs[key]["enc_image.model.features.0.weight"][1]))
    g.load_state_dict(W.make_state(specs[key]))
    return g.cuda()


def _dis_cla(specs):
    dis, cla = M.DisModel(), M.WriterClaModel(O.NUM_WRITERS)
    dis.load_state_dict(W.make_state(specs["dis"]))
    cla.load_state_dict(W.make_state(specs["cla"]))
    return dis.cuda().train(), cla.cuda().train()


def _cuda(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


def _maxabs(a, b):
    return float((a.detach().float().cpu() - b.detach().float().cpu()).abs().max())


def test_generator_forward_matches_reference(mode, specs, golden):
    gold = golden("gen_fwd_c15_b4.npz")
    sd = W.make_state(specs["gen_c15"])
    gen = _gen(specs).train()
    cpu = O.synthetic_batch(4, 15)
    batch = _cuda(cpu)
    res = gen.enc_image(batch["tr_img"])
    for i in range(6):
        assert [res[i].shape[j] for j in (0, 2, 3)] == [int(v) for v in gold[f"result{i}.shape"][[0, 2, 3]]]
    f_xt, f_embed = gen.enc_text(batch["label_xt"], res[-1].shape)
    xg = gen.decode(gen.mix(res, f_embed), res, f_embed, f_xt)
    assert xg.shape == (4, 1, 64, 216) and xg.dtype == torch.float32
    err_ref = _maxabs(xg, torch.from_numpy(gold["xg"]))
    if mode == "fp32":
        print(f"\n[fp32] generated image max-abs error vs reference: {err_ref:.3e}")
        assert err_ref <= 1e-4
        for i in range(6):
            assert abs(float(res[i].float().abs().mean()) - float(gold[f"result{i}.abs_mean"])) <= 1e-4
        assert rel_err(f_xt, torch.from_numpy(gold["f_xt"])) <= 1e-4
        stat_tol = 1e-4
    elif mode == "bf16":
        rms = float((xg.detach().cpu() - torch.from_numpy(gold["xg"])).pow(2).mean().sqrt())
        print(f"\n[bf16] generated image error vs reference: max-abs {err_ref:.3e}, rms {rms:.3e}")
        assert err_ref <= 2e-2
        assert rel_err(f_xt, torch.from_numpy(gold["f_xt"])) <= 2e-3
        stat_tol = 2e-3
    else:
        with torch.no_grad(), O.storage_model("bf16"):
            model = O.gen_forward(cpu["tr_img"], cpu["label_xt"], sd)
        ref = torch.from_numpy(gold["xg"])
        cost, cost_rms = _maxabs(model, ref), float((model - ref).pow(2).mean().sqrt())
        rms = float((xg.detach().cpu() - ref).pow(2).mean().sqrt())
        print(f"\n[bf16x1] image error vs fp32 reference: max-abs {err_ref:.3e}, rms {rms:.3e}; the bf16 storage model itself "
              f"(CPU oracle, same rounding points): max-abs {cost:.3e}, rms {cost_rms:.3e}; ours vs that model: {_maxabs(xg, model):.3e}")
        # bf16 mode must not be worse than what bf16 storage inherently costs on this network (see module docstring)
        assert rms <= 1.5 * cost_rms + 1e-3
        assert err_ref <= 1.5 * cost + 2e-2
        stat_tol = 5e-2
    sd_now = gen.state_dict()
    for k in gold.files:
        if k.startswith("post."):
            ref = torch.from_numpy(gold[k])
            if ref.dtype == torch.int64:
                assert int(sd_now[k[5:]]) == int(ref), k
            else:
                assert rel_err(sd_now[k[5:]], ref) <= stat_tol, k
    A.check_device_errors()


def test_generator_eval_mode_batch_one(mode, specs, golden):
    """tt.* generation scripts: model.eval(), batch 1 (BatchNorm running stats, instance stats elsewhere)."""
    gold = golden("gen_fwd_c15_b4.npz")
    gen = _gen(specs).eval()
    cpu = O.synthetic_batch(4, 15)
    with torch.no_grad():
        xg = gen(cpu["tr_img"][:1].cuda(), cpu["label_xt"][:1].cuda())
    err_ref = _maxabs(xg, torch.from_numpy(gold["xg_eval_b1"]))
    if mode == "fp32":
        print(f"\n[fp32] eval-mode batch-1 image max-abs error vs reference: {err_ref:.3e}")
        assert err_ref <= 1e-4
    elif mode == "bf16":
        print(f"\n[bf16] eval-mode batch-1 image max-abs error vs reference: {err_ref:.3e}")
        assert err_ref <= 2e-2
    else:
        with torch.no_grad(), O.storage_model("bf16"):
            model = O.gen_forward(cpu["tr_img"][:1], cpu["label_xt"][:1], W.make_state(specs["gen_c15"]), training=False)
        ref = torch.from_numpy(gold["xg_eval_b1"])
        cost = _maxabs(model, ref)
        print(f"\n[bf16x1] eval-mode batch-1 image max-abs vs fp32 reference {err_ref:.3e} (bf16 storage model: {cost:.3e})")
        assert err_ref <= 1.5 * cost + 2e-2


def test_dis_cla_forward_and_losses(mode, specs, golden):
    gold, gg = golden("dis_cla_b4.npz"), golden("gen_fwd_c15_b4.npz")
    dis, cla = _dis_cla(specs)
    cpu = O.synthetic_batch(4, 15)
    batch = _cuda(cpu)
    xg = torch.from_numpy(gg["xg"]).cuda()
    out = dis(xg)
    vals = dict(real=float(dis.calc_dis_real_loss(batch["img_xt"])), fake=float(dis.calc_dis_fake_loss(xg)),
                gen=float(dis.calc_gen_loss(xg)), cla=float(cla(batch["img_xt"], batch["tr_wid"])))
    if mode in ("fp32", "bf16"):
        tol = 1e-4 if mode == "fp32" else 2e-3
        print(f"\n[{mode}] dis logits vs reference: {rel_err(out, torch.from_numpy(gold['dis.out'])):.3e}")
        assert rel_err(out, torch.from_numpy(gold["dis.out"])) <= tol
        for k, ref in (("real", "dis.real_loss"), ("fake", "dis.fake_loss"), ("gen", "dis.gen_loss"), ("cla", "cla.loss")):
            assert abs(vals[k] - float(gold[ref])) <= tol * max(1.0, abs(float(gold[ref]))), k
    else:
        dsd, csd = W.make_state(specs["dis"]), W.make_state(specs["cla"])
        with torch.no_grad(), O.storage_model("bf16"):
            m_out = O.dis_forward(torch.from_numpy(gg["xg"]), dsd)
            m_cla = float(O.cla_loss(cpu["img_xt"], cpu["tr_wid"], csd))
        print(f"\n[bf16x1] dis logits: vs bf16-storage oracle {rel_err(out, m_out):.3e}, vs fp32 reference "
              f"{rel_err(out, torch.from_numpy(gold['dis.out'])):.3e}")
        assert rel_err(out, m_out) <= 2e-2
        assert abs(vals["cla"] - m_cla) <= 2e-2 * max(1.0, abs(m_cla))
        assert abs(vals["real"] - float(gold["dis.real_loss"])) <= 5e-2 * max(1.0, abs(float(gold["dis.real_loss"])))
    A.check_device_errors()


def test_out_of_range_writer_id_is_flagged(specs):
    A.set_precision("fp32")
    _, cla = _dis_cla(specs)
    batch = _cuda(O.synthetic_batch(2, 15))
    cla(batch["img_xt"], torch.tensor([3, 500], device="cuda"))
    with pytest.raises(RuntimeError):
        A.check_device_errors()


def _oracle_gen_update(specs, batch, kind):
    full = {}
    for pre, key in (("gen.", "gen_c15"), ("dis.", "dis"), ("cla.", "cla")):
        for k, v in W.make_state(specs[key]).items():
            full[pre + k] = v.clone().requires_grad_(v.is_floating_point())
    with O.storage_model(kind):
        lt, ld, lc, xg, xgs = O.gen_update(batch, full)
        lt.backward()
    return full, float(ld), float(lc)


def test_gen_update_gradients(mode, specs, golden):
    """network_tro.py:57-103 without the recogniser term: l_total = l_dis + l_cla, backward into the generator."""
    gd = golden("grads_c15_b4.npz")
    cpu_batch = O.synthetic_batch(4, 15)
    full, ld_o, lc_o = _oracle_gen_update(specs, cpu_batch, "bf16" if mode == "bf16x1" else "fp32")
    gen = _gen(specs).train()
    dis, cla = _dis_cla(specs)
    batch = _cuda(cpu_batch)
    res = gen.enc_image(batch["tr_img"])
    outs = []
    for lab in (batch["label_xt"], batch["label_xt_swap"]):
        f_xt, f_embed = gen.enc_text(lab, res[-1].shape)
        outs.append(gen.decode(gen.mix(res, f_embed), res, f_embed, f_xt))
    l_dis = (dis.calc_gen_loss(outs[0]) + dis.calc_gen_loss(outs[1])) / 2
    l_cla = (cla(outs[0], batch["tr_wid"]) + cla(outs[1], batch["tr_wid"])) / 2
    (l_dis + l_cla).backward()
    ltol = {"fp32": 1e-4, "bf16": 5e-3, "bf16x1": 5e-2}[mode]
    assert abs(float(l_dis) - ld_o) <= ltol * max(1.0, abs(ld_o)) and abs(float(l_cla) - lc_o) <= ltol * max(1.0, abs(lc_o))
    if mode == "fp32":
        assert abs(float(l_dis) - float(gd["gen.l_dis"])) <= 1e-4 and abs(float(l_cla) - float(gd["gen.l_cla"])) <= 1e-3
    noise = set(gd["gen.noise_keys"].tolist())
    ref_norm = dict(zip(gd["gen.keys"].tolist(), gd["gen.norms"].tolist()))
    rows, dead = [], 0
    dots = na = nb = 0.0
    for k, p in gen.named_parameters():
        go = full["gen." + k].grad
        if ref_norm[k] < 0:                       # never receives a gradient in the reference
            assert p.grad is None, k
            dead += 1
            continue
        assert p.grad is not None, k
        if k in noise:                            # bias in front of a norm: exact gradient is zero, only rounding noise
            continue
        a, b = p.grad.double().cpu().reshape(-1), go.double().reshape(-1)
        dots += float(a @ b); na += float(a @ a); nb += float(b @ b)
        rows.append((cosine(p.grad, go), float(p.grad.norm()) / max(ref_norm[k], 1e-30), k))
    glob = dots / (na ** 0.5 * nb ** 0.5)
    rows.sort()
    print(f"\n[{mode}] gen_update gradients: global cosine {glob:.6f}; {dead} tensors without gradient; worst tensors:")
    for c, r, k in rows[:4]:
        print(f"      cos {c:.6f}  |g|/|g_ref| {r:.4f}  {k}")
    assert dead == 96
    if mode == "fp32":
        assert glob >= 0.999 and rows[0][0] >= 0.999
        assert all(abs(r - 1.0) <= 2e-3 for _, r, _ in rows)
    elif mode == "bf16":
        assert glob >= 0.999 and rows[0][0] >= 0.995
        assert all(abs(r - 1.0) <= 3e-2 for _, r, _ in rows), sorted(rows, key=lambda t: -abs(t[1] - 1))[:3]
    else:
        # two independent single-pass bf16 realisations of this network agree to ~0.92-0.96 (the storage model against itself with a
        # different summation order behaves the same); kernel-level gradient accuracy is pinned block by block in
        # tests/test_gpu_conv_tc.py and tests/test_gpu_blocks.py at cosine >= 0.999 / 0.99
        assert glob >= 0.85 and rows[0][0] >= 0.7
        assert sum(abs(r - 1.0) > 0.1 for _, r, _ in rows) <= len(rows) // 10, sorted(rows, key=lambda t: -abs(t[1] - 1))[:3]


def test_dis_and_cla_update_gradients(mode, specs, golden):
    gd = golden("grads_c15_b4.npz")
    dis, cla = _dis_cla(specs)
    cpu = O.synthetic_batch(4, 15)
    batch = _cuda(cpu)
    xg, xgs = torch.from_numpy(gd["xg"]).cuda(), torch.from_numpy(gd["xg_swap"]).cuda()
    im1 = batch["tr_img"][:, 0:1].clone().requires_grad_()
    l_real = (dis.calc_dis_real_loss(im1) + dis.calc_dis_real_loss(batch["tr_img"][:, 1:2])) / 2
    l_real.backward(retain_graph=True)                  # network_tro.py:113
    l_fake = (dis.calc_dis_fake_loss(xg) + dis.calc_dis_fake_loss(xgs)) / 2
    l_fake.backward()
    l_c = cla(batch["tr_img"][:, 0:1], batch["tr_wid"])
    l_c.backward()
    ltol = {"fp32": 1e-4, "bf16": 2e-3, "bf16x1": 5e-2}[mode]
    for val, key in ((l_real, "dis.l_real"), (l_fake, "dis.l_fake"), (l_c, "cla.loss")):
        assert abs(float(val) - float(gd[key])) <= ltol * max(1.0, abs(float(gd[key]))), key
    lim = {"fp32": 2e-3, "bf16": 1e-2, "bf16x1": 0.1}[mode]
    assert abs(float(im1.grad.norm()) - float(gd["dis.dimg_norm"])) <= lim * float(gd["dis.dimg_norm"])
    worst = 0.0
    for net, name in ((dis, "dis"), (cla, "cla")):
        ref_norm = dict(zip(gd[name + ".keys"].tolist(), gd[name + ".norms"].tolist()))
        heads = dict(zip(gd[name + ".keys"].tolist(), gd[name + ".heads"]))
        for k, p in net.named_parameters():
            assert p.grad is not None, k
            r = abs(float(p.grad.norm()) - ref_norm[k]) / max(ref_norm[k], 1e-12)
            worst = max(worst, r)
            assert r <= lim, (name, k, r)
            if mode == "fp32":
                h = p.grad.reshape(-1)[:8].float().cpu().numpy()
                ref_h = heads[k][:h.size]
                assert np.abs(h - ref_h).max() <= lim * max(float(np.abs(ref_h).max()), ref_norm[k] / max(1.0, p.numel() ** 0.5)) + 1e-7, (name, k)
    print(f"\n[{mode}] dis/cla update: worst relative gradient-norm deviation vs reference {worst:.3e}")


# (full-size properties at 50 style planes / batch 64 live in tests/test_gpu_parity_c50.py)


def test_cuda_graph_replay_matches_eager_iterations(specs):
    """Trainer(cuda_graph=True) - three captured sub-steps, eager gradient exchange + Adam in between - must walk the
    same trajectory as the eager trainer (main_run.py:146-167 order).  The trajectory itself is chaotic with respect to the
    order of the fp32 atomics in the statistics / weight-gradient reductions (two EAGER trainers drift apart in the 4th
    digit of the generator loss within three iterations), so the bound is the eager-vs-eager drift measured alongside."""
    from affganwriting_b200.trainer import Trainer
    import bench
    from affganwriting_b200 import load_data as LD
    A.set_precision("f16")
    try:
        dev = torch.device("cuda", 0)
        batch = LD.batch_to_device(bench.synthetic_batch(4, 50, 7), torch.device('cuda'))
        torch.manual_seed(0)
        a = Trainer(num_writers=500, device=dev)
        b = Trainer(num_writers=500, device=dev, wgrad_stream=False)       # single-stream weight gradients
        g = Trainer(num_writers=500, device=dev, cuda_graph=True)
        g.GRAPH_WARMUP = 1                     # capture at iteration 1, pure replays from iteration 2 on
        # same, with the generator's / classifier's exchange + Adam on a side stream under the next graph replay
        o = Trainer(num_writers=500, device=dev, cuda_graph=True, overlap_exchange=True)
        o.GRAPH_WARMUP = 1
        b.model.load_state_dict(a.model.state_dict())
        g.model.load_state_dict(a.model.state_dict())
        o.model.load_state_dict(a.model.state_dict())
        drift_ee = drift_eg = drift_eo = 0.0
        first = None
        for it in range(7):
            la, lb, lg, lo = a.train_step(batch), b.train_step(batch), g.train_step(batch), o.train_step(batch)
            first = first or {k: float(v) for k, v in la.items()}
            for k in la:
                drift_ee = max(drift_ee, abs(float(la[k]) - float(lb[k])))
                drift_eg = max(drift_eg, abs(float(la[k]) - float(lg[k])))
                drift_eo = max(drift_eo, abs(float(la[k]) - float(lo[k])))
            if it < 3:          # eager, capture pass, first pure replay: before the divergence has had time to grow
                for lx in (lg, lo):
                    # (1e-3: the run-to-run noise of one forward in this mode is 8e-4 on the image, tests/test_gpu_parity_c50.py)
                    assert all(abs(float(la[k]) - float(lx[k])) <= 1e-3 * max(1.0, abs(float(la[k]))) for k in la), \
                        (it, {k: (float(la[k]), float(lx[k])) for k in la})
        assert g.graph_launches > 1000 and g._graphs is not None and g._eager_steps == 1
        # the optimiser steps must reach the kernels (packed-weight cache invalidation, ops.weights_updated): the writer
        # classifier's loss falls by ~0.007 per iteration at lr 1e-5 on a fixed batch
        assert float(la["cla"]) < first["cla"] - 0.02 and float(lg["cla"]) < first["cla"] - 0.02

        def weight_drift(x, y):
            sx, sy = x.model.state_dict(), y.model.state_dict()
            for k, v in sx.items():
                if not v.is_floating_point():
                    assert torch.equal(v, sy[k]), k                      # num_batches_tracked
            return max(float((v - sy[k]).abs().max()) for k, v in sx.items() if v.is_floating_point())
        assert o._pending and o.overlap_exchange      # the generator's step of the last iteration is still on the side stream
        o.join()
        assert not o._pending
        w_ee, w_eg, w_eo = weight_drift(a, b), weight_drift(a, g), weight_drift(a, o)
        print(f"\nafter 7 iterations: loss drift eager/eager {drift_ee:.2e}, eager/graph {drift_eg:.2e}, eager/overlapped "
              f"{drift_eo:.2e}; weight+buffer drift eager/eager {w_ee:.2e}, eager/graph {w_eg:.2e}, eager/overlapped {w_eo:.2e}")
        assert drift_eg <= 10 * drift_ee + 2e-2 and drift_eo <= 10 * drift_ee + 2e-2
        assert w_eg <= 10 * w_ee + 0.3 and w_eo <= 10 * w_ee + 0.3
        assert a.model.iter_num == g.model.iter_num == o.model.iter_num
        assert float(lo["cla"]) < first["cla"] - 0.02
    finally:
        A.set_precision("fp32")


def test_short_final_batch_after_capture_runs_eagerly(specs):
    """The reference's loaders keep the last, shorter batch of an epoch (main_run.py:123-130): after the graphs have been
    captured at one batch size, a batch of another size must take the eager path (not raise in copy_, not broadcast a single
    sample over the captured batch), and the next full batch must replay again."""
    from affganwriting_b200.trainer import Trainer
    import bench
    from affganwriting_b200 import load_data as LD
    A.set_precision("f16")
    try:
        dev = torch.device("cuda", 0)
        full = LD.batch_to_device(bench.synthetic_batch(4, 50, 7), dev)
        short = LD.batch_to_device(bench.synthetic_batch(2, 50, 8), dev)
        one = LD.batch_to_device(bench.synthetic_batch(1, 50, 9), dev)
        torch.manual_seed(0)
        t = Trainer(num_writers=500, device=dev, cuda_graph=True, overlap_exchange=True)
        e = Trainer(num_writers=500, device=dev)
        e.model.load_state_dict(t.model.state_dict())
        t.GRAPH_WARMUP = 1
        for _ in range(3):
            t.train_step(full); e.train_step(full)
        assert t._graphs is not None
        n0 = t.model.iter_num
        ls, le = t.train_step(short), e.train_step(short)           # eager fallback inside the graphed trainer
        assert t.model.iter_num == n0 + 1
        for k in ls:
            assert torch.isfinite(ls[k]) and abs(float(ls[k]) - float(le[k])) <= 2e-2 * max(1.0, abs(float(le[k]))), k
        with pytest.raises(ValueError):                              # train-mode BatchNorm refuses one sample, like the reference
            t.train_step(one)
        lf = t.train_step(full)                                      # replays again
        assert all(torch.isfinite(v) for v in lf.values())
        kept = lf["gen"].clone()
        t.train_step(full)
        assert float(kept) == float(lf["gen"])                      # returned losses are copies, not views of graph memory
        sd = t.state_dict()                                          # joins the side stream first
        assert not t._pending and all(torch.isfinite(v).all() for v in sd.values() if v.is_floating_point())
    finally:
        A.set_precision("fp32")


def test_accumulate_into_grad_matches_autograd_accumulation():
    """ops.accumulate_into_grad: a second use of a parameter adds into its existing .grad inside the weight-gradient kernel
    instead of through autograd's ATen add - same gradients (dis_update's two backward calls, gen_update's two decodes)."""
    from affganwriting_b200 import ops
    A.set_precision("f16")
    try:
        g = torch.Generator(device="cuda").manual_seed(9)
        x1 = ops.to_internal(torch.randn(2, 64, 8, 27, device="cuda", generator=g))
        x2 = ops.to_internal(torch.randn(2, 64, 8, 27, device="cuda", generator=g))
        out = {}
        for flag in (False, True):
            w = torch.nn.Parameter(torch.randn(96, 64, 3, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 0.05)
            b = torch.nn.Parameter(torch.zeros(96, device="cuda"))
            with ops.accumulate_into_grad(flag):
                ops.conv2d(x1, w, b, pad=1, pad_mode="reflect").square().mean().backward()
                first = w.grad.data_ptr()
                (ops.conv2d(x2, w, b, pad=1, pad_mode="reflect").square().mean() * 3).backward()     # second use: accumulates
                ops.linear(x1.reshape(-1, 64)[:, :64], w[:, :, 0, 0].detach().clone().requires_grad_(), None).sum().backward()
            out[flag] = (w.grad.clone(), b.grad.clone(), w.grad.data_ptr() == first)
        assert out[True][2]                                       # accumulated in place
        assert float((out[True][0] - out[False][0]).abs().max()) <= 1e-6 * float(out[False][0].abs().max())
        assert float((out[True][1] - out[False][1]).abs().max()) <= 1e-5 * float(out[False][1].abs().max())
    finally:
        A.set_precision("fp32")


def test_wgrad_side_stream_matches_main_stream():
    """ops.wgrad_side_stream: weight-gradient GEMMs forked onto a second stream (first use creates .grad, second use accumulates
    into it inside the kernel) give the gradients of the single-stream path once the context has exited."""
    from affganwriting_b200 import ops
    A.set_precision("f16")
    try:
        g = torch.Generator(device="cuda").manual_seed(9)
        x1 = ops.to_internal(torch.randn(4, 64, 16, 54, device="cuda", generator=g))
        x2 = ops.to_internal(torch.randn(4, 64, 16, 54, device="cuda", generator=g))
        out = {}
        for flag in (False, True):
            gw = torch.Generator(device="cuda").manual_seed(1)
            w = torch.nn.Parameter(torch.randn(96, 64, 3, 3, device="cuda", generator=gw) * 0.05)
            w2 = torch.nn.Parameter(torch.randn(64, 96, 3, 3, device="cuda", generator=gw) * 0.05)
            wl = torch.nn.Parameter(torch.randn(32, 64, device="cuda", generator=gw) * 0.05)
            b = torch.nn.Parameter(torch.zeros(96, device="cuda"))
            xin = x1.clone().requires_grad_()
            with ops.wgrad_side_stream(flag):
                for rep in range(3):                                # enough work that the side stream runs behind
                    h = ops.conv2d(xin, w, b, pad=1, pad_mode="reflect")
                    h = ops.conv2d(ops.instance_norm(h, act="relu"), w2, None, pad=1)
                    (h.square().mean() * (rep + 1)).backward()
                (ops.conv2d(x2, w, b, pad=1, pad_mode="reflect").square().mean() * 3).backward()
                ops.linear(x1.permute(0, 2, 3, 1).reshape(-1, 64), wl, None).square().mean().backward()
            out[flag] = [t.grad.clone() for t in (w, w2, wl, b, xin)]
        torch.cuda.synchronize()
        for a, r in zip(out[True], out[False]):
            # (order of the fp32 atomics / of the accumulation into .grad: ~6e-5 of the largest element, measured)
            assert float((a - r).abs().max()) <= 5e-4 * float(r.abs().max())
    finally:
        A.set_precision("fp32")


def test_trainer_wgrad_stream_matches_single_stream(specs):
    """Trainer(wgrad_stream=True) (default: weight-gradient GEMMs on a second stream) against wgrad_stream=False: the same
    first iteration (losses, and the weights after the three Adam steps) up to the order of the fp32 atomics."""
    from affganwriting_b200.trainer import Trainer
    import bench
    from affganwriting_b200 import load_data as LD
    A.set_precision("f16")
    try:
        dev = torch.device("cuda", 0)
        batch = LD.batch_to_device(bench.synthetic_batch(4, 50, 7), torch.device('cuda'))
        torch.manual_seed(0)
        a = Trainer(num_writers=500, device=dev, wgrad_stream=False)
        b = Trainer(num_writers=500, device=dev, wgrad_stream=False)
        s = Trainer(num_writers=500, device=dev, wgrad_stream=True)
        assert s.wgrad_stream and not a.wgrad_stream
        b.model.load_state_dict(a.model.state_dict())
        s.model.load_state_dict(a.model.state_dict())
        la, lb, ls = a.train_step(batch), b.train_step(batch), s.train_step(batch)
        torch.cuda.synchronize()
        for k in la:
            assert abs(float(la[k]) - float(ls[k])) <= 1e-3 * max(1.0, abs(float(la[k]))), (k, float(la[k]), float(ls[k]))

        def rel(x, y):
            sx, sy = x.model.state_dict(), y.model.state_dict()
            return max(float((v - sy[k]).norm()) / max(float(v.norm()), 1e-20) for k, v in sx.items()
                       if v.is_floating_point() and v.numel() > 1)
        noise, diff = rel(a, b), rel(a, s)
        print(f"\nweights after one iteration: single/single {noise:.2e}, single/side-stream {diff:.2e}")
        assert diff <= 10 * noise + 1e-4
        # every parameter that has a gradient in the single-stream trainer has one in the side-stream trainer
        ga = {n for n, p in a.model.named_parameters() if p.grad is not None}
        gs = {n for n, p in s.model.named_parameters() if p.grad is not None}
        assert ga == gs
    finally:
        A.set_precision("fp32")


def test_adam_step_invalidates_packed_weight_cache():
    """optim.Adam updates parameters through raw pointers (no autograd version bump): the step itself must invalidate the
    packed bf16 operand copies the convolutions cache per parameter."""
    from affganwriting_b200 import ops
    from affganwriting_b200.optim import Adam
    A.set_precision("f16")
    try:
        x = ops.to_internal(torch.randn(2, 64, 8, 27, device="cuda"))
        w = torch.nn.Parameter(torch.randn(64, 64, 3, 3, device="cuda") * 0.05)
        opt = Adam([w], lr=1e-2)
        y0 = ops.conv2d(x, w, None, pad=1)
        y0.square().mean().backward()
        opt.step()
        y1 = ops.conv2d(x, w, None, pad=1).detach()
        fresh = ops.conv2d(x, w.detach().clone(), None, pad=1)
        assert float((y1 - y0.detach()).abs().max()) > 1e-3          # the stepped weights are used ...
        assert float((y1 - fresh).abs().max()) <= 1e-5               # ... and they are exactly the current ones
    finally:
        A.set_precision("fp32")


def test_graphed_generator_matches_eager(specs):
    """inference.GraphedGenerator: the captured forward returns what the eager forward returns, and follows a weight update."""
    from affganwriting_b200.inference import GraphedGenerator
    from affganwriting_b200 import ops
    A.set_precision("f16")
    try:
        gen = _gen(specs).eval()
        batch = _cuda(O.synthetic_batch(4, 15))
        fast = GraphedGenerator(gen)
        with torch.no_grad():
            ref = gen(batch["tr_img"], batch["label_xt"]).clone()
            for _ in range(4):
                out = fast(batch["tr_img"], batch["label_xt"])
            assert fast._graph is not None
            assert float((out - ref).abs().max()) <= 2e-3           # run-to-run noise of the fp32 atomics is ~3e-4
            other = gen(batch["tr_img"].flip(0), batch["label_xt_swap"]).clone()
            out2 = fast(batch["tr_img"].flip(0), batch["label_xt_swap"])
            assert float((out2 - other).abs().max()) <= 2e-3
            gen.dec.model[7].conv.weight.mul_(1.5)           # a weight update between calls must be seen by the replay
            ops.weights_updated(gen)
            moved = gen(batch["tr_img"], batch["label_xt"]).clone()
            out3 = fast(batch["tr_img"], batch["label_xt"])
            assert float((out3 - moved).abs().max()) <= 2e-3 and float((out3 - ref).abs().max()) > 1e-2
    finally:
        A.set_precision("fp32")


def test_contran_model_modes_match_oracle(specs):
    """ConTranModel.forward (network_tro.py:39-138): cla_update / dis_update / gen_update compose the same losses and
    gradients as the oracle's per-pair formulation - including the batching of the paired discriminator / classifier
    evaluations into one 2B pass (an identity, see network_tro.ConTranModel._pair)."""
    from affganwriting_b200.network_tro import ConTranModel
    A.set_precision("fp32")
    cpu = O.synthetic_batch(2, 50)
    full = {}
    for pre, key in (("gen.", "gen_c50"), ("dis.", "dis"), ("cla.", "cla")):
        for k, v in W.make_state(specs[key]).items():
            full[pre + k] = v.clone().requires_grad_(v.is_floating_point())
    lt, ld, lc, _, _ = O.gen_update(cpu, full)
    lt.backward()
    gen_grads = {k[4:]: v.grad.clone() for k, v in full.items() if k.startswith("gen.") and v.grad is not None}
    for v in full.values():
        v.grad = None
    l_real, l_fake = O.dis_update(cpu, full)
    (l_real + l_fake).backward()
    dis_grads = {k[4:]: v.grad.clone() for k, v in full.items() if k.startswith("dis.") and v.grad is not None}
    l_c = O.cla_update(cpu, full)

    model = ConTranModel(O.NUM_WRITERS, device=torch.device("cuda", 0))
    model.load_state_dict({k: v.detach() for k, v in full.items()})
    model.train()
    batch = (None, cpu["tr_wid"], None, cpu["tr_img"], None, None, cpu["img_xt"], cpu["label_xt"], cpu["label_xt_swap"])
    got_c = model(batch, 0, "cla_update")
    assert abs(float(got_c) - float(l_c)) <= 1e-4 * max(1.0, abs(float(l_c)))
    model.zero_grad()
    model.load_state_dict({k: v.detach() for k, v in full.items()})         # BatchNorm buffers back to the start
    got_d = model(batch, 0, "dis_update")
    assert abs(float(got_d) - float(l_real + l_fake)) <= 2e-4 * max(1.0, abs(float(l_real + l_fake)))
    dots = na = nb = 0.0
    for k, p in model.dis.named_parameters():
        a, b = p.grad.double().cpu().reshape(-1), dis_grads[k].double().reshape(-1)
        dots += float(a @ b); na += float(a @ a); nb += float(b @ b)
    assert dots / (na ** 0.5 * nb ** 0.5) >= 0.9999
    model.zero_grad()
    model.load_state_dict({k: v.detach() for k, v in full.items()})
    got_t, got_ld, got_lc, _, _ = model(batch, 0, "gen_update")
    assert abs(float(got_ld) - float(ld)) <= 2e-4 * max(1.0, abs(float(ld)))
    assert abs(float(got_lc) - float(lc)) <= 2e-4 * max(1.0, abs(float(lc)))
    assert abs(float(got_t) - float(lt)) <= 2e-4 * max(1.0, abs(float(lt)))
    dots = na = nb = 0.0
    for k, p in model.gen.named_parameters():
        if p.grad is None or k not in gen_grads:
            continue
        a, b = p.grad.double().cpu().reshape(-1), gen_grads[k].double().reshape(-1)
        dots += float(a @ b); na += float(a @ a); nb += float(b @ b)
    cos = dots / (na ** 0.5 * nb ** 0.5)
    print(f"\nConTranModel gen_update vs oracle: l_total {float(got_t):.6f} / {float(lt):.6f}, gradient cosine {cos:.6f}")
    assert cos >= 0.999


def test_graph_replay_sees_weights_loaded_between_iterations(specs):
    """The captured sub-steps read the packed operand copies of the weights in place and do not re-pack them (the optimiser
    step does): weights written behind the trainer's back - load_state_dict between two iterations - must still reach the
    next replay."""
    from affganwriting_b200.trainer import Trainer
    import bench
    from affganwriting_b200 import load_data as LD
    A.set_precision("f16")
    try:
        dev = torch.device("cuda", 0)
        batch = LD.batch_to_device(bench.synthetic_batch(4, 50, 7), torch.device('cuda'))
        torch.manual_seed(0)
        t = Trainer(num_writers=500, device=dev, cuda_graph=True, overlap_exchange=True)
        t.GRAPH_WARMUP = 1
        e = Trainer(num_writers=500, device=dev)
        for _ in range(3):
            t.train_step(batch)                                       # eager, capture, replay
        assert t._graphs is not None and all(t._packed[n] for n in t.names)
        n_graph = t.graph_launches
        sd = {k: v.clone() for k, v in e.model.state_dict().items()}   # other weights
        t.join()
        t.model.load_state_dict(sd)
        lt = t.train_step(batch)                                      # a replay on the loaded weights
        le = e.train_step(batch)                                      # the eager trainer that owns them
        for k in le:
            assert abs(float(lt[k]) - float(le[k])) <= 2e-3 * max(1.0, abs(float(le[k]))), (k, float(lt[k]), float(le[k]))
        assert n_graph == t.graph_launches
    finally:
        A.set_precision("fp32")


def test_early_generator_forward_matches_the_single_stream_iteration(specs):
    """Trainer(early_generator_forward=True, concurrent_gen_heads=True) - the defaults of the replayed configuration:
    gen_update's generator forward is issued inside dis_update, beside the discriminator pass on a second stream, and in
    gen_update the classifier's pass over the generated pair runs beside the discriminator's.  Same launches, same losses, same gradients
    and BatchNorm buffers as the single-stream iteration; checked eagerly in fp32 (run-to-run noise ~1e-6) and as CUDA-graph
    replays in the benchmarked mode against an eager trainer with the option off."""
    from affganwriting_b200.trainer import Trainer
    import bench
    from affganwriting_b200 import load_data as LD
    dev = torch.device("cuda", 0)
    batch = LD.batch_to_device(bench.synthetic_batch(4, 50, 7), torch.device('cuda'))

    def cos(x, y, sub):
        gx = torch.cat([p.grad.flatten() for p in getattr(x.model, sub).parameters() if p.grad is not None])
        gy = torch.cat([p.grad.flatten() for p in getattr(y.model, sub).parameters() if p.grad is not None])
        assert gx.numel() == gy.numel()
        return float(torch.dot(gx.double(), gy.double()) / (gx.double().norm() * gy.double().norm()))

    A.set_precision("fp32")
    torch.manual_seed(0)
    a = Trainer(num_writers=500, device=dev)
    b = Trainer(num_writers=500, device=dev)
    s = Trainer(num_writers=500, device=dev, early_generator_forward=True, concurrent_gen_heads=True)
    assert s.early_generator_forward and s.concurrent_gen_heads
    s.model.side_text_encoder = True                              # (the Trainer switches it on with overlap_exchange)
    assert not a.early_generator_forward and not a.concurrent_gen_heads and not a.share_generator_forward
    s.model.load_state_dict(a.model.state_dict())
    b.model.load_state_dict(a.model.state_dict())
    b.train_step(batch)                                           # a second default run: the noise yardstick
    n0 = A.launch_count()
    la = a.train_step(batch)
    n_a = A.launch_count() - n0
    n0 = A.launch_count()
    ls = s.train_step(batch)
    n_s = A.launch_count() - n0
    assert n_s == n_a                                             # nothing skipped: both generator forwards run
    assert "pair" not in s._shared                                # consumed by gen_update
    torch.cuda.synchronize()
    for k in la:
        assert abs(float(la[k]) - float(ls[k])) <= 1e-5 * max(1.0, abs(float(la[k]))), (k, float(la[k]), float(ls[k]))
    for sub in ("gen", "dis", "cla"):
        c, c_noise = cos(a, s, sub), cos(a, b, sub)
        print(f"\nfp32 {sub}: gradient cosine default/default {c_noise:.9f}, default / early generator forward {c:.9f}")
        assert 1.0 - c <= 10 * (1.0 - c_noise) + 1e-5
    sa, ss = a.state_dict(), s.state_dict()
    for k, v in sa.items():
        if not v.is_floating_point():
            assert torch.equal(v, ss[k]), k
        elif "running_" in k:
            assert float((v - ss[k]).abs().max()) <= 1e-5 * max(1.0, float(v.abs().max())), k
    del a, b, s

    A.set_precision("f16")
    try:
        torch.manual_seed(0)
        a = Trainer(num_writers=500, device=dev)
        b = Trainer(num_writers=500, device=dev)
        g = Trainer(num_writers=500, device=dev, cuda_graph=True, overlap_exchange=True)
        assert g.early_generator_forward and g.concurrent_gen_heads and not a.early_generator_forward
        assert g.model.side_text_encoder and not a.model.side_text_encoder
        g.GRAPH_WARMUP = 1
        for t in (b, g):
            t.model.load_state_dict(a.model.state_dict())
        drift = drift_ee = 0.0
        for it in range(5):                                           # eager, capture pass, then pure replays
            la, lb, lg = a.train_step(batch), b.train_step(batch), g.train_step(batch)
            drift = max(drift, max(abs(float(la[k]) - float(lg[k])) for k in la))
            drift_ee = max(drift_ee, max(abs(float(la[k]) - float(lb[k])) for k in la))
            if it < 3:
                assert all(abs(float(la[k]) - float(lg[k])) <= 2e-3 * max(1.0, abs(float(la[k]))) for k in la), \
                    (it, {k: (float(la[k]), float(lg[k])) for k in la})
        assert g._graphs is not None and all(torch.isfinite(v) for v in lg.values())
        print(f"f16: loss drift over 5 iterations eager/eager {drift_ee:.2e}, eager / graph with early generator forward {drift:.2e}")
        assert drift <= 10 * drift_ee + 2e-2

        def absdrift(x, y):
            sx, sy = x.state_dict(), y.state_dict()
            for k, v in sx.items():
                if not v.is_floating_point():
                    assert torch.equal(v, sy[k]), k
            return max(float((v - sy[k]).abs().max()) for k, v in sx.items() if v.is_floating_point())
        g.join()
        assert absdrift(a, g) <= 10 * absdrift(a, b) + 0.3
    finally:
        A.set_precision("fp32")


def test_shared_generator_forward_matches_the_reference_composition(specs):
    """Trainer(share_generator_forward=True): dis_update generates the fake pair once WITH its autograd graph and gen_update
    back-propagates through it, instead of the reference's two identical generator forwards per iteration
    (network_tro.py:59-63,117-118).  Same losses, same gradients, same BatchNorm buffers (running statistics advanced twice):
    checked in fp32 mode, where run-to-run noise is ~1e-6, and as CUDA-graph replays in the benchmarked mode (the generator's
    autograd graph then spans two captured graphs)."""
    from affganwriting_b200.trainer import Trainer
    import bench
    from affganwriting_b200 import load_data as LD
    dev = torch.device("cuda", 0)
    batch = LD.batch_to_device(bench.synthetic_batch(4, 50, 7), torch.device('cuda'))

    def cos(x, y, sub):
        gx = torch.cat([p.grad.flatten() for p in getattr(x.model, sub).parameters() if p.grad is not None])
        gy = torch.cat([p.grad.flatten() for p in getattr(y.model, sub).parameters() if p.grad is not None])
        assert gx.numel() == gy.numel()
        return float(torch.dot(gx.double(), gy.double()) / (gx.double().norm() * gy.double().norm()))

    A.set_precision("fp32")
    torch.manual_seed(0)
    a = Trainer(num_writers=500, device=dev)
    b = Trainer(num_writers=500, device=dev)
    s = Trainer(num_writers=500, device=dev, share_generator_forward=True)
    assert s.share_generator_forward and not a.share_generator_forward
    s.model.load_state_dict(a.model.state_dict())
    b.model.load_state_dict(a.model.state_dict())
    b.train_step(batch)                                           # a second default run: the noise yardstick
    n0 = A.launch_count()
    la = a.train_step(batch)
    n_a = A.launch_count() - n0
    n0 = A.launch_count()
    ls = s.train_step(batch)
    n_s = A.launch_count() - n0
    assert n_s < n_a - 200                                        # one generator forward less
    assert not s._shared                                          # consumed by gen_update
    for k in la:
        assert abs(float(la[k]) - float(ls[k])) <= 1e-5 * max(1.0, abs(float(la[k]))), (k, float(la[k]), float(ls[k]))
    for sub in ("gen", "dis", "cla"):
        c, c_noise = cos(a, s, sub), cos(a, b, sub)
        print(f"\nfp32 {sub}: gradient cosine default/default {c_noise:.9f}, default / shared-forward {c:.9f}")
        # (fp32 run-to-run noise: order of the fp32 atomics in the statistics / weight-gradient reductions, amplified by the
        # un-normalised discriminator on a batch of 4)
        assert 1.0 - c <= 10 * (1.0 - c_noise) + 1e-5
    sa, ss = a.state_dict(), s.state_dict()
    for k, v in sa.items():
        if not v.is_floating_point():
            assert torch.equal(v, ss[k]), k                       # num_batches_tracked: advanced twice per iteration
        elif "running_" in k:
            assert float((v - ss[k]).abs().max()) <= 1e-5 * max(1.0, float(v.abs().max())), k
    del a, b, s

    A.set_precision("f16")
    try:
        torch.manual_seed(0)
        a = Trainer(num_writers=500, device=dev)
        b = Trainer(num_writers=500, device=dev)
        g = Trainer(num_writers=500, device=dev, cuda_graph=True, overlap_exchange=True, share_generator_forward=True)
        g.GRAPH_WARMUP = 1
        for t in (b, g):
            t.model.load_state_dict(a.model.state_dict())
        drift = drift_ee = 0.0
        for it in range(5):                                           # eager, capture pass, then pure replays
            la, lb, lg = a.train_step(batch), b.train_step(batch), g.train_step(batch)
            drift = max(drift, max(abs(float(la[k]) - float(lg[k])) for k in la))
            drift_ee = max(drift_ee, max(abs(float(la[k]) - float(lb[k])) for k in la))
            if it < 3:
                assert all(abs(float(la[k]) - float(lg[k])) <= 2e-3 * max(1.0, abs(float(la[k]))) for k in la), \
                    (it, {k: (float(la[k]), float(lg[k])) for k in la})
        assert g._graphs is not None and all(torch.isfinite(v) for v in lg.values())
        print(f"f16: loss drift over 5 iterations eager/eager {drift_ee:.2e}, eager / graph shared-forward {drift:.2e}")
        assert drift <= 10 * drift_ee + 2e-2

        def absdrift(x, y):                                           # (Adam turns noise-level gradients into +-lr steps: absolute)
            sx, sy = x.state_dict(), y.state_dict()
            for k, v in sx.items():
                if not v.is_floating_point():
                    assert torch.equal(v, sy[k]), k
            return max(float((v - sy[k]).abs().max()) for k, v in sx.items() if v.is_floating_point())
        g.join()
        assert absdrift(a, g) <= 10 * absdrift(a, b) + 0.3
    finally:
        A.set_precision("fp32")
