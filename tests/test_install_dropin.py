"""The drop-in claim of SURVEY.md §8(b), exercised: after affganwriting_b200.install.install() the REFERENCE's own
network_tro.ConTranModel (network_tro.py:17-26) is built from this package's classes, loads a state_dict the reference's
classes wrote, key for key and in the same order - and, on the GPU box (from the staged copy under oracle/_ref/), its own
forward() runs rec_update / cla_update / dis_update / gen_update through libaffgw kernels with the reference's unmodified
recogniser attached."""
import numpy as np
import pytest
import torch

from oracle import ref_bootstrap as rb

needs_ref = pytest.mark.skipif(not rb.available(), reason="reference tree neither at /root/reference nor staged in oracle/_ref")


@pytest.fixture()
def installed():
    import affganwriting_b200.install as inst
    ns = rb.load_network(50)
    inst.install(recogniser=False)
    try:
        yield ns, inst
    finally:
        inst.uninstall()


def _reference_state(ns):
    """state_dict written by the REFERENCE's classes (VGG wiring of modules_tro.py:211)."""
    torch.manual_seed(3)
    m = ns.modules_tro
    parts = {"gen": ns.Gen(12), "cla": m.WriterClaModel(500), "dis": m.DisModel(), "rec": m.RecModel(pretrain=False)}
    sd = {}
    for pre, mod in parts.items():
        for k, v in mod.state_dict().items():
            sd[f"{pre}.{k}"] = v
    return sd


@needs_ref
def test_reference_contran_model_is_built_from_dropin_classes():
    ns = rb.load_network(50)
    ref_sd = _reference_state(ns)          # before install(): the reference's own classes
    import affganwriting_b200.install as inst
    done = inst.install(recogniser=False)
    try:
        assert ("network_tro", "GenModel_FC") in done and ("modules_tro", "Conv2dBlock") in done and ("blocks", "iAFF") in done
        nt = ns.network_tro
        model = nt.ConTranModel(500, 500, True)
        for sub in ("gen", "dis", "cla"):
            assert type(getattr(model, sub)).__module__.startswith("affganwriting_b200."), sub
        assert type(model.rec).__module__ == "modules_tro"            # the recogniser stays the reference's
        assert type(model.gen.dec.model[0].model[0].model[0]).__module__ == "affganwriting_b200.blocks"
        own = model.state_dict()
        assert list(own.keys()) == list(ref_sd.keys())               # same keys, same registration order
        for k, v in ref_sd.items():
            assert own[k].shape == v.shape and own[k].dtype == v.dtype, k
        missing, unexpected = model.load_state_dict(ref_sd, strict=True)
        assert not missing and not unexpected
        k = "gen.dec.model.0.model.1.model.1.norm.iAff.global_att.2.running_var"
        assert torch.equal(model.state_dict()[k].cpu(), ref_sd[k])
    finally:
        inst.uninstall()
    assert ns.modules_tro.DisModel.__module__ == "modules_tro" and ns.network_tro.GenModel_FC.__module__ == "modules_tro"
    # default install(): the recogniser is replaced too, with the same keys
    inst.install()
    try:
        model = ns.network_tro.ConTranModel(500, 500, True)
        assert type(model.rec).__module__ == "affganwriting_b200.recognizer"
        assert list(model.state_dict().keys()) == list(ref_sd.keys())
        model.load_state_dict(ref_sd, strict=True)
    finally:
        inst.uninstall()
    assert ns.modules_tro.RecModel.__module__ == "modules_tro"


@needs_ref
@pytest.mark.gpu
def test_reference_contran_forward_runs_on_libaffgw(installed, specs):
    """One full iteration of the reference driver (main_run.py:146-167) through the REFERENCE's ConTranModel.forward after
    install(): losses finite, every sub-network receives gradients, libaffgw kernels did the generator / discriminator /
    classifier work, and the generator image equals the CPU oracle's on the same weights."""
    import affganwriting_b200 as A
    import loss_tro
    from oracle import affgw_oracle as O
    from oracle import weights as W
    ns, _ = installed
    A.set_precision("fp32")
    nt = ns.network_tro
    model = nt.ConTranModel(500, 500, True)
    model.gen.load_state_dict(W.make_state(specs["gen_c50"]))
    model.dis.load_state_dict(W.make_state(specs["dis"]))
    model.cla.load_state_dict(W.make_state(specs["cla"]))
    model.train()
    cpu = O.synthetic_batch(4, 50)
    tr_label = cpu["label_xt"].unsqueeze(1).repeat(1, 50, 1)
    batch = (np.zeros(4, dtype=np.int64), cpu["tr_wid"], np.arange(4), cpu["tr_img"], torch.full((4, 50), 216), tr_label,
             cpu["img_xt"], cpu["label_xt"], cpu["label_xt_swap"])
    n0 = A.launch_count()
    cer = loss_tro.CER()
    l_rec = model(batch, 0, "rec_update", cer)
    assert any(p.grad is not None for p in model.rec.seq2seq.decoder.parameters())
    model.zero_grad()
    l_cla = model(batch, 0, "cla_update")
    assert all(p.grad is not None for p in model.cla.parameters())
    model.zero_grad()
    model.iter_num = 1                                       # skip the PNG dump of iteration 0 (network_tro.py:132-137)
    l_dis = model(batch, 0, "dis_update")
    assert all(p.grad is not None for p in model.dis.parameters())
    model.zero_grad()
    l_total, l_d, l_c, l_l1, l_r = model(batch, 0, "gen_update", [loss_tro.CER(), loss_tro.CER()])
    assert A.launch_count() - n0 > 1000
    live = [k for k, p in model.gen.named_parameters() if p.grad is not None]
    assert len(live) == len(list(model.gen.parameters())) - 96
    with torch.no_grad():
        ref_c = float(O.cla_update(cpu, {"cla." + k: v for k, v in W.make_state(specs["cla"]).items()}))
    assert abs(float(l_cla) - ref_c) <= 1e-4 * max(1.0, abs(ref_c))
    for v in (l_cla, l_dis, l_d, l_c):
        assert torch.isfinite(v)
    assert float(l_total) == float(l_total) or torch.isnan(l_r)       # l_rec is NaN-prone on random-init logits (rec_oracle.py header)
    # the image the reference's forward produced through our generator == CPU oracle on the same weights
    with torch.no_grad():
        model.eval()
        model.gen.load_state_dict(W.make_state(specs["gen_c50"]))
        f = model.gen.enc_image(cpu["tr_img"].cuda())
        ft, fe = model.gen.enc_text(cpu["label_xt"].cuda(), f[-1].shape)
        xg = model.gen.decode(model.gen.mix(f, fe), f, fe, ft)
        ref = O.gen_forward(cpu["tr_img"], cpu["label_xt"], W.make_state(specs["gen_c50"]), training=False)
    assert float((xg.cpu() - ref).abs().max()) <= 1e-4
    A.check_device_errors()
