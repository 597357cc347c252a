"""Stage-by-stage comparison against the CPU oracle (debug aid, run under gpurun)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affganwriting_b200 as A
from affganwriting_b200 import modules_tro as M, load_data, ops
from oracle import affgw_oracle as O, weights as W

spec = json.load(open(os.path.join(ROOT, "tests/golden/state_spec.json")))
sd = W.make_state(spec["gen_c15"])

def err(a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    return float((a - b).abs().max()), float(b.abs().max())

def build():
    load_data.NUM_CHANNEL = 15
    g = M.GenModel_FC(12); load_data.NUM_CHANNEL = 50
    g.load_state_dict(sd)
    return g.cuda()

for mode in ("fp32", "bf16"):
    for training, B in ((True, 4), (False, 1), (False, 4)):
        A.set_precision(mode)
        gen = build().train(training)
        batch = O.synthetic_batch(4, 15)
        img, lab = batch["tr_img"][:B], batch["label_xt"][:B]
        with torch.no_grad():
            ro = O.image_encoder(img, sd)
            fxt_o, fe_o = O.text_encoder(lab, ro[-1].shape, sd, "enc_text.", training)
            fm_o = O.mix(ro, fe_o, sd)
            xo = O.decoder(fm_o, ro, fxt_o, sd, "dec.", training)
        for rep in range(2):
            gen.load_state_dict(sd)
            with torch.no_grad():
                r = gen.enc_image(img.cuda())
                fxt, fe = gen.enc_text(lab.cuda(), r[-1].shape)
                fm = gen.mix(r, fe)
                x = gen.decode(fm, r, fe, fxt)
            print(f"[{mode} train={training} B={B} rep={rep}] enc5 {err(r[5], ro[5])} enc3 {err(r[3], ro[3])} f_xt {err(fxt, fxt_o)} "
                  f"f_embed {err(fe, fe_o)} f_mix {err(fm, fm_o)} xg {err(x, xo)}")
A.set_precision("fp32")
