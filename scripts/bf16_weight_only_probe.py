"""CPU experiment (oracle only): image error when ONLY the weights, or ONLY the activations, of every convolution /
linear are rounded to bf16 - i.e. what a 2-pass split (a_hi*w_hi + a_lo*w_hi, or a_hi*w_hi + a_hi*w_lo) would cost."""
import json, os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import affgw_oracle as O, weights as W
spec = json.load(open(os.path.join(ROOT, "tests/golden/state_spec.json")))["gen_c15"]
sd = W.make_state(spec)
batch = O.synthetic_batch(4, 15)
r = lambda t: t.bfloat16().float()
with torch.no_grad():
    ref = O.gen_forward(batch["tr_img"], batch["label_xt"], sd)
    for name, rw, ra in (("weights only", True, False), ("activations only", False, True), ("both", True, True)):
        O._conv = lambda x, w, b=None, **kw: F.conv2d(r(x) if ra else x, r(w) if rw else w, b, **kw)
        O._linear = lambda x, w, b=None: F.linear(r(x) if ra else x, r(w) if rw else w, b)
        y = O.gen_forward(batch["tr_img"], batch["label_xt"], sd)
        d = y - ref
        print(f"{name:18s} image max-abs {float(d.abs().max()):.3e} rms {float(d.square().mean().sqrt()):.3e}")
