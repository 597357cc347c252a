"""CPU experiment: which rounding points of the bf16 storage model dominate the image error (oracle only)."""
import json, os, sys, itertools
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import affgw_oracle as O, weights as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
C = int(sys.argv[2]) if len(sys.argv) > 2 else 15
spec = json.load(open(os.path.join(ROOT, "tests/golden/state_spec.json")))["gen_c%d" % C]
sd = W.make_state(spec)
batch = O.synthetic_batch(B, C)
region = ["none"]
enabled = set()
orig_q = O.q
def q(t):
    if region[-1] in enabled:
        return t + (t.detach().bfloat16().float() - t.detach()) if t.is_floating_point() else t
    return t
O.q = q
def wrap(name, reg):
    f = getattr(O, name)
    def g(*a, **k):
        region.append(reg)
        try:
            return f(*a, **k)
        finally:
            region.pop()
    setattr(O, name, g)
wrap("image_encoder", "enc"); wrap("text_encoder", "text"); wrap("mix", "mix"); wrap("decoder", "dec"); wrap("iaff", "iaff"); wrap("get_key", "key")
with torch.no_grad():
    ref = O.gen_forward(batch["tr_img"], batch["label_xt"], sd)
    for en in (["enc"], ["text"], ["mix"], ["dec"], ["iaff"], ["key"], ["enc", "text", "mix", "dec", "key"], ["enc", "text", "mix", "dec", "key", "iaff"]):
        enabled.clear(); enabled.update(en)
        y = O.gen_forward(batch["tr_img"], batch["label_xt"], sd)
        d = (y - ref)
        print(f"{'+'.join(en):40s} max-abs {float(d.abs().max()):.3e} rms {float(d.square().mean().sqrt()):.3e}")
