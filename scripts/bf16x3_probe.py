"""CPU experiment: image error of the generator when every conv/linear is the 3-term split-bf16 product
a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with activations stored as hi+lo bf16 pairs (oracle only)."""
import json, os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import affgw_oracle as O, weights as W
B, C = 4, 15
spec = json.load(open(os.path.join(ROOT, "tests/golden/state_spec.json")))["gen_c%d" % C]
sd = W.make_state(spec)
batch = O.synthetic_batch(B, C)
def hi(t): return t.bfloat16().float()
def split(t):
    h = hi(t); return h, hi(t - h)
mode = {"m": "fp32"}
def conv(x, w, b=None, **kw):
    if mode["m"] == "fp32": return F.conv2d(x, w, b, **kw)
    xh, xl = split(x); wh, wl = split(w)
    if mode["m"] == "x1": return F.conv2d(xh, wh, b, **kw)
    if mode["m"] == "x2": return F.conv2d(xh, wh, b, **kw) + F.conv2d(xl, wh, None, **kw)
    return F.conv2d(xh, wh, b, **kw) + F.conv2d(xl, wh, None, **kw) + F.conv2d(xh, wl, None, **kw)
def lin(x, w, b=None):
    if mode["m"] == "fp32": return F.linear(x, w, b)
    xh, xl = split(x); wh, wl = split(w)
    if mode["m"] == "x1": return F.linear(xh, wh, b)
    if mode["m"] == "x2": return F.linear(xh, wh, b) + F.linear(xl, wh)
    return F.linear(xh, wh, b) + F.linear(xl, wh) + F.linear(xh, wl)
O._conv = conv; O._linear = lin
O.q = lambda t: t
with torch.no_grad():
    ref = O.gen_forward(batch["tr_img"], batch["label_xt"], sd)
    for m in ("x1", "x2", "x3"):
        mode["m"] = m
        y = O.gen_forward(batch["tr_img"], batch["label_xt"], sd)
        d = y - ref
        print(m, f"max-abs {float(d.abs().max()):.3e} rms {float(d.square().mean().sqrt()):.3e}")
