"""CPU experiment: growth of the bf16-rounding error through the VGG-IN style encoder (oracle only)."""
import json, os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import affgw_oracle as O, weights as W
B, C = 4, 15
spec = json.load(open(os.path.join(ROOT, "tests/golden/state_spec.json")))["gen_c%d" % C]
sd = W.make_state(spec)
batch = O.synthetic_batch(B, C)
def r(t): return t.bfloat16().float()
def run(round_act, round_w, round_in, round_pre):
    x = batch["tr_img"]
    if round_in: x = r(x)
    outs = []
    for kind, idx, _ in O.vgg_layers():
        if kind == "conv":
            w = sd[f"enc_image.model.features.{idx}.weight"]
            x = F.conv2d(x, r(w) if round_w else w, sd[f"enc_image.model.features.{idx}.bias"], padding=1)
            if round_pre: x = r(x)
        elif kind == "in":
            x = O.instance_norm(x)
        elif kind == "relu":
            x = torch.relu(x)
            if round_act: x = r(x)
            outs.append(x)
        else:
            x = F.max_pool2d(x, 2, 2)
    return outs
with torch.no_grad():
    ref = run(0, 0, 0, 0)
    for name, cfg in (("all", (1, 1, 1, 1)), ("act only", (1, 0, 0, 0)), ("w only", (0, 1, 0, 0)), ("input only", (0, 0, 1, 0)), ("pre only", (0, 0, 0, 1))):
        got = run(*cfg)
        print(name, " ".join(f"{float((g - f).square().mean().sqrt() / f.square().mean().sqrt()):.1e}" for g, f in zip(got, ref)))
    # min per-(n,c) std of the conv outputs per layer (instance norm amplification)
    x = batch["tr_img"]; 
    for kind, idx, _ in O.vgg_layers():
        if kind == "conv":
            x = F.conv2d(x, sd[f"enc_image.model.features.{idx}.weight"], sd[f"enc_image.model.features.{idx}.bias"], padding=1)
            s = x.std(dim=(2, 3)); m = x.mean(dim=(2, 3)).abs()
            print(f"layer {idx}: per-(n,c) std min {float(s.min()):.2e} median {float(s.median()):.2e}; |mean|/std median {float((m / s).median()):.2f} max {float((m / s).max()):.1f}")
        elif kind == "in": x = O.instance_norm(x)
        elif kind == "relu": x = torch.relu(x)
        else: x = F.max_pool2d(x, 2, 2)
