"""CPU experiment (oracle only): which convolutions tolerate fewer tensor-core passes / which operand format.

Every convolution / linear of the oracle is replaced by an autograd Function that rounds its GEMM operands the way the
tcgen05 kernels see them, separately for the three GEMMs of a layer:

    forward   y  = conv(R(x, fx, pf), R(w, fw, pf))
    dgrad     dx = conv^T(R(dy, fg, pd), R(w, fw, pd))
    wgrad     dw = corr(R(x, fx, pw), R(dy, fg, pw))

R(t, fmt, passes): passes = 1 -> round to fmt ("bf16" 8 significant bits, "f16" 11 bits); passes = 3 -> hi + lo planes of
that format (the a_lo * w_lo product the kernels drop is below 2^-16 relative and ignored here); passes = 0 -> exact fp32.

Usage: python scripts/precision_sweep.py [B] [C_s] [experiment]
"""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import affgw_oracle as O, weights as W   # noqa: E402

W_SCALE = 256.0     # power-of-two scale applied to fp16 weight planes (undone exactly in the epilogue)


def rnd(t, fmt, passes, scale=1.0):
    if passes == 0:
        return t
    if scale == "amax":         # dynamic power-of-two scale: amax * s in [2^13, 2^14)  (dY planes in fp16)
        amax = float(t.abs().max())
        scale = 2.0 ** int(torch.floor(torch.log2(torch.tensor(16384.0 / max(amax, 1e-30)))))
    cast = (lambda v: v.bfloat16().float()) if fmt == "bf16" else (lambda v: v.clamp(-65504, 65504).half().float())
    s = t * scale
    hi = cast(s)
    if passes == 3:
        hi = hi + cast(s - hi)
    return hi / scale


def e4m3(t, scale):
    """saturating round to fp8 e4m3 of t * scale (scale a power of two), returned unscaled"""
    return (t * scale).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float() / scale


def pow2_scale(t, top):
    amax = float(t.abs().max())
    return 2.0 ** int(torch.floor(torch.log2(torch.tensor(top / max(amax, 1e-30)))))


def fp8_corrected_conv(x, w, b, stride, padding, mode):
    """a*w ~= f16(a)*f16(w) [kind::f16, 1 MMA] + e4m3(a_lo)*e4m3(w) + e4m3(a)*e4m3(w_lo) [kind::f8f6f4: 2 MMAs at twice the rate].
    mode "f8c": per-tensor amax scales for all four fp8 planes; "f8cfix": fixed activation scales (no amax pass):
    a * 8 (saturates above 56), a_lo * 2^14."""
    half = lambda v: v.clamp(-65504, 65504).half().float()
    x16, w16 = half(x), half(w * W_SCALE) / W_SCALE
    xl, wl = x - x16, w - w16
    if mode == "f8cfix":
        sx, sxl = 8.0, 8.0 * 2048.0
    else:
        sx, sxl = pow2_scale(x, 256.0), pow2_scale(xl, 256.0)
    sw, swl = pow2_scale(w, 256.0), pow2_scale(wl, 256.0)
    y = F.conv2d(x16, w16, b, stride=stride, padding=padding)
    y = y + F.conv2d(e4m3(xl, sxl), e4m3(w, sw), None, stride=stride, padding=padding)
    y = y + F.conv2d(e4m3(x, sx), e4m3(wl, swl), None, stride=stride, padding=padding)
    return y


class SimConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding, spec):
        fx, fw, fg, pf, pd, pw = spec
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, spec, b is not None)
        if isinstance(pf, str):                                   # "f8c<bits>": one fp16 MMA + the two correction products in fp8
            return fp8_corrected_conv(x, w, b, stride, padding, pf)
        pfx, pfw = pf if isinstance(pf, tuple) else (pf, pf)      # (planes of x, planes of w): (3, 1) = two MMAs, x split only
        return F.conv2d(rnd(x, fx, pfx), rnd(w, fw, pfw, W_SCALE if fw == "f16" else 1.0), b, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, padding, (fx, fw, fg, pf, pd, pw), has_b = ctx.cfg
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.nn.grad.conv2d_input(x.shape, rnd(w, fw, pd, W_SCALE if fw == "f16" else 1.0), rnd(dy, fg, pd, "amax" if fg == "f16" else 1.0),
                                            stride=stride, padding=padding)
        if ctx.needs_input_grad[1]:
            dw = torch.nn.grad.conv2d_weight(rnd(x, fx, pw), w.shape, rnd(dy, fg, pw, "amax" if fg == "f16" else 1.0), stride=stride, padding=padding)
        if has_b and ctx.needs_input_grad[2]:
            db = dy.sum(dim=(0, 2, 3))
        return dx, dw, db, None, None, None


class Policy:
    """name -> (fx, fw, fg, pf, pd, pw); `default` for unnamed layers."""

    def __init__(self, default, overrides=None):
        self.default, self.overrides = default, dict(overrides or {})
        self.names = {}
        self.seen = []

    def bind(self, sd, prefix=""):
        for k, v in sd.items():
            self.names[id(v)] = prefix + k

    def spec(self, w):
        name = self.names.get(id(w), "?")
        if name not in self.seen:
            self.seen.append(name)
        for pat, sp in self.overrides.items():
            if name.startswith(pat):
                return sp
        return self.default


POLICY = [None]


def _conv(x, w, b=None, stride=1, padding=0):
    sp = POLICY[0].spec(w)
    return SimConv.apply(x, w, b, stride, padding, sp)


def _linear(x, w, b=None):
    sp = POLICY[0].spec(w)
    lead = x.shape[:-1]
    y = SimConv.apply(x.reshape(-1, x.shape[-1], 1, 1), w.view(w.shape[0], w.shape[1], 1, 1), b, 1, 0, sp)
    return y.reshape(*lead, w.shape[0])


O._conv, O._linear = _conv, _linear


def cosine(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def run(full, batch, policy, dis_policy=None):
    """gen_update + dis_update + cla_update under `policy` (dis_update under `dis_policy` when given); returns images and
    gradients."""
    POLICY[0] = policy
    policy.bind(full)
    if dis_policy is not None:
        dis_policy.bind(full)
    for p in full.values():
        p.grad = None
    lt, ld, lc, xg, xgs = O.gen_update(batch, full)
    lt.backward()
    gg = {k: v.grad.clone() for k, v in full.items() if k.startswith("gen.") and v.grad is not None}
    for p in full.values():
        p.grad = None
    if dis_policy is not None:
        POLICY[0] = dis_policy
    l_real, l_fake = O.dis_update(batch, full)
    POLICY[0] = policy
    (l_real + l_fake).backward()
    dg = {k: v.grad.clone() for k, v in full.items() if k.startswith("dis.") and v.grad is not None}
    for p in full.values():
        p.grad = None
    O.cla_update(batch, full).backward()
    cg = {k: v.grad.clone() for k, v in full.items() if k.startswith("cla.") and v.grad is not None}
    return dict(xg=xg.detach(), xgs=xgs.detach(), gen=gg, dis=dg, cla=cg, losses=(float(ld), float(lc), float(l_real), float(l_fake)))


def compare(ref, got, noise_keys=()):
    img = max(float((ref["xg"] - got["xg"]).abs().max()), float((ref["xgs"] - got["xgs"]).abs().max()))
    out = {"img": img}
    for net in ("gen", "dis", "cla"):
        dots = na = nb = 0.0
        worst, worst_k = 1.0, None
        for k, g in ref[net].items():
            if k in noise_keys or float(g.norm()) < 1e-12:
                continue
            a, b = got[net][k].double().reshape(-1), g.double().reshape(-1)
            dots += float(a @ b); na += float(a @ a); nb += float(b @ b)
            c = cosine(got[net][k], g)
            if c < worst:
                worst, worst_k = c, k
        out[net] = (dots / (na ** 0.5 * nb ** 0.5), worst, worst_k)
    return out


def fmt(res):
    return ("img %.2e | gen cos %.6f (worst %.6f %s) | dis %.6f (%.6f) | cla %.6f (%.6f)" %
            (res["img"], res["gen"][0], res["gen"][1], (res["gen"][2] or "")[-40:], res["dis"][0], res["dis"][1],
             res["cla"][0], res["cla"][1]))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 15
    exp = sys.argv[3] if len(sys.argv) > 3 else "global"
    torch.set_num_threads(os.cpu_count())
    spec = json.load(open(os.path.join(ROOT, "tests/golden/state_spec.json")))
    full = {}
    for pre, key in (("gen.", "gen_c%d" % C), ("dis.", "dis"), ("cla.", "cla")):
        for k, v in W.make_state(spec[key]).items():
            full[pre + k] = v.clone().requires_grad_(v.is_floating_point())
    batch = O.synthetic_batch(B, C)
    exact = ("f16", "f16", "bf16", 0, 0, 0)
    t0 = time.time()
    ref = run(full, batch, Policy(exact))
    print("fp32 reference: %.1f s; losses %s" % (time.time() - t0, ref["losses"]))
    # biases in front of a normalisation have an exactly-zero gradient: rounding noise only
    noise = {k for k, g in ref["gen"].items() if k.endswith(".bias") and float(g.norm()) < 1e-6}

    def show(tag, policy):
        r = compare(ref, run(full, batch, policy), noise)
        print("%-44s %s" % (tag, fmt(r)), flush=True)
        return r

    if exp == "global":
        show("bf16 3/3/3", Policy(("bf16", "bf16", "bf16", 3, 3, 3)))
        show("bf16 1/1/1", Policy(("bf16", "bf16", "bf16", 1, 1, 1)))
        show("f16 x,w 3/3/3 (dy bf16)", Policy(("f16", "f16", "bf16", 3, 3, 3)))
        show("f16 x,w 1/1/1 (dy bf16 1)", Policy(("f16", "f16", "bf16", 1, 1, 1)))
        show("f16 x,w fwd 3, bwd 1/1", Policy(("f16", "f16", "bf16", 3, 1, 1)))
        show("f16 x,w fwd 1, bwd exact", Policy(("f16", "f16", "bf16", 1, 0, 0)))
        show("bf16 fwd 3, bwd 1/1", Policy(("bf16", "bf16", "bf16", 3, 1, 1)))
        show("bf16 fwd 3, dgrad 1, wgrad 3", Policy(("bf16", "bf16", "bf16", 3, 1, 3)))
        show("bf16 fwd 3, dgrad 3, wgrad 1", Policy(("bf16", "bf16", "bf16", 3, 3, 1)))
    elif exp == "layers":
        # one layer at a time forward-1-pass (f16), everything else exact
        p = Policy(exact)
        run(full, batch, p)
        names = [n for n in p.seen if n.endswith("weight")]
        for fmt_name in ("f16", "bf16"):
            for n in names:
                show("%s fwd1 only %s" % (fmt_name, n[-36:]), Policy(exact, {n: (fmt_name, fmt_name, "bf16", 1, 0, 0)}))
    elif exp == "plans":
        vgg = ["gen.enc_image.model.features.%d." % i for i in (0, 3, 6, 9, 13, 16, 19, 22, 26, 29, 32, 35, 39, 42, 45, 48)]
        for fm in ("f16",):
            for k in (0, 2, 4, 6, 8, 10, 12, 16):
                ov = {n: (fm, fm, "bf16", 3, 1, 1) for n in vgg[:k]}
                show("%s: first %d VGG convs fwd 3, rest 1; bwd 1" % (fm, k), Policy((fm, fm, "bf16", 1, 1, 1), ov))
    elif exp == "plans2":
        vgg = ["gen.enc_image.model.features.%d." % i for i in (0, 3, 6, 9, 13, 16, 19, 22, 26, 29, 32, 35, 39, 42, 45, 48)]
        dec_res = ["gen.dec.model.0.model.%d.model.%d.conv." % (i, j) for i in (0, 1) for j in (0, 1)]
        dec_up = ["gen.dec.model.%d.conv." % i for i in (2, 4, 6)]
        f = "f16"
        hi, lo = (f, f, "bf16", 3, 1, 1), (f, f, "bf16", 1, 1, 1)
        for tag, nv, ones in (("A: vgg8+ dec", 8, dec_res + dec_up), ("B: vgg12+ dec", 12, dec_res + dec_up),
                              ("C: vgg8+ ups only", 8, dec_up), ("D: vgg4+ dec", 4, dec_res + dec_up),
                              ("E: ups only", 16, dec_up), ("F: vgg10+ dec", 10, dec_res + dec_up)):
            ov = {n: lo for n in vgg[nv:] + ones}
            show(tag + " fwd1, rest fwd3; bwd 1", Policy(hi, ov))
        ov = {n: lo for n in vgg[8:] + dec_res + dec_up}
        ov.update({"dis.": (f, f, "bf16", 3, 1, 3), "cla.": (f, f, "bf16", 3, 1, 3)})
        show("A + dis/cla wgrad 3", Policy(hi, ov))
        ov = {n: ("bf16", "bf16", "bf16", 1, 1, 1) for n in dec_up}
        show("bf16 all: ups fwd1, rest 3; bwd 1", Policy(("bf16", "bf16", "bf16", 3, 1, 1), ov))
    elif exp == "plans3":
        dec_res = ["gen.dec.model.0.model.%d.model.%d.conv." % (i, j) for i in (0, 1) for j in (0, 1)]
        dec_up = ["gen.dec.model.%d.conv." % i for i in (2, 4, 6)]
        b3, f1 = ("bf16", "bf16", "bf16", 3, 1, 1), ("f16", "f16", "f16", 1, 1, 1)
        show("bf16 3/1/1 everywhere (shipping)", Policy(b3))
        show("ups f16 1/1/1 (dy f16), rest bf16 3/1/1", Policy(b3, {n: f1 for n in dec_up}))
        show("ups+res f16 1/1/1, rest bf16 3/1/1", Policy(b3, {n: f1 for n in dec_up + dec_res}))
        show("last 2 ups f16 1/1/1, rest bf16 3/1/1", Policy(b3, {n: f1 for n in dec_up[1:]}))
        show("ups f16 fwd 1, bwd 3 ; rest bf16 3/1/1", Policy(b3, {n: ("f16", "f16", "f16", 1, 3, 3) for n in dec_up}))
    elif exp == "plans4":
        dec_res = ["gen.dec.model.0.model.%d.model.%d.conv." % (i, j) for i in (0, 1) for j in (0, 1)]
        dec_up = ["gen.dec.model.%d.conv." % i for i in (2, 4, 6)]
        vgg = ["gen.enc_image.model.features.%d." % i for i in (0, 3, 6, 9, 13, 16, 19, 22, 26, 29, 32, 35, 39, 42, 45, 48)]
        f3, f1 = ("f16", "f16", "f16", 3, 1, 1), ("f16", "f16", "f16", 1, 1, 1)
        show("all f16 (dY f16 scaled) 3/1/1", Policy(f3))
        show("all f16 3/1/1, decoder res+ups 1/1/1", Policy(f3, {n: f1 for n in dec_up + dec_res}))
        show("  + dis/cla fwd 1", Policy(f3, {**{n: f1 for n in dec_up + dec_res}, "dis.": f1, "cla.": f1}))
        show("  + last 4 VGG fwd 1", Policy(f3, {n: f1 for n in dec_up + dec_res + vgg[12:]}))
        show("  + last 8 VGG fwd 1", Policy(f3, {n: f1 for n in dec_up + dec_res + vgg[8:]}))
    elif exp == "disfwd":
        # the generator forward of dis_update runs under no_grad and only feeds the discriminator: how few passes does IT need?
        dec_up = ["gen.dec.model.%d.conv." % i for i in (2, 4, 6)]
        f3, f1 = ("f16", "f16", "f16", 3, 1, 1), ("f16", "f16", "f16", 1, 1, 1)
        base = {n: f1 for n in dec_up}
        ship = Policy(f3, base)
        def show2(tag, dis_policy):
            r = compare(ref, run(full, batch, ship, dis_policy), noise)
            print("%-44s %s" % (tag, fmt(r)), flush=True)
        show2("shipping everywhere", None)
        if len(sys.argv) <= 4:
            show2("dis_update: generator fwd 1 pass (f16)", Policy(f3, {"gen.": f1}))
            show2("dis_update: generator fwd 1 pass (bf16)", Policy(f3, {"gen.": ("bf16", "bf16", "bf16", 1, 1, 1)}))
        vgg = ["gen.enc_image.model.features.%d." % i for i in (0, 3, 6, 9, 13, 16, 19, 22, 26, 29, 32, 35, 39, 42, 45, 48)]
        show2("dis_update: VGG fwd 1 pass, decoder as shipped", Policy(f3, {**base, **{n: f1 for n in vgg}}))
        dec_res = ["gen.dec.model.0.model.%d.model.%d.conv." % (i, j) for i in (0, 1) for j in (0, 1)]
        if len(sys.argv) > 5 and sys.argv[5] == "disnet":
            # the discriminator's own forward inside dis_update (its weight gradients are what the sub-step produces)
            rel = {**base, **{n: f1 for n in vgg[4:] + dec_res}}
            show2("shipped relaxed generator forward", Policy(f3, rel))
            show2("  + discriminator forward 1 pass", Policy(f3, {**rel, "dis.": f1}))
            show2("  + discriminator fwd 1, wgrad 3", Policy(f3, {**rel, "dis.": ("f16", "f16", "f16", 1, 1, 3)}))
            return
        if len(sys.argv) > 5:
            for k in (2, 4, 6, 8):
                show2("dis_update: VGG[%d:] + decoder ResBlocks fwd 1 pass" % k, Policy(f3, {**base, **{n: f1 for n in vgg[k:] + dec_res}}))
            return
        for k in (8, 12):
            show2("dis_update: VGG[%d:] fwd 1 pass" % k, Policy(f3, {**base, **{n: f1 for n in vgg[k:]}}))
        show2("dis_update: decoder ResBlocks fwd 1 pass", Policy(f3, {**base, **{n: f1 for n in dec_res}}))
    elif exp == "genimg":
        # generation only (no gradients): the bar is the image, 2e-2 max-abs
        dec_up = ["gen.dec.model.%d.conv." % i for i in (2, 4, 6)]
        dec_res = ["gen.dec.model.0.model.%d.model.%d.conv." % (i, j) for i in (0, 1) for j in (0, 1)]
        vgg = ["gen.enc_image.model.features.%d." % i for i in (0, 3, 6, 9, 13, 16, 19, 22, 26, 29, 32, 35, 39, 42, 45, 48)]
        f3, f1 = ("f16", "f16", "f16", 3, 1, 1), ("f16", "f16", "f16", 1, 1, 1)
        base = {n: f1 for n in dec_up}
        show("shipping", Policy(f3, base))
        for k in (12, 8, 4, 2, 0):
            show("VGG[%d:] + decoder ResBlocks fwd 1 pass" % k, Policy(f3, {**base, **{n: f1 for n in vgg[k:] + dec_res}}))
        show("every generator layer fwd 1 pass", Policy(f3, {"gen.": f1}))
    elif exp == "fp8corr":
        dec_up = ["gen.dec.model.%d.conv." % i for i in (2, 4, 6)]
        vgg = ["gen.enc_image.model.features.%d." % i for i in (0, 3, 6, 9, 13, 16, 19, 22, 26, 29, 32, 35, 39, 42, 45, 48)]
        f3, f1 = ("f16", "f16", "f16", 3, 1, 1), ("f16", "f16", "f16", 1, 1, 1)
        c8, c8f = ("f16", "f16", "f16", "f8c", 1, 1), ("f16", "f16", "f16", "f8cfix", 1, 1)
        base = {n: f1 for n in dec_up}
        show("shipping f16: 3/1/1, ups 1", Policy(f3, base))
        show("  all other fwd: f16 + 2 fp8 corrections (amax)", Policy(c8, base))
        show("  all other fwd: f16 + 2 fp8 corrections (fixed)", Policy(c8f, base))
        show("  VGG only fp8 corrections (amax)", Policy(f3, {**base, **{n: c8 for n in vgg}}))
        show("  VGG only fp8 corrections (fixed)", Policy(f3, {**base, **{n: c8f for n in vgg}}))
    elif exp == "plans5":
        dec_up = ["gen.dec.model.%d.conv." % i for i in (2, 4, 6)]
        dec_res = ["gen.dec.model.0.model.%d.model.%d.conv." % (i, j) for i in (0, 1) for j in (0, 1)]
        vgg = ["gen.enc_image.model.features.%d." % i for i in (0, 3, 6, 9, 13, 16, 19, 22, 26, 29, 32, 35, 39, 42, 45, 48)]
        f3, f1 = ("f16", "f16", "f16", 3, 1, 1), ("f16", "f16", "f16", 1, 1, 1)
        xw = ("f16", "f16", "f16", (3, 1), 1, 1)      # x split, w single plane: 2 MMAs
        wx = ("f16", "f16", "f16", (1, 3), 1, 1)      # w split, x single plane: 2 MMAs
        base = {n: f1 for n in dec_up}
        show("shipping f16: 3/1/1, ups 1", Policy(f3, base))
        show("  + all VGG fwd 2 MMAs (x split)", Policy(f3, {**base, **{n: xw for n in vgg}}))
        show("  + all VGG fwd 2 MMAs (w split)", Policy(f3, {**base, **{n: wx for n in vgg}}))
        show("  + VGG[8:] fwd 2 MMAs (x split)", Policy(f3, {**base, **{n: xw for n in vgg[8:]}}))
        show("  + dec res fwd 2 MMAs (x split)", Policy(f3, {**base, **{n: xw for n in dec_res}}))
        show("  + everything but ups fwd 2 MMAs (x split)", Policy(xw, base))


if __name__ == "__main__":
    main()
