"""CPU oracle: a functional, state-dict-driven restatement of the reference hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  affganwriting_b200/ never does.

Parity pinning: the reference ships no tests or golden vectors for this path, so the
pins are outputs of the reference itself, run in the build container by oracle/make_golden.py and
committed under tests/golden/ (tests/test_oracle_golden.py checks this file against them).

Everything is fp32 torch on CPU, written as pure functions over a flat {key: tensor} state dict with the
reference's checkpoint key names.  Because it is composed of differentiable torch ops, reference
gradients come from autograd over this restatement.  Each function cites the reference lines it follows
(paths relative to /root/reference/GAN_word/).
"""
import string

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# constants and label handling (load_data.py:11-19, 31-40, 169-179)
# ----------------------------------------------------------------------------------------------
IMG_HEIGHT = 64
IMG_WIDTH = 216
MAX_CHARS = 10
OUTPUT_MAX_LEN = MAX_CHARS + 2
NUM_WRITERS = 500
TOKENS = {"GO_TOKEN": 0, "END_TOKEN": 1, "PAD_TOKEN": 2}
NUM_TOKENS = 3
_LETTERS = string.ascii_lowercase + string.ascii_uppercase
LETTER2INDEX = {c: i for i, c in enumerate(_LETTERS)}
VOCAB_SIZE = len(_LETTERS) + NUM_TOKENS  # 55


def label_padding(word, output_max_len=OUTPUT_MAX_LEN):
    """load_data.py:169-179 — GO + (letter index + 3) + END, PAD to output_max_len."""
    ids = [LETTER2INDEX[c] + NUM_TOKENS for c in word]
    ids = [TOKENS["GO_TOKEN"]] + ids + [TOKENS["END_TOKEN"]]
    ids += [TOKENS["PAD_TOKEN"]] * (output_max_len - len(ids))
    return ids


def text_column_map(width, ts=OUTPUT_MAX_LEN):
    """modules_tro.py:295-313 — which token feeds feature-map column c (-1 = PAD embedding)."""
    reps = max(1, width // ts)
    cols = []
    for i in range(ts):
        cols += [i] * reps
    cols += [-1] * (width % ts)
    return cols


# ----------------------------------------------------------------------------------------------
# storage-precision model.  The product's bf16 mode stores every activation, and feeds every convolution / linear
# operand, in bfloat16 (fp32 accumulation and fp32 statistics); `storage_model("bf16")` makes the oracle round at exactly
# those points (straight-through in backward) so that a bf16 run can be checked tightly, while the distance between the
# two oracle modes measures what bf16 storage itself costs on this network.
# ----------------------------------------------------------------------------------------------
_MODEL = {"bf16": False}


class storage_model:
    def __init__(self, kind):
        assert kind in ("fp32", "bf16")
        self.kind = kind

    def __enter__(self):
        self.prev = _MODEL["bf16"]
        _MODEL["bf16"] = self.kind == "bf16"

    def __exit__(self, *a):
        _MODEL["bf16"] = self.prev


def q(t):
    """Round to the storage dtype of the active model (identity in fp32), gradient passes straight through."""
    if not _MODEL["bf16"] or not t.is_floating_point():
        return t
    return t + (t.detach().bfloat16().float() - t.detach())


def _conv(x, w, b=None, **kw):
    return F.conv2d(x, q(w), b, **kw)


def _linear(x, w, b=None):
    return F.linear(x, q(w), b)


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _act(x, kind):
    if kind == "relu":
        return torch.relu(x)
    if kind == "lrelu":
        return F.leaky_relu(x, 0.2)
    if kind == "tanh":
        return torch.tanh(x)
    assert kind == "none", kind
    return x


def _pad(x, p, kind):
    if p == 0:
        return x
    if kind == "reflect":
        return F.pad(x, (p, p, p, p), mode="reflect")
    if kind == "replicate":
        return F.pad(x, (p, p, p, p), mode="replicate")
    assert kind == "zero", kind
    return F.pad(x, (p, p, p, p))


def instance_norm(x, eps=1e-5):
    """nn.InstanceNorm2d(affine=False): biased variance over H*W per (n, c)."""
    m = x.mean(dim=(2, 3), keepdim=True)
    v = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - m) / torch.sqrt(v + eps)


def batch_norm(x, sd, prefix, training, stats=None, momentum=0.1, eps=1e-5):
    """nn.BatchNorm{1,2}d.  In training mode uses batch statistics (biased var) and, if `stats` is a dict,
    records the running-stat update the reference would apply (unbiased var, momentum 0.1)."""
    dims = [d for d in range(x.dim()) if d != 1]
    shape = [1, -1] + [1] * (x.dim() - 2)
    w, b = sd[prefix + "weight"], sd[prefix + "bias"]
    if training:
        n = x.numel() // x.shape[1]
        if n <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
        m = x.mean(dim=dims)
        v = x.var(dim=dims, unbiased=False)
        if stats is not None:
            rm = stats.get(prefix + "running_mean", sd[prefix + "running_mean"])
            rv = stats.get(prefix + "running_var", sd[prefix + "running_var"])
            nb = stats.get(prefix + "num_batches_tracked", sd[prefix + "num_batches_tracked"])
            stats[prefix + "running_mean"] = (1 - momentum) * rm + momentum * m.detach()
            stats[prefix + "running_var"] = (1 - momentum) * rv + momentum * v.detach() * n / (n - 1)
            stats[prefix + "num_batches_tracked"] = nb + 1
    else:
        m, v = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    return (x - m.view(shape)) / torch.sqrt(v.view(shape) + eps) * w.view(shape) + b.view(shape)


# ----------------------------------------------------------------------------------------------
# blocks.py
# ----------------------------------------------------------------------------------------------
def conv2d_block(x, sd, prefix, ks, st, padding=0, norm="none", activation="relu", pad_type="zero",
                 activation_first=False, adain=None, residual=None, addend=None, store=True):
    """blocks.py:150-163.  `adain` = dict(weight, bias, input, training, stats) when norm == 'adain'."""
    w = sd[prefix + "conv.weight"]
    b = sd.get(prefix + "conv.bias")
    if activation_first:
        x = _act(x, activation)
    x = _conv(_pad(x, padding, pad_type), w, b, stride=st)
    if addend is not None:
        x = x + addend
    post = "none" if activation_first else activation
    if norm == "in":
        x = q(_act(instance_norm(q(x)), post))
    elif norm == "adain":
        x = adaptive_instance_norm(q(x), sd, prefix + "norm.", act=post, residual=residual, **adain)
        residual = None
    else:
        assert norm == "none", norm
        x = _act(x, post)
        x = q(x) if store else x
    if residual is not None:
        x = q(x + residual)
    return x


def calc_mean_std(feat, eps=1e-5):
    """blocks.py:227-235 — unbiased variance, eps added before the square root."""
    n, c = feat.shape[:2]
    var = feat.reshape(n, c, -1).var(dim=2) + eps
    return feat.reshape(n, c, -1).mean(dim=2).view(n, c, 1, 1), var.sqrt().view(n, c, 1, 1)


def get_key(feats, feat):
    """blocks.py:210-223 — nearest resize of the style map to feats' (h, w), then mean/variance norm."""
    h, w = feats.shape[2:]
    r = F.interpolate(feat, (h, w))
    m, s = calc_mean_std(r)
    return q((r - m) / s)


def _att_branch(x, sd, prefix, training, stats, glob):
    """blocks.py:246-281 — (GAP) conv1x1 C->C/4, BN, ReLU, conv1x1 C/4->C, BN."""
    o = 1 if glob else 0
    if glob:
        x = q(x.mean(dim=(2, 3), keepdim=True))
    x = q(_conv(x, sd[f"{prefix}{o}.weight"], sd[f"{prefix}{o}.bias"]))
    x = q(torch.relu(batch_norm(x, sd, f"{prefix}{o + 1}.", training, stats)))
    x = q(_conv(x, sd[f"{prefix}{o + 3}.weight"], sd[f"{prefix}{o + 3}.bias"]))
    return q(batch_norm(x, sd, f"{prefix}{o + 4}.", training, stats))


def iaff(x, residual, sd, prefix, training=True, stats=None):
    """blocks.py:286-299.  Round two reuses global_att, not global_att2 (line 295)."""
    xa = q(x + residual)
    xl = _att_branch(xa, sd, prefix + "local_att.", training, stats, False)
    xg = _att_branch(xa, sd, prefix + "global_att.", training, stats, True)
    wei = torch.sigmoid(xl + xg)
    xi = q(x * wei + residual * (1 - wei))
    xl2 = _att_branch(xi, sd, prefix + "local_att2.", training, stats, False)
    xg2 = _att_branch(xi, sd, prefix + "global_att.", training, stats, True)
    wei2 = torch.sigmoid(xl2 + xg2)
    return q(x * wei2 + residual * (1 - wei2))


def adaptive_instance_norm(x, sd, prefix, weight, bias, input=None, training=True, stats=None, eps=1e-5, act="none",
                           residual=None):
    """blocks.py:188-204.  Instance statistics always (F.batch_norm(..., training=True)), biased variance;
    weight/bias are per-(n, c) activations of length B*C."""
    if input is not None:
        x = iaff(x, get_key(x, input), sd, prefix + "iAff.", training, stats)
    b, c = x.shape[:2]
    y = _act(instance_norm(x, eps) * weight.view(b, c, 1, 1) + bias.view(b, c, 1, 1), act)
    return q(y if residual is None else y + residual)


def res_block(x, sd, prefix, norm, activation, pad_type, adain0=None, adain1=None):
    """blocks.py:21-39."""
    y = conv2d_block(x, sd, prefix + "model.0.", 3, 1, 1, norm, activation, pad_type, adain=adain0)
    return conv2d_block(y, sd, prefix + "model.1.", 3, 1, 1, norm, "none", pad_type, adain=adain1, residual=x)


def act_first_res_block(x, sd, prefix, fin, fout):
    """blocks.py:42-65 with activation='lrelu', norm='none'."""
    xs = x
    if fin != fout:
        xs = conv2d_block(x, sd, prefix + "conv_s.", 1, 1, activation="none")
    dx = conv2d_block(x, sd, prefix + "conv_0.", 3, 1, 1, "none", "lrelu", "reflect", activation_first=True)
    return conv2d_block(dx, sd, prefix + "conv_1.", 3, 1, 1, "none", "lrelu", "reflect", activation_first=True, addend=xs)


def linear_block(x, sd, prefix, norm="none", activation="relu", training=True, stats=None):
    """blocks.py:68-103."""
    x = _linear(x, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])
    if norm == "bn":
        x = batch_norm(q(x), sd, prefix + "norm.", training, stats)
    elif norm == "in":
        x = F.instance_norm(x.unsqueeze(0)).squeeze(0) if x.dim() == 2 else F.instance_norm(x)
    return q(_act(x, activation))


def mlp(x, sd, prefix, n_blk=3, norm="none", activ="relu", training=True, stats=None):
    """modules_tro.py:684-697."""
    x = x.reshape(x.shape[0], -1)
    for i in range(n_blk - 1):
        x = linear_block(x, sd, f"{prefix}model.{i}.", norm, activ, training, stats)
    return linear_block(x, sd, f"{prefix}model.{n_blk - 1}.", "none", "none")


# ----------------------------------------------------------------------------------------------
# style encoder: VGG-19 'E' variant with InstanceNorm (vgg_tro_channel3_modi.py:40-67, modules_tro.py:331-375)
# ----------------------------------------------------------------------------------------------
VGG_CFG = [64, 64, 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512]
VGG_SLICE_ENDS = (3, 9, 16, 29, 42, 51)


def vgg_layers():
    """[(kind, features-index, out_channels)] in nn.Sequential order: conv, in, relu / pool."""
    layers = []
    for v in VGG_CFG:
        if v == "M":
            layers.append(("pool", len(layers), None))
        else:
            layers.append(("conv", len(layers), v))
            layers.append(("in", len(layers), v))
            layers.append(("relu", len(layers), v))
    return layers


def image_encoder(x, sd, prefix="enc_image."):
    """modules_tro.py:360-374 — six intermediate maps at the slice ends 3/9/16/29/42/end."""
    outs = []
    x = q(x)
    for kind, idx, _ in vgg_layers():
        if kind == "conv":
            x = q(_conv(x, sd[f"{prefix}model.features.{idx}.weight"], sd[f"{prefix}model.features.{idx}.bias"],
                        padding=1))
        elif kind == "in":
            x = instance_norm(x)
        elif kind == "relu":
            x = q(torch.relu(x))
        else:
            x = F.max_pool2d(x, 2, 2)
        if idx + 1 in VGG_SLICE_ENDS:
            outs.append(x)
    return outs


# ----------------------------------------------------------------------------------------------
# text encoder, mix, decoder, generator (modules_tro.py:208-317, 586-607)
# ----------------------------------------------------------------------------------------------
RESNET_PLAN = {"resnet18": ("basic", (2, 2, 2, 2), (64, 64, 128, 256, 512)),
               "resnet50": ("bottleneck", (3, 4, 6, 3), (64, 256, 512, 1024, 2048))}


def resnet_encoder(x, sd, prefix="enc_image.", arch="resnet50", training=True, stats=None):
    """modules_tro.py:464-533 (ResNet-50, the encoder active in the reference's GenModel_FC) and modules_tro2.py:447-516
    (ResNet-18): torchvision ResNet trunk - conv7x7/s2 + BN + ReLU (feat1), MaxPool2d(3, 2, 1), four stages of
    BasicBlock / Bottleneck (feat2..feat5; stride on conv1 of a BasicBlock, on conv2 of a Bottleneck) - then 1x1
    `reduce_layers` to 512 channels and a bilinear (align_corners=False) resize of the last map to (8, 27)."""
    kind, depths, _ = RESNET_PLAN[arch]
    m = prefix + "model."

    def bn(t, name, relu=False):
        t = batch_norm(t, sd, name, training, stats)
        return q(torch.relu(t)) if relu else q(t)

    x = q(x)
    x = bn(q(_conv(x, sd[m + "conv1.weight"], None, stride=2, padding=3)), m + "bn1.", relu=True)
    feats = [x]
    x = F.max_pool2d(x, 3, 2, 1)
    for li, depth in enumerate(depths, start=1):
        for bi in range(depth):
            b = f"{m}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            identity = x
            if kind == "basic":
                out = bn(q(_conv(x, sd[b + "conv1.weight"], None, stride=stride, padding=1)), b + "bn1.", relu=True)
                out = bn(q(_conv(out, sd[b + "conv2.weight"], None, padding=1)), b + "bn2.")
            else:
                out = bn(q(_conv(x, sd[b + "conv1.weight"], None)), b + "bn1.", relu=True)
                out = bn(q(_conv(out, sd[b + "conv2.weight"], None, stride=stride, padding=1)), b + "bn2.", relu=True)
                out = bn(q(_conv(out, sd[b + "conv3.weight"], None)), b + "bn3.")
            if b + "downsample.0.weight" in sd:
                identity = bn(q(_conv(x, sd[b + "downsample.0.weight"], None, stride=stride)), b + "downsample.1.")
            x = q(torch.relu(out + identity))
        feats.append(x)
    results = [q(_conv(f, sd[f"{prefix}reduce_layers.{i}.weight"], sd[f"{prefix}reduce_layers.{i}.bias"]))
               for i, f in enumerate(feats)]
    results[-1] = F.interpolate(results[-1], size=(8, 27), mode="bilinear", align_corners=False)
    return results


def resnet18_standalone(x, sd, prefix="", training=True, stats=None):
    """Resnet18.py:38-88 (`ResNet18`, `BasicBlock` :9-36): conv3x3 stride (2, 1) + BN + ReLU, MaxPool2d(3, (2, 1), 1) -> map 0;
    three stages of two BasicBlocks (first block of each stage: stride 2 on conv1 and a conv1x1/s2 + BN shortcut) -> maps
    1..3; MaxPool2d(3, 1, 1) of the last -> map 4.  No biases, BatchNorm eps 1e-5, widths nb/4, nb/4, nb/2, nb."""
    def bn(t, name, relu=False):
        t = batch_norm(t, sd, name, training, stats)
        return q(torch.relu(t)) if relu else q(t)

    x = q(x)
    x = bn(q(_conv(x, sd[prefix + "conv1.weight"], None, stride=(2, 1), padding=1)), prefix + "bn1.", relu=True)
    x = F.max_pool2d(x, 3, (2, 1), 1)
    results = [x]
    for li in (1, 2, 3):
        for bi in range(2):
            b = f"{prefix}layer{li}.{bi}."
            stride = 2 if bi == 0 else 1
            out = bn(q(_conv(x, sd[b + "conv1.weight"], None, stride=stride, padding=1)), b + "bn1.", relu=True)
            out = bn(q(_conv(out, sd[b + "conv2.weight"], None, padding=1)), b + "bn2.")
            identity = x
            if b + "downsample.0.weight" in sd:
                identity = bn(q(_conv(x, sd[b + "downsample.0.weight"], None, stride=stride)), b + "downsample.1.")
            x = q(torch.relu(out + identity))
        results.append(x)
    results.append(F.max_pool2d(x, 3, 1, 1))
    return results


def text_encoder(label, f_xs_shape, sd, prefix="enc_text.", training=True, stats=None):
    """modules_tro.py:285-317 -> (adain params [B,4096], content map [B,512,h,w])."""
    emb = q(sd[prefix + "embed.weight"][label])                        # b, t, 64
    b, ts = emb.shape[:2]
    h = q(_linear(emb.reshape(b, -1), sd[prefix + "fc.0.weight"], sd[prefix + "fc.0.bias"]))
    h = q(torch.relu(batch_norm(h, sd, prefix + "fc.1.", training, stats)))
    h = q(_linear(h, sd[prefix + "fc.3.weight"], sd[prefix + "fc.3.bias"]))
    h = q(torch.relu(batch_norm(h, sd, prefix + "fc.4.", training, stats)))
    out = _linear(h, sd[prefix + "fc.6.weight"], sd[prefix + "fc.6.bias"])          # AdaIN parameters stay fp32
    chars = q(_linear(emb, sd[prefix + "linear.weight"], sd[prefix + "linear.bias"]))  # b, t, 512
    pad = q(_linear(q(sd[prefix + "embed.weight"][TOKENS["PAD_TOKEN"]]), sd[prefix + "linear.weight"],
                    sd[prefix + "linear.bias"]))
    cols = [chars[:, c] if c >= 0 else pad.unsqueeze(0).expand(b, -1) for c in text_column_map(f_xs_shape[-1], ts)]
    row = torch.stack(cols, dim=2)                                     # b, 512, w
    return out, row.unsqueeze(2).expand(-1, -1, f_xs_shape[-2], -1).contiguous()


def mix(results, feat_embed, sd, prefix=""):
    """modules_tro.py:252-259 — per-pixel Linear(1024, 512) over cat(style, content)."""
    f = torch.cat([results[-1], feat_embed], dim=1).permute(0, 2, 3, 1)
    return q(_linear(f, sd[prefix + "linear_mix.weight"], sd[prefix + "linear_mix.bias"])).permute(0, 3, 1, 2)


def decoder(content, results, adain_params, sd, prefix="dec.", training=True, stats=None, max_pooled3=None):
    """modules_tro.py:226-249 (parameter assignment) + 586-607 (layers)."""
    x = content
    style3 = F.max_pool2d(results[3], 2, 2) if max_pooled3 is None else max_pooled3
    inputs = {1: style3, 3: results[4]}
    k = 0
    for blk in range(2):
        ad = []
        for _ in range(2):
            ad.append(dict(bias=adain_params[:, 1024 * k:1024 * k + 512].contiguous().view(-1),
                           weight=adain_params[:, 1024 * k + 512:1024 * k + 1024].contiguous().view(-1),
                           input=inputs.get(k), training=training, stats=stats))
            k += 1
        x = res_block(x, sd, f"{prefix}model.0.model.{blk}.", "adain", "relu", "reflect", ad[0], ad[1])
    for i, idx in enumerate((2, 4, 6)):
        x = F.interpolate(x, scale_factor=2)
        x = conv2d_block(x, sd, f"{prefix}model.{idx}.", 5, 1, 2, "in", "relu", "reflect")
    return conv2d_block(x, sd, f"{prefix}model.7.", 7, 1, 3, "none", "tanh", "reflect", store=False)   # fp32 image


def gen_forward(tr_img, label, sd, prefix="", training=True, stats=None, results=None):
    """network_tro.py:60-66: enc_image -> enc_text -> mix -> decode."""
    if results is None:
        results = image_encoder(tr_img, sd, prefix + "enc_image.")
    f_xt, f_embed = text_encoder(label, results[-1].shape, sd, prefix + "enc_text.", training, stats)
    f_mix = mix(results, f_embed, sd, prefix)
    return decoder(f_mix, results, f_xt, sd, prefix + "dec.", training, stats)


# ----------------------------------------------------------------------------------------------
# discriminator / writer classifier (modules_tro.py:119-201)
# ----------------------------------------------------------------------------------------------
def dis_cla_plan(n_layers=6):
    """[(kind, Sequential index, fin, fout)] of cnn_f after the stem."""
    plan, nf, idx = [], 16, 1
    for _ in range(n_layers - 1):
        nf_out = min(nf * 2, 1024)
        plan += [("res", idx, nf, nf), ("res", idx + 1, nf, nf_out), ("pool", idx + 2, None, None)]
        idx += 4
        nf = nf_out
    nf_out = min(nf * 2, 1024)
    plan += [("res", idx, nf, nf), ("res", idx + 1, nf, nf_out)]
    return plan


def dis_features(x, sd, prefix=""):
    x = conv2d_block(q(x), sd, prefix + "cnn_f.0.", 7, 1, 3, "none", "none", "reflect")
    for kind, idx, fin, fout in dis_cla_plan():
        if kind == "res":
            x = act_first_res_block(x, sd, f"{prefix}cnn_f.{idx}.", fin, fout)
        else:
            x = q(F.avg_pool2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), 3, 2))
    # head: kernel IMG_HEIGHT//32 = 2, stride IMG_WIDTH//32+1 = 7, lrelu first (modules_tro.py:139-142); fp32 logits
    x = conv2d_block(x, sd, prefix + "cnn_c.0.", 2, 7, 0, "none", "lrelu", "zero", activation_first=True, store=False)
    return x.squeeze(-1).squeeze(-1)


def dis_forward(x, sd, prefix=""):
    return dis_features(x, sd, prefix)


def dis_loss(x, sd, prefix="", target=1.0):
    """modules_tro.py:152-168 — BCE-with-logits against all-ones (real / gen) or all-zeros (fake)."""
    r = dis_features(x, sd, prefix)
    return F.binary_cross_entropy_with_logits(r, torch.full_like(r, target))


def cla_loss(x, y, sd, prefix=""):
    """modules_tro.py:195-201."""
    return F.cross_entropy(dis_features(x, sd, prefix), y)


# ----------------------------------------------------------------------------------------------
# step composition without the recogniser (network_tro.py:50-138, loss weights :10-13)
# ----------------------------------------------------------------------------------------------
def cla_update(batch, sd):
    return cla_loss(batch["tr_img"][:, 0:1], batch["tr_wid"], _sub(sd, "cla."))


def dis_update(batch, sd, stats=None):
    d = _sub(sd, "dis.")
    l_real = (dis_loss(batch["tr_img"][:, 0:1], d, target=1.0) + dis_loss(batch["tr_img"][:, 1:2], d, target=1.0)) / 2
    with torch.no_grad():
        g = _sub(sd, "gen.")
        res = image_encoder(batch["tr_img"], g)
        xg = gen_forward(None, batch["label_xt"], g, stats=stats, results=res)
        xg_swap = gen_forward(None, batch["label_xt_swap"], g, stats=stats, results=res)
    l_fake = (dis_loss(xg, d, target=0.0) + dis_loss(xg_swap, d, target=0.0)) / 2
    return l_real, l_fake


def gen_update(batch, sd, stats=None):
    """l_total = l_dis + l_cla (w_l1 = 0; recogniser term out of scope)."""
    g, d, c = _sub(sd, "gen."), _sub(sd, "dis."), _sub(sd, "cla.")
    res = image_encoder(batch["tr_img"], g)
    xg = gen_forward(None, batch["label_xt"], g, stats=stats, results=res)
    xg_swap = gen_forward(None, batch["label_xt_swap"], g, stats=stats, results=res)
    l_dis = (dis_loss(xg, d, target=1.0) + dis_loss(xg_swap, d, target=1.0)) / 2
    l_cla = (cla_loss(xg, batch["tr_wid"], c) + cla_loss(xg_swap, batch["tr_wid"], c)) / 2
    return l_dis + l_cla, l_dis, l_cla, xg, xg_swap


# ----------------------------------------------------------------------------------------------
# synthetic inputs — shared by tests, smoke() and bench.py so every arm sees the same data
# ----------------------------------------------------------------------------------------------
def synthetic_batch(batch, num_channel=50, seed=1234):
    g = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(seed)

    def imgs(n, c):
        x = torch.rand(n, c, IMG_HEIGHT, IMG_WIDTH, generator=g) * 2 - 1
        widths = torch.randint(40, IMG_WIDTH + 1, (n, c), generator=g)
        cols = torch.arange(IMG_WIDTH).view(1, 1, 1, -1)
        return torch.where(cols >= widths.view(n, c, 1, 1), torch.full_like(x, -1.0), x)

    def words(n):
        out = []
        for _ in range(n):
            ln = rng.randint(1, MAX_CHARS + 1)
            out.append("".join(_LETTERS[i] for i in rng.randint(0, len(_LETTERS), ln)))
        return out

    w1, w2 = words(batch), words(batch)
    return dict(
        tr_img=imgs(batch, num_channel),
        img_xt=imgs(batch, 1),
        tr_wid=torch.randint(0, NUM_WRITERS, (batch,), generator=g),
        label_xt=torch.tensor([label_padding(w) for w in w1], dtype=torch.int64),
        label_xt_swap=torch.tensor([label_padding(w) for w in w2], dtype=torch.int64),
        words=w1, words_swap=w2,
    )


# ----------------------------------------------------------------------------------------------
# uint8 wire format of the images
# ----------------------------------------------------------------------------------------------
def normalize_resized_u8(img_u8, height=64, width=216):
    """load_data.py:152-166, the arithmetic of `read_image_single` AFTER cv2.resize (:151): `img/255.` and `1. - img` in
    float64, crop to `width` columns or paste into a float32 zero canvas, `(canvas - 0.5) / 0.5` in float32.
    img_u8: numpy uint8 [height, w].  Returns (float32 [height, width], img_width)."""
    import numpy as np
    img = img_u8 / 255.
    img = 1. - img
    img_width = img.shape[-1]
    if img_width > width:
        out = img[:, :width]
        img_width = width
    else:
        out = np.zeros((height, width), dtype="float32")
        out[:, :img_width] = img
    out = out.astype("float32")
    return (out - 0.5) / 0.5, img_width


def pad_resized_u8(img_u8, height=64, width=216):
    """The same image as ONE uint8 canvas (the wire format): crop to `width` columns or right-pad with 255, the grey level
    whose normalisation is the canvas background (1 - 255/255 = 0 -> -1)."""
    import numpy as np
    out = np.full((height, width), 255, dtype=np.uint8)
    w = min(width, img_u8.shape[-1])
    out[:, :w] = img_u8[:, :w]
    return out


def decode_u8(u8):
    """Wire format -> normalised float32 (numpy, any shape): element-wise restatement of load_data.py:152-166."""
    import numpy as np
    canvas = (1. - u8 / 255.).astype("float32")
    return (canvas - np.float32(0.5)) / np.float32(0.5)
