"""TEST INFRASTRUCTURE - CPU restatement of the reference's line-level generator (BASELINE.json configs[4]):

    SpacedGenerator.forward   line_generation/model/pure_gen.py:12-50
    StyledConvBlock           pure_gen.py:140-216     AdaptiveInstanceNorm :52-69     NoiseInjection :72-79
    Blur                      pure_gen.py:123-136     FusedUpsample :250-279          EqualLR :218-247     PixelNorm :306-311

Functional over a state_dict with the reference's keys (`conv.N.*`; `gen.*` aliases the same tensors, pure_gen.py:40).  The
reference draws `torch.randn_like(out)` twice per block inside forward (:199,:205): here the ten noise tensors are an argument.
Pinned against the unmodified reference by oracle/make_golden_linegen.py (tests/golden/linegen.npz, tests/test_oracle_golden.py)."""
from math import sqrt

import torch
import torch.nn.functional as F

BLUR = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]) / 16.


def blur(x):
    c = x.shape[1]
    return F.conv2d(x, BLUR.view(1, 1, 3, 3).repeat(c, 1, 1, 1).to(x.dtype), padding=1, groups=c)


def fused_upsample(x, w, b, pad=1):
    """pure_gen.py:268-279: the 3x3 kernel (x sqrt(2 / fan_in)) zero-padded to 5x5 and averaged over its four unit shifts -> 4x4,
    then a stride-2 transposed convolution."""
    cin, cout, k, _ = w.shape
    wp = F.pad(w * sqrt(2 / (cin * k * k)), [1, 1, 1, 1])
    w4 = (wp[:, :, 1:, 1:] + wp[:, :, :-1, 1:] + wp[:, :, 1:, :-1] + wp[:, :, :-1, :-1]) / 4
    return F.conv_transpose2d(x, w4, b, stride=2, padding=pad)


def adain(x, style, sd, p):
    s = F.linear(style, sd[p + "style.weight"], sd[p + "style.bias"])
    gamma, beta = s[:, :, None, None].chunk(2, 1)
    return gamma * F.instance_norm(x, eps=1e-5) + beta


def noise_inject(x, noise, sd, p):
    w = sd[p + "weight_orig"]
    return x + w * sqrt(2 / (w.shape[1] * w[0][0].numel())) * noise            # equal_lr: fan_in = size(1) * numel(w[0][0])


def styled_block(x, style, sd, p, kind, noises):
    if kind == "initial":
        x = F.conv_transpose2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=(0, 1))
    elif kind == "vertical":
        x = F.interpolate(x, scale_factor=(2, 1), mode="nearest")
        x = blur(F.conv2d(x, sd[p + "conv1.1.weight"], sd[p + "conv1.1.bias"], padding=1))
    else:
        x = blur(fused_upsample(x, sd[p + "conv1.0.weight"], sd[p + "conv1.0.bias"]))
    x = adain(F.leaky_relu(noise_inject(x, noises[0], sd, p + "noise1."), 0.2), style, sd, p + "adain1.")
    x = F.conv2d(x, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    return adain(F.leaky_relu(noise_inject(x, noises[1], sd, p + "noise2."), 0.2), style, sd, p + "adain2.")


KINDS = ("initial", "vertical", "vertical", "fused", "fused")


def noise_shapes(batch, T, dim=256):
    """shapes of the ten noise tensors for a [T, batch, n_class] content"""
    out, h, w, c = [], 4, T, dim
    for i, kind in enumerate(KINDS):
        if kind == "vertical":
            h, c = h * 2, c // 2
        elif kind == "fused":
            h, w, c = h * 2, w * 2, c // 2
        out += [(batch, c, h, w)] * 2
    return out


def spaced_generator(content, style, sd, noises, append_style=True, n_style_trans=6):
    """content [T, b, n_class], style [b, style_size] -> image [b, 1, 64, 4 T]   (pure_gen.py:42-50)"""
    x = content.permute(1, 2, 0)
    x = x.reshape(x.size(0), x.size(1), 1, x.size(2))
    s = style / torch.sqrt(torch.mean(style ** 2, dim=1, keepdim=True) + 1e-8)          # PixelNorm
    for i in range(n_style_trans):
        s = F.leaky_relu(F.linear(s, sd[f"style_emb.{1 + 2 * i}.weight"], sd[f"style_emb.{1 + 2 * i}.bias"]), 0.2)
    if append_style:
        x = torch.cat((x, s[:, :, None, None].expand(-1, -1, 1, x.size(3))), dim=1)
    for i, kind in enumerate(KINDS):
        x = styled_block(x, s, sd, f"conv.{i}.", kind, noises[2 * i:2 * i + 2])
    w = sd["out.0.conv.weight_orig"]
    return torch.tanh(F.conv2d(x, w * sqrt(2 / (w.shape[1] * w[0][0].numel())), sd["out.0.conv.bias"]))


def synthetic_inputs(batch, T, n_class=80, style_size=128, seed=77):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, n_class, (T, batch), generator=g)
    content = torch.zeros(T, batch, n_class)
    content.scatter_(2, idx.unsqueeze(2), 1.0)
    style = torch.randn(batch, style_size, generator=g)
    noises = [torch.randn(s, generator=g) for s in noise_shapes(batch, T)]
    return content, style, noises
