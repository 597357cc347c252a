"""Line-level generator (BASELINE.json configs[4]: `line_generation/generate.py` on 64 x 1024 lines):
drop-in for the reference's `line_generation/model/pure_gen.py`

    SpacedGenerator :12-50    StyledConvBlock :140-216    AdaptiveInstanceNorm :52-69    NoiseInjection :72-79
    Blur :123-136             FusedUpsample :250-279      EqualConv2d :281-291           PixelNorm :306-311

with the same constructor arguments, module tree and `state_dict` keys (`conv.N.*`, the `gen.*` alias, `*.weight_orig` of the
equal-learning-rate layers, the Blur buffers), every forward on libaffgw kernels: the transposed / up-sampling / 3x3 / 1x1
convolutions and the style linears on the tcgen05 kernels (`FusedUpsample`'s stride-2 transposed convolution runs as what it
is - the input gradient of a stride-2 convolution - on the dgrad kernels), instance norm + style affine, noise + LeakyReLU,
the depthwise blur and PixelNorm on streaming kernels.

GENERATION ONLY (what `generate.py:796-850` does under `torch.no_grad()`); training the line model is out of scope.  The
reference draws `torch.randn_like(out)` twice per block inside forward (pure_gen.py:199,205); `forward(..., noise=[...])`
takes the ten tensors instead (parity tests), otherwise they are drawn on the device.
"""
from math import sqrt

import torch
from torch import nn

from . import ops, rec_ops


class PixelNorm(nn.Module):
    def forward(self, x):
        return ops.pixel_norm(x)


class _EqualLRWeight(nn.Module):
    """Parameter container of the reference's `equal_lr` wrapper (pure_gen.py:218-247): the stored tensor is `weight_orig`, the
    weight used in forward is weight_orig * sqrt(2 / fan_in), fan_in = size(1) * numel(weight[0][0])."""

    def scaled(self):
        w = self.weight_orig
        return w * sqrt(2 / (w.size(1) * w[0][0].numel()))


class NoiseInjection(_EqualLRWeight):
    def __init__(self, channel):
        super().__init__()
        self.weight_orig = nn.Parameter(torch.ones(1, channel, 1, 1) * 0.01)

    def forward(self, image, noise):
        """-> LeakyReLU(image + weight * noise, 0.2): the activation that follows every injection (pure_gen.py:199-201) is fused"""
        n, c = image.shape[:2]
        scaled = rec_ops.scale_nc(ops.input_to_internal(noise), self.scaled().reshape(1, c).expand(n, c))
        return ops.add_act(image, scaled, "lrelu")


class AdaptiveInstanceNorm(nn.Module):
    def __init__(self, in_channel, style_dim):
        super().__init__()
        self.norm = nn.InstanceNorm2d(in_channel)
        self.style = nn.Linear(style_dim, in_channel * 2)
        self.style.bias.data[:in_channel] = 1
        self.style.bias.data[in_channel:] = 0

    def forward(self, x, style):
        c = x.shape[1]
        s = ops.linear(style, self.style.weight, self.style.bias)                # [b, 2 C]: gamma | beta
        return ops.instance_norm(x, gamma=s[:, :c].contiguous().reshape(-1), beta=s[:, c:].contiguous().reshape(-1),
                                 eps=self.norm.eps)


class Blur(nn.Module):
    def __init__(self, channel):
        super().__init__()
        weight = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32).view(1, 1, 3, 3)
        weight = weight / weight.sum()
        self.register_buffer("weight", weight.repeat(channel, 1, 1, 1))
        self.register_buffer("weight_flip", torch.flip(weight, [2, 3]).repeat(channel, 1, 1, 1))

    def forward(self, x):
        return ops.blur3(x)            # the fixed binomial kernel (the buffers are kept for the checkpoint keys only)


class FusedUpsample(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, padding=0, only_vertical=False):
        super().__init__()
        if only_vertical:
            raise NotImplementedError("FusedUpsample(only_vertical=True) is not used by SpacedGenerator (pure_gen.py:21-27)")
        self.stride = 2
        self.multiplier = sqrt(2 / (in_channel * kernel_size * kernel_size))
        self.weight = nn.Parameter(torch.randn(in_channel, out_channel, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.zeros(out_channel))
        self.pad = padding

    def forward(self, x):
        # pure_gen.py:268-276: the scaled kernel zero-padded by one and averaged over its four unit shifts (3x3 -> 4x4)
        w = torch.nn.functional.pad(self.weight * self.multiplier, [1, 1, 1, 1])
        w = (w[:, :, 1:, 1:] + w[:, :, :-1, 1:] + w[:, :, 1:, :-1] + w[:, :, :-1, :-1]) / 4
        return ops.conv_transpose2d(x, w.contiguous(), self.bias, stride=self.stride, pad=self.pad)


class EqualConv2d(nn.Module):
    class _Conv(_EqualLRWeight):
        def __init__(self, cin, cout, k):
            super().__init__()
            self.bias = nn.Parameter(torch.zeros(cout))                          # (registered before weight_orig, like equal_lr leaves it)
            self.weight_orig = nn.Parameter(torch.randn(cout, cin, k, k))

    def __init__(self, in_channel, out_channel, kernel_size):
        super().__init__()
        self.conv = EqualConv2d._Conv(in_channel, out_channel, kernel_size)

    def forward(self, x, post_act="none"):
        return ops.conv2d(x, self.conv.scaled(), self.conv.bias, post_act=post_act)


class StyledConvBlock(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size=3, padding=1, style_dim=512, initial=False, upsample=False,
                 only_vertical=False, fused=False):
        super().__init__()
        self.kind = "conv"
        if initial:
            self.conv1 = nn.ConvTranspose2d(in_channel, out_channel, (4, 3), padding=(0, 1))
            self.kind = "initial"
        elif upsample and fused:
            self.conv1 = nn.Sequential(FusedUpsample(in_channel, out_channel, kernel_size, padding=padding,
                                                     only_vertical=only_vertical), Blur(out_channel))
            self.kind = "fused"
        elif upsample:
            self.conv1 = nn.Sequential(nn.Upsample(scale_factor=(2, 1) if only_vertical else 2, mode="nearest"),
                                       nn.Conv2d(in_channel, out_channel, kernel_size, padding=padding), Blur(out_channel))
            self.kind = "vertical" if only_vertical else "up"
        else:
            self.conv1 = nn.Conv2d(in_channel, out_channel, kernel_size, padding=padding)
        self.pad = padding
        self.noise1 = NoiseInjection(out_channel)
        self.adain1 = AdaptiveInstanceNorm(out_channel, style_dim)
        self.lrelu1 = nn.LeakyReLU(0.2)
        self.conv2 = nn.Conv2d(out_channel, out_channel, kernel_size, padding=padding)
        self.noise2 = NoiseInjection(out_channel)
        self.adain2 = AdaptiveInstanceNorm(out_channel, style_dim)
        self.lrelu2 = nn.LeakyReLU(0.2)

    def _conv1(self, x):
        if self.kind == "initial":
            # ConvTranspose2d(kernel (4, 3), padding (0, 1)) on a one-row input [b, C, 1, T]: output row h, column t is
            # sum_kw x[:, t + 1 - kw] . W[:, :, h, kw] - one GEMM over the 3-tap windows, [b T, 3 C] x [3 C, 4 Cout]
            m = self.conv1
            b, c, one, t = x.shape
            assert one == 1
            xs = torch.nn.functional.pad(x.reshape(b, c, t), (1, 1))                           # b, C, T + 2
            win = torch.stack([xs[:, :, 2 - kw:2 - kw + t] for kw in range(3)], 1)             # b, kw, C, T  (x[t + 1 - kw])
            win = win.permute(0, 3, 1, 2).reshape(b * t, 3 * c).contiguous()
            co = m.weight.shape[1]
            wm = m.weight.permute(2, 1, 3, 0).reshape(4 * co, 3 * c).contiguous()              # (h, co) x (kw, ci)
            y = ops.linear(win, wm, m.bias.repeat(4))                                          # b T, 4 Cout
            return y.reshape(b, t, 4, co).permute(0, 3, 2, 1)                                  # NCHW view of [b, 4, T, Cout] storage
        if self.kind == "fused":
            return self.conv1[1](self.conv1[0](x))
        if self.kind in ("vertical", "up"):
            n, c, h, w = x.shape
            sh, sw = (2, 1) if self.kind == "vertical" else (2, 2)
            x = ops.resize_nearest(x, h * sh, w * sw)
            return self.conv1[2](ops.conv2d(x, self.conv1[1].weight, self.conv1[1].bias, pad=self.pad, pad_mode="zero"))
        return ops.conv2d(x, self.conv1.weight, self.conv1.bias, pad=self.pad, pad_mode="zero")

    def forward(self, inp, noise=None):
        x, style = inp
        out = ops.input_to_internal(self._conv1(ops.input_to_internal(x)))
        out = self.adain1(self.noise1(out, noise[0] if noise is not None else torch.randn_like(out)), style)
        out = ops.conv2d(out, self.conv2.weight, self.conv2.bias, pad=self.pad, pad_mode="zero")
        out = self.adain2(self.noise2(out, noise[1] if noise is not None else torch.randn_like(out)), style)
        return out, style


class SpacedGenerator(nn.Module):
    def __init__(self, n_class, style_size, dim=256, output_dim=1, n_style_trans=6, emb_dropout=False, append_style=False,
                 small=False):
        super().__init__()
        self.append_style = append_style
        in_ch = n_class + style_size if append_style else n_class
        self.conv = nn.Sequential(
            StyledConvBlock(in_ch, dim, upsample=False, style_dim=style_size, initial=True),
            StyledConvBlock(dim, dim // 2, upsample=True, only_vertical=True, fused=False, style_dim=style_size),
            StyledConvBlock(dim // 2, dim // 4, upsample=True, only_vertical=True, fused=False, style_dim=style_size),
            StyledConvBlock(dim // 4, dim // 8, upsample=True, only_vertical=False, fused=True, style_dim=style_size),
            StyledConvBlock(dim // 8, dim // 16, upsample=not small, only_vertical=False, fused=True, style_dim=style_size))
        self.out = nn.Sequential(EqualConv2d(dim // 16, output_dim, 1), nn.Tanh())
        layers = [PixelNorm()]
        drop = emb_dropout if type(emb_dropout) is float else 0.5
        for i in range(n_style_trans):
            layers.append(nn.Linear(style_size, style_size))
            if emb_dropout and i < n_style_trans - 1:
                layers.append(nn.Dropout(drop, True))
            layers.append(nn.LeakyReLU(0.2, True))
        self.style_emb = nn.Sequential(*layers)
        self.gen = self.conv

    @torch.no_grad()
    def forward(self, content, style, return_intermediate=False, noise=None):
        """content [T, b, n_class], style [b, style_size] -> image [b, output_dim, 64, 4 T]   (pure_gen.py:42-50)"""
        x = content.permute(1, 2, 0)
        x = x.reshape(x.size(0), x.size(1), 1, x.size(2)).float()
        s = None
        for m in self.style_emb:
            if isinstance(m, PixelNorm):
                s = m(style)
            elif isinstance(m, nn.Linear):
                s = ops.linear(s, m.weight, m.bias, post_act="lrelu")           # every Linear is followed by LeakyReLU(0.2)
            elif isinstance(m, nn.Dropout) and self.training:
                raise NotImplementedError("style_emb dropout is a training-time feature; generation runs under eval()")
        if self.append_style:
            x = torch.cat((x, s[:, :, None, None].expand(-1, -1, 1, x.size(3))), dim=1)
        for i, blk in enumerate(self.conv):
            x, _ = blk((x, s), None if noise is None else noise[2 * i:2 * i + 2])
        return self.out[0](x, post_act="tanh")
