"""Drop-in replacements for the reference's model assembly (reference GAN_word/modules_tro.py).

  DisModel :119-168   WriterClaModel :170-201   GenModel_FC :208-259   TextEncoder_FC :268-317
  ImageEncoder (VGG) :331-375   Decoder :586-607   MLP :684-697   get_num_adain_params :110-116
  RecModel :610-638 (affganwriting_b200.recognizer)

Same constructor signatures, attribute names, method names and state_dict keys; forwards run libaffgw kernels.
GenModel_FC takes the style encoder the reference selects by editing source (modules_tro.py:211-219) as an
explicit `encoder=` argument and defaults to the VGG ImageEncoder (the north-star path).
"""
import numpy as np
import torch
from torch import nn

from . import ops
from .blocks import ActFirstResBlock, Conv2dBlock, LinearBlock, ResBlocks
from .load_data import IMG_HEIGHT, IMG_WIDTH, OUTPUT_MAX_LEN, tokens, vocab_size
from .recognizer import RecModel  # noqa: F401  (modules_tro.py:610-638)
from .vgg_tro_channel3_modi import _pad64, vgg19_bn


def get_num_adain_params(model):
    n = 0
    for m in model.modules():
        if m.__class__.__name__ == "AdaptiveInstanceNorm2d":
            n += 2 * m.num_features
    return n


def _dis_cla_trunk(n_layers, final_dim):
    """modules_tro.py:121-144 / 172-192 (shared by DisModel and WriterClaModel)."""
    nf = 16
    cnn_f = [Conv2dBlock(1, nf, 7, 1, 3, pad_type="reflect", norm="none", activation="none")]
    for _ in range(n_layers - 1):
        nf_out = int(np.min([nf * 2, 1024]))
        cnn_f += [ActFirstResBlock(nf, nf, None, "lrelu", "none")]
        cnn_f += [ActFirstResBlock(nf, nf_out, None, "lrelu", "none")]
        cnn_f += [nn.ReflectionPad2d(1)]
        cnn_f += [nn.AvgPool2d(kernel_size=3, stride=2)]
        nf = int(np.min([nf * 2, 1024]))
    nf_out = int(np.min([nf * 2, 1024]))
    cnn_f += [ActFirstResBlock(nf, nf, None, "lrelu", "none")]
    cnn_f += [ActFirstResBlock(nf, nf_out, None, "lrelu", "none")]
    # kernel IMG_HEIGHT // 32 = 2, stride IMG_WIDTH // 32 + 1 = 7: positional (ks, st) quirk of the reference
    cnn_c = [Conv2dBlock(nf_out, final_dim, IMG_HEIGHT // (2 ** (n_layers - 1)), IMG_WIDTH // (2 ** (n_layers - 1)) + 1,
                         norm="none", activation="lrelu", activation_first=True)]
    return nn.Sequential(*cnn_f), nn.Sequential(*cnn_c)


import os as _os
_DECODER_F16 = _os.environ.get("AFFGW_DECODER_F16", "0") != "0"
_UPCONV_PASSES = int(_os.environ.get("AFFGW_UPCONV_PASSES", "1"))
_TRUNK_WGRAD_PASSES = int(_os.environ.get("AFFGW_TRUNK_WGRAD_PASSES", "1"))


def _run_trunk(cnn_f, cnn_c, x):
    # weight gradients of these trunks sum B*H*W strongly cancelling terms of a mean-reduced loss: single-pass operands cost them
    # ~3e-4 of per-tensor cosine (0.99956 instead of 0.99982 for the worst discriminator tensor at batch 8, measured on B200 and
    # predicted by scripts/precision_sweep.py) - still 2.3x inside the 0.999 bar, for 2 ms of the step.
    # AFFGW_TRUNK_WGRAD_PASSES=3 restores the split operands for these layers only.
    with ops.wgrad_passes(_TRUNK_WGRAD_PASSES):
        return _run_trunk_inner(cnn_f, cnn_c, x)


def _run_trunk_inner(cnn_f, cnn_c, x):
    x = ops.input_to_internal(x)
    i, n = 0, len(cnn_f)
    while i < n:
        m = cnn_f[i]
        if isinstance(m, nn.ReflectionPad2d):
            assert isinstance(cnn_f[i + 1], nn.AvgPool2d)
            x = ops.avg_pool3s2_reflect(x)
            i += 2
        else:
            x = m(x)
            i += 1
    out = cnn_c[0](x, out_dtype=torch.float32)     # logits in fp32 for the losses
    return out.squeeze(-1).squeeze(-1)


class DisModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.n_layers = 6
        self.final_size = 1024
        self.cnn_f, self.cnn_c = _dis_cla_trunk(self.n_layers, self.final_size)
        self.bce = nn.BCEWithLogitsLoss()

    def forward(self, x):
        return _run_trunk(self.cnn_f, self.cnn_c, x)

    def calc_dis_fake_loss(self, input_fake):
        return ops.bce_with_logits_const(self.forward(input_fake), 0.0)

    def calc_dis_real_loss(self, input_real):
        return ops.bce_with_logits_const(self.forward(input_real), 1.0)

    def calc_gen_loss(self, input_fake):
        return ops.bce_with_logits_const(self.forward(input_fake), 1.0)


class WriterClaModel(nn.Module):
    def __init__(self, num_writers):
        super().__init__()
        self.n_layers = 6
        self.cnn_f, self.cnn_c = _dis_cla_trunk(self.n_layers, num_writers)
        self.cross_entropy = nn.CrossEntropyLoss()

    def forward(self, x, y):
        return ops.cross_entropy(_run_trunk(self.cnn_f, self.cnn_c, x), y)


class TextEncoder_FC(nn.Module):
    def __init__(self, text_max_len):
        super().__init__()
        embed_size = 64
        self.embed = nn.Embedding(vocab_size, embed_size)
        self.fc = nn.Sequential(
            nn.Linear(text_max_len * embed_size, 1024), nn.BatchNorm1d(1024), nn.ReLU(inplace=False),
            nn.Linear(1024, 2048), nn.BatchNorm1d(2048), nn.ReLU(inplace=False),
            nn.Linear(2048, 4096))
        self.linear = nn.Linear(embed_size, 512)

    def forward(self, x, f_xs_shape):
        """labels [B, ts] int64 -> (AdaIN parameters [B, 4096] fp32, content map [B, 512, h, w])."""
        b, ts = x.shape
        h, w = int(f_xs_shape[-2]), int(f_xs_shape[-1])
        reps = max(1, w // ts)
        w_out = ts * reps + (w % ts)
        # one extra slot per sample carries the PAD token so that its embedding flows through `linear` like the
        # reference's embedded_padding_char (modules_tro.py:306-310); integer handling stays exact
        pad_col = torch.full((b, 1), tokens["PAD_TOKEN"], dtype=torch.int64, device=x.device)
        ids = torch.cat([x, pad_col], dim=1)
        emb = ops.embedding(ids, self.embed.weight)                       # b, ts+1, 64
        flat = emb.view(b, -1)[:, :ts * emb.shape[-1]]                    # b, ts*64 (row pitch (ts+1)*64)
        fc = self.fc
        hdn = ops.batch_norm(ops.linear(flat, fc[0].weight, fc[0].bias), fc[1], act="relu")
        hdn = ops.batch_norm(ops.linear(hdn, fc[3].weight, fc[3].bias), fc[4], act="relu")
        out = ops.linear(hdn, fc[6].weight, fc[6].bias, out_dtype=torch.float32)
        chars = ops.linear(emb, self.linear.weight, self.linear.bias)     # b, ts+1, 512
        return out, ops.text_tile(chars, h, w_out, reps)


class ImageEncoder(nn.Module):
    """VGG19-IN style encoder returning the six intermediate maps (reference modules_tro.py:331-375).
    The reference wraps the six slices in nn.DataParallel; here data parallelism is process-per-GPU
    (affganwriting_b200.parallel), so the slices are plain index ranges."""
    SLICES = ((0, 3), (3, 9), (9, 16), (16, 29), (29, 42), (42, None))

    def __init__(self):
        super().__init__()
        self.model = vgg19_bn(False)
        self.output_dim = 512

    def encode_with_intermediate(self, input_img):
        x = ops.input_to_internal(input_img, c_pad=_pad64(input_img.shape[1]))
        results = []
        n = len(self.model.features)
        for a, b in self.SLICES:
            x = self.model.run(x, a, n if b is None else b)
            results.append(x)
        return results

    def forward(self, x):
        return self.encode_with_intermediate(x)


class Decoder(nn.Module):
    def __init__(self, ups=3, n_res=2, dim=512, out_dim=1, res_norm="adain", activ="relu", pad_type="reflect"):
        super().__init__()
        model = [ResBlocks(n_res, dim, res_norm, activ, pad_type=pad_type)]
        for _ in range(ups):
            model += [nn.Upsample(scale_factor=2),
                      Conv2dBlock(dim, dim // 2, 5, 1, 2, norm="in", activation=activ, pad_type=pad_type)]
            dim //= 2
        model += [Conv2dBlock(dim, out_dim, 7, 1, 3, norm="none", activation="tanh", pad_type=pad_type)]
        self.model = nn.Sequential(*model)

    def forward(self, x):
        # AFFGW_DECODER_F16=1: in mode 'bf16', run the decoder's 3x3 / 5x5 convolutions on fp16 operand planes with one pass
        # per GEMM (ops.operand_format) - superseded by mode 'f16', kept for A/B measurements
        with ops.operand_format("f16" if _DECODER_F16 else None):
            return self._forward(x)

    def _forward(self, x):
        x = ops.input_to_internal(x)
        mods = list(self.model)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Upsample):
                blk = mods[i + 1]
                blk.upsample = 2                      # nearest x2 folded into the conv's operand gather
                # the three 5x5 up-convolutions are the last dense layers in front of the image, where forward rounding is
                # amplified least: in mode 'f16' their FORWARD GEMM runs one fp16 pass instead of three (image 1.5e-3, worst
                # per-tensor gradient cosine ~0.9998 in scripts/precision_sweep.py "plans4"; AFFGW_UPCONV_PASSES=3 disables)
                n_fwd = _UPCONV_PASSES if ops.precision() == "f16" else None
                with ops.conv_passes(fwd=n_fwd, dgrad=1 if n_fwd else None, wgrad=1 if n_fwd else None):
                    x = blk(x)
                i += 2
            elif i == len(mods) - 1:
                x = m(x, out_dtype=torch.float32)     # tanh image in fp32
                i += 1
            else:
                x = m(x)
                i += 1
        return x


def make_encoder(name, in_channels=None):
    """'vgg' (ImageEncoder, the north-star path), 'resnet50' (the encoder active in the reference, modules_tro.py:219),
    'resnet18' (modules_tro2.py:447-516) or 'dino' (dinomodel.py: DINOv2 ViT-L/14, generation only)."""
    from .load_data import NUM_CHANNEL
    from .resnet_encoder import ImageEncoderResNet18, ImageEncoderResNet50
    c = NUM_CHANNEL if in_channels is None else in_channels
    if name == "vgg":
        return ImageEncoder()
    if name == "resnet50":
        return ImageEncoderResNet50(weight_path=None, in_channels=c)
    if name == "resnet18":
        return ImageEncoderResNet18(weight_path=None, in_channels=c)
    if name in ("dino", "dinov2"):                    # modules_tro.py:214 (commented line): generation only
        from .dinomodel import ImageEncoderDINOv2
        return ImageEncoderDINOv2(None, arch="vitl14", ckpt_path=None, in_channels=c, final_size=(8, 27), tap_blocks=[4, 8, 16, 23])
    raise ValueError(f"unknown style encoder {name!r}")


class GenModel_FC(nn.Module):
    def __init__(self, text_max_len=OUTPUT_MAX_LEN, encoder=None):
        super().__init__()
        if isinstance(encoder, str):
            encoder = make_encoder(encoder)
        self.enc_image = encoder if encoder is not None else ImageEncoder()
        self.enc_text = TextEncoder_FC(text_max_len)
        self.dec = Decoder()
        self.linear_mix = nn.Linear(1024, 512)
        self.max_conv = nn.MaxPool2d(kernel_size=2, stride=2)

    def assign_adain_params(self, adain_params, results, embed):
        # modules_tro.py:226-242; `input` persists between calls exactly like the reference's attribute
        i = 0
        for m in self.dec.modules():
            if m.__class__.__name__ == "AdaptiveInstanceNorm2d":
                mean = adain_params[:, :m.num_features]
                std = adain_params[:, m.num_features:2 * m.num_features]
                m.bias = mean.contiguous().view(-1)
                m.weight = std.contiguous().view(-1)
                m.con = embed
                if i == 1:
                    m.input = ops.max_pool2(results[3])
                elif i == 3:
                    m.input = results[4]
                if adain_params.size(1) > 2 * m.num_features:
                    adain_params = adain_params[:, 2 * m.num_features:]
                i += 1

    def decode(self, content, results, embed, adain_params):
        self.assign_adain_params(adain_params, results, embed)
        return self.dec(content)

    def mix(self, results, feat_embed):
        # cat along channels then per-pixel Linear(1024, 512) == 1x1 convolution in NHWC (modules_tro.py:252-259)
        style = ops.input_to_internal(results[-1])
        feat_embed = ops.input_to_internal(feat_embed)
        if style.shape[1] != 512:                       # channel-padded views never reach here, guard anyway
            style = style[:, :512]
        f = _concat_channels(style, feat_embed)
        return ops.conv2d(f, self.linear_mix.weight, self.linear_mix.bias)

    def forward(self, tr_img, label):
        """network_tro.py:60-66 composition, exposed as one call for generation."""
        f_xss = self.enc_image(tr_img)
        f_xt, f_embed = self.enc_text(label, f_xss[-1].shape)
        f_mix = self.mix(f_xss, f_embed)
        return self.decode(f_mix, f_xss, f_embed, f_xt)


class _ConcatC(torch.autograd.Function):
    """torch.cat([style, content], dim=1) on NHWC tensors (modules_tro.py:256) and its split backward."""

    @staticmethod
    def forward(ctx, a, b):
        n, ca, h, w = a.shape
        cb = b.shape[1]
        out = ops.empty_cl(n, ca + cb, h, w, a.dtype, a.device)
        ops.L.call("affgw_concat_channels", a.data_ptr(), b.data_ptr(), out.data_ptr(), ops.L.dt(a), n * h * w, ca, cb,
                   ops.L.stream())
        ctx.dims = (n, ca, cb, h, w)
        return out

    @staticmethod
    def backward(ctx, g):
        n, ca, cb, h, w = ctx.dims
        g = ops._dense_cl(g)
        da = ops.empty_cl(n, ca, h, w, g.dtype, g.device)
        db = ops.empty_cl(n, cb, h, w, g.dtype, g.device)
        ops.L.call("affgw_split_channels", g.data_ptr(), da.data_ptr(), db.data_ptr(), ops.L.dt(g), n * h * w, ca, cb,
                   ops.L.stream())
        return da, db


def _concat_channels(a, b):
    return _ConcatC.apply(ops._dense_cl(a), ops._dense_cl(b, a.dtype))


class MLP(nn.Module):
    def __init__(self, in_dim=64, out_dim=4096, dim=256, n_blk=3, norm="none", activ="relu"):
        super().__init__()
        model = [LinearBlock(in_dim, dim, norm=norm, activation=activ)]
        for _ in range(n_blk - 2):
            model += [LinearBlock(dim, dim, norm=norm, activation=activ)]
        model += [LinearBlock(dim, out_dim, norm="none", activation="none")]
        self.model = nn.Sequential(*model)

    def forward(self, x):
        x = ops.input_to_internal(x.reshape(x.size(0), -1))
        for blk in self.model:
            x = blk(x)
        return x
