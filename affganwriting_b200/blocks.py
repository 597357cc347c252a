"""Drop-in replacements for the reference's building blocks (reference GAN_word/blocks.py).

Same class names, constructor signatures, attribute names and state_dict keys; the torch.nn leaf modules are
kept only as parameter / buffer containers (so checkpoints load unchanged and default initialisation is the
reference's) while every forward runs libaffgw kernels through affganwriting_b200.ops.  Tensors keep the
reference's logical NCHW shapes and are stored channels-last.

  Conv2dBlock            blocks.py:106-163      ResBlock / ResBlocks   blocks.py:6-39
  ActFirstResBlock       blocks.py:42-65        LinearBlock            blocks.py:68-103
  AdaptiveInstanceNorm2d blocks.py:166-207      iAFF                   blocks.py:238-299
  get_key / mean_variance_norm / calc_mean_std   blocks.py:210-235
"""
import torch
from torch import nn

from . import ops

import os as _os
_RELAXED_RES = _os.environ.get("AFFGW_RELAXED_RES", "1") != "0"

_ACTS = ("relu", "lrelu", "tanh", "none")
_PADS = ("reflect", "replicate", "zero")


def _act_module(activation):
    # containers only (kept so that `.activation` has the reference's type and truthiness)
    if activation == "relu":
        return nn.ReLU(inplace=False)
    if activation == "lrelu":
        return nn.LeakyReLU(0.2, inplace=False)
    if activation == "tanh":
        return nn.Tanh()
    if activation == "none":
        return None
    assert 0, "Unsupported activation: {}".format(activation)


class ResBlocks(nn.Module):
    def __init__(self, num_blocks, dim, norm, activation, pad_type):
        super().__init__()
        self.model = nn.Sequential(*[ResBlock(dim, norm=norm, activation=activation, pad_type=pad_type)
                                     for _ in range(num_blocks)])

    def forward(self, x):
        x = ops.input_to_internal(x)
        for blk in self.model:
            x = blk(x)
        return x


class ResBlock(nn.Module):
    def __init__(self, dim, norm="in", activation="relu", pad_type="zero"):
        super().__init__()
        self.model = nn.Sequential(
            Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation=activation, pad_type=pad_type),
            Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation="none", pad_type=pad_type))

    def forward(self, x):
        x = ops.input_to_internal(x)
        if ops.relaxed() and _RELAXED_RES:          # ops.relaxed_forward: the two 3x3 convolutions (not the iAFF 1x1s) single-pass
            self.model[0].fwd_passes = self.model[1].fwd_passes = 1
        try:
            h = self.model[0](x)
            # `out += residual` (blocks.py:38) is fused into the second block's normalisation epilogue
            return self.model[1](h, residual=x)
        finally:
            self.model[0].fwd_passes = self.model[1].fwd_passes = None


class ActFirstResBlock(nn.Module):
    def __init__(self, fin, fout, fhid=None, activation="lrelu", norm="none"):
        super().__init__()
        self.learned_shortcut = (fin != fout)
        self.fin = fin
        self.fout = fout
        self.fhid = min(fin, fout) if fhid is None else fhid
        self.conv_0 = Conv2dBlock(self.fin, self.fhid, 3, 1, padding=1, pad_type="reflect", norm=norm,
                                  activation=activation, activation_first=True)
        self.conv_1 = Conv2dBlock(self.fhid, self.fout, 3, 1, padding=1, pad_type="reflect", norm=norm,
                                  activation=activation, activation_first=True)
        if self.learned_shortcut:
            self.conv_s = Conv2dBlock(self.fin, self.fout, 1, 1, activation="none", use_bias=False)

    def forward(self, x):
        x = ops.input_to_internal(x)
        x_s = self.conv_s(x) if self.learned_shortcut else x
        dx = self.conv_0(x)
        # `x_s + dx` (blocks.py:64) rides in conv_1's epilogue when no norm follows the conv
        if self.conv_1.norm is None:
            return self.conv_1(dx, addend=x_s)
        return ops.add(x_s, self.conv_1(dx))


class LinearBlock(nn.Module):
    def __init__(self, in_dim, out_dim, norm="none", activation="relu"):
        super().__init__()
        self.fc = nn.Linear(in_dim, out_dim, bias=True)
        if norm == "bn":
            self.norm = nn.BatchNorm1d(out_dim)
        elif norm == "in":
            self.norm = nn.InstanceNorm1d(out_dim)
        elif norm == "none":
            self.norm = None
        else:
            assert 0, "Unsupported normalization: {}".format(norm)
        self._norm_kind, self._act = norm, activation
        self.activation = _act_module(activation)

    def forward(self, x):
        x = ops.input_to_internal(x)
        act = self._act
        if self._norm_kind == "none":
            return ops.linear(x, self.fc.weight, self.fc.bias, post_act=act)
        out = ops.linear(x, self.fc.weight, self.fc.bias)
        if self._norm_kind == "bn":
            if act == "tanh":
                out = ops.batch_norm(out, self.norm)
                return _standalone_act(out, act)
            return ops.batch_norm(out, self.norm, act=act)
        # nn.InstanceNorm1d on a 2-D [B, F] input (blocks.py:78-79,99-100): torch reads it as ONE unbatched (C = B, L = F)
        # sample, i.e. every row is normalised over its F features (biased variance, eps 1e-5, no affine, no running stats)
        b, f = out.shape
        y = ops.instance_norm(out.reshape(b, 1, f, 1), act="none" if act == "tanh" else act, eps=self.norm.eps)
        y = y.reshape(b, f)
        return _standalone_act(y, act) if act == "tanh" else y


def _standalone_act(x, act):
    raise NotImplementedError("tanh after a normalisation layer does not occur on the reference path")


class Conv2dBlock(nn.Module):
    def __init__(self, in_dim, out_dim, ks, st, padding=0, norm="none", activation="relu", pad_type="zero",
                 use_bias=True, activation_first=False):
        super().__init__()
        self.use_bias = use_bias
        self.activation_first = activation_first
        if pad_type == "reflect":
            self.pad = nn.ReflectionPad2d(padding)
        elif pad_type == "replicate":
            self.pad = nn.ReplicationPad2d(padding)
        elif pad_type == "zero":
            self.pad = nn.ZeroPad2d(padding)
        else:
            assert 0, "Unsupported padding type: {}".format(pad_type)
        if norm == "bn":
            self.norm = nn.BatchNorm2d(out_dim)
        elif norm == "in":
            self.norm = nn.InstanceNorm2d(out_dim)
        elif norm == "adain":
            self.norm = AdaptiveInstanceNorm2d(out_dim)
        elif norm == "none":
            self.norm = None
        else:
            assert 0, "Unsupported normalization: {}".format(norm)
        self.activation = _act_module(activation)
        self.conv = nn.Conv2d(in_dim, out_dim, ks, st, bias=self.use_bias)
        self._pad, self._pad_type, self._st = int(padding), pad_type, int(st)
        self._norm_kind, self._act = norm, activation
        self.upsample = 1   # Decoder folds the preceding nn.Upsample(scale_factor=2) into the gather
        self.fwd_passes = None   # set to 1 by ResBlock inside ops.relaxed_forward()

    def forward(self, x, residual=None, addend=None, out_dtype=None):
        """residual: added after norm+activation (ResBlock); addend: added to the raw conv result (shortcut)."""
        x = ops.input_to_internal(x)
        act, first = self._act, self.activation_first
        fuse_post = (self._norm_kind == "none" and not first)
        with ops.conv_passes(fwd=self.fwd_passes):      # None keeps the mode's pass count
            y = ops.conv2d(x, self.conv.weight, self.conv.bias, stride=self._st, pad=self._pad, pad_mode=self._pad_type,
                           upsample=self.upsample, pre_act=act if first else "none",
                           post_act=act if fuse_post else "none", addend=addend, out_dtype=out_dtype)
        post = "none" if first else act
        if self._norm_kind == "in":
            y = ops.instance_norm(y, act=post, residual=residual, eps=self.norm.eps)
        elif self._norm_kind == "adain":
            y = self.norm(y, act=post, residual=residual)
        elif self._norm_kind == "bn":
            y = ops.batch_norm(y, self.norm, act=post)
            if residual is not None:
                y = ops.add(y, residual)
        elif residual is not None:
            y = ops.add(y, residual)
        return y


class AdaptiveInstanceNorm2d(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.weight = None
        self.bias = None
        self.input = None
        self.con = None
        # never used in forward (reference blocks.py:176-178): kept for checkpoint compatibility
        self.conv = nn.Conv2d(1024, 512, kernel_size=3, padding=1)
        self.linear_mix = nn.Linear(1024, 512)
        self.iAff = iAFF(512)
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def mix(self, feat_xs, feat_embed):
        f = torch.cat([ops.input_to_internal(feat_xs), ops.input_to_internal(feat_embed)], dim=1)
        return ops.conv2d(ops.to_internal(f), self.linear_mix.weight, self.linear_mix.bias)

    def forward(self, x1, act="none", residual=None):
        assert self.weight is not None and self.bias is not None, "Please assign AdaIN weight first"
        x1 = ops.input_to_internal(x1)
        if self.input is not None:
            x = self.iAff(x1, get_key(x1, self.input))
        else:
            x = x1
        # per-(n, c) statistics always (F.batch_norm(..., training=True), blocks.py:201-203); the registered
        # running_mean / running_var buffers never change in the reference (only their .repeat(b) copies do)
        return ops.instance_norm(x, act=act, gamma=self.weight, beta=self.bias, residual=residual, eps=self.eps)

    def __repr__(self):
        return self.__class__.__name__ + "(" + str(self.num_features) + ")"


def get_key(feats, feat):
    _, _, h, w = feats.shape
    return mean_variance_norm(ops.resize_nearest(ops.input_to_internal(feat), h, w))


def mean_variance_norm(feat):
    # (feat - mean) / sqrt(var_unbiased + 1e-5)   (blocks.py:218-235)
    return ops.instance_norm(ops.input_to_internal(feat), eps=1e-5, unbiased=True)


def calc_mean_std(feat, eps=1e-5):
    """Statistics helper kept for API completeness (torch reductions; not on the accelerated path)."""
    n, c = feat.shape[:2]
    f = feat.float().reshape(n, c, -1)
    return f.mean(dim=2).view(n, c, 1, 1), (f.var(dim=2) + eps).sqrt().view(n, c, 1, 1)


class iAFF(nn.Module):
    """Iterative attentional feature fusion.  Sequential members are parameter containers with the reference's
    indices (local: 0 conv, 1 bn, 3 conv, 4 bn; global: 1 conv, 2 bn, 4 conv, 5 bn)."""

    def __init__(self, channels=512, r=4):
        super().__init__()
        inter = int(channels // r)

        def local():
            return nn.Sequential(nn.Conv2d(channels, inter, 1), nn.BatchNorm2d(inter), nn.ReLU(inplace=True),
                                 nn.Conv2d(inter, channels, 1), nn.BatchNorm2d(channels))

        def glob():
            return nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(channels, inter, 1), nn.BatchNorm2d(inter),
                                 nn.ReLU(inplace=True), nn.Conv2d(inter, channels, 1), nn.BatchNorm2d(channels))
        self.local_att = local()
        self.global_att = glob()
        self.local_att2 = local()
        self.global_att2 = glob()   # allocated but never called (reference blocks.py:295 reuses global_att)
        self.sigmoid = nn.Sigmoid()

    @staticmethod
    def _branch(x, seq, o):
        h = ops.conv2d(x, seq[o].weight, seq[o].bias)
        h = ops.batch_norm(h, seq[o + 1], act="relu")
        h = ops.conv2d(h, seq[o + 3].weight, seq[o + 3].bias)
        return ops.batch_norm(h, seq[o + 4])

    def forward(self, x, residual):
        x, residual = ops.input_to_internal(x), ops.input_to_internal(residual)
        xa = ops.add(x, residual)
        xl = self._branch(xa, self.local_att, 0)
        xg = self._branch(ops.global_avg_pool(xa), self.global_att, 1)
        xi = ops.iaff_gate(x, residual, xl, xg)
        xl2 = self._branch(xi, self.local_att2, 0)
        xg2 = self._branch(ops.global_avg_pool(xi), self.global_att, 1)
        return ops.iaff_gate(x, residual, xl2, xg2)
