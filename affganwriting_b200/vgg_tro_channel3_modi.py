"""VGG-19 'E' style encoder with InstanceNorm (reference GAN_word/vgg_tro_channel3_modi.py:40-90).

`features` is an nn.Sequential with the reference's indices (conv at 0,3,6,9,13,...) so checkpoint keys match;
the InstanceNorm2d / ReLU / MaxPool2d members are placeholders - forward fuses conv(+bias) -> instance-norm+ReLU
and max-pool into libaffgw kernels.
"""
import torch.nn as nn

from . import load_data, ops

import os as _os

# Convolutions sit at `features` indices 0, 3, 6, 9 | 13, 16, 19, 22 | 26, 29, 32, 35 | 39, 42, 45, 48.  Inside
# ops.relaxed_forward() the ones at index >= RELAXED_FROM run single-pass: 14 -> from index 16 on, i.e. convolutions 6-16
# (75 % of the encoder's FLOPs; the first five keep three passes).  Measured on B200 at 50 planes, batch 8 (tests/test_gpu_parity_c50.py, complete dis_update,
# worst discriminator tensor): from 26 (convolution 9 on) -> 0.999985, 19 (7 on) -> 0.99992, 14 (6 on) -> 0.99994, 7 (4 on) -> 0.99969,
# 0 (every layer) -> 0.99938;
# a value past the last convolution (48) switches the relaxation off
RELAXED_FROM = int(_os.environ.get("AFFGW_RELAXED_VGG_FROM", "14"))

cfg = {
    "E": [64, 64, 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512],
}


def make_layers(cfg_list, batch_norm=False, in_channels=None):
    layers = []
    in_channels = load_data.NUM_CHANNEL if in_channels is None else in_channels
    for v in cfg_list:
        if v == "M":
            layers += [nn.MaxPool2d(kernel_size=2, stride=2)]
        else:
            conv2d = nn.Conv2d(in_channels, v, kernel_size=3, padding=1)
            layers += [conv2d, nn.InstanceNorm2d(v), nn.ReLU(inplace=True)] if batch_norm else [conv2d, nn.ReLU(inplace=True)]
            in_channels = v
    return nn.Sequential(*layers)


class VGG(nn.Module):
    def __init__(self, features, init_weights=True):
        super().__init__()
        self.features = features
        if init_weights:
            self._initialize_weights()

    def _initialize_weights(self):
        # vgg_tro_channel3_modi.py:29-37
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def run(self, x, start, stop):
        """Execute features[start:stop] with fused kernels; x is an internal (channels-last) tensor."""
        feats = self.features
        i = start
        while i < stop:
            m = feats[i]
            if isinstance(m, nn.Conv2d):
                if ops.relaxed() and i >= ops.relaxed_from(RELAXED_FROM):         # ops.relaxed_forward: one fp16 pass for the deep layers
                    with ops.conv_passes(fwd=1):
                        x = ops.conv2d(x, m.weight, m.bias, stride=1, pad=1, pad_mode="zero")
                else:
                    x = ops.conv2d(x, m.weight, m.bias, stride=1, pad=1, pad_mode="zero")
                nxt = feats[i + 1] if i + 1 < stop else None
                if isinstance(nxt, nn.InstanceNorm2d):
                    has_relu = i + 2 < stop and isinstance(feats[i + 2], nn.ReLU)
                    x = ops.instance_norm(x, act="relu" if has_relu else "none", eps=nxt.eps)
                    i += 3 if has_relu else 2
                    continue
                i += 1
            elif isinstance(m, nn.MaxPool2d):
                x = ops.max_pool2(x)
                i += 1
            elif isinstance(m, nn.InstanceNorm2d):
                x = ops.instance_norm(x, eps=m.eps)
                i += 1
            elif isinstance(m, nn.ReLU):
                raise RuntimeError("VGG slice starts on a bare ReLU; slices must begin at a conv or pool")
            else:
                raise RuntimeError(f"unexpected VGG member {type(m).__name__}")
        return x

    def forward(self, x):
        x = ops.input_to_internal(x, c_pad=_pad64(x.shape[1]))
        return self.run(x, 0, len(self.features))


def _pad64(c):
    # channel padding to the tensor-core gather granule now happens inside the convolution (affgw_split_planes)
    return c


def vgg19_bn(pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("pretrained VGG weights are not shipped; load a checkpoint with load_state_dict")
    return VGG(make_layers(cfg["E"], batch_norm=True), **kwargs)
