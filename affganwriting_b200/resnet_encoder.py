"""torchvision-ResNet style encoders of the generator (reference GAN_word/modules_tro.py:464-533 - the encoder that is
ACTIVE in the reference's GenModel_FC, ResNet-50 - and its ResNet-18 twin, modules_tro2.py:447-516).

The torchvision model object and the fx feature extractor are kept only as parameter / buffer containers, so the
state_dict keys (`model.*`, `extractor.*`, `reduce_layers.*`) are exactly the reference's; every forward runs libaffgw
kernels: strided / 1x1 / 3x3 / 7x7 convolutions on the tcgen05 kernels, BatchNorm with fused ReLU, the 3x3/s2 max pool,
act(a + b) residual tails, 1x1 reducers to 512 channels and the bilinear resize of the last map to 8 x 27.
"""
import torch
import torch.nn.functional as F  # noqa: F401  (kept for parity with the reference's imports)
from torch import nn

from . import ops


def _tv():
    try:
        from torchvision.models import resnet18, resnet50
        from torchvision.models.feature_extraction import create_feature_extractor
    except Exception as e:  # pragma: no cover
        raise RuntimeError("the ResNet style encoders need torchvision (as the reference does): %s" % e)
    return resnet18, resnet50, create_feature_extractor


def _conv(x, m, post_act="none"):
    return ops.conv2d(x, m.weight, m.bias, stride=m.stride[0], pad=m.padding[0], pad_mode="zero", post_act=post_act)


def _block(x, blk):
    """BasicBlock / Bottleneck forward (torchvision.models.resnet): conv-bn-relu chain, `out += identity`, relu."""
    identity = x
    if hasattr(blk, "conv3"):       # Bottleneck
        out = ops.batch_norm(_conv(x, blk.conv1), blk.bn1, act="relu")
        out = ops.batch_norm(_conv(out, blk.conv2), blk.bn2, act="relu")
        out = ops.batch_norm(_conv(out, blk.conv3), blk.bn3)
    else:                           # BasicBlock
        out = ops.batch_norm(_conv(x, blk.conv1), blk.bn1, act="relu")
        out = ops.batch_norm(_conv(out, blk.conv2), blk.bn2)
    if blk.downsample is not None:
        identity = ops.batch_norm(_conv(x, blk.downsample[0]), blk.downsample[1])
    return ops.add_act(out, identity, "relu")


class _ImageEncoderResNet(nn.Module):
    ARCH = None
    REDUCE_IN = None

    def __init__(self, weight_path=None, in_channels=50):
        super().__init__()
        resnet18, resnet50, create_feature_extractor = _tv()
        self.output_dim = 512
        self.model = {"resnet18": resnet18, "resnet50": resnet50}[self.ARCH](weights=None)
        if weight_path:
            self.model.load_state_dict(torch.load(weight_path, map_location="cpu"))
        # first convolution re-shaped to `in_channels` planes exactly like the reference (modules_tro.py:478-493)
        original_conv = self.model.conv1
        new_conv = nn.Conv2d(in_channels, original_conv.out_channels, kernel_size=original_conv.kernel_size,
                             stride=original_conv.stride, padding=original_conv.padding, bias=original_conv.bias is not None)
        with torch.no_grad():
            new_conv.weight[:, :3] = original_conv.weight
            if in_channels > 3:
                new_conv.weight[:, 3:] = original_conv.weight[:, :1].repeat(1, in_channels - 3, 1, 1)
        self.model.conv1 = new_conv
        return_nodes = {"relu": "feat1", "layer1": "feat2", "layer2": "feat3", "layer3": "feat4", "layer4": "feat5"}
        self.extractor = create_feature_extractor(self.model, return_nodes=return_nodes)    # key compatibility only
        self.reduce_layers = nn.ModuleList([nn.Conv2d(c, 512, kernel_size=1) for c in self.REDUCE_IN])

    def _features(self, x):
        m = self.model
        x = ops.input_to_internal(x)
        x = ops.batch_norm(_conv(x, m.conv1), m.bn1, act="relu")
        feats = [x]
        x = ops.max_pool3s2(x)
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                x = _block(x, blk)
            feats.append(x)
        return feats

    def encode_with_intermediate(self, x):
        results = [_conv(f, r) for f, r in zip(self._features(x), self.reduce_layers)]
        results[-1] = ops.resize_bilinear(results[-1], 8, 27)
        return results

    def forward(self, x):
        return self.encode_with_intermediate(x)

    def train(self, mode=True):
        # the fx GraphModule shares the BatchNorm modules with `model`; keep both views in the same mode
        super().train(mode)
        return self


class ImageEncoderResNet50(_ImageEncoderResNet):
    """modules_tro.py:464-533."""
    ARCH = "resnet50"
    REDUCE_IN = (64, 256, 512, 1024, 2048)


class ImageEncoderResNet18(_ImageEncoderResNet):
    """modules_tro2.py:447-516 (the class is called ImageEncoderResNet50 there too, but wraps resnet18)."""
    ARCH = "resnet18"
    REDUCE_IN = (64, 64, 128, 256, 512)
