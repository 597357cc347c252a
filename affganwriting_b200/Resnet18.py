"""Drop-in for the reference's stand-alone ResNet-18 (GAN_word/Resnet18.py:4-88): `conv3x3`, `BasicBlock`, `ResNet18`.

Same constructor arguments, attribute names and state_dict keys (`conv1.weight`, `bn1.*`, `layer{1,2,3}.{0,1}.conv{1,2}.weight`,
`...bn{1,2}.*`, `layer*.0.downsample.{0,1}.*`); the parameter containers are plain torch modules, every forward runs
libaffgw kernels.  What is specific to this network: the stem convolution and the first max pool stride the ROWS by 2 and
the COLUMNS by 1 (Resnet18.py:43,45), which is the `stride_w` field of `affgw_conv_desc` and `affgw_maxpool3_*`; the
last map goes through a stride-1 3x3 max pool (:46, :85).

The reference cannot run this class inside GenModel_FC (the decoder's AdaIN keys expect 512 channels and 8 x 27 maps),
so it is provided, like there, as a stand-alone module; parity: tests/test_gpu_resnet.py against tests/golden/resnet18_standalone.npz.
"""
from torch import nn

from . import ops


def conv3x3(in_planes, out_planes, stride=1):
    """3 x 3, padding 1, no bias (Resnet18.py:4-7); a parameter container - `_conv` runs it."""
    return nn.Conv2d(in_planes, out_planes, 3, stride=stride, padding=1, bias=False)


def _conv(x, m):
    return ops.conv2d(x, m.weight, m.bias, stride=m.stride, pad=m.padding[0], pad_mode="zero")


def _bn(planes):
    return nn.BatchNorm2d(planes, eps=1e-05)


class BasicBlock(nn.Module):
    """Resnet18.py:9-36: conv-bn-relu, conv-bn, (+ 1 x 1 strided shortcut), add, relu."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.stride = stride
        self.conv1, self.bn1 = conv3x3(inplanes, planes, stride), _bn(planes)
        self.relu = nn.ReLU(inplace=True)                       # kept for attribute / repr compatibility only
        self.conv2, self.bn2 = conv3x3(planes, planes), _bn(planes)
        self.downsample = downsample

    def forward(self, x):
        x = ops.input_to_internal(x)
        main = ops.batch_norm(_conv(x, self.conv1), self.bn1, act="relu")
        main = ops.batch_norm(_conv(main, self.conv2), self.bn2)
        skip = x if self.downsample is None else ops.batch_norm(_conv(x, self.downsample[0]), self.downsample[1])
        return ops.add_act(main, skip, "relu")                  # `out += residual; relu(out)` in one kernel


class ResNet18(nn.Module):
    """Resnet18.py:38-88.  Maps returned (channels, rows, columns for an H x W input):
    (nb/4, H/4, W), (nb/4, H/8, W/2), (nb/2, H/16, W/4), (nb, H/32, W/8) and its 3 x 3 / stride-1 max pool."""
    STAGES = (("layer1", 4, (2, 2)), ("layer2", 2, 2), ("layer3", 1, 2))      # attribute, nb_feat divisor, stride

    def __init__(self, nb_feat=384, in_channels=3):
        super().__init__()
        self.inplanes = nb_feat // 4
        # registration order = the reference's, so state_dict() lists the keys in the same order
        self.conv1 = nn.Conv2d(in_channels, self.inplanes, 3, stride=(2, 1), padding=1, bias=False)
        self.bn1 = _bn(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool1 = nn.MaxPool2d(3, stride=(2, 1), padding=1)
        self.maxpool2 = nn.MaxPool2d(3, stride=(1, 1), padding=1)
        for name, div, stride in self.STAGES:
            setattr(self, name, self._make_layer(BasicBlock, nb_feat // div, 2, stride=stride))

    def _make_layer(self, block, planes, blocks, stride=1):
        out_planes = planes * block.expansion
        shortcut = None
        if stride != 1 or self.inplanes != out_planes:
            shortcut = nn.Sequential(nn.Conv2d(self.inplanes, out_planes, 1, stride=stride, bias=False), _bn(out_planes))
        stage = [block(self.inplanes, planes, stride, shortcut)] + [block(out_planes, planes) for _ in range(blocks - 1)]
        self.inplanes = out_planes
        return nn.Sequential(*stage)

    def forward(self, x):
        x = ops.input_to_internal(x)
        x = ops.max_pool3(ops.batch_norm(_conv(x, self.conv1), self.bn1, act="relu"), self.maxpool1.stride)
        maps = [x]
        for name, _, _ in self.STAGES:
            for blk in getattr(self, name):
                x = blk(x)
            maps.append(x)
        maps.append(ops.max_pool3(x, self.maxpool2.stride))
        return maps
