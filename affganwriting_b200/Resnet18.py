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
    """Resnet18.py:4-7 (parameter container; see `_conv`)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def _conv(x, m):
    return ops.conv2d(x, m.weight, m.bias, stride=m.stride, pad=m.padding[0], pad_mode="zero")


class BasicBlock(nn.Module):
    """Resnet18.py:9-36."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes, eps=1e-05)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes, eps=1e-05)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        x = ops.input_to_internal(x)
        out = ops.batch_norm(_conv(x, self.conv1), self.bn1, act="relu")
        out = ops.batch_norm(_conv(out, self.conv2), self.bn2)
        residual = x
        if self.downsample is not None:
            residual = ops.batch_norm(_conv(x, self.downsample[0]), self.downsample[1])
        return ops.add_act(out, residual, "relu")       # out += residual; relu


class ResNet18(nn.Module):
    """Resnet18.py:38-88: five maps, [B, nb/4, H/4, W] [B, nb/4, H/8, W/2] [B, nb/2, H/16, W/4] [B, nb, H/32, W/8] x 2."""

    def __init__(self, nb_feat=384, in_channels=3):
        super().__init__()
        self.inplanes = nb_feat // 4
        self.conv1 = nn.Conv2d(in_channels, self.inplanes, kernel_size=3, stride=(2, 1), padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(self.inplanes, eps=1e-05)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool1 = nn.MaxPool2d(kernel_size=3, stride=(2, 1), padding=1)
        self.maxpool2 = nn.MaxPool2d(kernel_size=3, stride=(1, 1), padding=1)
        self.layer1 = self._make_layer(BasicBlock, nb_feat // 4, 2, stride=(2, 2))
        self.layer2 = self._make_layer(BasicBlock, nb_feat // 2, 2, stride=2)
        self.layer3 = self._make_layer(BasicBlock, nb_feat, 2, stride=2)

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion, eps=1e-05))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = ops.input_to_internal(x)
        x = ops.batch_norm(_conv(x, self.conv1), self.bn1, act="relu")
        x = ops.max_pool3(x, self.maxpool1.stride)
        results = [x]
        for layer in (self.layer1, self.layer2, self.layer3):
            for blk in layer:
                x = blk(x)
            results.append(x)
        results.append(ops.max_pool3(x, self.maxpool2.stride))
        return results
