import torch.nn as nn
from torchvision.models.resnet import Bottleneck, BasicBlock
from torchvision.models.resnet import ResNet as _TVResNet

_CFG = {
    "resnet18": (BasicBlock, [2, 2, 2, 2]),
    "resnet34": (BasicBlock, [3, 4, 6, 3]),
    "resnet50": (Bottleneck, [3, 4, 6, 3]),
    "resnet101": (Bottleneck, [3, 4, 23, 3]),
}


class _Pool(nn.Module):
    def __init__(self):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool2d(1)

    def forward(self, x):
        return self.pool(x).flatten(1)


class ResNet(_TVResNet):
    """torchvision ResNet with the attribute names the reference's monkey-patched forward uses
    (spark/resnet.py:13-46: act1, global_pool, drop_rate, fc)."""

    def __init__(self, block, layers, in_chans=3, num_classes=1000):
        super().__init__(block, layers, num_classes=num_classes)
        if in_chans != 3:
            self.conv1 = nn.Conv2d(in_chans, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.drop_rate = 0.0

    @property
    def act1(self):
        return self.relu

    @property
    def global_pool(self):
        return lambda x: self.avgpool(x).flatten(1)


def make_resnet(name, in_chans=3, num_classes=1000):
    block, layers = _CFG[name]
    return ResNet(block, layers, in_chans=in_chans, num_classes=num_classes)
