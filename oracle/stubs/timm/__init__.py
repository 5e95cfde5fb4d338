"""Stand-in for the un-vendored timm==0.6.7 (environment.yml:72): resnet50 built on torchvision's ResNet.

Same v1.5 bottleneck topology (stride on the 3x3), same parameter names (conv1, bn1, layerN.i.{conv,bn}{1,2,3},
downsample.{0,1}, fc), eval-mode arithmetic identical.  Differences are init-only / train-only (zero_init_last,
DropPath) and irrelevant to parity on explicitly supplied weights.  PARITY UNPINNED against real timm.
"""


def create_model(name, pretrained=False, in_chans=3, num_classes=1000, **kwargs):
    from timm.models.resnet import make_resnet

    return make_resnet(name, in_chans=in_chans, num_classes=num_classes)
