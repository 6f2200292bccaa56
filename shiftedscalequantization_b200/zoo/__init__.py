"""Calibration fixtures: the stock FP architectures the quantised block wrappers dispatch on
(quant_block.specials). Random-init only — checkpoints are not available offline. Attribute names follow
the upstream model files (models/resnet.py, models/mobilenetv2.py, models/regnet.py) because the block
wrappers and BRECQ state_dicts address sub-modules by those names."""
from .resnet import BasicBlock, Bottleneck, ResNet, resnet18, resnet34, resnet50, resnet101
from .mobilenetv2 import InvertedResidual, MobileNetV2, mobilenetv2
from .regnet import ResBottleneckBlock, RegNet, regnetx_600m, regnetx_3200m

ARCHS = {
    'resnet18': resnet18, 'resnet34': resnet34, 'resnet50': resnet50, 'resnet101': resnet101,
    'mobilenetv2': mobilenetv2, 'regnetx_600m': regnetx_600m, 'regnetx_3200m': regnetx_3200m,
}


def build(arch: str, **kwargs):
    return ARCHS[arch](**kwargs)
