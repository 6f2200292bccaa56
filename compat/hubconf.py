"""model entry points of upstream trash/hubconf.py; checkpoints cannot be downloaded offline"""
import warnings

from shiftedscalequantization_b200 import zoo

dependencies = ['torch']


def _make(arch):
    def build(pretrained=False, **kwargs):
        if pretrained:
            warnings.warn(f'{arch}: BRECQ checkpoints are not available offline; using the seeded random init')
        return zoo.build(arch, **kwargs)
    build.__name__ = arch
    return build


resnet18, resnet50 = _make('resnet18'), _make('resnet50')
mobilenetv2 = _make('mobilenetv2')
regnetx_600m, regnetx_3200m = _make('regnetx_600m'), _make('regnetx_3200m')
