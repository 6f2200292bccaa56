"""stand-in for the (unvendored) pretrained.PyTorch_CIFAR10 package the CIFAR driver imports (main_cifar10.py:6,
quant_block.py:11): CIFAR ResNets built from the zoo's ResNet with 10 classes (BASELINE config 1)."""
import warnings

from shiftedscalequantization_b200.quant.quant_block import BasicBlockCIFAR  # noqa: F401
from shiftedscalequantization_b200 import zoo


def _make(arch):
    def build(pretrained=False, progress=True, device='cpu', **kwargs):
        if pretrained:
            warnings.warn(f'{arch}: CIFAR-10 checkpoints are not available offline; using the seeded random init')
        return zoo.build(arch, num_classes=10, **kwargs).to(device)
    return build


resnet18, resnet34, resnet50 = _make('resnet18'), _make('resnet34'), _make('resnet50')
