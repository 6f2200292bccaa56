from shiftedscalequantization_b200.zoo.resnet import BasicBlock, Bottleneck, ResNet, resnet18, resnet34, resnet50, resnet101  # noqa: F401
