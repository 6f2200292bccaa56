from shiftedscalequantization_b200.zoo.mobilenetv2 import InvertedResidual, MobileNetV2, mobilenetv2  # noqa: F401
