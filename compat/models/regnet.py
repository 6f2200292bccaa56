from shiftedscalequantization_b200.zoo.regnet import (AnyHead, AnyStage, BottleneckTransform, RegNet, ResBottleneckBlock,  # noqa: F401
                                                       SimpleStemIN, regnetx_600m, regnetx_3200m)
