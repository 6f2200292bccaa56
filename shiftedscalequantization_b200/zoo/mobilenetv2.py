"""MobileNetV2 (structure of upstream models/mobilenetv2.py:24-136; conv Sequential indices 0/3/6 are what
QuantInvertedResidual picks up)."""
import math

import torch.nn as nn


def _cbr(cin, cout, k, stride, groups=1, act=True):
    mods = [nn.Conv2d(cin, cout, k, stride, k // 2, groups=groups, bias=False), nn.BatchNorm2d(cout)]
    if act:
        mods.append(nn.ReLU6(inplace=True))
    return mods


class InvertedResidual(nn.Module):
    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        assert stride in [1, 2]
        self.stride = stride
        hidden = round(inp * expand_ratio)
        self.use_res_connect = self.stride == 1 and inp == oup
        self.expand_ratio = expand_ratio
        mods = [] if expand_ratio == 1 else _cbr(inp, hidden, 1, 1)            # pw
        mods += _cbr(hidden, hidden, 3, stride, groups=hidden)                 # dw
        mods += _cbr(hidden, oup, 1, 1, act=False)                             # pw-linear
        self.conv = nn.Sequential(*mods)

    def forward(self, x):
        return x + self.conv(x) if self.use_res_connect else self.conv(x)


class MobileNetV2(nn.Module):
    SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]

    def __init__(self, n_class=1000, input_size=224, width_mult=1., dropout=0.0):
        super().__init__()
        assert input_size % 32 == 0
        cin = int(32 * width_mult)
        self.last_channel = int(1280 * width_mult) if width_mult > 1.0 else 1280
        feats = [nn.Sequential(*_cbr(3, cin, 3, 2))]
        for t, c, n, s in self.SETTING:
            cout = int(c * width_mult)
            for i in range(n):
                feats.append(InvertedResidual(cin, cout, s if i == 0 else 1, expand_ratio=t))
                cin = cout
        feats.append(nn.Sequential(*_cbr(cin, self.last_channel, 1, 1)))
        self.features = nn.Sequential(*feats)
        self.classifier = nn.Sequential(nn.Dropout(dropout), nn.Linear(self.last_channel, n_class))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, math.sqrt(2. / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    def forward(self, x):
        return self.classifier(self.features(x).mean([2, 3]))


def mobilenetv2(**kw):
    return MobileNetV2(**kw)
