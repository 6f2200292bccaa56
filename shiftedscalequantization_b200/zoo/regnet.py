"""RegNetX (structure of upstream models/regnet.py:30-330: stem.conv/bn/relu, s{i}.b{j} blocks with
f.a/a_bn/a_relu/b/b_bn/b_relu/c/c_bn, optional proj/bn, relu; head.fc)."""
import math

import numpy as np
import torch.nn as nn

CONFIGS = {
    'regnetx_600m': dict(WA=36.97, W0=48, WM=2.24, DEPTH=16, GROUP_W=24),
    'regnetx_3200m': dict(WA=26.31, W0=88, WM=2.25, DEPTH=25, GROUP_W=48),
}


class SimpleStemIN(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=2, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(True)

    def forward(self, x):
        return self.relu(self.bn(self.conv(x)))


class BottleneckTransform(nn.Module):
    def __init__(self, w_in, w_out, stride, bm, gw, se_r=None):
        super().__init__()
        w_b = int(round(w_out * bm))
        self.a = nn.Conv2d(w_in, w_b, 1, bias=False)
        self.a_bn = nn.BatchNorm2d(w_b)
        self.a_relu = nn.ReLU(True)
        self.b = nn.Conv2d(w_b, w_b, 3, stride=stride, padding=1, groups=w_b // gw, bias=False)
        self.b_bn = nn.BatchNorm2d(w_b)
        self.b_relu = nn.ReLU(True)
        self.c = nn.Conv2d(w_b, w_out, 1, bias=False)
        self.c_bn = nn.BatchNorm2d(w_out)
        self.c_bn.final_bn = True

    def forward(self, x):
        for layer in self.children():
            x = layer(x)
        return x


class ResBottleneckBlock(nn.Module):
    def __init__(self, w_in, w_out, stride, bm=1.0, gw=1, se_r=None):
        super().__init__()
        self.proj_block = (w_in != w_out) or (stride != 1)
        if self.proj_block:
            self.proj = nn.Conv2d(w_in, w_out, 1, stride=stride, bias=False)
            self.bn = nn.BatchNorm2d(w_out)
        self.f = BottleneckTransform(w_in, w_out, stride, bm, gw, se_r)
        self.relu = nn.ReLU(True)

    def forward(self, x):
        skip = self.bn(self.proj(x)) if self.proj_block else x
        return self.relu(skip + self.f(x))


class AnyStage(nn.Module):
    def __init__(self, w_in, w_out, stride, d, bm, gw):
        super().__init__()
        for i in range(d):
            self.add_module(f'b{i + 1}', ResBottleneckBlock(w_in if i == 0 else w_out, w_out, stride if i == 0 else 1, bm, gw))

    def forward(self, x):
        for blk in self.children():
            x = blk(x)
        return x


class AnyHead(nn.Module):
    def __init__(self, w_in, nc):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(w_in, nc, bias=True)

    def forward(self, x):
        return self.fc(self.avg_pool(x).flatten(1))


def _stage_plan(cfg, q=8):
    """per-stage (width, depth) from the RegNet linear parameterisation (arXiv:2003.13678 eq. 2-4)"""
    ws_cont = np.arange(cfg['DEPTH']) * cfg['WA'] + cfg['W0']
    ks = np.round(np.log(ws_cont / cfg['W0']) / np.log(cfg['WM']))
    ws = (np.round(cfg['W0'] * np.power(cfg['WM'], ks) / q) * q).astype(int).tolist()
    widths, depths = [], []
    for w in ws:
        if widths and widths[-1] == w:
            depths[-1] += 1
        else:
            widths.append(w); depths.append(1)
    gws = [min(cfg['GROUP_W'], w) for w in widths]
    widths = [int(round(w / g) * g) for w, g in zip(widths, gws)]      # widths divisible by the group width
    return widths, depths, gws


class RegNet(nn.Module):
    def __init__(self, cfg, nc=1000):
        super().__init__()
        widths, depths, gws = _stage_plan(cfg)
        self.stem = SimpleStemIN(3, 32)
        prev = 32
        for i, (w, d, g) in enumerate(zip(widths, depths, gws)):
            self.add_module(f's{i + 1}', AnyStage(prev, w, 2, d, 1.0, g))
            prev = w
        self.head = AnyHead(prev, nc)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0.0, math.sqrt(2.0 / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 1.0 / float(m.weight.size(1)))
                m.bias.data.zero_()

    def forward(self, x):
        for module in self.children():
            x = module(x)
        return x


def regnetx_600m(**kw):
    return RegNet(CONFIGS['regnetx_600m'], **kw)


def regnetx_3200m(**kw):
    return RegNet(CONFIGS['regnetx_3200m'], **kw)
