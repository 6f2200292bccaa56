"""Quantised residual-block wrappers — mirror of the reference's quant/quant_block.py (BaseQuantBlock :13-73,
QuantBasicBlock :76-130, QuantBottleneck :133-166, QuantResBottleneckBlock :169-202,
QuantInvertedResidual :205-239, specials :242-248).

A block is a reconstruction unit: its QuantModules share one calibration loop, and its tail
(residual add -> activation -> activation quantiser) lives here because of the branch structure.
Unlike upstream, setPathName/toggleHardTarget exist on the base class, so ResNet-50, RegNetX and
MobileNetV2 construct without the AttributeError upstream raises at quant_model.py:30.
"""
import torch.nn as nn

from ..zoo.mobilenetv2 import InvertedResidual
from ..zoo.regnet import ResBottleneckBlock
from ..zoo.resnet import BasicBlock, Bottleneck
from .quant_layer import QuantModule, StraightThrough, UniformAffineQuantizer


class BasicBlockCIFAR(BasicBlock):
    """stand-in for pretrained.PyTorch_CIFAR10.cifar10_models.resnet.BasicBlockCIFAR (quant_block.py:11):
    the CIFAR config builds its ResNet from the same BasicBlock."""


class BaseQuantBlock(nn.Module):
    """Shared state of all block wrappers: quant-state flags, the block-tail activation quantiser and the
    feature caches used by the shifted-scale loops."""

    def __init__(self, act_quant_params: dict = {}):
        super().__init__()
        self.use_weight_quant = False
        self.use_act_quant = False
        act_quant_params['disable_act_quant'] = False
        self.act_quantizer = UniformAffineQuantizer(**act_quant_params)
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False
        self.cache_features = 'none'
        self.cached_inp_features = []
        self.cached_out_features = []
        self.selectionInited = False
        self.pathName = ''

    def quant_modules(self):
        return [m for m in self.modules() if isinstance(m, QuantModule)]

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.quant_modules():
            m.set_quant_state(weight_quant, act_quant)

    def set_quant_init_state(self):
        for m in self.quant_modules():
            m.set_quant_init_state()

    def set_quant_state_block(self, state, act=False):
        for m in self.quant_modules():
            if act:
                m.use_act_quant = state
            else:
                m.use_weight_quant = state

    # upstream defines these twice; the later, block-local definitions win (quant_block.py:68-73)
    def disable_cache_features(self):
        self.cache_features = 'none'

    def clear_cached_features(self):
        self.cached_inp_features = []
        self.cached_out_features = []

    def setPathName(self, curName):
        self.pathName = curName
        for name, m in self.named_modules():
            if isinstance(m, QuantModule):
                m.pathName = curName + '.' + name

    def toggleHardTarget(self):
        for m in self.quant_modules():
            m.weight_quantizer.hard_targets = not m.weight_quantizer.hard_targets

    # block tail shared by every residual wrapper
    def _tail(self, out):
        out = self.activation_function(out)
        if self.use_act_quant:
            out = self.act_quantizer(out)
        return out

    def _cache_in(self, x):
        if self.cache_features == 'if':
            self.cached_inp_features += [x.to('cpu').clone().detach()]

    def _cache_out(self, out):
        if self.cache_features == 'of':
            self.cached_out_features += [out.to('cpu').clone().detach()]


def _qm(conv, wq, aq, act=None, last=False):
    m = QuantModule(conv, wq, aq, disable_act_quant=last)
    if act is not None:
        m.activation_function = act
    return m


class QuantBasicBlock(BaseQuantBlock):
    """ResNet-18/34 block: conv1(+relu) -> conv2, + (downsample(x) | x), relu, act-quant."""

    def __init__(self, basic_block: BasicBlock, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.conv1 = _qm(basic_block.conv1, weight_quant_params, act_quant_params, basic_block.relu1)
        self.conv2 = _qm(basic_block.conv2, weight_quant_params, act_quant_params, last=True)
        self.activation_function = basic_block.relu2
        self.downsample = None if basic_block.downsample is None else \
            _qm(basic_block.downsample[0], weight_quant_params, act_quant_params, last=True)
        self.stride = basic_block.stride
        self.dump_cnt = 0

    def forward(self, x):
        self._cache_in(x)
        residual = x if self.downsample is None else self.downsample(x)
        out = self.conv2(self.conv1(x))
        out += residual
        out = self._tail(out)
        self._cache_out(out)
        return out

    def setPathName(self, curName):
        self.pathName = curName
        self.conv1.pathName = curName + '.conv1'
        self.conv2.pathName = curName + '.conv2'
        if self.downsample is not None:
            self.downsample.pathName = curName + '.downsample'


class QuantBottleneck(BaseQuantBlock):
    """ResNet-50/101/152 block: 1x1 -> 3x3 -> 1x1."""

    def __init__(self, bottleneck: Bottleneck, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.conv1 = _qm(bottleneck.conv1, weight_quant_params, act_quant_params, bottleneck.relu1)
        self.conv2 = _qm(bottleneck.conv2, weight_quant_params, act_quant_params, bottleneck.relu2)
        self.conv3 = _qm(bottleneck.conv3, weight_quant_params, act_quant_params, last=True)
        self.activation_function = bottleneck.relu3
        self.downsample = None if bottleneck.downsample is None else \
            _qm(bottleneck.downsample[0], weight_quant_params, act_quant_params, last=True)
        self.stride = bottleneck.stride

    def forward(self, x):
        self._cache_in(x)               # upstream caches block features in QuantBasicBlock only (quant_block.py:100,115)
        residual = x if self.downsample is None else self.downsample(x)
        out = self.conv3(self.conv2(self.conv1(x)))
        out += residual
        out = self._tail(out)
        self._cache_out(out)
        return out


class QuantResBottleneckBlock(BaseQuantBlock):
    """RegNetX block (no SE): f.a -> f.b (grouped 3x3) -> f.c, optional projection on the skip."""

    def __init__(self, bottleneck: ResBottleneckBlock, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.conv1 = _qm(bottleneck.f.a, weight_quant_params, act_quant_params, bottleneck.f.a_relu)
        self.conv2 = _qm(bottleneck.f.b, weight_quant_params, act_quant_params, bottleneck.f.b_relu)
        self.conv3 = _qm(bottleneck.f.c, weight_quant_params, act_quant_params, last=True)
        self.activation_function = bottleneck.relu
        self.proj_block = bottleneck.proj_block
        self.downsample = _qm(bottleneck.proj, weight_quant_params, act_quant_params, last=True) if self.proj_block else None

    def forward(self, x):
        self._cache_in(x)
        residual = self.downsample(x) if self.proj_block else x
        out = self.conv3(self.conv2(self.conv1(x)))
        out += residual
        out = self._tail(out)
        self._cache_out(out)
        return out


class QuantInvertedResidual(BaseQuantBlock):
    """MobileNetV2 block; no activation after the (optional) residual add."""

    def __init__(self, inv_res: InvertedResidual, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.use_res_connect = inv_res.use_res_connect
        self.expand_ratio = inv_res.expand_ratio
        convs = [inv_res.conv[i] for i in ((0, 3) if self.expand_ratio == 1 else (0, 3, 6))]
        mods = [_qm(c, weight_quant_params, act_quant_params, None if i == len(convs) - 1 else nn.ReLU6(),
                    last=(i == len(convs) - 1)) for i, c in enumerate(convs)]
        self.conv = nn.Sequential(*mods)

    def forward(self, x):
        self._cache_in(x)
        out = x + self.conv(x) if self.use_res_connect else self.conv(x)
        out = self._tail(out)
        self._cache_out(out)
        return out


specials = {
    BasicBlock: QuantBasicBlock,
    BasicBlockCIFAR: QuantBasicBlock,
    Bottleneck: QuantBottleneck,
    ResBottleneckBlock: QuantResBottleneckBlock,
    InvertedResidual: QuantInvertedResidual,
}
