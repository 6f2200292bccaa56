"""QuantModel — mirror of the reference's quant/quant_model.py:7-106: BN folding, recursive replacement of
Conv2d/Linear by QuantModule and of residual blocks by their quantised wrappers, quant-state toggles."""
import torch.nn as nn

from .fold_bn import search_fold_and_remove_bn
from .quant_block import BaseQuantBlock, specials
from .quant_layer import QuantModule, StraightThrough

_UNITS = (QuantModule, BaseQuantBlock)


class QuantModel(nn.Module):
    def __init__(self, model: nn.Module, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        super().__init__()
        search_fold_and_remove_bn(model)
        self.model = model
        self.quant_module_refactor(self.model, weight_quant_params, act_quant_params)
        self.qState = []

    def quant_module_refactor(self, module: nn.Module, weight_quant_params: dict = {}, act_quant_params: dict = {},
                              depth=0, moduleName=''):
        """Walk the children once: special blocks -> wrapper, conv/linear -> QuantModule, a ReLU/ReLU6 that
        follows a QuantModule is absorbed into it (quant_model.py:15-44; children literally named 'relu2' are
        skipped as upstream does)."""
        last_qm = None
        for name, child in module.named_children():
            path = moduleName + '.' + name
            if name in ['relu2']:
                continue
            if type(child) in specials:
                wrapped = specials[type(child)](child, weight_quant_params, act_quant_params)
                setattr(module, name, wrapped)
                wrapped.setPathName(path)
            elif isinstance(child, (nn.Conv2d, nn.Linear)):
                last_qm = QuantModule(child, weight_quant_params, act_quant_params)
                last_qm.pathName = path
                setattr(module, name, last_qm)
            elif isinstance(child, (nn.ReLU, nn.ReLU6)):
                if last_qm is not None:
                    last_qm.activation_function = child
                    setattr(module, name, StraightThrough())
            elif isinstance(child, StraightThrough):
                continue
            else:
                self.quant_module_refactor(child, weight_quant_params, act_quant_params, depth + 1, moduleName=path)

    def _units(self):
        return [m for m in self.model.modules() if isinstance(m, _UNITS)]

    def _quant_modules(self):
        return [m for m in self.model.modules() if isinstance(m, QuantModule)]

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        for m in self._units():
            m.set_quant_state(weight_quant, act_quant)

    def set_quant_init_state(self):
        for m in self._units():
            m.set_quant_init_state()

    def forward(self, input):
        return self.model(input)

    def set_first_last_layer_to_8bit(self):
        """8-bit stem/head (quant_model.py:59-69): first layer weights+acts, last layer weights, and the
        activation feeding the last layer; the first layer is excluded from reconstruction."""
        mods = self._quant_modules()
        mods[0].weight_quantizer.bitwidth_refactor(8)
        mods[0].act_quantizer.bitwidth_refactor(8)
        mods[-1].weight_quantizer.bitwidth_refactor(8)
        mods[-2].act_quantizer.bitwidth_refactor(8)
        mods[0].ignore_reconstruction = True

    def disable_network_output_quantization(self):
        self._quant_modules()[-1].disable_act_quant = True

    def synchorize_activation_statistics(self):
        """all-average of the activation step sizes across ranks (upstream intent, commented out at
        quant_model.py:78-83 but still called by Brecq/main_imagenet_dist.py:211)."""
        from ..dist import all_average_
        for m in self._quant_modules():
            if m.act_quantizer.delta is not None:
                all_average_(m.act_quantizer.delta.data)

    def disable_cache_features(self):
        for m in self._units():
            m.disable_cache_features()

    def clear_cached_features(self):
        for m in self._units():
            m.clear_cached_features()

    def store_quantization_state(self):
        self.qState = [m.use_weight_quant for m in self.modules() if isinstance(m, QuantModule)]

    def restore_quantization_state(self):
        for m, state in zip([m for m in self.modules() if isinstance(m, QuantModule)], self.qState):
            m.use_weight_quant = state
