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
        self._init_weight_scales_together()
        return self.model(input)

    def _init_weight_scales_together(self):
        """The per-channel MSE scale search of a layer depends on its weights only (quant_layer.py:100-166), so the searches of
        every layer that is about to initialise itself in this forward are issued together, round-robin on side streams: a layer
        has 64-512 output channels, i.e. one launch fills a fraction of the GPU. Each quantiser ends up exactly as its own lazy
        initialisation would leave it (same kernel, same inputs: bit-identical delta / zero point). One host check for all
        layers instead of one per layer; anything unusual (an all-zero channel, several ranks, per-tensor or `max` scales, rows
        too long for the row kernel) is left to the lazy per-layer path, which raises what upstream raises."""
        import torch
        from .. import dist as ssq_dist, ops
        from .quant_layer import UniformAffineQuantizer
        todo = []
        for m in self._quant_modules():
            q = m.weight_quantizer
            if (m.use_weight_quant and m.cache_features == 'none' and m._engine_weight is None and type(q) is UniformAffineQuantizer
                    and q.inited is False and q.scale_method == 'mse' and q.channel_wise and m.weight.is_cuda
                    and m.weight.dtype == torch.float32 and 0 < m.weight[0].numel() <= 48 * 1024):
                todo.append(m)
        if len(todo) < 2 or ssq_dist.world_size() > 1:
            return
        outs = ops.mse_scale_search_many([(m.weight.detach().reshape(m.weight.shape[0], -1), m.weight_quantizer.n_levels,
                                           m.weight_quantizer.sym) for m in todo])
        if bool(torch.stack([(o[4] < 0).any() for o in outs]).any()):
            return                                              # an all-NaN-score channel somewhere: the lazy path raises at that layer
        for m, (delta, zp, raw, _score, _idx) in zip(todo, outs):
            q = m.weight_quantizer
            shape = (-1, 1, 1, 1) if m.weight.dim() == 4 else (-1, 1)
            q.delta = nn.Parameter(delta.view(shape))
            q.zero_point = nn.Parameter(zp.view(shape))
            q.raw_zero_point = raw.view(shape)
            q.inited = True

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
        """all-average of the activation step sizes across ranks (upstream intent, commented out at quant_model.py:78-83 but
        still called by Brecq/main_imagenet_dist.py:211). Upstream's text walks the QuantModules only; a block's own output
        quantiser (quant_block.py:32) is initialised from the rank's images in exactly the same way, and replicas that are to
        stay identical through the activation phase need it averaged too — so every initialised activation quantiser of the
        model is covered, and its zero point (an integer per rank: post-ReLU tensors give 0 everywhere) is averaged and
        rounded."""
        from ..dist import all_average_
        from .quant_layer import UniformAffineQuantizer
        seen = set()
        for m in self._units():
            q = getattr(m, 'act_quantizer', None)
            if isinstance(q, UniformAffineQuantizer) and q.delta is not None and q.inited and id(q) not in seen:
                seen.add(id(q))
                all_average_(q.delta.data)
                if q.zero_point is not None:
                    all_average_(q.zero_point.data)
                    q.zero_point.data.round_()

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
