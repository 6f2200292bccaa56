"""`quant` as the reference's drivers import it: every sub-module name resolves to the B200-native mirror."""
import importlib
import sys

_PKG = "shiftedscalequantization_b200.quant"
for _name in ("quant_layer", "quant_block", "quant_model", "adaptive_rounding", "channelQuant", "channelQuantMSE",
              "channelQuantAct", "fold_bn", "data_utils", "block_recon", "layer_recon", "layer_recon_shiftedScale",
              "layer_recon_fused_shiftedScale"):
    sys.modules[f"{__name__}.{_name}"] = importlib.import_module(f"{_PKG}.{_name}")

from shiftedscalequantization_b200.quant import (BaseQuantBlock, QuantModel, QuantModule,  # noqa: E402,F401
                                                   block_reconstruction, layer_reconstruction)
