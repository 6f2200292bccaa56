"""Drop-in `quant` package (same exports as the reference's quant/__init__.py:1-5)."""
from .block_recon import block_reconstruction
from .layer_recon import layer_reconstruction
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quant_model import QuantModel
