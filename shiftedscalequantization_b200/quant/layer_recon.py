"""layer_reconstruction — mirror of the reference's quant/layer_recon.py:10-104; shares the loop engine and
the loss with block_recon (upstream duplicates LossFunction verbatim at layer_recon.py:107-168)."""
import torch

from .block_recon import LinearTempDecay, LossFunction, reconstruct_unit  # noqa: F401  (re-exported names)
from .quant_layer import QuantModule
from .quant_model import QuantModel


def layer_reconstruction(model: QuantModel, layer: QuantModule, cali_data: torch.Tensor,
                         batch_size: int = 32, iters: int = 20000, weight: float = 0.001, opt_mode: str = 'mse',
                         asym: bool = False, include_act_func: bool = True, b_range: tuple = (20, 2),
                         warmup: float = 0.0, act_quant: bool = False, lr: float = 4e-5, p: float = 2.0,
                         multi_gpu: bool = False, eval: bool = False, bias_cal: bool = False, scaling: str = 'weak',
                         host_resident: bool = False):
    """Single-layer variant of block_reconstruction (first/last layers and layers outside any block)."""
    reconstruct_unit(model, layer, cali_data, is_block=False, batch_size=batch_size, iters=iters, weight=weight,
                     opt_mode=opt_mode, asym=asym, include_act_func=include_act_func, b_range=b_range, warmup=warmup,
                     act_quant=act_quant, lr=lr, p=p, multi_gpu=multi_gpu, eval=eval, bias_cal=bias_cal, scaling=scaling,
                     host_resident=host_resident)
