"""names the reference's ShiftedScaleQuant.py pulls in with `from myScaledMethods import *` (upstream myScaledMethods.py)"""
from quant.quant_model import QuantModel  # noqa: F401
from quant.quant_layer import QuantModule, UniformAffineQuantizer  # noqa: F401
from quant.quant_block import BaseQuantBlock, QuantBasicBlock  # noqa: F401
from quant.channelQuant import ChannelQuant  # noqa: F401
from quant.channelQuantMSE import ChannelQuantMSE  # noqa: F401
from common import *  # noqa: F401,F403
from shiftedscalequantization_b200.scaled_methods import (  # noqa: F401
    QuantRecursiveShiftRecon, build_ShiftedChannelQuant, build_ShiftedChannelQuantBlock, build_ShiftedChannelQuantLayer,
    build_ShiftedChannelQuantMSE, build_ShiftedChannelQuantMSEBlock, build_ShiftedChannelQuantMSELayer,
    build_qnn_from_model, channelShift_wLoss_flow, channelShift_wMSE_flow, run_ShiftReconFused, set_cache_state,
    set_quant_state_block, toggle_hardTarget)
