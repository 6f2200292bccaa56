"""Callers of the shifted-scale quantisers — the builder / cache-state helpers of the reference's myScaledMethods.py
(:17-120, :263-307) and the two calibration flows of ShiftedScaleQuant.py (channelShift_wMSE :119-183,
channelShift_wLoss :185-286, QuantRecursiveShiftRecon :12-29, run_ShiftReconFused :48-59), without the experiment
glue around them (Telegram bot, hard-coded checkpoint paths and devices, pickled intermediate files)."""
import torch
import torch.nn as nn

from .quant.channelQuant import ChannelQuant
from .quant.channelQuantMSE import ChannelQuantMSE
from .quant.layer_recon_fused_shiftedScale import block_recon_fused_shiftedScale, layer_recon_fused_shiftedScale
from .quant.quant_block import BaseQuantBlock, QuantBasicBlock
from .quant.quant_layer import QuantModule, UniformAffineQuantizer
from .quant.quant_model import QuantModel


# ------------------------------------------------------------------------------------------------ builders
def build_ShiftedChannelQuantMSELayer(model, curName, layer, delta=1.0, **kwargs):
    layer.weight_quantizer = ChannelQuantMSE(delta, uaq=layer.weight_quantizer, weight_tensor=layer.org_weight.data,
                                             shiftTarget=kwargs['shiftTarget'], opt_mode=kwargs['opt_mode'],
                                             level=kwargs['level'], threshold=kwargs['threshold'], name=curName)
    layer.use_weight_quant = True
    layer.cache_features = 'none'
    layer.weight_quantizer.init_scale(layer.org_weight.data)


def build_ShiftedChannelQuantMSEBlock(model, prv_name, block, delta=1.0, **kwargs):
    for name, layer in block.named_children():
        if isinstance(layer, QuantModule) and isinstance(layer.weight_quantizer, UniformAffineQuantizer):
            build_ShiftedChannelQuantMSELayer(model, prv_name + '.' + name, layer, delta, **kwargs)


def build_ShiftedChannelQuantMSE(model: nn.Module, layerDisabled, prv_name="", delta=1.0, **kwargs):
    """every reconstructable QuantModule not listed in layerDisabled gets a ChannelQuantMSE with its input scale
    searched (ShiftedScaleQuant.py:154-174; blocks listed in layerDisabled are built whole, as upstream)"""
    for name, module in model.named_children():
        curName = prv_name + '.' + name
        if isinstance(module, QuantModule):
            if module.ignore_reconstruction is True:
                continue
            if curName not in layerDisabled:
                build_ShiftedChannelQuantMSELayer(model, curName, module, delta, **kwargs)
        elif isinstance(module, QuantBasicBlock):
            if module.ignore_reconstruction is True:
                continue
            if curName in layerDisabled:
                build_ShiftedChannelQuantMSEBlock(model, curName, module, delta, **kwargs)
            else:
                build_ShiftedChannelQuantMSE(module, layerDisabled, curName, delta, **kwargs)
        else:
            build_ShiftedChannelQuantMSE(module, layerDisabled, curName, delta, **kwargs)


def _shift_target(curName, kwargs):
    skip = tuple(kwargs.get('skipShiftLayer', ()))
    return kwargs['shiftTarget'] if not (skip and curName.startswith(skip)) else [2 / 2]


def build_ShiftedChannelQuantLayer(model, curName, layer, delta=1.0, **kwargs):
    layer.weight_quantizer = ChannelQuant(delta, uaq=layer.weight_quantizer, weight_tensor=layer.org_weight.data,
                                          shiftTarget=_shift_target(curName, kwargs), name=curName)
    layer.use_weight_quant = True
    layer.cache_features = 'none'


def build_ShiftedChannelQuantBlock(model, prv_name, block, delta=1.0, **kwargs):
    for name, layer in block.named_children():
        if isinstance(layer, QuantModule) and isinstance(layer.weight_quantizer, UniformAffineQuantizer):
            build_ShiftedChannelQuantLayer(model, prv_name + '.' + name, layer, delta, **kwargs)


def build_ShiftedChannelQuant(model: nn.Module, layerEnabled, prv_name="", delta=1.0, **kwargs):
    for name, module in model.named_children():
        curName = prv_name + '.' + name
        if isinstance(module, QuantModule):
            if module.ignore_reconstruction is True:
                continue
            if curName in layerEnabled:
                build_ShiftedChannelQuantLayer(model, curName, module, delta, **kwargs)
        elif isinstance(module, QuantBasicBlock):
            if module.ignore_reconstruction is True:
                continue
            if curName in layerEnabled:
                build_ShiftedChannelQuantBlock(model, curName, module, delta, **kwargs)
            else:
                build_ShiftedChannelQuant(module, layerEnabled, curName, delta, **kwargs)
        else:
            build_ShiftedChannelQuant(module, layerEnabled, curName, delta, **kwargs)


# ------------------------------------------------------------------------------------------------ state helpers
def set_quant_state_block(model, layers, prv_name='', state=False, act=False):
    for name, module in model.named_children():
        curName = prv_name + '.' + name
        if isinstance(module, QuantModule):
            if module.ignore_reconstruction is True:
                continue
            if curName in layers:
                if act:
                    module.use_act_quant = state
                else:
                    module.use_weight_quant = state
        elif isinstance(module, BaseQuantBlock):
            if module.ignore_reconstruction is True:
                continue
            if curName in layers:
                module.set_quant_state_block(state, act)
        else:
            set_quant_state_block(module, layers, curName, state, act)


def set_cache_state(model, layers, prv_name='', state='none'):
    for name, module in model.named_children():
        curName = prv_name + '.' + name
        if curName in layers:
            if module.ignore_reconstruction is True:
                continue
            module.cache_features = state
        elif isinstance(module, QuantModule):
            continue
        else:
            set_cache_state(module, layers, curName, state)


def toggle_hardTarget(model, curName, layer, **kwargs):
    layer.weight_quantizer.hard_targets = not layer.weight_quantizer.hard_targets


# ------------------------------------------------------------------------------------------------ recon dispatch
def run_ShiftReconFused(model, curName, module, qnn, test_loader, act=False, **kwargs):
    iters = kwargs['iters']
    if isinstance(module, QuantModule):
        loss = layer_recon_fused_shiftedScale(module, iters, (0.01, kwargs['lmda']), qnn, test_loader, act=act)
    elif isinstance(module, QuantBasicBlock):
        loss = block_recon_fused_shiftedScale(module, iters, (0.01, kwargs['lmda']), qnn, test_loader, act=act)
    else:
        raise ValueError('Not supported reconstruction module type: {}'.format(type(module)))
    return [loss]


def QuantRecursiveShiftRecon(model: nn.Module, layerEnabled, qnn, test_loader, prv_name="", ret=dict(), act=False, **kwargs):
    for name, module in model.named_children():
        curName = prv_name + '.' + name
        if isinstance(module, (QuantModule, BaseQuantBlock)):
            if module.ignore_reconstruction is True:
                continue
            if curName in layerEnabled:
                ret[curName] = run_ShiftReconFused(model, curName, module, qnn, test_loader, act, **kwargs)
            elif isinstance(module, BaseQuantBlock):
                QuantRecursiveShiftRecon(module, layerEnabled, qnn, test_loader, curName, ret, act, **kwargs)
        else:
            QuantRecursiveShiftRecon(module, layerEnabled, qnn, test_loader, curName, ret, act, **kwargs)
    return ret


# ------------------------------------------------------------------------------------------------ flows
def build_qnn_from_model(cnn, n_bits_w=2, n_bits_a=4, channel_wise=True, w_scale_method='max', a_scale_method='mse',
                         disable_8bit_head_stem=False):
    """the quantiser configuration of myScaledMethods.build_qnn (:279-286) around an already-built FP model"""
    wq = {'n_bits': n_bits_w, 'channel_wise': channel_wise, 'scale_method': w_scale_method, 'tune_delta_zero': False, 'symmetric': False}
    aq = {'n_bits': n_bits_a, 'channel_wise': False, 'scale_method': a_scale_method, 'tune_delta_zero': False, 'leaf_param': True, 'symmetric': False}
    dev = next(cnn.parameters()).device
    qnn = QuantModel(model=cnn, weight_quant_params=wq, act_quant_params=aq).to(dev).eval()
    if not disable_8bit_head_stem:
        qnn.set_first_last_layer_to_8bit()
    return qnn


@torch.no_grad()
def _forward_all(qnn, cali_data, batch_size):
    dev = next(qnn.parameters()).device
    for i in range(len(cali_data) // batch_size):
        qnn(cali_data[i * batch_size:(i + 1) * batch_size].to(dev))


def channelShift_wMSE_flow(qnn, cali_data, level=1, threshold=1.0, shift_quant_mode='max', shiftTarget=2,
                           layerDisabled=('.model.fc',)):
    """ShiftedScaleQuant.channelShift_wMSE without the dataset/validation glue: scale init, then every enabled layer
    gets a ChannelQuantMSE with its input scale searched. Returns the model with weight quantisation on."""
    dev = next(qnn.parameters()).device
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali_data[:64].to(dev))
    build_ShiftedChannelQuantMSE(qnn, list(layerDisabled), '', delta=1.0, shiftTarget=shiftTarget, level=level,
                                 threshold=threshold, opt_mode=shift_quant_mode)
    return qnn


def channelShift_wLoss_flow(qnn, cali_data, layerEnabled, iters, lmda, shiftTarget, batch_size=64, skipShiftLayer=()):
    """ShiftedScaleQuant.channelShift_wLoss (:185-286) without validation: per enabled unit, cache its quantised-path
    inputs ('if') and FP outputs ('of'), then run the fused shift+round reconstruction on them."""
    dev = next(qnn.parameters()).device
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali_data[:64].to(dev))
    kwargs = dict(iters=iters, lmda=lmda, shiftTarget=shiftTarget, skipShiftLayer=list(skipShiftLayer))
    build_ShiftedChannelQuant(qnn, layerEnabled, '', **kwargs)
    qnn.set_quant_state(False, False)
    losses = {}
    for layer in layerEnabled:
        set_cache_state(qnn, [layer], prv_name='', state='if')
        _forward_all(qnn, cali_data, batch_size)
        qnn.store_quantization_state()
        qnn.set_quant_state(False, False)
        set_cache_state(qnn, [layer], prv_name='', state='of')
        _forward_all(qnn, cali_data, batch_size)
        qnn.restore_quantization_state()
        set_cache_state(qnn, [layer], prv_name='', state='none')
        set_quant_state_block(qnn, [layer], '', True)
        QuantRecursiveShiftRecon(qnn, [layer], qnn, None, '', losses, **kwargs)
        qnn.clear_cached_features()
    return qnn, losses
