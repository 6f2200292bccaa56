"""BatchNorm folding, host-side one-off (reference: quant/fold_bn.py:14-92). Plain torch on purpose:
it runs once per model, before any quantiser exists."""
import torch
import torch.nn as nn
import torch.nn.init as init


from .quant_layer import StraightThrough  # noqa: E402  (upstream defines a second identical class here)


def _fold_bn(conv_module, bn_module):
    """(w, b) of conv∘bn as a single conv  (fold_bn.py:14-34)"""
    w = conv_module.weight.data
    mean, var = bn_module.running_mean, bn_module.running_var
    safe_std = torch.sqrt(var + bn_module.eps)
    per_oc = (conv_module.out_channels, 1, 1, 1)
    conv_bias = conv_module.bias
    if bn_module.affine:
        gain = bn_module.weight / safe_std
        weight = w * gain.view(per_oc)
        shift = bn_module.bias - bn_module.weight * mean / safe_std
        bias = shift if conv_bias is None else bn_module.weight * conv_bias / safe_std + shift
    else:
        weight = w / safe_std.view(per_oc)
        shift = -mean / safe_std
        bias = shift if conv_bias is None else conv_bias / safe_std + shift
    return weight, bias


def fold_bn_into_conv(conv_module, bn_module):
    w, b = _fold_bn(conv_module, bn_module)
    if conv_module.bias is None:
        conv_module.bias = nn.Parameter(b)
    else:
        conv_module.bias.data = b
    conv_module.weight.data = w
    # leave the BN an identity-equivalent in case it is still called (fold_bn.py:44-46)
    bn_module.running_mean = bn_module.bias.data
    bn_module.running_var = bn_module.weight.data ** 2


def reset_bn(module: nn.BatchNorm2d):
    if module.track_running_stats:
        module.running_mean.zero_()
        module.running_var.fill_(1 - module.eps)
    if module.affine:
        init.ones_(module.weight)
        init.zeros_(module.bias)


def is_bn(m):
    return isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d))


def is_absorbing(m):
    return isinstance(m, (nn.Conv2d, nn.Linear))


def search_fold_and_remove_bn(model):
    """depth-first: a BN directly following a conv/linear among the same parent's children is folded and
    replaced by StraightThrough (fold_bn.py:67-79). Returns the last absorbing module seen."""
    model.eval()
    prev = None
    for name, child in model.named_children():
        if is_bn(child) and is_absorbing(prev):
            fold_bn_into_conv(prev, child)
            setattr(model, name, StraightThrough())
        elif is_absorbing(child):
            prev = child
        else:
            prev = search_fold_and_remove_bn(child)
    return prev


def search_fold_and_reset_bn(model):
    model.eval()
    prev = None
    for _name, child in model.named_children():
        if is_bn(child) and is_absorbing(prev):
            fold_bn_into_conv(prev, child)
        else:
            search_fold_and_reset_bn(child)
        prev = child
