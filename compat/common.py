"""`common` for the reference's drivers (upstream common.py:19-298): flags, seeding, meters, validation.
icecream is optional here (upstream imports it unconditionally and disables it)."""
import argparse
import os
import random
import time
from datetime import datetime

import numpy as np
import torch

try:  # upstream: `from icecream import ic; ic.disable()`
    from icecream import ic
    ic.configureOutput(includeContext=True)
    ic.disable()
except Exception:  # pragma: no cover
    def ic(*a, **k):
        return a[0] if a else None


def time_format():
    return f'{datetime.now():%Y-%m-%d %H:%M:%S}'


def loadArgments():
    """same flag names/defaults as upstream (note: `type=bool` flags treat any non-empty string as True there too)"""
    p = argparse.ArgumentParser(description='running parameters', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    add = p.add_argument
    add('--seed', default=1005, type=int, help='random seed for results reproduction')
    add('--arch', default='resnet18', type=str,
        choices=['resnet18', 'resnet50', 'mobilenetv2', 'regnetx_600m', 'regnetx_3200m', 'mnasnet'])
    add('--batch_size', default=64, type=int, help='mini-batch size for data loader')
    add('--workers', default=4, type=int, help='number of workers for data loader')
    add('--data_path', default='~/dataset/cifar10', type=str, required=False)
    add('--n_bits_w', default=2, type=int, help='bitwidth for weight quantization')
    add('--channel_wise', default=True, type=bool, help='apply channel_wise quantization for weights')
    add('--n_bits_a', default=4, type=int, help='bitwidth for activation quantization')
    add('--act_quant', default=True, type=bool, help='apply activation quantization')
    add('--disable_8bit_head_stem', default=False, type=bool)
    add('--test_before_calibration', default=True, type=bool)
    add('--num_samples', default=1024, type=int, help='size of the calibration dataset')
    add('--iters_w', default=20000, type=int, help='number of iteration for adaround')
    add('--weight', default=0.01, type=float, help='weight of rounding cost vs the reconstruction loss.')
    add('--sym', default=True, type=bool, help='symmetric reconstruction, not recommended')
    add('--b_start', default=20, type=int, help='temperature at the beginning of calibration')
    add('--b_end', default=2, type=int, help='temperature at the end of calibration')
    add('--warmup', default=0.2, type=float, help='in the warmup period no regularization is applied')
    add('--step', default=20, type=int, help='record snn output per step')
    add('--iters_a', default=5000, type=int, help='number of iteration for LSQ')
    add('--lr', default=4e-4, type=float, help='learning rate for LSQ')
    add('--p', default=2.4, type=float, help='L_p norm minimization for LSQ')
    add('--make_checkpoint', default=False, type=bool, help='generate checkpoint')
    add('--skip_test', default=False, type=bool, help='skip default test')
    add('--run_device', default='cuda:0', type=str, help='gpu usage')
    add('--msg_bot_enable', default=True, type=bool, help='use messaging bot for monitoring')
    add('--make_init_data', default=False, type=bool, help='Make Initiallize weight data')
    add('--dataset', default='cifar10', type=str, help='dataset name')
    add('--bypassChannelShift', default=False, type=bool, help='do not run channel shift function')
    add('--mse_level', default=1, type=int, help='1, 2, 4, ...')
    add('--mse_threshold', default=1.0, type=float, help='how much the rounding range is widened')
    add('--shift_quant_mode', default='max', type=str, help='mse or max')
    add('--w_scale_method', default='mse', type=str, help='mse or max')
    add('--a_scale_method', default='mse', type=str, help='mse or max')
    add('--test', default=False, type=bool, help='test')
    return p.parse_args()


def seed_all(seed=1029):
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True


class AverageMeter(object):
    def __init__(self, name, fmt=':f'):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return ('{name} {val' + self.fmt + '} ({avg' + self.fmt + '})').format(**self.__dict__)


class ProgressMeter(object):
    def __init__(self, num_batches, meters, prefix=""):
        digits = len(str(num_batches // 1))
        self.batch_fmtstr = '[{:' + str(digits) + 'd}/' + ('{:' + str(digits) + 'd}').format(num_batches) + ']'
        self.meters, self.prefix = meters, prefix

    def display(self, batch):
        print('\t'.join([self.prefix + self.batch_fmtstr.format(batch)] + [str(m) for m in self.meters]))


def accuracy(output, target, topk=(1,)):
    with torch.no_grad():
        maxk = max(topk)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
        return [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0)) for k in topk]


def get_train_samples(train_loader, num_samples):
    chunks, have = [], 0
    for batch in train_loader:
        chunks.append(batch[0])
        have += batch[0].size(0)
        if have >= num_samples:
            break
    return torch.cat(chunks, dim=0)[:num_samples]


@torch.no_grad()
def validate_model(val_loader, model, device=None, print_freq=100, print_result=False, simple=False, bit=-1):
    if device is None:
        device = next(model.parameters()).device
    else:
        model.to(device)
    batch_time, top1, top5 = AverageMeter('Time', ':6.3f'), AverageMeter('Acc@1', ':6.2f'), AverageMeter('Acc@5', ':6.2f')
    progress = ProgressMeter(len(val_loader), [batch_time, top1, top5], prefix='Test: ')
    model.eval()
    outputs, end = [], time.time()
    for i, (images, target) in enumerate(val_loader):
        images, target = images.to(device), target.to(device)
        output = model(images)
        if bit != -1:
            outputs.append(output)
        acc1, acc5 = accuracy(output, target, topk=(1, 5))
        top1.update(acc1[0], images.size(0))
        top5.update(acc5[0], images.size(0))
        batch_time.update(time.time() - end)
        end = time.time()
        if i % print_freq == 0 and print_result:
            progress.display(i)
        if simple and i > 5:
            break
    if bit != -1:   # needs ./output_loss/result_{bit}bit.pt, which upstream does not ship either
        ref_out = torch.load(f'./output_loss/result_{bit}bit.pt')
        print(f'MSE[{bit}] : {torch.mean(torch.square(torch.cat(outputs, dim=0) - ref_out)):.2e}')
    if print_result:
        print(' * Acc@1 {top1.avg:.3f}'.format(top1=top1))
    return top1.avg


@torch.no_grad()
def validate_with_loss(val_loader, model, device=None, print_freq=100, print_result=False, simple=False, bit=-1):
    return validate_model(val_loader, model, device, print_freq, print_result, simple, bit), 0


def print_model_hierarchy(model, depth=0):
    for name, child in model.named_children():
        print("--" * depth, name)
        print_model_hierarchy(child, depth + 1)
