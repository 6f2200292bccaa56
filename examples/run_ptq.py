#!/usr/bin/env python
"""End-to-end PTQ calibration on synthetic data through the reference's public API sequence
(the flow of Brecq/main_imagenet.py:171-243 / main_cifar10.py:10-104): build QuantModel -> 8-bit stem/head ->
weight-scale init -> block/layer reconstruction of every unit -> activation-scale init -> activation reconstruction.

    python examples/run_ptq.py --arch resnet18 --n_bits_w 2 --n_bits_a 4 --res 224 --num_samples 1024 --iters_w 20000

Several GPUs of one box (the flow of Brecq/main_imagenet_dist.py:156-221): one rank per GPU under torchrun; every rank builds
the same seeded model, keeps its shard of the calibration set (num_samples / N images), and the reconstruction calls run with
multi_gpu=True — the unit's gradients are summed over the ranks every iteration, so all replicas learn the same parameters:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/run_ptq.py --arch regnetx_3200m
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shiftedscalequantization_b200 import quant as Q, zoo          # noqa: E402
from shiftedscalequantization_b200.quant.quant_block import QuantBasicBlock  # noqa: E402
from shiftedscalequantization_b200.quant import (BaseQuantBlock, QuantModel, QuantModule,  # noqa: E402
                                                   block_reconstruction, layer_reconstruction)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--arch', default='resnet18', choices=sorted(zoo.ARCHS))
    ap.add_argument('--seed', default=1005, type=int)
    ap.add_argument('--res', default=224, type=int)
    ap.add_argument('--num_classes', default=1000, type=int)
    ap.add_argument('--n_bits_w', default=2, type=int)
    ap.add_argument('--n_bits_a', default=4, type=int)
    ap.add_argument('--channel_wise', default=1, type=int)
    ap.add_argument('--act_quant', default=1, type=int)
    ap.add_argument('--scale_method', default='mse')
    ap.add_argument('--num_samples', default=1024, type=int)
    ap.add_argument('--iters_w', default=20000, type=int)
    ap.add_argument('--iters_a', default=5000, type=int)
    ap.add_argument('--weight', default=0.01, type=float)
    ap.add_argument('--b_start', default=20, type=int)
    ap.add_argument('--b_end', default=2, type=int)
    ap.add_argument('--warmup', default=0.2, type=float)
    ap.add_argument('--lr', default=4e-4, type=float)
    ap.add_argument('--p', default=2.4, type=float)
    ap.add_argument('--batch_size', default=32, type=int)
    ap.add_argument('--max_units', default=0, type=int, help='reconstruct only the first N units (0 = all)')
    ap.add_argument('--tf32', default=0, type=int, help='1: let cuDNN/cuBLAS use TF32 (torch default); 0: true fp32 like the CPU reference')
    ap.add_argument('--cudnn_benchmark', default=1, type=int)
    # the reference README's flags (README.md:20,33-34)
    flag = lambda v: str(v).lower() in ('1', 'true', 'yes')
    ap.add_argument('--device_gpu', default='cuda:0')
    ap.add_argument('--host_resident', default=False, type=flag,
                    help="weight phase: keep the cached features in pinned host memory (upstream's keep_gpu=False) and pull every mini-batch over PCIe")
    ap.add_argument('--bias_cal', default=False, type=flag, help='learn the output-channel scale gamma^z and offset varphi^z')
    ap.add_argument('--bias_ch_quant', default=False, type=flag,
                    help='learn the input-channel group R: shifted-scale ChannelQuant + fused shift/rounding loop on BasicBlock units')
    args = ap.parse_args(argv)

    from shiftedscalequantization_b200 import dist as ssq_dist
    rank, local, world = ssq_dist.init_from_env()
    multi = world > 1
    dev = torch.device('cuda', local) if multi else torch.device(args.device_gpu)
    torch.cuda.set_device(dev)
    if rank != 0:                                              # one voice: the other ranks compute silently
        import builtins
        builtins.print = lambda *a, **k: None
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    torch.manual_seed(args.seed)
    kw = {} if args.arch.startswith('regnet') else ({'n_class': args.num_classes} if args.arch == 'mobilenetv2' else {'num_classes': args.num_classes})
    cnn = zoo.build(args.arch, **kw).to(dev).eval()
    wq = {'n_bits': args.n_bits_w, 'channel_wise': bool(args.channel_wise), 'scale_method': args.scale_method}
    aq = {'n_bits': args.n_bits_a, 'channel_wise': False, 'scale_method': args.scale_method, 'leaf_param': bool(args.act_quant)}
    qnn = QuantModel(model=cnn, weight_quant_params=wq, act_quant_params=aq).to(dev).eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(args.num_samples, 3, args.res, args.res)
    if multi:
        cali = ssq_dist.shard_calibration(cali)                # main_imagenet_dist.py:165: num_samples / ngpus images per rank
        from shiftedscalequantization_b200.quant import layer_recon_shiftedScale as LS
        LS.MULTI_GPU = True                                    # the shifted-scale loops' own switch

    t0 = time.time()
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali[:64].to(dev))
    torch.cuda.synchronize()
    print(f'weight scale init: {time.time() - t0:.2f}s')

    done = [0]

    def shifted_block(block, iters):
        """--bias_ch_quant: ChannelQuant swap, 'if'/'of' feature caches, fused shift + rounding loop
        (the flow of ShiftedScaleQuant.py:244-255,384-392 / myScaledMethods.py)"""
        from shiftedscalequantization_b200.quant.channelQuant import ChannelQuant
        from shiftedscalequantization_b200.quant.layer_recon_fused_shiftedScale import block_recon_fused_shiftedScale
        for m in block.modules():
            if isinstance(m, QuantModule):
                m.weight_quantizer = ChannelQuant(1.0, uaq=m.weight_quantizer, weight_tensor=m.org_weight.data,
                                                  shiftTarget=[0.96875, 1.03125, 1.0], name=m.pathName)
        block.clear_cached_features()
        for mode, wq_on in (('if', True), ('of', False)):
            qnn.set_quant_state(wq_on, False)
            block.cache_features = mode
            with torch.no_grad():
                for i in range(0, cali.shape[0], args.batch_size):
                    qnn(cali[i:i + args.batch_size].to(dev))
            block.cache_features = 'none'
        qnn.set_quant_state(False, False)
        block.set_quant_state(True, False)
        block_recon_fused_shiftedScale(block, iters=iters, lmda=[args.weight, args.weight], model=qnn, bias_cal=args.bias_cal)
        block.clear_cached_features()

    def recon_model(model: nn.Module, **kwargs):
        for name, module in model.named_children():
            if args.max_units and done[0] >= args.max_units:
                return
            if isinstance(module, QuantModule):
                if not module.ignore_reconstruction:
                    print(f'Reconstruction for layer {name}')
                    layer_reconstruction(qnn, module, **kwargs)
                    done[0] += 1
            elif isinstance(module, BaseQuantBlock):
                if not module.ignore_reconstruction:
                    print(f'Reconstruction for block {name}')
                    if args.bias_ch_quant and not kwargs.get('act_quant') and isinstance(module, QuantBasicBlock):
                        shifted_block(module, kwargs['iters'])
                    else:
                        block_reconstruction(qnn, module, **kwargs)
                    done[0] += 1
            else:
                recon_model(module, **kwargs)

    t0 = time.time()
    recon_model(qnn, cali_data=cali, iters=args.iters_w, weight=args.weight, asym=True, b_range=(args.b_start, args.b_end),
                warmup=args.warmup, act_quant=False, opt_mode='mse', batch_size=args.batch_size, bias_cal=args.bias_cal,
                host_resident=args.host_resident, multi_gpu=multi)
    torch.cuda.synchronize()
    print(f'weight reconstruction: {time.time() - t0:.2f}s for {done[0]} units x {args.iters_w} iterations')
    qnn.set_quant_state(weight_quant=True, act_quant=False)
    with torch.no_grad():
        out_w = qnn(cali[:32].to(dev))
    assert torch.isfinite(out_w).all()

    if args.act_quant:
        qnn.set_quant_state(True, True)
        with torch.no_grad():
            qnn(cali[:64].to(dev))
        qnn.disable_network_output_quantization()
        if multi:
            qnn.synchorize_activation_statistics()             # main_imagenet_dist.py:211: the ranks saw different images
        done[0] = 0
        t0 = time.time()
        recon_model(qnn, cali_data=cali, iters=args.iters_a, act_quant=True, opt_mode='mse', lr=args.lr, p=args.p,
                    batch_size=args.batch_size, multi_gpu=multi)
        torch.cuda.synchronize()
        print(f'activation reconstruction: {time.time() - t0:.2f}s for {done[0]} units x {args.iters_a} iterations')
        qnn.set_quant_state(weight_quant=True, act_quant=True)
        with torch.no_grad():
            out_wa = qnn(cali[:32].to(dev))
        assert torch.isfinite(out_wa).all()
    sd = qnn.state_dict()
    if multi:
        # replicas must have learned the same parameters bit for bit (summed gradients, identical Adam steps)
        import torch.distributed as td
        digest = torch.stack([v.detach().double().sum() for k, v in sorted(sd.items()) if v.is_floating_point() and v.is_cuda])
        both = [torch.empty_like(digest) for _ in range(world)]
        td.all_gather(both, digest)
        assert all(torch.equal(both[0], b) for b in both), 'replicas diverged'
        print(f'{world} replicas identical ({digest.numel()} tensors compared)')
    print(f'done: state_dict with {len(sd)} tensors; W{args.n_bits_w}A{args.n_bits_a if args.act_quant else 32}')
    if multi:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()
    return qnn


if __name__ == '__main__':
    main()
