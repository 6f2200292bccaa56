#!/usr/bin/env python
"""BASELINE.json configs[2] and configs[4] on N GPUs of one box (one rank per GPU, launched by torchrun):

    configs[4]  RegNetX-3200M W2A4 block reconstruction: every unit (blocks + fc) of the network, the calibration images
                sharded by rank, the flat AdaRound gradient of the unit summed across ranks every iteration
                (Brecq/main_imagenet_dist.py:156-221 is the reference's flow; quant/block_recon.py:100-102 its exchange)
    configs[2]  ResNet-50 W4A4 shifted-scale LAYER reconstruction (quant/layer_recon_shiftedScale.py:262-338 with the
                module switch MULTI_GPU, :141), the calibration batch sharded by rank

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 examples/scale_configs.py --config regnet
    python examples/scale_configs.py --config resnet50_shift          # N = 1

Prints ONE JSON line per run (rank 0): iterations/s over all ranks (weak scaling: every rank completes its own batch-32
iteration per step, as bench.py counts), ms per step as the max over ranks of the device time, the exchange path in use.
Synthetic randn images, random-init weights.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B                                                               # noqa: E402  (fd 1 -> stderr; B.emit prints the line)
from shiftedscalequantization_b200 import dist as D, ops, quant as Q, zoo      # noqa: E402

AQ = {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}


def regnet(args, rank, local, world, dev):
    import torch.distributed as td
    torch.manual_seed(1005)
    cnn = zoo.build(args.arch).to(dev).eval()
    qnn = Q.QuantModel(cnn, {'n_bits': 2, 'channel_wise': True, 'scale_method': 'mse'}, dict(AQ)).to(dev).eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(args.images, 3, 224, 224)
    lo, hi = D.shard_range(args.images, rank, world)
    cali = cali[lo:hi]
    qnn.set_quant_state(True, False)
    torch.cuda.synchronize(dev); t0 = time.perf_counter()
    with torch.no_grad():
        qnn(cali[:64].to(dev))
    torch.cuda.synchronize(dev); search_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    engines, _ = B.make_engines(Q, qnn, cali, dev, act_quant=False, multi_gpu=world > 1)
    torch.cuda.synchronize(dev); setup_s = time.perf_counter() - t0
    ms = B.timed_steps(engines, args.steps, args.warmup, dev, world)
    n_units = len(engines)
    alpha = sum(int(e.flat.numel()) for e in engines)
    exch = "none (single GPU)" if world == 1 else ("peer-memory kernel (ssq_grad_exchange_adam)" if all(getattr(e, "sym", None) is not None for e in engines)
                                                   else "NCCL all_reduce + ssq_adam_step")
    mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    B.release(engines)
    if rank == 0:
        B.emit(json.dumps({
            "config": f"configs[4]: {args.arch} W2A4 block reconstruction, {n_units} units, {args.images} randn 224x224 images sharded over {world} rank(s), "
                      "mini-batch 32 per rank, weight-rounding phase, fp32 convolutions",
            "n_gpus": world, "units": n_units, "alpha_elems": alpha, "ms_per_step": ms, "iters_per_s": world * n_units / (ms * 1e-3),
            "step": f"one iteration on each of the {n_units} units", "steps": args.steps, "warmup": args.warmup,
            "exchange": exch, "grad_bytes_per_step": 4 * alpha, "weight_scale_search_s": round(search_s, 3),
            "capture_and_engine_setup_s": round(setup_s, 2), "peak_hbm_gib": round(mem, 2)}))
    if world > 1:
        td.barrier()


def resnet50_shift(args, rank, local, world, dev):
    import torch.distributed as td
    from shiftedscalequantization_b200.quant import layer_recon_shiftedScale as LS
    from shiftedscalequantization_b200.quant.channelQuant import ChannelQuant
    out = {}
    for path in args.layers.split(","):
        torch.manual_seed(1005)
        cnn = zoo.build("resnet50").to(dev).eval()
        qnn = Q.QuantModel(cnn, {'n_bits': 4, 'channel_wise': True, 'scale_method': 'max'}, dict(AQ)).to(dev).eval()
        qnn.set_first_last_layer_to_8bit()
        cali = torch.randn(args.images, 3, 224, 224)
        qnn.set_quant_state(True, False)
        with torch.no_grad():
            qnn(cali[:32].to(dev))
        layer = qnn
        for part in path.split('.'):
            layer = layer[int(part)] if part.isdigit() else getattr(layer, part)
        layer.weight_quantizer = ChannelQuant(1.0, uaq=layer.weight_quantizer, weight_tensor=layer.org_weight.data,
                                              shiftTarget=[0.96875, 1.03125, 1.0], name=layer.pathName)
        lo, hi = D.shard_range(args.images, rank, world)
        data = cali[lo:hi]
        for mode, wq_on in (('if', True), ('of', False)):
            qnn.set_quant_state(wq_on, False)
            layer.cache_features = mode
            with torch.no_grad():
                for i in range(0, data.shape[0], 32):
                    qnn(data[i:i + 32].to(dev))
            layer.cache_features = 'none'
        qnn.set_quant_state(False, False)
        layer.set_quant_state(True, False)
        LS.MULTI_GPU = world > 1
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            LS.layer_recon_shiftedScale(layer, iters=args.iters, lmda=0.01, model=qnn)
        LS.MULTI_GPU = False
        st = dict(LS.LAST_LOOP_STATS)
        t = torch.tensor([st["loop_ms"]], device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        out[path] = {"iters_per_s": world * st["iters"] / (float(t) * 1e-3), "us_per_iter": 1e3 * float(t) / st["iters"],
                     "iters": st["iters"], "captured": bool(st.get("captured")), "ssq_launches_per_iter": st.get("launches_per_iter")}
        del qnn, cnn, layer
        torch.cuda.empty_cache()
    if rank == 0:
        B.emit(json.dumps({
            "config": f"configs[2]: ResNet-50 W4A4 shifted-scale layer reconstruction (layer_recon_shiftedScale), {args.images} randn 224x224 "
                      f"images sharded over {world} rank(s), mini-batch 32 per rank, fp32 convolutions",
            "n_gpus": world, "layers": out,
            "note": "iters_per_s counts the batch-32 iterations all ranks complete (weak scaling); loop time = max over ranks of the device "
                    "time of the iteration loop (ChannelQuant init, feature capture and read-outs excluded)"}))
    if world > 1:
        td.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="regnet", choices=["regnet", "resnet50_shift"])
    ap.add_argument("--arch", default="regnetx_3200m")
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--iters", type=int, default=600)
    ap.add_argument("--layers", default="model.layer1.0.conv2,model.layer3.2.conv2,model.layer4.2.conv2")
    args = ap.parse_args()
    rank, local, world = D.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    (regnet if args.config == "regnet" else resnet50_shift)(args, rank, local, world, dev)
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


if __name__ == "__main__":
    main()
