#!/usr/bin/env python
"""BASELINE.json configs[2], configs[3] and configs[4] on N GPUs of one box (one rank per GPU, launched by torchrun):

    configs[4]  RegNetX-3200M W2A4 block reconstruction: every unit (blocks + fc) of the network, the calibration images
                sharded by rank, the flat AdaRound gradient of the unit summed across ranks every iteration
                (Brecq/main_imagenet_dist.py:156-221 is the reference's flow; quant/block_recon.py:100-102 its exchange)
    configs[3]  MobileNetV2 W3A3: whole-model K2a / K2b timings, channelShift_wMSE_flow, block reconstruction of every unit
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


def mobilenetv2_mse(args, rank, local, world, dev):
    """configs[3]: MobileNetV2 W3A3 (depthwise-heavy) — whole-model MSE weight-scale search (K2a, rows of 9 / 16..960 / 1280),
    ChannelQuantMSE input-scale search of every layer (K2b, scaled_methods.channelShift_wMSE_flow), then the plain AdaRound block
    reconstruction of every unit (the most memory-bound fake-quant shapes: depthwise [C,1,3,3])."""
    from shiftedscalequantization_b200 import scaled_methods as SM
    from shiftedscalequantization_b200.quant.channelQuantMSE import ChannelQuantMSE
    torch.manual_seed(1005)
    cnn = zoo.mobilenetv2().to(dev).eval()
    qnn = Q.QuantModel(cnn, {'n_bits': 3, 'channel_wise': True, 'scale_method': 'mse'},
                       {'n_bits': 3, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}).to(dev).eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(args.images, 3, 224, 224)
    qnn.set_quant_state(False, False)
    with torch.no_grad():
        qnn(cali[:64].to(dev))                                # CUDA / cuDNN warm-up outside the timed scale search
    mods = [m for m in qnn.modules() if isinstance(m, Q.QuantModule)]
    rows = [(m.org_weight.reshape(m.org_weight.shape[0], -1).contiguous(), m.weight_quantizer.n_levels) for m in mods]

    def ev(fn, reps=5):
        fn(); torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps
    k2a_ms = ev(lambda: ops.mse_scale_search_many([(r, nl, False) for r, nl in rows]))        # as QuantModel.forward issues them
    qnn.set_quant_state(True, False)
    torch.cuda.synchronize(dev); t0 = time.perf_counter()
    with torch.no_grad():
        qnn(cali[:64].to(dev))
    torch.cuda.synchronize(dev); first_q_s = time.perf_counter() - t0
    # K2b on every layer, the kernels only (same arguments ChannelQuantMSE.init_scale builds, level 16, threshold 1.5)
    level, thr = 16, 1.5
    jobs = []
    for m in mods:
        q = m.weight_quantizer
        w = m.org_weight.reshape(m.org_weight.shape[0], -1).contiguous()
        L = q.n_levels
        cand = torch.tensor([i / level for i in range(level, 0, -1)], dtype=torch.float32, device=dev)
        lo = float(torch.tensor(0.0 - 0.5 / (L - 1) * thr, dtype=torch.float32)); hi = float(torch.tensor(1.0 + 0.5 / (L - 1) * thr, dtype=torch.float32))
        jobs.append((w, q.delta.reshape(-1).contiguous(), q.raw_zero_point.reshape(-1).contiguous(), cand, L - 1, lo, hi, torch.ones(w.shape[1], device=dev)))
    k2b_ms = ev(lambda: [ops.inp_scale_search(*j) for j in jobs])
    # the public flow (ChannelQuantMSE objects built and initialised for every layer)
    torch.cuda.synchronize(dev); t0 = time.perf_counter()
    SM.channelShift_wMSE_flow(qnn, cali[:64], level=level, threshold=thr, layerDisabled=('.model.classifier.1',))
    torch.cuda.synchronize(dev); flow_s = time.perf_counter() - t0
    built = sum(isinstance(m.weight_quantizer, ChannelQuantMSE) for m in mods)
    # AdaRound block reconstruction of every unit on the plain quantisers
    torch.manual_seed(1005)
    cnn = zoo.mobilenetv2().to(dev).eval()
    qnn = Q.QuantModel(cnn, {'n_bits': 3, 'channel_wise': True, 'scale_method': 'mse'},
                       {'n_bits': 3, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}).to(dev).eval()
    qnn.set_first_last_layer_to_8bit()
    lo_, hi_ = D.shard_range(args.images, rank, world)
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali[:64].to(dev))
    engines, _ = B.make_engines(Q, qnn, cali[lo_:hi_], dev, act_quant=False, multi_gpu=world > 1)
    ms = B.timed_steps(engines, args.steps, args.warmup, dev, world)
    n_units = len(engines)
    B.release(engines)
    if rank == 0:
        B.emit(json.dumps({
            "config": f"configs[3]: MobileNetV2 W3A3, {args.images} randn 224x224 images, {world} rank(s), fp32 convolutions",
            "n_gpus": world, "quant_modules": len(mods), "weight_rows": int(sum(r.shape[0] for r, _ in rows)), "weight_elems": int(sum(r.numel() for r, _ in rows)),
            "k2a_weight_scale_search_all_layers_ms": round(k2a_ms, 3), "first_quantised_forward_s": round(first_q_s, 3),
            "k2b_inp_scale_search_all_layers_ms": round(k2b_ms, 3), "k2b_level": level,
            "channelShift_wMSE_flow_s": round(flow_s, 3), "channelquantmse_layers_built": built,
            "block_recon": {"units": n_units, "ms_per_step": ms, "iters_per_s": world * n_units / (ms * 1e-3),
                            "step": f"one AdaRound iteration on each of the {n_units} units, mini-batch 32"}}))
    if world > 1:
        import torch.distributed as td
        td.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="regnet", choices=["regnet", "resnet50_shift", "mobilenetv2_mse"])
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
    {"regnet": regnet, "resnet50_shift": resnet50_shift, "mobilenetv2_mse": mobilenetv2_mse}[args.config](args, rank, local, world, dev)
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


if __name__ == "__main__":
    main()
