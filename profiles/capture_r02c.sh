# round-2 capture, part c: N GPUs of one box (N = first argument, default 8). Host-link probe, the bench line at N (value + e2e),
# BASELINE configs[4] (RegNetX-3200M W2A4, all units) and configs[2] (ResNet-50 W4A4 shifted-scale layer reconstruction) at N.
# Every command bounded; N-GPU box time is charged N-fold.
N=${1:-8}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 150 $TR --master-port 29701 scratch/host_link_probe.py > gpurun_out/r02_host_link_n$N.json 2> gpurun_out/r02_host_link_n$N.err; tail -2 gpurun_out/r02_host_link_n$N.err
timeout 240 $TR --master-port 29702 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -3 gpurun_out/r02_bench_n$N.err
timeout 240 $TR --master-port 29703 examples/scale_configs.py --config regnet --steps 20 > gpurun_out/r02_regnet_n$N.json 2> gpurun_out/r02_regnet_n$N.err; tail -3 gpurun_out/r02_regnet_n$N.err
timeout 200 $TR --master-port 29704 examples/scale_configs.py --config resnet50_shift > gpurun_out/r02_resnet50_shift_n$N.json 2> gpurun_out/r02_resnet50_shift_n$N.err; tail -3 gpurun_out/r02_resnet50_shift_n$N.err
if [ "$N" = "8" ]; then
  timeout 200 $TR --master-port 29705 bench.py --gpus $N --steps 10 --warmup 3 --host-stage dma > gpurun_out/r02_bench_n${N}_dma.json 2> gpurun_out/r02_bench_n${N}_dma.err; tail -2 gpurun_out/r02_bench_n${N}_dma.err
fi
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r02_pytest_multi_gpu.txt
fi
ls -la gpurun_out | tail -12
