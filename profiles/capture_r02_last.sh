# last round-2 run with the final code: the whole GPU suite and the bench line of both arms (no profiler)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -1 gpurun_out/r02_bench_reference.err
python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; tail -2 gpurun_out/r02_bench_full.err
timeout 200 python examples/scale_configs.py --config mobilenetv2_mse --steps 20 > gpurun_out/r02_mobilenetv2_n1.json 2> gpurun_out/r02_mobilenetv2_n1.err; cat gpurun_out/r02_mobilenetv2_n1.json
