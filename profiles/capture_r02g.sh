# round-2 capture, part g: K2 tests after the interval / idle-grid changes, batched weight-scale init, K2 numbers, MobileNetV2
set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_shift_modules_gpu.py tests/test_recon_gpu.py -m gpu -q -x -k "scale or k2 or mse or inp or fisher or together or golden" 2>&1 | tail -3 | tee gpurun_out/r02g_pytest.txt
python bench.py --k2-only > gpurun_out/r02_k2_plain.json 2> gpurun_out/r02_k2_plain.err; tail -1 gpurun_out/r02_k2_plain.err
timeout 300 python examples/scale_configs.py --config mobilenetv2_mse --steps 20 > gpurun_out/r02_mobilenetv2_n1.json 2> gpurun_out/r02_mobilenetv2_n1.err; cat gpurun_out/r02_mobilenetv2_n1.json
