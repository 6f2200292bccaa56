# round-2 capture, part h: K2b as four launches (epoch flag) — tests, K2 numbers, MobileNetV2, metric list of the K2b kernels only
set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_shift_modules_gpu.py -m gpu -q -x -k "scale or k2 or mse or inp or golden" 2>&1 | tail -3 | tee gpurun_out/r02h_pytest.txt
python bench.py --k2-only > gpurun_out/r02h_k2_plain.json 2> gpurun_out/r02h_k2_plain.err; tail -1 gpurun_out/r02h_k2_plain.err
timeout 200 python examples/scale_configs.py --config mobilenetv2_mse --steps 20 > gpurun_out/r02_mobilenetv2_n1.json 2> gpurun_out/r02_mobilenetv2_n1.err; cat gpurun_out/r02_mobilenetv2_n1.json
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum --clock-control none -k regex:"inp_scale" -c 130 --csv --log-file gpurun_out/r02h_k2b_ncu.csv python bench.py --k2-only > gpurun_out/r02h_ncu_k2b.log 2>&1
tail -1 gpurun_out/r02h_ncu_k2b.log | cut -c1-100
