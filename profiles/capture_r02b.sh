# round-2 capture, part b (every profiler run bounded by `timeout` and `-c`): K2 tests + K2 roofline numbers + K2 ncu metric list,
# launch list of the step, ncu --set full of launch 3 (fused backward + Adam) and of the K2b sweep, N=1 baselines of configs[2]/[4]
set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "scale or k2 or mse or inp" 2>&1 | tail -3 | tee gpurun_out/r02b_pytest_k2.txt
python bench.py --k2-only > gpurun_out/r02_k2_plain.json 2> gpurun_out/r02_k2_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum --clock-control none -k regex:"mse_search|mse_rank|mse_settle|inp_scale_sweep|inp_scale_fit|row_minmax" -c 400 --csv --log-file gpurun_out/r02_k2_ncu.csv python bench.py --k2-only > gpurun_out/r02_ncu_k2.log 2>&1
tail -2 gpurun_out/r02_ncu_k2.log
SMALL="--steps 3 --warmup 3 --images 64 --skip-e2e --skip-act --skip-cpu --skip-micro --skip-tf32 --skip-shift --skip-extras --unit-iters 3 --cudnn-benchmark 0"
python bench.py $SMALL > gpurun_out/r02_small_plain.json 2> gpurun_out/r02_small_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches.csv python bench.py $SMALL > gpurun_out/r02_small_ncu.json 2> gpurun_out/r02_small_ncu.err
wc -l gpurun_out/r02_launches.csv
python bench.py --micro-only > gpurun_out/r02_micro_plain.log 2>&1 && \
for spec in "ada_bwd_adam_mt_kernel 3 ada_bwd_adam_mt_kernel"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/r02_full_$3 python bench.py --micro-only > gpurun_out/r02_ncu_$3.log 2>&1
  tail -1 gpurun_out/r02_ncu_$3.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:inp_scale_sweep -s 22 -c 1 -f -o gpurun_out/r02_full_inp_scale_sweep_kernel python bench.py --k2-only > gpurun_out/r02_ncu_sweep.log 2>&1
tail -1 gpurun_out/r02_ncu_sweep.log
timeout 600 python examples/scale_configs.py --config resnet50_shift > gpurun_out/r02_resnet50_shift_n1.json 2> gpurun_out/r02_resnet50_shift_n1.err; tail -2 gpurun_out/r02_resnet50_shift_n1.err
timeout 900 python examples/scale_configs.py --config regnet --steps 20 > gpurun_out/r02_regnet_n1.json 2> gpurun_out/r02_regnet_n1.err; tail -2 gpurun_out/r02_regnet_n1.err
python scratch/host_link_probe.py > gpurun_out/r02_host_link_n1.json 2> gpurun_out/r02_host_link_n1.err; tail -2 gpurun_out/r02_host_link_n1.err
ls -la gpurun_out/ | tail -30
