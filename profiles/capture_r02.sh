# round-2 capture: GPU tests, full bench, launch list of the step (small config, after the plain run exited 0),
# ncu --set full of the kernels that are new this round (fused backward+Adam, iteration prologue, K2a rows / tensor, K2b sweep)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; tail -2 gpurun_out/r02_bench_full.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -2 gpurun_out/r02_bench_reference.err
SMALL="--steps 3 --warmup 3 --images 64 --skip-e2e --skip-act --skip-cpu --skip-micro --skip-tf32 --skip-shift --skip-extras --unit-iters 3 --cudnn-benchmark 0"
python bench.py $SMALL > gpurun_out/r02_small_plain.json 2> gpurun_out/r02_small_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches.csv python bench.py $SMALL > gpurun_out/r02_small_ncu.json 2> gpurun_out/r02_small_ncu.err
python bench.py --micro-only > gpurun_out/r02_micro_plain.log 2>&1 && \
for spec in "ada_bwd_adam_mt_kernel 3 ada_bwd_adam_mt_kernel" "ada_fwd_mt_kernel 3 ada_fwd_mt_kernel" "recon_loss_kernel 3 recon_loss_kernel"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/r02_full_$3 python bench.py --micro-only > gpurun_out/r02_ncu_$3.log 2>&1
  tail -1 gpurun_out/r02_ncu_$3.log
done
python bench.py --k2-only > gpurun_out/r02_k2_plain.json 2> gpurun_out/r02_k2_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum --clock-control none -k regex:"mse_search|mse_rank|mse_settle|inp_scale_sweep|inp_scale_fit|row_minmax" --csv --log-file gpurun_out/r02_k2_ncu.csv python bench.py --k2-only > gpurun_out/r02_ncu_k2.log 2>&1
tail -1 gpurun_out/r02_ncu_k2.log
ls -la gpurun_out/*.ncu-rep | tail -14
