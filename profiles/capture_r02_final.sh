# round-2 final capture on one B200 (every profiler run bounded): the whole GPU suite, the bench line of both arms, the launch list of
# the step, ncu --set full of launch 3 and of the K2b sweep, the K2 metric list. Summaries: python profiles/make_summary.py r02
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -1 gpurun_out/r02_bench_reference.err
python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; tail -2 gpurun_out/r02_bench_full.err
SMALL="--steps 3 --warmup 3 --images 64 --skip-e2e --skip-act --skip-cpu --skip-micro --skip-tf32 --skip-shift --skip-extras --unit-iters 3 --cudnn-benchmark 0"
python bench.py $SMALL > gpurun_out/r02_small_plain.json 2> gpurun_out/r02_small_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches.csv python bench.py $SMALL > gpurun_out/r02_small_ncu.json 2> gpurun_out/r02_small_ncu.err
wc -l gpurun_out/r02_launches.csv
python bench.py --micro-only > gpurun_out/r02_micro_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ada_bwd_adam_mt_kernel -s 3 -c 1 -f -o gpurun_out/r02_full_ada_bwd_adam_mt_kernel python bench.py --micro-only > gpurun_out/r02_ncu_ada_bwd_adam_mt_kernel.log 2>&1
tail -1 gpurun_out/r02_ncu_ada_bwd_adam_mt_kernel.log
python bench.py --k2-only > gpurun_out/r02_k2_plain.json 2> gpurun_out/r02_k2_plain.err && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:inp_scale_sweep -s 22 -c 1 -f -o gpurun_out/r02_full_inp_scale_sweep_kernel python bench.py --k2-only > gpurun_out/r02_ncu_sweep.log 2>&1
tail -1 gpurun_out/r02_ncu_sweep.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum --clock-control none -k regex:"mse_search|mse_rank|mse_settle|inp_scale|row_minmax" -c 600 --csv --log-file gpurun_out/r02_k2_ncu.csv python bench.py --k2-only > gpurun_out/r02_ncu_k2.log 2>&1
tail -1 gpurun_out/r02_ncu_k2.log | cut -c1-120
ls -la gpurun_out | tail -8
