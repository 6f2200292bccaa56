# round-2 capture, part f: whole GPU suite after the K2b / exchange / Fisher-test changes, K2 ncu metric list with the new K2b, e2e vs pull-CTA count, MobileNetV2 re-run
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.txt
python bench.py --k2-only > gpurun_out/r02_k2_plain.json 2> gpurun_out/r02_k2_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum --clock-control none -k regex:"mse_search|mse_rank|mse_settle|inp_scale|row_minmax" -c 500 --csv --log-file gpurun_out/r02_k2_ncu.csv python bench.py --k2-only > gpurun_out/r02_ncu_k2.log 2>&1
tail -1 gpurun_out/r02_ncu_k2.log | cut -c1-200
timeout 400 python scratch/e2e_ctas_probe.py 4 8 16 32 > gpurun_out/r02_e2e_ctas.json 2> gpurun_out/r02_e2e_ctas.err; cat gpurun_out/r02_e2e_ctas.json
timeout 300 python examples/scale_configs.py --config mobilenetv2_mse --steps 20 > gpurun_out/r02_mobilenetv2_n1.json 2> gpurun_out/r02_mobilenetv2_n1.err; cat gpurun_out/r02_mobilenetv2_n1.json
