set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/r01_bench_full.json 2> gpurun_out/r01_bench_full.err; tail -2 gpurun_out/r01_bench_full.err
SMALL="--steps 3 --warmup 3 --images 64 --skip-e2e --skip-act --skip-cpu --skip-micro --skip-tf32 --skip-shift --cudnn-benchmark 0"
python bench.py $SMALL > gpurun_out/r01_small_plain.json 2> gpurun_out/r01_small_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv python bench.py $SMALL > gpurun_out/r01_small_ncu.json 2> gpurun_out/r01_small_ncu.err
python bench.py --micro-only > gpurun_out/r01_micro_plain.log 2>&1 && \
for k in adam_kernel ada_fwd_kernel ada_bwd_kernel recon_loss_kernel fq_affine_fwd_vec fq_affine_bwd_kernel gather_rows_kernel export_vec_kernel import_vec_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r01_full_$k python bench.py --micro-only > gpurun_out/r01_ncu_$k.log 2>&1
  tail -1 gpurun_out/r01_ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep | tail -12
