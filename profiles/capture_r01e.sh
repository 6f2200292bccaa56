# final r01 capture (third session): tests, full bench, launch list of the step, ncu --set full of the kernels that changed
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/r01_bench_full.json 2> gpurun_out/r01_bench_full.err; tail -2 gpurun_out/r01_bench_full.err
SMALL="--steps 3 --warmup 3 --images 64 --skip-e2e --skip-act --skip-cpu --skip-micro --skip-tf32 --skip-shift --cudnn-benchmark 0"
python bench.py $SMALL > gpurun_out/r01_small_plain.json 2> gpurun_out/r01_small_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv python bench.py $SMALL > gpurun_out/r01_small_ncu.json 2> gpurun_out/r01_small_ncu.err
python bench.py --micro-only > gpurun_out/r01_micro_plain.log 2>&1 && \
for spec in "ada_fwd_kernel 3 ada_fwd_kernel" "fq_shift_fwd_vec 3 fq_shift_fwd_vec" "fq_shift_bwd_vec 3 fq_shift_bwd_vec" "fq_shift_fwd_vec 16 fq_shift_fwd_vec_deq" "fq_shift_bwd_vec 16 fq_shift_bwd_vec_deq"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/r01_full_$3 python bench.py --micro-only > gpurun_out/r01_ncu_$3.log 2>&1
  tail -1 gpurun_out/r01_ncu_$3.log
done
ls -la gpurun_out/r01_full_*.ncu-rep | tail -14
