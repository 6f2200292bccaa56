set -x
python -m pytest tests/test_kernels_gpu.py tests/test_shift_modules_gpu.py -m gpu -q -x -k "scale or k2 or mse or inp" 2>&1 | tail -3 | tee gpurun_out/r02e_pytest_k2.txt
python bench.py --k2-only > gpurun_out/r02e_k2_plain.json 2> gpurun_out/r02e_k2_plain.err; tail -1 gpurun_out/r02e_k2_plain.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:inp_scale_sweep -s 22 -c 1 -f -o gpurun_out/r02e_full_inp_scale_sweep_kernel python bench.py --k2-only > gpurun_out/r02e_ncu_sweep.log 2>&1
tail -1 gpurun_out/r02e_ncu_sweep.log
