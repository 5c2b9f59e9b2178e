#!/bin/bash
# Last round-2 measurement pass (one gpurun call) after the training-path GEMMs moved onto pnr_wgrad / pnr_linear:
# smoke, full GPU suite, bench line, train-step time + profile, timings and ncu --set full of the two new kernels.
#   gpurun --timeout 1500 -- 'bash tools/r2_final2.sh 2>&1 | tee gpurun_out/r2_final2.log'
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2_final2.log 2>&1; tail -3 gpurun_out/gpu_tests_r2_final2.log
timeout 200 python tools/time_train_step.py cfg3 2048 192 2>&1 | tail -1
timeout 200 python tools/time_train_step.py cfg2 4096 64 2>&1 | tail -1
timeout 200 python tools/profile_train_step.py cfg3 > gpurun_out/r02_train_step_profile_native.log 2>&1; head -30 gpurun_out/r02_train_step_profile_native.log | cut -c1-200
timeout 100 python tools/time_wgrad.py 2>&1 | tail -1
timeout 100 python tools/time_wgrad.py 393216 256 63 2>&1 | tail -1
timeout 100 python tools/time_linear.py 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err; tail -c 600 gpurun_out/bench_r2_final2.json; tail -2 gpurun_out/bench_r2_final2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 2 -c 1 -o gpurun_out/r02_wgrad \
    python tools/time_wgrad.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 2 -c 1 -o gpurun_out/r02_linear \
    python tools/time_linear.py 393216 256 256 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
