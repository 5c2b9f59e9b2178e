#!/bin/bash
# pnr_wgrad + pnr_linear: tests, timings against the library GEMMs they replace, the training step on them, ncu
# --set full of both kernels (one gpurun call):
#   gpurun --timeout 900 -- 'bash tools/r2_native.sh 2>&1 | tee gpurun_out/r2_native.log'
timeout 300 python -m pytest tests/test_gpu_backward.py -q -s -k "wgrad or linear3x" 2>&1 | grep -E "^\.?wgrad|linear3x|passed|failed|Error|error|assert" | head -60
timeout 600 python -m pytest tests/test_gpu_backward.py -q -x -k "network_backward or training_step or update_weights" 2>&1 | tail -3
timeout 100 python tools/time_wgrad.py 2>&1 | tail -1
timeout 100 python tools/time_wgrad.py 393216 256 63 2>&1 | tail -1
timeout 100 python tools/time_wgrad.py 4000000 256 256 2>&1 | tail -1
timeout 100 python tools/time_linear.py 2>&1 | tail -3
timeout 200 python tools/time_train_step.py cfg3 2048 192 2>&1 | tail -1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 2 -c 1 -o gpurun_out/r02_wgrad \
    python tools/time_wgrad.py > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 2 -c 1 -o gpurun_out/r02_linear \
    python tools/time_linear.py 393216 256 256 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
