#!/bin/bash
# pnr_wgrad + pnr_linear bring-up and the training step on them (one gpurun call):
#   gpurun --timeout 900 -- 'bash tools/r2_native.sh 2>&1 | tee gpurun_out/r2_native.log'
PNR_TEST_NEW_KERNELS=1 timeout 300 python -m pytest tests/test_gpu_backward.py -q -s -k "wgrad or linear3x" 2>&1 | grep -E "wgrad S|linear3x|passed|failed|Error|error|assert" | head -60
echo "== training path on the native GEMMs"
PNR_TRAIN_NATIVE=1 timeout 600 python -m pytest tests/test_gpu_backward.py -q -x -k "network_backward or training_step or update_weights" 2>&1 | tail -8
timeout 100 python tools/time_wgrad.py 2>&1 | tail -1
timeout 100 python tools/time_wgrad.py 393216 256 63 2>&1 | tail -1
echo "== train step: library GEMMs, then native"
timeout 200 python tools/time_train_step.py cfg3 2048 192 2>&1 | tail -1
PNR_TRAIN_NATIVE=1 timeout 200 python tools/time_train_step.py cfg3 2048 192 2>&1 | tail -1
