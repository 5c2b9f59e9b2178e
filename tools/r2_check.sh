#!/bin/bash
# Validation pass after a change to the training path (one gpurun call): the full GPU suite, the training steps,
# the step profile.   gpurun --timeout 900 -- 'bash tools/r2_check.sh 2>&1 | tee gpurun_out/r2_check.log'
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2_check.log 2>&1; tail -3 gpurun_out/gpu_tests_r2_check.log
grep -E "FAILED|Error" gpurun_out/gpu_tests_r2_check.log | head -10
timeout 200 python tools/time_train_step.py cfg3 2048 192 2>&1 | tail -1
timeout 200 python tools/time_train_step.py cfg2 4096 64 2>&1 | tail -1
timeout 200 python tools/profile_train_step.py cfg3 > gpurun_out/r02_train_step_profile_native2.log 2>&1; head -28 gpurun_out/r02_train_step_profile_native2.log | cut -c1-60,150-215
