#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x -k "update_weights or training_step or network_backward" 2>&1 | grep -E "^E  |passed|failed|rror" | head -20
timeout 200 python tools/time_train_step.py cfg3 2048 192 2>&1 | tail -2
timeout 200 python tools/time_train_step.py cfg2 4096 64 2>&1 | tail -2
timeout 200 python tools/profile_train_step.py cfg3 2>&1 | tail -32 | cut -c1-200 > gpurun_out/profile_train_step.log; tail -30 gpurun_out/profile_train_step.log | cut -c1-62,140-200
