#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_backward.py -q -k "network_backward or training_step or small_gradients" 2>&1 | grep -E "^E  |passed|failed" | head -20
timeout 200 python tools/time_backward.py cfg2 64 fp16x3 2>&1 | tail -1
timeout 200 python tools/time_train_step.py cfg3 2048 192 2>&1 | tail -2
timeout 200 python tools/time_train_step.py cfg2 4096 64 2>&1 | tail -2
echo "== view step in the upper accumulator half (both-heads programs): off / on"
PNR_VIEW_UPPER=0 timeout 300 python tools/time_render.py cfg3 2>&1 | grep cfg3
PNR_VIEW_UPPER=1 timeout 300 python tools/time_render.py cfg3 2>&1 | grep cfg3
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_mlp.py -q 2>&1 | tail -2
