#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_backward.py -q -k "network_backward or training_step or small_gradients" 2>&1 | grep -E "^E  |passed|failed" | head -20
timeout 200 python tools/time_backward.py cfg2 64 fp16x3 2>&1 | tail -1
