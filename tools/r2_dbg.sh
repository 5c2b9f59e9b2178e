#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_backward.py -q -k "network_backward or training_step" 2>&1 | grep -E "^E  |passed|failed" | head -20
echo "== view step as one N=128 half: off / on"
PNR_VIEW_ONE_HALF=0 timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 2>&1 | grep mlp
PNR_VIEW_ONE_HALF=1 timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 2>&1 | grep mlp
PNR_VIEW_ONE_HALF=0 PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py fp16x3 > gpurun_out/timeline_r2_view2.log 2>&1; head -1 gpurun_out/timeline_r2_view2.log
PNR_VIEW_ONE_HALF=1 PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py fp16x3 > gpurun_out/timeline_r2_view1.log 2>&1; head -1 gpurun_out/timeline_r2_view1.log
PNR_VIEW_ONE_HALF=0 timeout 300 python tools/time_render.py cfg3 2>&1 | grep cfg3
PNR_VIEW_ONE_HALF=1 timeout 300 python tools/time_render.py cfg3 2>&1 | grep cfg3
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_mlp.py tests/test_gpu_golden.py -q 2>&1 | tail -2
