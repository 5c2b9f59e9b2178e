#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_golden.py tests/test_gpu_fused.py -q -x 2>&1 | grep -E "^E  |passed|failed" | head -20
echo "== view epilogue on the producer warps: off / on"
PNR_VIEW_PRODUCERS=0 timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 2>&1 | grep mlp
PNR_VIEW_PRODUCERS=1 timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 bf16x3 2>&1 | grep mlp
PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py fp16x3 > gpurun_out/timeline_r2_vp.log 2>&1; head -1 gpurun_out/timeline_r2_vp.log
