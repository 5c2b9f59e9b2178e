#!/bin/bash
# GPU run of the MLP backward: parity tests (trunk kernel, full parameter gradients, training step), timing, timeline
timeout 900 python -m pytest tests/test_gpu_backward.py -q -k "mlp_backward or network_backward or training_step" 2>&1 | tail -15
timeout 200 python tools/time_backward.py cfg2 64 fp16x3 2>&1 | tail -3
PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py --backward fp16x3 > gpurun_out/timeline_r2_bwd.log 2>&1
head -1 gpurun_out/timeline_r2_bwd.log
