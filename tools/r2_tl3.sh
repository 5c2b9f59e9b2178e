#!/bin/bash
export PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so
timeout 100 python tools/timeline.py --composite fp16x3 cfg3 > gpurun_out/timeline_r2_comp_cfg3.log 2>&1; head -1 gpurun_out/timeline_r2_comp_cfg3.log
timeout 100 python tools/timeline.py fp16x3 cfg3 > gpurun_out/timeline_r2_fwd_cfg3.log 2>&1; head -1 gpurun_out/timeline_r2_fwd_cfg3.log
