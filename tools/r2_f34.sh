#!/bin/bash
# GPU run of the 8(f) rank 2 (loss side) / rank 3 / rank 4 pieces: parity tests, timing, one ncu capture of the new kernels
timeout 600 python -m pytest tests/test_gpu_panoptic.py tests/test_gpu_losses.py tests/test_gpu_scale.py -q -k "panoptic or hashgrid or generate_rays or loss" 2>&1 | tail -4
timeout 300 python tools/time_hashgrid.py 2>&1 | tail -6
