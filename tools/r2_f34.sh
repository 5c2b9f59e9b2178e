#!/bin/bash
# GPU run of the 8(f) rank 2 (loss side) / rank 3 / rank 4 pieces: parity tests, timing, one ncu capture of the new kernels
timeout 600 python -m pytest tests/test_gpu_panoptic.py tests/test_gpu_losses.py tests/test_gpu_scale.py -q -k "panoptic or hashgrid or generate_rays or loss" 2>&1 | tail -12
timeout 300 python tools/time_hashgrid.py 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none -k regex:'hashgrid_kernel|panoptic_fuse_kernel' -c 4 -o gpurun_out/r02_hashgrid_v2 \
    python tools/time_hashgrid.py 8 > /dev/null 2>&1
ls -la gpurun_out/r02_hashgrid_v2.ncu-rep
