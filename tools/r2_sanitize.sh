#!/bin/bash
# compute-sanitizer memcheck over the tests of the two new GEMM kernels (one gpurun call):
#   gpurun --timeout 400 -- 'bash tools/r2_sanitize.sh 2>&1 | tee gpurun_out/r2_sanitize.log'
timeout 170 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_backward.py -q -x \
    -k "(wgrad and not 14213 and not 20011) or (linear3x and not 37965) or stash_maxima" > gpurun_out/sanitizer3_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer3_memcheck.log | head -12
