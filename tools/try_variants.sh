#!/bin/bash
# Round-2 starter: build each staged kernel variant ON THE GPU BOX (the in-tree library of the snapshot is
# replaced there only), run the MLP parity tests and time the cfg2 frame.  Usage (from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/try_variants.sh 2>&1 | tee gpurun_out/variants.log'
# Variants: "" = product build; -DPNR_SPLIT_WAR = E0 released by two write-after-read barriers;
# -DPNR_WARP_ARRIVE = one e_done arrival per epilogue warp.  See DESIGN.md section 7.
for f in "" "-DPNR_SPLIT_WAR" "-DPNR_WARP_ARRIVE" "-DPNR_SPLIT_WAR -DPNR_WARP_ARRIVE"; do
  echo "=== variant: [$f]"
  rm -f panopticnerf_b200/libpnr.stamp
  PNR_NVCC_FLAGS="$f" timeout 500 python -m panopticnerf_b200._build > /dev/null 2>&1 || { echo "build failed"; continue; }
  timeout 400 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -2
  timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 2>&1 | grep mlp
done
rm -f panopticnerf_b200/libpnr.stamp
timeout 500 python -m panopticnerf_b200._build > /dev/null 2>&1   # leave the product build behind
