#!/bin/bash
# One-call evaluation of the epilogue-part options of the fused MLP (tuning aid): parity tests, then cfg2 kernel
# times and one-tile clock64 timelines for each combination of PNR_SPLIT_WAR / PNR_SPLIT_E1.
#   gpurun --timeout 1200 -- 'bash tools/r2_try.sh 2>&1 | tee gpurun_out/r2_try.log'
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2.log 2>&1; tail -12 gpurun_out/gpu_tests_r2.log
for w in 0 1; do for e in 0 1; do
  echo "=== PNR_SPLIT_WAR=$w PNR_SPLIT_E1=$e"
  PNR_SPLIT_WAR=$w PNR_SPLIT_E1=$e timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 2>&1 | grep mlp
  PNR_SPLIT_WAR=$w PNR_SPLIT_E1=$e PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py fp16x3 \
      > gpurun_out/timeline_r2_w${w}e${e}.log 2>&1
  head -1 gpurun_out/timeline_r2_w${w}e${e}.log
done; done
PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py fp16 > gpurun_out/timeline_r2_fp16.log 2>&1
head -1 gpurun_out/timeline_r2_fp16.log
for v in novmax; do
  [ -f panopticnerf_b200/libpnr_$v.so ] || continue
  echo "=== ablation: $v"
  PNR_LIB=$PWD/panopticnerf_b200/libpnr_$v.so timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 2>&1 | grep mlp
done
echo "=== view accumulator in the lower half (PNR_VIEW_UPPER=0)"
PNR_VIEW_UPPER=0 timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 2>&1 | grep mlp
PNR_VIEW_UPPER=0 PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py fp16x3 > gpurun_out/timeline_r2_viewlow.log 2>&1
head -1 gpurun_out/timeline_r2_viewlow.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; tail -c 3000 gpurun_out/bench_r2_a.json; tail -5 gpurun_out/bench_r2_a.err
