#!/bin/bash
# Round-2 measurement pass (one gpurun call): GPU tests, bench line, launch list of the bench command, one
# ncu --set full capture of the fused MLP on the cfg2 frame and of its compositing-epilogue variant inside a cfg3
# strip, and of the HBM-bound stage kernels at cfg3 sizes.
#   gpurun --timeout 2400 -- 'bash tools/r2_profile.sh 2>&1 | tee gpurun_out/r2_profile.log'
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2.log 2>&1; tail -4 gpurun_out/gpu_tests_r2.log
timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 bf16x3 2>&1 | grep mlp
echo "== accumulator flip off (PNR_ACC_FLIP=0)"
PNR_ACC_FLIP=0 timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 2>&1 | grep mlp
PNR_ACC_FLIP=0 PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py fp16x3 > gpurun_out/timeline_r2_noflip.log 2>&1
head -1 gpurun_out/timeline_r2_noflip.log
timeout 300 python tools/time_render.py cfg2 cfg3 2>&1 | grep cfg
PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py fp16x3 > gpurun_out/timeline_r2_final.log 2>&1
head -1 gpurun_out/timeline_r2_final.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; tail -c 600 gpurun_out/bench_r2_b.json; tail -3 gpurun_out/bench_r2_b.err
# launch list of the same command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-fast-mode > gpurun_out/bench_under_ncu.json 2> /dev/null
grep -c mlp_fused gpurun_out/r02_launches.csv
# the dominant kernel, once
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused -s 2 -c 1 -o gpurun_out/r02_mlp_fp16x3 \
    python tools/time_mlp.py cfg2 fp16x3 > /dev/null 2>&1
ls -la gpurun_out/r02_mlp_fp16x3.ncu-rep
# cfg3 strip: the MLP with the compositing epilogue (coarse N = 64 and fine N = 192 launches) + the stage kernels
timeout 600 ncu --set full --clock-control none -k regex:mlp_fused -s 4 -c 2 -o gpurun_out/r02_mlp_comp_cfg3 \
    python tools/time_render.py cfg3:96 > /dev/null 2>&1
ls -la gpurun_out/r02_mlp_comp_cfg3.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:'composite_kernel|sample_pdf_kernel|intersect_kernel|stratified_kernel|interval_kernel|tag_kernel|fixed_maps_kernel|label_tiles_kernel' \
    -c 12 -o gpurun_out/r02_stage_kernels python tools/time_render.py cfg3:32 > /dev/null 2>&1
ls -la gpurun_out/r02_stage_kernels.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_cfg3.csv \
    python tools/time_render.py cfg3:96 > /dev/null 2>&1
# sanitizers on the small configurations (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards of the
# compositing epilogue and the hand-off words; synccheck: barrier misuse)
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_fused.py -q -x \
      -k "compositing_epilogue_matches and (cfg1 or cfg3-over3)" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
bash tools/r2_scaling.sh 1
