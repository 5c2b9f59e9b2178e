#!/bin/bash
# Final round-2 measurement pass (one gpurun call): smoke, full GPU suite, kernel timings, bench line, launch list,
# ncu --set full of the fused MLP (cfg2 frame) and of its compositing-epilogue variant (cfg3 strip), sanitizers on the
# new kernels.   gpurun --timeout 2400 -- 'bash tools/r2_final.sh 2>&1 | tee gpurun_out/r2_final.log'
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2_final.log 2>&1; tail -3 gpurun_out/gpu_tests_r2_final.log
timeout 150 python tools/time_mlp.py cfg2 fp16x3 fp16 bf16x3 2>&1 | grep mlp
timeout 300 python tools/time_render.py cfg2 cfg3 2>&1 | grep cfg
PNR_LIB=$PWD/panopticnerf_b200/libpnr_timeline.so timeout 100 python tools/timeline.py fp16x3 > gpurun_out/timeline_r2_final2.log 2>&1; head -1 gpurun_out/timeline_r2_final2.log
timeout 200 python tools/time_backward.py cfg2 64 fp16x3 2>&1 | tail -1
timeout 200 python tools/time_train_step.py cfg3 2048 192 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 1500 gpurun_out/bench_r2_final.json; tail -3 gpurun_out/bench_r2_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_ref.json 2>/dev/null; tail -c 400 gpurun_out/bench_r2_final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-fast-mode > gpurun_out/bench_under_ncu.json 2> /dev/null
grep -c mlp_fused gpurun_out/r02_launches_final.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused -s 2 -c 1 -o gpurun_out/r02_mlp_fp16x3_final \
    python tools/time_mlp.py cfg2 fp16x3 > /dev/null 2>&1
ls -la gpurun_out/r02_mlp_fp16x3_final.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:mlp_fused -s 4 -c 2 -o gpurun_out/r02_mlp_comp_cfg3_final \
    python tools/time_render.py cfg3:96 > /dev/null 2>&1
ls -la gpurun_out/r02_mlp_comp_cfg3_final.ncu-rep
timeout 400 ncu --set full --clock-control none -k regex:mlp_fused -s 2 -c 1 -o gpurun_out/r02_mlp_bwd_final \
    python tools/time_backward.py cfg2 16 fp16x3 > /dev/null 2>&1
ls -la gpurun_out/r02_mlp_bwd_final.ncu-rep
for tool in memcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 12 python -m pytest tests/test_gpu_backward.py tests/test_gpu_mlp.py -q -x \
      -k "(mlp_backward_trunk_matches and cfg1) or unpadded or tile_tails" > gpurun_out/sanitizer2_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer2_$tool.log | tr '\n' ' ')"
done
