timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x > gpurun_out/gpu_tests_r2c.log 2>&1; tail -8 gpurun_out/gpu_tests_r2c.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2d.log 2>&1; tail -5 gpurun_out/gpu_tests_r2d.log
echo "== compositing epilogue on"; timeout 300 python tools/time_render.py cfg3 cfg2 2>&1 | grep cfg
echo "== two-kernel path (PNR_NO_COMP=1)"; PNR_NO_COMP=1 timeout 300 python tools/time_render.py cfg3 cfg2 2>&1 | grep cfg
timeout 150 python tools/time_mlp.py cfg2 fp16x3 2>&1 | grep mlp
