timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2b.log 2>&1; tail -4 gpurun_out/gpu_tests_r2b.log
timeout 300 python tools/time_render.py cfg3 cfg2 2>&1 | grep cfg
bash tools/r2_scaling.sh 1
