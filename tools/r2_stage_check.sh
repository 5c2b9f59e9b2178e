#!/bin/bash
# stage kernels after a change: bit-exact parity tests (also under the strict end-to-end floors) + per-stage times of a cfg3 strip
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_golden.py tests/test_gpu_fused.py tests/test_gpu_scale.py -q 2>&1 | tail -2
echo "== PNR_TEST_STRICT=1 (SURVEY 8(a) floors end to end)"
PNR_TEST_STRICT=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/stage_launches.csv \
    python tools/time_render.py cfg3:96 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/stage_launches.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[hdr+2:hdr+14]:
    if 'mlp_fused' not in r[ki]: print(f"{r[vi]:>12s} {r[ui]}  {r[ki][:70]}")
PY
