#!/bin/bash
# stage kernels after a change: bit-exact parity tests + per-stage times of a cfg3 frame
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_golden.py tests/test_gpu_fused.py -q 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/stage_launches.csv \
    python tools/time_render.py cfg3:96 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/stage_launches.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[hdr+2:hdr+40]:
    if 'mlp_fused' not in r[ki]: print(f"{r[vi]:>12s} {r[ui]}  {r[ki][:70]}")
PY
