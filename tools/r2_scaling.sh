#!/bin/bash
# Multi-GPU evidence: weak and strong scaling of the cfg2 frame and config 5, fp32 map tiles and
# label tiles, on N GPUs of one box.   gpurun --gpus N --timeout 1500 -- 'bash tools/r2_scaling.sh N'
N=${1:-2}
P=$((29500 + N))
run() {   # name, bench args...
  local name=$1; shift
  if [ "$N" = "1" ]; then timeout 400 python bench.py --gpus 1 "$@" > gpurun_out/scale_r2_${name}_n${N}.json 2> gpurun_out/scale_r2_${name}_n${N}.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
         bench.py --gpus $N "$@" > gpurun_out/scale_r2_${name}_n${N}.json 2> gpurun_out/scale_r2_${name}_n${N}.err; fi
  P=$((P + 7))
  python - "$name" "$N" <<'PY'
import json, sys
name, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/scale_r2_{name}_n{n}.json").read().strip().splitlines()[-1])
    print(f"{name:22s} N={n}: {d['value']/1e6:8.3f} Mrays/s  {d['ms_per_step']:8.2f} ms/step  e2e {d['e2e']['value']/1e6:8.3f}  "
          f"mlp {d['roofline']['kernel_ms']:7.2f} ms  gather: {d['config']['gather'][:70]}")
except Exception as e:
    print(name, n, "FAILED", e)
PY
}
if [ "$N" -ge 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "two_rank or second_device" 2>&1 | tail -3
fi
A="--steps 6 --warmup 3 --no-fast-mode --no-cpu-baseline --no-extra"
run weak_cfg2_maps     $A --scaling weak   --gather maps
run strong_cfg2_maps   $A --scaling strong --gather maps
run strong_cfg2_labels $A --scaling strong --gather labels
run strong_cfg5_labels $A --config cfg5 --scaling strong --gather labels
[ "$N" -ge 4 ] || run strong_cfg3_labels $A --config cfg3 --scaling strong --gather labels
