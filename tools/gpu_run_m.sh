#!/bin/bash
# r02 session M: staged tile kernels with separate interior / boundary tile loops
bash tools/gpu_ab.sh r02m --pytest \
  "plain|X=1|--steps 40 --warmup 3 --no-parity" \
  "ghost|CFEM_FORCE_GHOST=1|--steps 40 --warmup 3 --no-parity" \
  "kpp|X=1|--workload kpp --steps 20 --warmup 3 --no-parity"
python - <<'PY'
import json
for n in ("plain","ghost","kpp"):
    d=json.loads(open(f"gpurun_out/r02m_{n}.json").read().strip().splitlines()[-1])
    print(n, d["config"].get("solve_timing"))
PY
