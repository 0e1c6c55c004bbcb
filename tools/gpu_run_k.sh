#!/bin/bash
# r02 session K: barrier word placement / poll backoff, ghost-variant occupancy, Euler e2e through device staging
D=/root/repo/conservation-fem_b200/cfem_b200
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "euler" > gpurun_out/r02k_pytest.log 2>&1; tail -2 gpurun_out/r02k_pytest.log
bash tools/gpu_ab.sh r02k \
  "sep|X=1|--steps 40 --warmup 3 --no-parity" \
  "same|CFEM_LIB=$D/libcfem_b200_same.so|--steps 40 --warmup 3 --no-parity" \
  "bo|CFEM_LIB=$D/libcfem_b200_bo.so|--steps 40 --warmup 3 --no-parity" \
  "sep2|X=1|--steps 40 --warmup 3 --no-parity" \
  "ghost6|CFEM_FORCE_GHOST=1|--steps 40 --warmup 3 --no-parity" \
  "ghost5|CFEM_FORCE_GHOST=1 CFEM_LIB=$D/libcfem_b200_g5.so|--steps 40 --warmup 3 --no-parity" \
  "euler|X=1|--workload euler --steps 5 --warmup 3"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02k_euler.json").read().strip().splitlines()[-1])
print("euler ms/step %.2f e2e DoF/s %.3e -> e2e ms/step %.1f" % (d["ms_per_step"], d["e2e"]["value"], d["config"]["dofs"]/d["e2e"]["value"]*1e3))
PY
