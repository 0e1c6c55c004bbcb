#!/bin/bash
# r02 session N: one host round trip per Newton iteration (async persistent solve) vs CFEM_STEP_SYNC=1
bash tools/gpu_ab.sh r02n --pytest \
  "async|X=1|--steps 40 --warmup 3" \
  "sync|CFEM_STEP_SYNC=1|--steps 40 --warmup 3 --no-parity" \
  "kpp|X=1|--workload kpp --steps 20 --warmup 3 --no-parity" \
  "kpp_sync|CFEM_STEP_SYNC=1|--workload kpp --steps 20 --warmup 3 --no-parity"
