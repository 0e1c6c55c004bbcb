#!/bin/bash
# r02 session O: persistent BiCGStab with the verdict from the recurrence (two reductions per iteration)
bash tools/gpu_ab.sh r02o --pytest \
  "new|X=1|--steps 40 --warmup 3" \
  "kpp|X=1|--workload kpp --steps 20 --warmup 3 --no-parity"
