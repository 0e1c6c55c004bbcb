#!/bin/bash
# r02 session J: persistent BiCGStab with tile flags instead of the third grid barrier (flags), plus LL-tagged barrier totals (new)
L=/root/repo/conservation-fem_b200/cfem_b200/libcfem_b200_flags.so
bash tools/gpu_ab.sh r02j --pytest \
  "new|X=1|--steps 40 --warmup 3" \
  "flags|CFEM_LIB=$L|--steps 40 --warmup 3 --no-parity" \
  "ghost|CFEM_FORCE_GHOST=1|--steps 40 --warmup 3 --no-parity" \
  "kpp|X=1|--workload kpp --steps 20 --warmup 3 --no-parity" \
  "new2|X=1|--steps 40 --warmup 3 --no-parity"
