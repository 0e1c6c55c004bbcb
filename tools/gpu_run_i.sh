#!/bin/bash
# r02 session I: full GPU suite on the final build + A/B of the multi-pass assembly and the L2 set-aside policy
L=/root/repo/conservation-fem_b200/cfem_b200/libcfem_b200_oldasm.so
bash tools/gpu_ab.sh r02i --pytest \
  "new|X=1|--steps 40 --warmup 3" \
  "oldasm|CFEM_LIB=$L|--steps 40 --warmup 3" \
  "new2|X=1|--steps 40 --warmup 3" \
  "kpp_new|X=1|--workload kpp --steps 20 --warmup 3" \
  "kpp_oldasm|CFEM_LIB=$L|--workload kpp --steps 20 --warmup 3" \
  "euler|X=1|--workload euler --steps 5 --warmup 3"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02i_smoke.log 2>&1; echo "smoke rc=$?"
