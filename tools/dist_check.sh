#!/bin/bash
# usage: tools/dist_check.sh N   -> multi-GPU parity check in the three communication modes
N=$1
for mode in "CFEM_COMM=nccl" "CFEM_HALO=exchange" "CFEM_HALO=fused"; do
  echo "== $mode"
  env $mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_gpu_check.py > gpurun_out/dist_$N.log 2>&1
  grep "dist x\|DIST_\|CfemError" gpurun_out/dist_$N.log | grep -v "comm:" | sort | uniq | head -8
done
