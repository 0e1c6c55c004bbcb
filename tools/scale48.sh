#!/bin/bash
# usage: tools/scale48.sh N -> parity (default comm mode) + bench at N GPUs
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 300 $TR tests/dist_gpu_check.py > gpurun_out/parity_$N.log 2>&1; echo "parity rc=$?"
grep "dist x\|DIST_\|CfemError\|Error" gpurun_out/parity_$N.log | grep -v "comm:" | sort | uniq | head -12
tools/scale_run.sh $N --steps 40
