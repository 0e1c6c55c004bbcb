#!/bin/bash
# usage: tools/comm_ab.sh N  -> parity + comm latency + bench at N GPUs, new vs old comm kernels
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 300 $TR tests/dist_gpu_check.py > gpurun_out/ab_parity.log 2>&1; echo "parity rc=$?"
grep "dist x\|DIST_\|CfemError\|Error" gpurun_out/ab_parity.log | grep -v "comm:" | sort | uniq | head -12
timeout 200 $TR tools/comm_bench.py 2>&1 | grep "us per op"
CFEM_PUSH=kernel CFEM_ALLREDUCE=ticket timeout 200 $TR tools/comm_bench.py 2>&1 | grep "us per op"
tools/scale_run.sh $N --steps 40
CFEM_PUSH=kernel tools/scale_run.sh $N --steps 40
