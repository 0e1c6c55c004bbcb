#!/bin/bash
# N-GPU session: distributed parity, then weak-scaling points with the device-side wait accounting
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_gpu_check.py > gpurun_out/r02f_dist_$N.log 2>&1
grep "dist x\|DIST_\|Error" gpurun_out/r02f_dist_$N.log | grep -v "comm:" | sort | uniq | head -12
run() { tag=$1; shift; envs=$1; shift
  env $envs bash tools/scale_run.sh "$@"; n=$1
  cp gpurun_out/scale_$n.json gpurun_out/r02f_${tag}.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02f_${tag}.json").read().strip().splitlines()[-1]); print("   parity", d.get("parity_rel_l2")); print("   comm_wait", d["config"].get("comm_wait"))
PY
}
run n1 "A=1" 1 --steps 20 --warmup 3 --no-parity
run n1_chebchain "CFEM_CHEB=chain" 1 --steps 20 --warmup 3 --no-parity
run n1_forceghost "CFEM_FORCE_GHOST=1" 1 --steps 20 --warmup 3 --no-parity
run n$N "A=1" $N --steps 20 --warmup 3
run n${N}_chebchain "CFEM_CHEB=chain" $N --steps 20 --warmup 3 --no-parity
run n${N}_merged "CFEM_BICGSTAB=merged CFEM_CHEB=chain" $N --steps 20 --warmup 3 --no-parity
