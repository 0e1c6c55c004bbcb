#!/bin/bash
# 2-GPU session: distributed parity (default comm mode = fused halo + in-kernel all-reduce), weak-scaling point N=2, N=1 reference
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_gpu_check.py > gpurun_out/r02d_dist_$N.log 2>&1
grep "dist x\|DIST_\|Error" gpurun_out/r02d_dist_$N.log | grep -v "comm:" | sort | uniq | head -12
bash tools/scale_run.sh 1 --steps 20 --warmup 3 --no-parity
cp gpurun_out/scale_1.json gpurun_out/r02d_scale_1.json
bash tools/scale_run.sh $N --steps 20 --warmup 3
cp gpurun_out/scale_$N.json gpurun_out/r02d_scale_$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_$N.json").read().strip().splitlines()[-1]); print("parity", d.get("parity_rel_l2")); print("comm", d["config"].get("comm"))
PY
CFEM_BICGSTAB=5k bash tools/scale_run.sh $N --steps 20 --warmup 3 --no-parity
cp gpurun_out/scale_$N.json gpurun_out/r02d_scale_${N}_5k.json
