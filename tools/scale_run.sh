#!/bin/bash
# usage: tools/scale_run.sh N [extra bench args]  -> prints a one-line summary of bench.py at N GPUs
N=$1; shift
if [ "$N" = 1 ]; then python bench.py --no-cpu-baseline "$@" > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N "$@" > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; fi
tail -1 gpurun_out/scale_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d value=%.1fM ms/step=%.3f e2e=%.1fM' % (d['n_gpus'], d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))
print('   breakdown', {k: round(v,3) for k,v in r['breakdown_ms_per_step'].items()})
print('   launches', r['launches_per_step'], 'tiles', d['config'].get('tiles'))"
grep -i "error\|Traceback" gpurun_out/scale_$N.err | head -3
