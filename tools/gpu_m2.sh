#!/bin/bash
# usage: gpu_m2.sh N "<name>:<bench args>" ...   — bench.py at N GPUs for several argument sets, bounded in time
N=${1:-2}; shift
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%:*}; args=${v#*:}
  timeout ${RUN_TIMEOUT:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 $args > gpurun_out/multi_${name}_$N.json 2> gpurun_out/multi_${name}_$N.err
  tail -1 gpurun_out/multi_${name}_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
p=d.get('parity') or {}
print('$name N=$N ms/step=%.3f fill=%.3f march=%.3f (kernel %.3f) e2e_frame=%.3f slabs=%s timeouts=%s vs1gpu=%s oracle=%s' % (d['ms_per_step'], d['fill']['ms'], d['march']['ms'], d['march']['kernel_ms'], d['e2e']['frame_ms'], d['config']['slabs'], p.get('link_timeouts'), (p.get('vs_single_gpu') or {}).get('max_rel_err'), (p.get('vs_oracle') or {}).get('max_rel_err')))" || tail -8 gpurun_out/multi_${name}_$N.err
done
