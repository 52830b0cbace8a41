#!/bin/bash
# usage: gpu_multi_one.sh N [bench args]  — one bench.py run at N GPUs (default path), JSON line into gpurun_out/
N=${1:-8}; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/multi_final_$N.json 2> gpurun_out/multi_final_$N.err
tail -1 gpurun_out/multi_final_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=$N ms/step=%.3f fill=%.3f march=%.3f (kernel %.3f) e2e_frame=%.3f slabs=%s' % (d['ms_per_step'], d['fill']['ms'], d['march']['ms'], d['march']['kernel_ms'], d['e2e']['frame_ms'], d['config']['slabs']))
print(d['config']['parallelism']); print(json.dumps(d.get('parity'))[:400])" || tail -8 gpurun_out/multi_final_$N.err
