#!/bin/bash
# usage: gpu_multi.sh N [bench args]  — slab tests on >= 2 GPUs, then bench.py at N GPUs: default path, and the sweep after the density pass
N=${1:-2}; shift
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift; timeout ${RUN_TIMEOUT:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/multi_${name}_$N.json 2> gpurun_out/multi_${name}_$N.err; tail -1 gpurun_out/multi_${name}_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$name N=$N ms/step=%.3f fill=%.3f march=%.3f (kernel %.3f) e2e_frame=%.3f slabs=%s parity=%s' % (d['ms_per_step'], d['fill']['ms'], d['march']['ms'], d['march']['kernel_ms'], d['e2e']['frame_ms'], d['config']['slabs'], json.dumps(d.get('parity'))[:300]))" || tail -8 gpurun_out/multi_${name}_$N.err; }
run overlap "$@"
run nooverlap --no-sweep-overlap --no-cpu-baseline "$@"
