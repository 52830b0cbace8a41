#!/bin/bash
# usage: gpu_multi.sh N   — slab tests on >= 2 GPUs, then bench.py at N GPUs: linked sweep vs NCCL band pipeline
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 $EXTRA_ARGS > gpurun_out/multi_${name}_$N.json 2> gpurun_out/multi_${name}_$N.err; tail -1 gpurun_out/multi_${name}_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$name N=$N ms/step=%.3f fill=%.3f march=%.3f (kernel %.3f) e2e_frame=%.3f' % (d['ms_per_step'], d['fill']['ms'], d['march']['ms'], d['march']['kernel_ms'], d['e2e']['frame_ms']))" || tail -5 gpurun_out/multi_${name}_$N.err; }
run linked VPE_X=1
run ncclimage VPE_SLAB_NCCL_IMAGE=1
[ -n "$WITH_NCCL" ] && run nccl VPE_SLAB_NCCL_SWEEP=1
