#!/bin/bash
# usage: gpu_trace.sh N [cfg] [modes...]  — per-stage timeline of one frame at N GPUs (tools/slab_trace.py)
N=${1:-2}; cfg=${2:-cfg3}; shift; shift
for mode in ${@:-overlap nooverlap}; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/slab_trace.py $cfg $mode 2>&1 | grep " rank " | sort
done
