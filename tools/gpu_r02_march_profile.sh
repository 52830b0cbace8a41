#!/bin/bash
# One B200: bench line + one ncu --set full capture of the full-image march launch (source-level counters).
# usage: tools/gpu_r02_march_profile.sh <tag> [kernel regex] [extra bench args]
tag=${1:-r02}; pat=${2:-k_march}; extra=$3
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline $extra > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); print('$tag', 'march kern %.3f ms'%d['march']['kernel_ms'], 'fill kern %.3f ms'%d['fill']['kernel_ms'], 'e2e march %.3f'%d['e2e']['march_ms'])"
# launches of the march kernel: 3 warm-up + 1 timed full-image launches, then the host path's bands: capture the timed one
ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 3 -c 1 -f -o gpurun_out/${tag}_march python bench.py --steps 1 --warmup 3 --no-cpu-baseline $extra > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
