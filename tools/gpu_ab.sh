#!/bin/bash
# A/B on one box: pytest -m gpu, then bench lines for several march kernels. usage: gpu_ab.sh [pytest-args]
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q $1 2>&1 | tail -5
for mk in 0 2; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --march-kernel $mk 2> gpurun_out/ab_$mk.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('march-kernel=$mk', 'march kern %.3f ms'%d['march']['kernel_ms'], 'fill kern %.3f ms'%d['fill']['kernel_ms'], 'fill %.3f ms'%d['fill']['ms'], 'samples', d['march']['ray_samples'])"; tail -3 gpurun_out/ab_$mk.err; done
