#!/bin/bash
# usage: gpu_sweep_build2.sh "<name>:<EXTRA flags>" ...   (rebuilds the library per variant on the GPU box; march/fill kernel ms only)
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  make -C volumetric-particles-for-unity_b200/csrc clean >/dev/null; make -C volumetric-particles-for-unity_b200/csrc EXTRA="$flags" > gpurun_out/build_$name.log 2>&1 || { echo "build $name failed"; tail -5 gpurun_out/build_$name.log; continue; }
  grep -A2 "Function properties" gpurun_out/build_$name.log | grep -A2 "k_march_flatILi32ELb1ELb1ELb1E" | tail -2 | cut -c1-120
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/sweep_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', 'march kern %.3f ms'%d['march']['kernel_ms'], 'fill kern %.3f ms'%d['fill']['kernel_ms'], 'fill %.3f ms'%d['fill']['ms'])"
done
