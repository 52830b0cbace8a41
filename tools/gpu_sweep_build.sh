#!/bin/bash
# usage: gpu_sweep_build.sh "<name>:<EXTRA flags>" ...   (rebuilds the library per variant on the GPU box; bench only)
mkdir -p gpurun_out; rm -f gpurun_out/sweep_*.json
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  make -C volumetric-particles-for-unity_b200/csrc clean >/dev/null; make -C volumetric-particles-for-unity_b200/csrc EXTRA="$flags" > gpurun_out/build_$name.log 2>&1 || { echo "build $name failed"; tail -5 gpurun_out/build_$name.log; continue; }
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sweep_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'march_ms=%.3f kern=%.3f fill_ms=%.3f fill_kern=%.3f frac=%.4f' % (d['march']['ms'], d['march']['kernel_ms'], d['fill']['ms'], d['fill']['kernel_ms'], d['roofline']['frac']))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
