#!/bin/bash
# usage: gpu_sweep_env.sh "<name>:<ENV=VAL ...>" ...   (same library, different run-time switches)
mkdir -p gpurun_out; rm -f gpurun_out/sweep_*.json
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sweep_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'march_ms=%.3f kern=%.3f fill_ms=%.3f fill_kern=%.3f frac=%.4f e2e_march=%.2f e2e_fill=%.2f' % (d['march']['ms'], d['march']['kernel_ms'], d['fill']['ms'], d['fill']['kernel_ms'], d['roofline']['frac'], d['e2e']['march_ms'], d['e2e']['fill_ms']))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
