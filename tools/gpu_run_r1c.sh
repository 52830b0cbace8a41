#!/bin/bash
# parity tests + bench + one ncu --set full capture of the two hot kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/sweep_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/sweep_pytest.log
tail -5 gpurun_out/sweep_pytest.log
rm -f gpurun_out/sweep_*.json
run() { name=$1; shift; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; }
run default VPE_X=1
ncu --set full --clock-control none --import-source on -k regex:'k_march|k_fill_columns' -s 4 -c 4 -f -o gpurun_out/r01c python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r01c.log 2>&1
tail -3 gpurun_out/ncu_r01c.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sweep_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'march_ms=%.3f kern=%.3f fill_ms=%.3f fill_kern=%.3f samples=%d frac=%.4f e2e_march=%.2f e2e_fill=%.2f' % (d['march']['ms'], d['march']['kernel_ms'], d['fill']['ms'], d['fill']['kernel_ms'], d['march']['ray_samples'], d['roofline']['frac'], d['e2e']['march_ms'], d['e2e']['fill_ms']))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
