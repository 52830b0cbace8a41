#!/bin/bash
# GPU experiment: parity tests, then cfg3 bench for march variants (legacy loop, fast loop, warp tile shapes).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/sweep_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/sweep_pytest.log
tail -5 gpurun_out/sweep_pytest.log
VPE_MARCH_LEGACY=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_legacy.json 2> gpurun_out/sweep_legacy.err
for lw in 0 2 3 4 5; do
  VPE_MARCH_TILE_LOG2W=$lw python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_fast_lw$lw.json 2> gpurun_out/sweep_fast_lw$lw.err
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_fast_auto.json 2> gpurun_out/sweep_fast_auto.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sweep_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'march_ms=%.3f kern=%.3f fill_ms=%.3f samples=%d frac=%.4f' % (d['march']['ms'], d['march']['kernel_ms'], d['fill']['ms'], d['march']['ray_samples'], d['roofline']['frac']))
    except Exception as e:
        print(f, 'ERR', e)
PY
