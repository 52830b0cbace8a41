#!/bin/bash
# usage: tools/gpu_n1_quick.sh <tag> [pytest args]   — GPU parity tests, then one N=1 bench line without the CPU legs
tag=${1:-quick}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-general-paths > gpurun_out/n1_${tag}.json 2> gpurun_out/n1_${tag}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/n1_${tag}.json").read().strip().splitlines()[-1])
print("${tag}: ms/step %.3f fill %.3f (kernel %.3f) march kernel %.3f skipped %.16f launches %s" % (
    d["ms_per_step"], d["fill"]["ms"], d["fill"]["kernel_ms"], d["march"]["kernel_ms"], d["march"]["skipped_sample_frac"], d["gpu_launches"]))
PY
tail -3 gpurun_out/n1_${tag}.err
