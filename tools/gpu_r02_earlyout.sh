#!/bin/bash
# march kernel time against marchEarlyOutTransmittance (NOT reference semantics: sample counts change; profiles/r02_march.md)
for eo in 0 1e-7 1e-5 1e-3; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --early-out $eo | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eo=$eo', 'march kern %.3f ms'%d['march']['kernel_ms'], 'samples', d['march']['ray_samples'])"; done
