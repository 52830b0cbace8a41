for t in 1 2 3 4 5; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tile-log2w $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tile_log2w=$t', 'march kern %.3f ms'%d['march']['kernel_ms'])"; done
