# Dev tool (GPU box): occupancy-focused check.
timeout 600 python -m pytest tests -m gpu -x -q -k "occupancy or repeated or f16f8_holds or channels_last or mismatched" 2>&1 | tail -8
for w in occupancy; do timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', d['config']['precision'], '%.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'])"; done
