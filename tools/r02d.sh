# Dev tool (GPU box): parity tests + device-resident throughput of all four decoders.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('image', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks'])"
for w in occupancy video nerf; do timeout 200 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', d['config']['precision'], '%.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'])"; done
