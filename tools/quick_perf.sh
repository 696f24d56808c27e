# Dev tool (GPU box): parity tests + in-kernel profile of the image kernel + device-resident throughput of all four decoders.
timeout 500 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -2
timeout 200 python tools/profile_image.py 2>&1 | tail -2
for w in occupancy video nerf; do timeout 200 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', d['config']['precision'], '%.4g' % d['value'])"; done
