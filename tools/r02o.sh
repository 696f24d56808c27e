export DDMI_B200_IMAGE_TS=1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('TS image', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks'], d['parity'], d['roofline']['frac'])"
