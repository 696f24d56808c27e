# Dev tool (GPU box): image parity tests + device-resident timing (shipping and profiling builds) + the bench line
timeout 600 python -m pytest tests -m gpu -x -q -k "image or repeated" 2>&1 | tail -3
timeout 200 python tools/profile_image.py 2>&1 | tail -2 | head -1
DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so timeout 200 python tools/profile_image.py 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('image', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks'], d['parity']['max_abs_vs_fp32_kernel'], 'frac %.4f' % d['roofline']['frac'])"
