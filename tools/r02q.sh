export DDMI_B200_IMAGE_TS=1
timeout 600 python -m pytest tests -m gpu -x -q -k "image or repeated" 2>&1 | tail -4
timeout 200 python tools/profile_image.py 2>&1 | tail -2 | head -1
DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so timeout 200 python tools/profile_image.py 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('TS image', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks'], d['parity'], d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:image_umma -s 1 -c 1 -o gpurun_out/r02q_image_ts python bench.py --batch 8 --res 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02q_ncu.log 2>&1
DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so timeout 200 python tools/profile_timeline.py > gpurun_out/r02q_timeline_ts.log 2>&1
