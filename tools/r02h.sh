python - <<'PY' 2>&1 | tail -12
import sys, torch
sys.path.insert(0, '.')
from ddmi_b200 import _lib
dev='cuda:0'
L=_lib.lib()
B,C,H,W=2,64,128,128
plane = torch.randn(B,C,H,W, device=dev)
out = torch.zeros(64,2,64, device=dev)
mapd = torch.zeros(64, dtype=torch.int32, device=dev)
for variant in (1, 0):
    for (x,y,c) in ((0,0,0),(36,5,64),(100,127,64)):
        out.zero_()
        try:
            _lib.check(L.ddmi_selftest_tma(plane.data_ptr(), B,C,H,W, x,y,c, variant, mapd.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            ref = torch.zeros(64,2,64, device=dev)
            xs, ys = min(64, W-x), min(2, H-y)
            ref[:, :ys, :xs] = plane.reshape(B*C,H,W)[c:c+64, y:y+ys, x:x+xs]
            print('variant', variant, (x,y,c), 'max diff', float((out-ref).abs().max()))
        except Exception as e:
            print('variant', variant, (x,y,c), 'ERR', str(e)[:100]); break
PY

timeout 900 python -m pytest tests -m gpu -x -q -k "image or repeated" 2>&1 | tail -4
DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so timeout 200 python tools/profile_image.py 2>&1 | tail -2
DDMI_B200_NO_TMA_PATCH=1 DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so timeout 200 python tools/profile_image.py 2>&1 | tail -2
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('image', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks'], d['parity'])"
