export DDMI_B200_IMAGE_TS=1 DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so
for ns in full mixed; do
  echo "=== TS NSPLIT=$ns"
  DDMI_B200_NSPLIT=$ns timeout 200 python tools/profile_image.py 2>&1 | tail -2
done
DDMI_B200_NSPLIT=full timeout 200 python tools/profile_timeline.py > gpurun_out/r02l_timeline_ts_full.log 2>&1
python - <<'PY'
import os, sys
sys.path.insert(0, '.')
import bench
from ddmi_b200 import packing, _lib
m = bench.build_mlp()
p = packing.pack_image(m, 0.25, _lib.PREC_F16F8)
print('ts program:', p.ts, 'ops', len(p.program_host), 'stream MB', p.gemm.numel() / 1e6)
PY
