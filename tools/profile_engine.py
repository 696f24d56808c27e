"""Dev tool: MMA-thread cycle counters (CTA 0) for any engine kernel: python tools/profile_engine.py {video|nerf|occupancy}"""
import ctypes, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ddmi_b200 import _lib
kind = sys.argv[1]
buf = (ctypes.c_uint64 * 8)()
class A: pass
a = A(); a.workload = kind; a.batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4; a.steps = 1; a.warmup = 1; a.precision = os.environ.get("DDMI_B200_PRECISION", "f16f8"); a.no_cpu_baseline = True; a.e2e_chunk = 8; a.cpu_coords = 131072
import io, contextlib
bench.ARGS = a
_lib.check(_lib.lib().ddmi_debug_set(int(os.environ.get('DBG', '0'))))   # what-if switches (1: no epilogue work, 2: no MMAs, 4: no gathers)
with contextlib.redirect_stdout(io.StringIO()):
    bench.run_other(a)            # warm-up + 1 step
torch.cuda.synchronize()
_lib.check(_lib.lib().ddmi_debug_profile(buf, 1))
a.warmup = 0
s = io.StringIO()
with contextlib.redirect_stdout(s):
    bench.run_other(a)
_lib.check(_lib.lib().ddmi_debug_profile(buf, 1))
d = json.loads(s.getvalue().strip().splitlines()[-1])
v = list(buf); tiles = max(v[6], 1)
print(f"{kind}: {d['value']:.3e} coords/s; per pair-tile cycles (CTA 0, {v[6]} tiles): MMA wait-operands {v[3]/tiles:.0f}  "
      f"wait-weights {v[4]/tiles:.0f}  MMA total {v[5]/tiles:.0f}")
