"""Dev tool (GPU box): sustained L2 -> shared-memory bulk-copy rate per SM vs bytes in flight (csrc/microbench.cu::ringbench_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddmi_b200 import _lib
dev = 'cuda:0'
span = 4 << 20                                  # 4 MB window: the size of one packed weight image, L2-resident
src = torch.randint(0, 255, (span,), dtype=torch.uint8, device=dev)
out = torch.zeros(2, dtype=torch.int64, device=dev)
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
print("ctas slot_KB nslots in-flight_KB | cycles/slot  B/clk/SM  chip B/clk | GB/s chip")
for ctas in (1, 148):
    for slot_kb, nslots in ((8, 2), (8, 4), (8, 8), (8, 16), (8, 24), (16, 4), (16, 8), (16, 12), (32, 2), (32, 4), (32, 6)):
        iters = 4000
        for _ in range(2):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            _lib.check(L.ddmi_debug_ringbench(src.data_ptr(), span, slot_kb * 1024, nslots, iters, ctas, out.data_ptr(), st))
            t1.record()
            torch.cuda.synchronize()
        cyc = out.cpu().tolist()[0]
        ms = t0.elapsed_time(t1)
        bpc = slot_kb * 1024 * iters / cyc
        print(f"{ctas:4d} {slot_kb:6d} {nslots:6d} {slot_kb * nslots:10d} | {cyc / iters:10.1f} {bpc:9.1f} {bpc * ctas:10.0f} | "
              f"{ctas * slot_kb * 1024 * iters / ms / 1e6:8.1f}")
