"""Dev tool (GPU box): per-call times of the plane-producer tail kernels against eager torch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddmi_b200
from oracle import plane_tail_oracle as po
torch.set_grad_enabled(False)
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
dev = 'cuda:0'; B = 16
t = ddmi_b200.PlaneTail(128, 64, (None, 256, 512)).to(dev)
sd = {k: v.detach() for k, v in t.state_dict().items()}
hs = [torch.randn(B, 512, 64, 64, device=dev), torch.randn(B, 256, 128, 128, device=dev), torch.randn(B, 128, 256, 256, device=dev)]
def tm(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / 5
for name, ours, ref, flop in (('head 512->64 @64^2', lambda: t.head(2, hs[0]), lambda: po.head(sd, 2, hs[0]), 2 * B * 64 * 64 * 512 * 64),
                              ('head 256->64 @128^2', lambda: t.head(1, hs[1]), lambda: po.head(sd, 1, hs[1]), 2 * B * 128 * 128 * 256 * 64),
                              ('tail 128->64 3x3 @256^2', lambda: t.tail(hs[2]), lambda: po.tail(sd, hs[2]), 2 * B * 256 * 256 * 128 * 9 * 64),
                              ('tail, channels-last out', lambda: t.tail(hs[2], True), lambda: po.tail(sd, hs[2]), 2 * B * 256 * 256 * 128 * 9 * 64)):
    a, b = tm(ours), tm(ref)
    print(f"{name}: ours {a:.3f} ms ({flop / a / 1e9:.1f} TFLOP/s)   eager torch {b:.3f} ms ({flop / b / 1e9:.1f} TFLOP/s)")
