import sys, math, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, ddmi_oracle as orc
from ddmi_b200 import packing
torch.set_grad_enabled(False)
m = cases.build_module('image')
coords, planes, si = cases.image_inputs(batch=2, sizes=(16, 32, 64), res=96)
f = packing.fold_image(m, si)
# features per pixel (float64)
b = planes[0].shape[0]
grid = coords.repeat(b, 1, 1, 1).permute(0, 2, 3, 1).contiguous()
X = [torch.nn.functional.grid_sample(p, grid, padding_mode='border', align_corners=False).permute(0, 2, 3, 1).reshape(-1, 64).double() for p in planes]

def q(x, fmt):
    if fmt == 'exact': return x
    dt = {'bf16': torch.bfloat16, 'fp16': torch.float16}[fmt]
    return x.float().to(dt).double()
def split(x, fmt):
    hi = q(x, fmt); lo = q(x - hi, fmt); return hi, lo
def gemm(A, W, scheme):
    # returns A @ W.T with operand rounding per scheme
    if scheme == 'exact': return A @ W.t()
    fmt, mode = scheme
    Ah, Al = split(A, fmt); Wh, Wl = split(W, fmt)
    if mode == '3': return Ah @ Wh.t() + Al @ Wh.t() + Ah @ Wl.t()
    if mode == '2A': return Ah @ Wh.t() + Al @ Wh.t()          # activations split, weights single
    if mode == '2W': return Ah @ Wh.t() + Ah @ Wl.t()          # weights split, activations single
    if mode == '1': return Ah @ Wh.t()
    if mode == '4': return (Ah + Al) @ (Wh + Wl).t()
def run(scheme):
    lr = lambda v: torch.nn.functional.leaky_relu(v, 0.2)
    g = math.sqrt(2.0)
    h = None
    for i, d in enumerate(f['blocks']):
        x = X[i] if i < 3 else None
        inp = x if h is None else (torch.cat([h, x], 1) if x is not None else h)
        c = g * lr(gemm(inp, d['W1'], scheme) + d['b1'])
        c = g * lr(gemm(c, d['W2'], scheme) + d['b2'])
        c = lr(gemm(c, d['W3'], scheme) + d['b3'])
        sk = gemm(inp, d['Ws'], scheme) + d['cs'] if d['Ws'] is not None else inp / g
        h = c + sk
    return gemm(h, f['Wrgb'], scheme) + f['brgb']
ref = run('exact')
print('|out| max', float(ref.abs().max()))
for sch in [('bf16','3'), ('fp16','3'), ('fp16','2A'), ('fp16','2W'), ('fp16','1'), ('bf16','2A'), ('bf16','2W'), ('bf16','1')]:
    o = run(sch)
    e = (o - ref).abs()
    print(sch, 'max %.3e mean %.3e' % (float(e.max()), float(e.mean())))
