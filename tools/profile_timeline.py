"""Dev tool (GPU box): timeline of ONE tile iteration of CTA 0 of the tcgen05 image kernel.
Needs the profiling build:  make -C ddmi_b200/csrc prof && DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so python tools/profile_timeline.py
Event ids (csrc/umma_engine.cuh, decode_umma.cu): E thread 0: 0x01 parks on the MMA barrier, 0x02 woke, 0x03 accumulator drained,
0x10+q quarter q published, 0x20 / 0x21 gather begin / end; MMA lane: 0x100+pc UNIT starts issuing, 0x200+pc WAIT begins,
0x300+pc WAIT satisfied, 0x400+pc COMMIT issued."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ddmi_b200 import _lib, convert_to_coord_format_2d, get_scale_injection
torch.set_grad_enabled(False)
B, R = 16, 1024
dev = 'cuda:0'
kind = os.environ.get('KIND', 'image')
g = torch.Generator().manual_seed(1)
if kind == 'image':
    m = bench.build_mlp().to(dev)
    planes = [torch.randn(B, 64, s, s, generator=g).to(dev) for s in (64, 128, 256)]
    e = (R - 1) / R
    c = convert_to_coord_format_2d(1, R, R, hstart=-e, hend=e, wstart=-e, wend=e).to(dev)
    run = lambda: m(c, hdbf=planes, si=get_scale_injection(R))
elif kind == 'nerf':
    import numpy as np
    import ddmi_b200
    from ddmi_b200 import nerf_helpers as nh
    m = ddmi_b200.MLPNeRF(D=6, W=256, in_channels_xyz=159, skips=[2, 4], in_channels_dir=27).to(dev)
    fea = {k: torch.randn(4, 32, 64, 64, generator=g).to(dev) for k in ('xy', 'yz', 'xz')}
    H = W = 128
    focal = .5 * W / np.tan(.5 * 0.6911112070083618)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    ro, rd = nh.get_rays(H, W, K, nh.pose_spherical(40.0, -20, 5)[:3, :4], dev)
    vd = (rd / torch.norm(rd, dim=-1, keepdim=True)).reshape(-1, 3)
    rays = torch.cat([ro.reshape(-1, 3), rd.reshape(-1, 3), 2. * torch.ones(H * W, 1, device=dev), 6. * torch.ones(H * W, 1, device=dev), vd], -1)
    run = lambda: nh.render_rays_fused(rays, fea, m, 128, True)
elif kind == 'video':
    import ddmi_b200
    m = ddmi_b200.MLPVideo(in_ch=2, latent_dim=64, out_ch=3, ch=256).to(dev)
    hd = [[torch.randn(4, 64, s, s, generator=g).to(dev) for s in (64, 128, 256)],
          [torch.randn(4, 64, 16, s, generator=g).to(dev) for s in (64, 128, 256)],
          [torch.randn(4, 64, 16, s, generator=g).to(dev) for s in (64, 128, 256)]]
    cv = ddmi_b200.convert_to_coord_format_3d(1, 256, 256, 16, hstart=-255 / 256, hend=255 / 256, wstart=-255 / 256,
                                              wend=255 / 256, tstart=-15 / 16, tend=15 / 16)
    cv = {k: v.to(dev) for k, v in cv.items()}
    run = lambda: m(cv, hd)
else:
    import ddmi_b200
    m = ddmi_b200.MLP3D(in_ch=3, latent_dim=64, out_ch=1, ch=256).to(dev)
    hdbf = tuple([torch.randn(4, 64, s, s, generator=g).to(dev) for s in (16, 32, 64)] for _ in range(3))
    if os.environ.get('PTS', 'grid') == 'lattice':
        ax = (1.1 * torch.linspace(-0.5, 0.5, 128)).to(dev)
        run = lambda: m.decode_logits_lattice((ax, ax, ax), hdbf)
    else:
        if os.environ.get('PTS', 'grid') == 'grid':
            pts = (1.1 * ddmi_b200.make_3d_grid((-.5,) * 3, (.5,) * 3, (128,) * 3)).to(dev)
        else:
            pts = ((torch.rand(2000000, 3, generator=g) - 0.5) * 1.1).to(dev)
        run = lambda: m(pts[None].expand(4, -1, -1), hdbf).logits
m.precision = os.environ.get('PREC', 'f16f8')
L = _lib.lib()
run()
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * 8192)()
n = ctypes.c_int32()
_lib.check(L.ddmi_debug_trace(buf, 8192, ctypes.byref(n), 1))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); run(); t1.record()
torch.cuda.synchronize()
print(f"{kind}: {t0.elapsed_time(t1):.2f} ms")
_lib.check(L.ddmi_debug_trace(buf, 8192, ctypes.byref(n), 1))
ev = sorted(((buf[i] & ((1 << 48) - 1)), buf[i] >> 48) for i in range(n.value) if buf[i])
if not ev:
    print("no trace records: load the profiling build (DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so)")
    sys.exit(1)
t0 = ev[0][0]
names = {0x05: 'E tile done', 0x01: 'E park', 0x02: 'E woke (COMMIT 0)', 0x03: 'E drained', 0x04: 'E COMMIT 1 seen', 0x14: 'E acc cols 128.. drained', 0x15: 'E acc cols 0..127 drained',
         0x20: 'E gather begin', 0x21: 'E gather end'}
prev = t0
for t, i in ev:
    if i >= 0x700: nm = {0x700: 'B wait X free', 0x710: 'B A4 (raw X ready)', 0x720: 'B D1 seen', 0x730: 'B A5 (relu X ready)'}.get(i & 0xFF0, hex(i)) + f' s{i & 15}'
    elif i >= 0x600: nm = f'B blended s{(i - 0x600) // 64} c{(i - 0x600) % 64}'
    elif i in (0x500, 0x501, 0x502): nm = {0x500: 'M unit begins', 0x501: 'M  MMAs issued', 0x502: 'M  slots released'}[i]
    elif i >= 0x500: nm = f'I issued s{(i - 0x500) // 64} c{(i - 0x500) % 64}'
    elif i >= 0x400: nm = f'M COMMIT pc{i - 0x400}'
    elif i >= 0x300: nm = f'M WAIT ok pc{i - 0x300}'
    elif i >= 0x200: nm = f'M WAIT .. pc{i - 0x200}'
    elif i >= 0x100: nm = f'M UNIT pc{i - 0x100}'
    elif 0x10 <= i < 0x14: nm = f'E publish q{i - 0x10}'
    else: nm = names.get(i, hex(i))
    print(f"{t - t0:8d} (+{t - prev:6d})  {nm}")
    prev = t
