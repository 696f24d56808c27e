"""CPU oracle of the D2C-VAE decode path.  TEST INFRASTRUCTURE ONLY.

A restatement of the reference's algorithm in plain torch CPU ops, taking a
state dict with the reference's parameter names.  Only tests/,
__graft_entry__.smoke() and bench.py's baseline legs (cpu_baseline, --impl reference, gpu_eager_baseline) may
import this module; the product package (ddmi_b200/) never does.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md §4, F9), so the oracle is pinned against the reference ITSELF:
oracle/make_golden.py imports /root/reference (models/d2c_vae/mlp.py,
utils/nerf_helpers.py) in the build container, runs it on seeded inputs and
commits the outputs under tests/golden/; tests/test_oracle_golden.py checks this
file against those outputs on every run.

The third-party arithmetic the reference itself calls (torch's grid_sample,
linear, gelu, softplus, sigmoid, cumprod) is used here as-is: those ops are the
reference's arithmetic by execution.
"""
import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# sampling (utils/general_utils.py:71-94, 115-148)
# ---------------------------------------------------------------------------
def normalize_coordinate(p, plane, padding=0.1):
    """general_utils.py:71-94: pick two axes, map to [0, 1), clamp outliers."""
    sel = {'xz': [0, 2], 'xy': [0, 1], 'yz': [1, 2]}[plane]
    xy = p[:, :, sel] / (1 + padding + 10e-6) + 0.5
    xy = torch.where(xy >= 1, torch.full_like(xy, 1 - 10e-6), xy)
    xy = torch.where(xy < 0, torch.zeros_like(xy), xy)
    return xy


def sample_plane_feature(p, plane):
    """general_utils.py:115-119 -> grid (B,N,1,2) in [-1,1]."""
    return 2.0 * normalize_coordinate(p.clone(), plane)[:, :, None].float() - 1.0


def _gs(plane, grid, align):
    return F.grid_sample(plane, grid, padding_mode='border', align_corners=align, mode='bilinear')


def triplane_add(p1, p2, p3, c1, c2, c3):
    """general_utils.py:126-131."""
    x = _gs(p1, c1, True).squeeze(-1)
    x = x + _gs(p2, c2, True).squeeze(-1)
    x = x + _gs(p3, c3, True).squeeze(-1)
    return x


def triplane_concat(p1, p2, p3, c1, c2, c3):
    """general_utils.py:134-145: rows ordered (t,h,w), channels [xy,yt,xt]."""
    x1, x2, x3 = _gs(p1, c1, True), _gs(p2, c2, True), _gs(p3, c3, True)
    b, c, h, w = x1.shape
    t = x2.shape[2]
    x1 = x1[:, :, None].expand(b, c, t, h, w)
    x2 = x2[..., None].expand(b, c, t, h, w)
    x3 = x3[:, :, :, None].expand(b, c, t, h, w)
    x = torch.cat((x1, x2, x3), dim=1).reshape(b, 3 * c, -1)
    return x.permute(0, 2, 1).reshape(-1, 3 * c)


# ---------------------------------------------------------------------------
# image MLP (models/d2c_vae/mlp.py:34-66; blocks.py:11-23,139-173,187-283,
# 286-297,312-356,390-412,604-638; op/fused_act.py:75-88)
# ---------------------------------------------------------------------------
def _sinusoidal(x, dim):
    half = dim // 2
    e = torch.exp(torch.arange(half, dtype=x.dtype, device=x.device) * -(math.log(10000) / (half - 1)))
    e = x[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def _equal_linear(sd, pre, x):
    w = sd[pre + '.weight']
    return F.linear(x, w * (1 / math.sqrt(w.shape[1])), sd[pre + '.bias'])


def _modconv(sd, pre, x, style, demodulate):
    """ModulatedConv2d 1x1 (blocks.py:242-283), per-sample weights."""
    b, cin, h, w = x.shape
    W = sd[pre + '.weight']                                  # (1,out,in,1,1)
    s = _equal_linear(sd, pre + '.modulation', style).view(b, 1, cin, 1, 1)
    weight = (1 / math.sqrt(cin)) * W * s
    if demodulate:
        weight = weight * torch.rsqrt(weight.pow(2).sum([2, 3, 4]) + 1e-8).view(b, -1, 1, 1, 1)
    out = torch.einsum('boi,bihw->bohw', weight[:, :, :, 0, 0], x)
    return out


def _styled_conv(sd, pre, x, style, noise=None):
    """StyledConv.forward (blocks.py:349-356): modulated conv -> NoiseInjection (blocks.py:292-297: image + weight * noise)
    -> FusedLeakyReLU.  The reference draws the noise inside forward; parity is defined for noise.weight == 0 or for an
    EXPLICIT noise tensor (b,1,h,w), the `noise=` argument NoiseInjection.forward already has."""
    out = _modconv(sd, pre + '.conv', x, style, True)
    nw = sd[pre + '.noise.weight']
    if float(nw.abs().max()) != 0.0:
        if noise is None:
            raise ValueError("noise.weight != 0 needs an explicit noise tensor (the reference's own draw is not reproducible)")
        out = out + nw * noise.to(out.dtype)
    return F.leaky_relu(out + sd[pre + '.activate.bias'].view(1, -1, 1, 1), 0.2) * math.sqrt(2)


def _styled_res_block(sd, pre, x, style, noise=(None, None, None)):
    out = _styled_conv(sd, pre + '.conv1', x, style, noise[0])
    out = _styled_conv(sd, pre + '.conv2', out, style, noise[1])
    out = _styled_conv(sd, pre + '.conv3', out, style, noise[2])
    key = pre + '.skip.0.weight'
    if key in sd:
        w = sd[key]
        skip = F.conv2d(x, w * (1 / math.sqrt(w.shape[1])))
    else:
        skip = x
    return (out + skip) / math.sqrt(2)


# ---- the documented noise stream of the drop-in (ddmi_b200/csrc/common.cuh::philox_normal4) -------------------------
def philox4x32_10(c, k):
    """Philox-4x32-10 (Salmon et al., SC'11).  c: (..., 4) uint32 counters, k: (2,) uint32 key -> (..., 4) uint32."""
    import numpy as np
    c = [c[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    M0, M1, W0, W1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & mask]
        k0, k1 = (k0 + W0) & mask, (k1 + W1) & mask
    return np.stack(c, -1).astype(np.uint32)


def philox_noise(seed, batch, n):
    """The noise tensors MLP.forward(..., noise=<int seed>) uses, as 12 tensors (batch,1,n): layer l (0..11, order
    net_res1.conv1, conv2, conv3, net_res2.conv1, ...), item b, coordinate g:
      x = Philox4x32-10(key = (seed & 0xffffffff, seed >> 32), counter = (g & 0xffffffff, g >> 32, b, l // 3)),
      u_i = ((x_i >> 9) + 0.5) * 2^-23,  r = sqrt(-2 ln u_0),  noise = r cos(2 pi u_1), r sin(2 pi u_1), sqrt(-2 ln u_2) cos(2 pi u_3)
      for conv1, conv2, conv3 of block l // 3 (fp32 arithmetic)."""
    import numpy as np
    g = np.arange(n, dtype=np.uint64)
    out = []
    for blk in range(4):
        c = np.zeros((batch, n, 4), dtype=np.uint32)
        c[..., 0] = (g & np.uint64(0xFFFFFFFF)).astype(np.uint32)[None]
        c[..., 1] = (g >> np.uint64(32)).astype(np.uint32)[None]
        c[..., 2] = np.arange(batch, dtype=np.uint32)[:, None]
        c[..., 3] = blk
        x = philox4x32_10(c, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
        u = ((x >> 9).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
        r0 = np.sqrt(np.float32(-2.0) * np.log(u[..., 0]))
        r1 = np.sqrt(np.float32(-2.0) * np.log(u[..., 2]))
        a0, a1 = np.float32(2 * math.pi) * u[..., 1], np.float32(2 * math.pi) * u[..., 3]
        for v in (r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1)):
            out.append(torch.from_numpy(v.astype(np.float32)).reshape(batch, 1, n))
    return out


def image_decode(sd, coords, hdbf, si=1.0, noise=None):
    """MLP.forward (mlp.py:34-66).  coords (1,2,h,w); hdbf 3 planes (b,64,S,S).  noise: None, or the 12 explicit
    (b,1,h,w) tensors the 12 NoiseInjection modules add (order net_res1.conv1 .. net_res4.conv3)."""
    b = hdbf[0].shape[0]
    dt = hdbf[0].dtype
    coords = coords.to(dt).repeat(b, 1, 1, 1)
    sip = torch.ones_like(coords) * si
    grid = coords.permute(0, 2, 3, 1).contiguous()
    style = _sinusoidal(torch.ones(b, dtype=dt, device=hdbf[0].device) * si, 64)
    style = F.linear(style, sd['time_mlp.1.weight'], sd['time_mlp.1.bias'])
    style = F.linear(F.gelu(style), sd['time_mlp.3.weight'], sd['time_mlp.3.bias'])
    feats = [torch.cat((_gs(p, grid, False), sip), dim=1) for p in hdbf]
    h, w = coords.shape[2:]
    nz = [None] * 12 if noise is None else [t.reshape(b, 1, h, w) for t in noise]
    x = _styled_res_block(sd, 'net_res1', feats[0], style, nz[0:3])
    x = _styled_res_block(sd, 'net_res2', torch.cat((x, feats[1]), dim=1), style, nz[3:6])
    x = _styled_res_block(sd, 'net_res3', torch.cat((x, feats[2]), dim=1), style, nz[6:9])
    x = _styled_res_block(sd, 'net_res4', x, style, nz[9:12])
    return _modconv(sd, 'torgb.conv', x, style, False) + sd['torgb.bias']


# ---------------------------------------------------------------------------
# ResnetBlockFC decoders (blocks.py:673-716)
# ---------------------------------------------------------------------------
def _resnet_fc(sd, pre, x):
    net = F.linear(F.relu(x), sd[pre + '.fc_0.weight'], sd[pre + '.fc_0.bias'])
    dx = F.linear(F.relu(net), sd[pre + '.fc_1.weight'], sd[pre + '.fc_1.bias'])
    key = pre + '.shortcut.weight'
    xs = F.linear(x, sd[key]) if key in sd else x
    return xs + dx


def occupancy_logits(sd, coords, hdbf):
    """MLP3D.forward (mlp.py:82-111) -> logits (B,N)."""
    xy, yz, xz = hdbf
    g = [sample_plane_feature(coords, a) for a in ('xy', 'yz', 'xz')]
    g = [t.to(coords.dtype) for t in g]
    f = [triplane_add(xy[s], yz[s], xz[s], g[0], g[1], g[2]).transpose(1, 2) for s in range(3)]
    x = F.linear(coords, sd['net_p.weight'], sd['net_p.bias']) + _resnet_fc(sd, 'net_res1', f[0])
    x = _resnet_fc(sd, 'net_res2', torch.cat((x, f[1]), dim=-1))
    x = _resnet_fc(sd, 'net_res3', torch.cat((x, f[2]), dim=-1))
    x = _resnet_fc(sd, 'net_res4', x)
    return F.linear(x, sd['net_out.weight'], sd['net_out.bias']).squeeze(-1)


def video_decode(sd, coords, hdbf, thw=None):
    """MLPVideo.forward (mlp.py:128-157) -> (b,3,t,h,w).  The reference takes (t,h,w) of the output from the finest planes
    (:135-136); `thw` overrides them so tests can decode a row band of the query volume."""
    xy, yt, xt = hdbf
    b, _, h, w = xy[-1].shape
    t = yt[-1].shape[2]
    if thw is not None:
        t, h, w = thw
    dt = xy[-1].dtype
    cg = {k: coords[k].to(dt).repeat(b, 1, 1, 1).permute(0, 2, 3, 1).contiguous() for k in ('xy', 'yt', 'xt')}
    f = [triplane_concat(xy[s], yt[s], xt[s], cg['xy'], cg['yt'], cg['xt']) for s in range(3)]
    x = _resnet_fc(sd, 'net_res1', f[0])
    x = _resnet_fc(sd, 'net_res2', torch.cat((x, f[1]), dim=1))
    x = _resnet_fc(sd, 'net_res3', torch.cat((x, f[2]), dim=1))
    x = _resnet_fc(sd, 'net_res4', x)
    x = F.linear(F.leaky_relu(x, 0.2), sd['net_out.weight'], sd['net_out.bias'])
    return x.reshape(b, -1, 3).permute(0, 2, 1).reshape(b, 3, t, h, w)


# ---------------------------------------------------------------------------
# NeRF (mlp.py:241-281; nerf_helpers.py:82-112,134-143,211-279,296-530)
# ---------------------------------------------------------------------------
def nerf_mlp(sd, x, slope=1.0, sigma_only=False, D=6, skips=(2, 4), n_xyz=159):
    """MLPNeRF.forward; LeakyReLU(True) == slope 1.0."""
    inp = x[:, :n_xyz]
    h = inp
    for i in range(D):
        if i in skips:
            h = torch.cat([inp, h], -1)
        h = F.leaky_relu(F.linear(h, sd[f'xyz_encoding_{i + 1}.0.weight'], sd[f'xyz_encoding_{i + 1}.0.bias']), slope)
    sigma = F.linear(h, sd['sigma.weight'], sd['sigma.bias'])
    if sigma_only:
        return sigma
    fin = F.linear(h, sd['xyz_encoding_final.weight'], sd['xyz_encoding_final.bias'])
    d = F.leaky_relu(F.linear(torch.cat([fin, x[:, n_xyz:]], -1), sd['dir_encoding.0.weight'], sd['dir_encoding.0.bias']), slope)
    rgb = torch.sigmoid(F.linear(d, sd['rgb.0.weight'], sd['rgb.0.bias']))
    return torch.cat([rgb, sigma], -1)


def embed(x, multires):
    """Embedder.embed (nerf_helpers.py:82-112)."""
    out = [x]
    for f in 2. ** torch.linspace(0., multires - 1, steps=multires):
        out += [torch.sin(x * f.to(x.dtype)), torch.cos(x * f.to(x.dtype))]
    return torch.cat(out, -1)


def nerf_sample_depths(rays, n_samples, perturb=0., lindisp=False):
    """z_vals of render_rays (utils/nerf_helpers.py:356-380): linear in depth or (lindisp) disparity, optionally
    stratified with one torch.rand draw per interval."""
    dt = rays.dtype
    near, far = rays[:, 6:7], rays[:, 7:8]
    t = torch.linspace(0., 1., steps=n_samples).to(device=rays.device, dtype=dt)
    z = near * (1. - t) + far * t if not lindisp else 1. / (1. / near * (1. - t) + 1. / far * t)
    z = z.expand(rays.shape[0], n_samples)
    if perturb > 0.:
        mids = .5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mids, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mids], -1)
        z = lower + (upper - lower) * torch.rand(z.shape).to(device=rays.device, dtype=dt)
    return z


def sample_pdf(bins, weights, u):
    """sample_pdf (utils/nerf_helpers.py:166-209) for given uniform numbers u (N_rays, N_samples)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], -1)
    shp = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shp), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(shp), 2, inds_g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    return bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])


def nerf_render_rays(sd, rays, fea, n_samples, white_bkgd=True, slope=1.0, return_raw=False, perturb=0., lindisp=False):
    """render_rays + run_network + raw2outputs (N_importance=0, raw_noise_std=0).
    rays (N,11) [o d near far viewdir]; fea planes (1,32,R,R)."""
    dt = rays.dtype
    o, d, near, far, vd = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8], rays[:, 8:11]
    z = nerf_sample_depths(rays, n_samples, perturb, lindisp)       # (N,S)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]             # (N,S,3)
    npts = pts / 3.5
    lat = torch.cat((_gs(fea['xy'], npts[:, :, :2][None], True),
                     _gs(fea['yz'], npts[:, :, 1:][None], True),
                     _gs(fea['xz'], npts[:, :, [0, 2]][None], True)), dim=1).squeeze(0).permute(1, 2, 0)
    n = rays.shape[0]
    x = torch.cat([lat.reshape(n * n_samples, -1), embed(pts.reshape(-1, 3), 10),
                   embed(vd[:, None].expand(n, n_samples, 3).reshape(-1, 3), 4)], -1)
    raw = nerf_mlp(sd, x, slope).reshape(n, n_samples, 4)
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)], -1) * torch.norm(d[:, None, :], dim=-1)
    alpha = 1. - torch.exp(-F.softplus(raw[..., 3]) * dists)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb = torch.sum(weights[..., None] * raw[..., :3], -2)
    if white_bkgd:
        rgb = rgb + (1. - weights.sum(-1))[..., None]
    return (rgb, raw) if return_raw else rgb


# ---------------------------------------------------------------------------
# caller epilogues of the decoded image / video signal
# ---------------------------------------------------------------------------
def store_clamp(x):
    """`fake.clamp(-1., 1.)` -- evals/eval.py:162,226; tools/ldm/image.py:246."""
    return x.clamp(-1., 1.)


def store_u8_channels_last(x):
    """`rearrange((fake.clamp(-1,1) + 1) * 127.5, 'b c t h w -> b t h w c').type(torch.uint8)` -- evals/eval.py:289,336-337
    (for images the same with 'b c h w -> b h w c')."""
    y = ((x.clamp(-1, 1) + 1) * 127.5)
    return y.movedim(1, -1).contiguous().type(torch.uint8)
