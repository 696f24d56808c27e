"""Golden vectors of the plane-producer tail: a small instance of the REFERENCE's Decoder
(/root/reference/models/d2c_vae/autoencoder_unet.py:702-832) run on a seeded latent, with forward hooks capturing the inputs and
outputs of its `up[i].hdbf[0]` heads and of `norm_out` / `conv_out`.  Build container only:  python oracle/make_golden_plane_tail.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
import models.d2c_vae.autoencoder_unet as au  # noqa: E402
from oracle import plane_tail_oracle as po  # noqa: E402

torch.manual_seed(777)
d = {}
for tag, tanh_out in (('plain', False), ('tanh', True)):
    dec = au.Decoder(ch=32, out_ch=64, ch_mult=(1, 2, 4), num_res_blocks=1, attn_resolutions=[], in_channels=3, resolution=32,
                     z_channels=4, hdbf_resolutions=[16, 8], attn_type='none', tanh_out=tanh_out).eval()
    for p in dec.parameters():          # default init leaves the tail nearly linear: spread the statistics
        p.data.mul_(1.5)
    cap = {}
    hooks = [dec.norm_out.register_forward_hook(lambda m, i, o: cap.__setitem__('tail_in', i[0].detach().clone()))]
    for lvl, up in enumerate(dec.up):
        if len(up.hdbf):
            hooks.append(up.hdbf[0].register_forward_hook(
                lambda m, i, o, lvl=lvl: cap.__setitem__(f'head{lvl}', (i[0].detach().clone(), o.detach().clone()))))
    with torch.no_grad():
        planes = dec(torch.randn(1 if tanh_out else 2, 4, 8, 8))
    for h in hooks:
        h.remove()
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()
          if k.startswith('norm_out') or k.startswith('conv_out') or '.hdbf.' in k}
    d[tag] = {'sd': sd, 'tail_in': cap['tail_in'], 'tail_out': planes[-1].detach().clone(), 'tanh_out': tanh_out,
              'heads': {k: v for k, v in cap.items() if k.startswith('head')}}
    # the restatement against the reference's own forward
    e = float((po.tail(sd, cap['tail_in'], 32, tanh_out) - planes[-1]).abs().max())
    for k, (hin, hout) in d[tag]['heads'].items():
        e = max(e, float((po.head(sd, int(k[4:]), hin) - hout).abs().max()))
    print(tag, 'oracle vs reference', e, [tuple(p.shape) for p in planes])
torch.save(d, os.path.join(ROOT, 'tests', 'golden', 'plane_tail.pt'))
print('wrote', os.path.getsize(os.path.join(ROOT, 'tests', 'golden', 'plane_tail.pt')), 'bytes')
