"""CPU oracle of the plane-producer tail.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's baseline legs).
Restates models/d2c_vae/autoencoder_unet.py:812-814 (hdbf head) and :822-827 (norm_out -> x * sigmoid(x) -> conv_out [-> tanh])
with the torch ops the reference's modules execute (nn.Conv2d, GroupNorm(32, eps=1e-6): the reference's arithmetic by execution).
Pin: tests/golden/plane_tail.pt = inputs and outputs of these layers captured with forward hooks inside the REFERENCE's own
Decoder.forward (oracle/make_golden_plane_tail.py)."""
import torch
import torch.nn.functional as F


def head(sd, i_level, h):
    return F.conv2d(h, sd[f'up.{i_level}.hdbf.0.weight'], sd[f'up.{i_level}.hdbf.0.bias'])


def tail(sd, h, num_groups=32, tanh_out=False):
    x = F.group_norm(h, num_groups, sd['norm_out.weight'], sd['norm_out.bias'], eps=1e-6)
    x = x * torch.sigmoid(x)
    x = F.conv2d(x, sd['conv_out.weight'], sd['conv_out.bias'], stride=1, padding=1)
    return torch.tanh(x) if tanh_out else x
