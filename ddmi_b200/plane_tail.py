"""Plane-producer tail: the layers of the reference's D2C-VAE decoders that emit the PE planes --
`up[i].hdbf[0]` (nn.Conv2d(block_in, out_ch, 1), models/d2c_vae/autoencoder_unet.py:770-771, applied :812-814) and
`norm_out -> nonlinearity -> conv_out [-> tanh]` (:788-794, applied :822-827; the video / triplane decoders apply the same
modules to each of their planes, :1111-1142, :1531-1562) -- behind the reference's parameter names, fused into the library's
kernels (ddmi_plane_head / ddmi_plane_tail) and writing every plane ONCE in the layout its consumer reads:

    tail = PlaneTail.from_decoder(vae.decoder)            # or load_state_dict of the decoder's own keys (strict=False)
    planes = [tail.head(i, h_i) for ...] + [tail.tail(h)] # NCHW for MLP / MLPVideo
    planes = tail.tail(h, channels_last=True)             # (B,C,H,W) tensor with torch.channels_last strides: MLP3D and the
                                                          # NeRF renderer gather from it directly, no transposition pass

Inference only (no autograd), CUDA tensors only."""
import torch
from torch import nn

from . import _lib


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _prep(h, name):
    if not (torch.is_tensor(h) and h.is_cuda):
        raise RuntimeError(f"{name} must be a CUDA tensor (ddmi_b200 has no CPU path)")
    if h.dim() != 4:
        raise RuntimeError(f"{name} must be (B,C,H,W), got {tuple(h.shape)}")
    if torch.is_grad_enabled() and h.requires_grad:
        raise RuntimeError("ddmi_b200.plane_tail is inference only: call it under torch.no_grad()")
    return h.detach().to(torch.float32).contiguous()


def _out(b, c, hh, ww, dev, channels_last):
    if channels_last:       # (B,H,W,C) memory seen as a (B,C,H,W) tensor: torch.channels_last
        return torch.empty((b, hh, ww, c), device=dev, dtype=torch.float32).permute(0, 3, 1, 2)
    return torch.empty((b, c, hh, ww), device=dev, dtype=torch.float32)


class PlaneTail(nn.Module):
    """Same submodule names as the reference decoders: `norm_out`, `conv_out`, `up.{i}.hdbf.0` (levels without a head keep an
    empty list), so `load_state_dict(decoder.state_dict(), strict=False)` picks exactly these tensors up."""

    def __init__(self, block_in, out_ch=64, hdbf_in_channels=(), tanh_out=False, num_groups=32):
        super().__init__()
        if out_ch not in (32, 64):
            raise NotImplementedError("plane tails emit 64 channels (32 for the srn-cars planes)")
        self.norm_out = nn.GroupNorm(num_groups=num_groups, num_channels=block_in, eps=1e-6, affine=True)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)
        self.up = nn.ModuleList()
        for cin in hdbf_in_channels:          # index = i_level, None = no head at that level
            lvl = nn.Module()
            lvl.hdbf = nn.ModuleList([nn.Conv2d(cin, out_ch, 1)] if cin else [])
            self.up.append(lvl)
        self.tanh_out = tanh_out
        self.out_ch = out_ch

    @classmethod
    def from_decoder(cls, decoder):
        """Build from a reference Decoder / VideoDecoder_light / Decoder_triplane instance (shares no storage: copies)."""
        ins = [lvl.hdbf[0].in_channels if len(lvl.hdbf) else None for lvl in decoder.up]
        t = cls(decoder.conv_out.in_channels, decoder.conv_out.out_channels, ins, bool(getattr(decoder, 'tanh_out', False)),
                decoder.norm_out.num_groups)
        t.load_state_dict({k: v for k, v in decoder.state_dict().items() if k in t.state_dict()}, strict=True)
        return t.to(decoder.conv_out.weight.device)

    def head(self, i_level, h, channels_last=False):
        """`self.up[i_level].hdbf[0](h)` (autoencoder_unet.py:812-814)."""
        conv = self.up[i_level].hdbf[0]
        x = _prep(h, 'h')
        b, c, hh, ww = x.shape
        if c != conv.in_channels:
            raise RuntimeError(f"h has {c} channels, the level-{i_level} head takes {conv.in_channels}")
        out = _out(b, self.out_ch, hh, ww, x.device, channels_last)
        w = conv.weight.detach().to(torch.float32).reshape(self.out_ch, c).t().contiguous()       # (C_in, 1, C_out)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ddmi_plane_head(x.data_ptr(), b, c, hh, ww, w.data_ptr(), conv.bias.detach().float().data_ptr(),
                                                  self.out_ch, 1 if channels_last else 0, out.data_ptr(), _stream_ptr(x.device)))
        return out

    def tail(self, h, channels_last=False):
        """`conv_out(nonlinearity(norm_out(h)))` [+ tanh] (autoencoder_unet.py:822-827)."""
        x = _prep(h, 'h')
        b, c, hh, ww = x.shape
        if c != self.conv_out.in_channels:
            raise RuntimeError(f"h has {c} channels, conv_out takes {self.conv_out.in_channels}")
        out = _out(b, self.out_ch, hh, ww, x.device, channels_last)
        g = self.norm_out.num_groups
        stats = torch.empty(b * g * 2, device=x.device, dtype=torch.float32)
        f = lambda p: p.detach().to(torch.float32).contiguous()
        gw, gb, bias = f(self.norm_out.weight), f(self.norm_out.bias), f(self.conv_out.bias)
        w = self.conv_out.weight.detach().to(torch.float32).permute(1, 2, 3, 0).contiguous()       # (C_in, 3 * 3, C_out)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ddmi_plane_tail(x.data_ptr(), b, c, hh, ww, gw.data_ptr(), gb.data_ptr(), g, float(self.norm_out.eps),
                                                  w.data_ptr(), bias.data_ptr(), self.out_ch, 1 if self.tanh_out else 0,
                                                  1 if channels_last else 0, stats.data_ptr(), out.data_ptr(), _stream_ptr(x.device)))
        return out
