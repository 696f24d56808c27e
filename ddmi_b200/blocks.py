"""Parameter containers of the INR decoders.

The decode math of these blocks lives in the CUDA kernels (``csrc/``) and in the
host-side weight folding (``packing.py``); the classes here only own the
parameters, under exactly the names / shapes / default initialisations of the
reference's ``models/d2c_vae/blocks.py`` so that ``load_state_dict`` of a
reference checkpoint works (SURVEY.md §5 "Checkpoint / resume", §8a inventory).
None of them has a ``forward``: calling one is a bug, the owning decoder in
``mlp.py`` launches one fused kernel instead.
"""
import math

import torch
from torch import nn


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container; the decode is fused "
            "into the owning ddmi_b200 decoder (there is no eager fallback)")


class SinusoidalPosEmb(_ParamsOnly):
    """Sin/cos embedding of the scale-injection scalar (blocks.py:11-23)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class EqualLinear(_ParamsOnly):
    """weight ~ N(0,1)/lr_mul, runtime scale 1/sqrt(in)*lr_mul (blocks.py:139-173)."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul


class ModulatedConv2d(_ParamsOnly):
    """1x1 modulated (optionally demodulated) conv (blocks.py:187-283)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True):
        super().__init__()
        if kernel_size != 1:
            raise ValueError("the INR decode path only uses 1x1 modulated convs")
        self.eps = 1e-8
        self.in_channel, self.out_channel, self.kernel_size = in_channel, out_channel, kernel_size
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate


class NoiseInjection(_ParamsOnly):
    """Per-pixel noise scale, zero-initialised (blocks.py:286-297)."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))


class FusedLeakyReLU(_ParamsOnly):
    """Bias of ``lrelu(x + b, 0.2) * sqrt(2)`` (op/fused_act.py:75-88); the op
    itself is an MMA epilogue in the kernels."""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope, self.scale = negative_slope, scale


class StyledConv(_ParamsOnly):
    """conv -> noise -> fused bias+lrelu (blocks.py:312-356)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)


class EqualConv2d(_ParamsOnly):
    """Plain conv with runtime 1/sqrt(fan_in) scale (blocks.py:102-136)."""

    def __init__(self, in_channel, out_channel, kernel_size, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None


class ConvLayer(nn.Sequential):
    """The 1x1, bias-free, activation-free skip of StyledResBlock
    (blocks.py:453-534 with ``activate=False, bias=False``): key ``0.weight``."""

    def __init__(self, in_channel, out_channel, kernel_size):
        super().__init__(EqualConv2d(in_channel, out_channel, kernel_size, bias=False))


class StyledResBlock(_ParamsOnly):
    """(conv3(conv2(conv1(x))) + skip(x)) / sqrt(2) (blocks.py:604-638)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True):
        super().__init__()
        self.conv1 = StyledConv(in_channel, out_channel, 1, style_dim, demodulate=demodulate)
        self.conv2 = StyledConv(out_channel, out_channel, kernel_size, style_dim, demodulate=demodulate)
        self.conv3 = StyledConv(out_channel, out_channel, 1, style_dim, demodulate=demodulate)
        self.skip = ConvLayer(in_channel, out_channel, 1) if in_channel != out_channel else None


class ToRGB(_ParamsOnly):
    """Modulated (not demodulated) 1x1 conv + bias (blocks.py:390-412)."""

    def __init__(self, in_channel, out_channel, style_dim):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, out_channel, 1, 1))


class ResnetBlockFC(_ParamsOnly):
    """shortcut(x) + fc_1(relu(fc_0(relu(x)))), fc_1.weight zero-init
    (blocks.py:673-716)."""

    def __init__(self, size_in, size_out=None, size_h=None):
        super().__init__()
        size_out = size_in if size_out is None else size_out
        size_h = min(size_in, size_out) if size_h is None else size_h
        self.size_in, self.size_h, self.size_out = size_in, size_h, size_out
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        self.shortcut = None if size_in == size_out else nn.Linear(size_in, size_out, bias=False)
        nn.init.zeros_(self.fc_1.weight)
