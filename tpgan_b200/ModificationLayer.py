"""Drop-in for the reference's layer factories (ModificationLayer.py:5-302 in PandaKenWei/TP-GAN).

Same public names, argument meaning, `.out_channels` propagation, parameter structure (state_dict keys/shapes) and
random-number consumption order as the reference, so `torch.manual_seed(s); Generator(...)` yields bit-identical
weights and checkpoints load either way.  The modules built here are *parameter containers*: the Generator and
Discriminator of tpgan_b200.D_and_G_model trace them into the CUDA engine (tcgen05 kernels); nothing on the training
path calls their ATen forward.  Called stand-alone, a conv()/deconv() stack still runs through the same C-ABI kernels
via TCConv2d/TCConvTranspose2d.

Reference defects handled the way SURVEY.md 2.2 documents (F1: init receives the module; F2: activation=None in
Sequential; F4: isinstance(activation(), ...)): the behaviour the reference intends is implemented directly.
"""
from __future__ import annotations

import torch
import torch.nn as nn

__all__ = ["sequential", "weight_initialization", "conv", "deconv", "linear", "ResidualBlock", "TCConv2d",
           "TCConvTranspose2d"]


def _negative_slope(activation):
    """Slope of the fused epilogue for an activation module: LeakyReLU -> its slope, ReLU -> 0, None -> None."""
    if activation is None:
        return None
    if isinstance(activation, nn.LeakyReLU):
        return float(activation.negative_slope)
    if isinstance(activation, nn.ReLU):
        return 0.0
    raise NotImplementedError(f"activation {type(activation).__name__} has no fused tensor-core epilogue")


class TCConv2d(nn.Conv2d):
    """nn.Conv2d parameters; stand-alone forward runs the tcgen05 implicit-GEMM kernel (no ATen convolution)."""

    def forward(self, x):
        from .functional import conv2d_standalone
        return conv2d_standalone(self, x, transposed=False)


class TCConvTranspose2d(nn.ConvTranspose2d):
    def forward(self, x, output_size=None):
        from .functional import conv2d_standalone
        return conv2d_standalone(self, x, transposed=True)


def sequential(*mods):
    """nn.Sequential that inherits `.out_channels` from its last member that has one (reference :5-24)."""
    seq = nn.Sequential(*mods)
    for m in reversed(mods):
        width = getattr(m, "out_channels", None)
        if width is None:
            width = getattr(m, "out_features", None)
        if width is not None:
            seq.out_channels = width
            break
    return seq


def weight_initialization(weight, init, activation):
    """'kaiming' (a = the activation's negative slope, 0 for ReLU/None) or 'xavier' normal init (reference :26-52)."""
    if init is None:
        return
    w = weight.weight if isinstance(weight, nn.Module) else weight
    if init == "kaiming":
        nn.init.kaiming_normal_(w, a=getattr(activation, "negative_slope", 0))
    elif init == "xavier":
        nn.init.xavier_normal_(w)


def _norm_and_act(channels, activation, use_batchnorm):
    mods = []
    if use_batchnorm:
        bn = nn.BatchNorm2d(channels)
        mods = [activation, bn] if isinstance(activation, (nn.Sigmoid, nn.Tanh)) else [bn, activation]
    else:
        mods = [activation]
    return [m for m in mods if m is not None]


def conv(in_channels, out_channels, kernel_size, stride=1, padding=0, init="kaiming", activation=nn.ReLU(),
         use_batchnorm=False, pre_activation=False):
    """[ReflectionPad2d] -> Conv2d(bias = not BN) -> [BN] -> [activation]  (reference :54-123)."""
    mods = []
    if isinstance(padding, list):
        assert len(padding) != 3
        if len(padding) == 4:
            mods.append(nn.ReflectionPad2d(padding))
            padding = 0
    layer = TCConv2d(in_channels, out_channels, kernel_size, stride, padding, bias=not use_batchnorm)
    weight_initialization(layer, init, activation)
    mods.append(layer)
    extra = _norm_and_act(in_channels if pre_activation else out_channels, activation, use_batchnorm)
    mods = extra + mods if pre_activation else mods + extra
    seq = nn.Sequential(*mods)
    seq.out_channels = out_channels
    return seq


def deconv(in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, init="kaiming",
           activation=nn.ReLU(), use_batchnorm=False, pre_activation=False):
    """ConvTranspose2d(bias = not BN) -> [BN] -> [activation]  (reference :158-202)."""
    layer = TCConvTranspose2d(in_channels, out_channels, kernel_size, stride, padding, output_padding,
                              bias=not use_batchnorm)
    weight_initialization(layer, init, activation)
    extra = _norm_and_act(in_channels if pre_activation else out_channels, activation, use_batchnorm)
    mods = extra + [layer] if pre_activation else [layer] + extra
    seq = nn.Sequential(*mods)
    seq.out_channels = out_channels
    return seq


def linear(in_channels, out_channels, activation=None, use_batchnorm=False):
    """Linear(bias = not BN) -> [BatchNorm1d] -> [activation]  (reference :204-231)."""
    mods = [nn.Linear(in_channels, out_channels, bias=not use_batchnorm)]
    if use_batchnorm:
        mods.append(nn.BatchNorm1d(out_channels))
    if activation is not None:
        mods.append(activation)
    return nn.Sequential(*mods)


class ResidualBlock(nn.Module):
    """act( conv_k(act(conv_k(x))) + scaling_factor * shortcut(x) )  (reference :233-302).

    As in the reference, the shortcut is built from the *argument* `use_projection` (so it is the identity at every call
    site of the model), the second conv keeps PyTorch's default init and has no activation, and `padding` may be the
    4-list of a ReflectionPad2d."""

    def __init__(self, in_channels, out_channels=None, kernel_size=3, stride=1, padding=None, weight_init="kaiming",
                 activation=nn.ReLU(), is_bottleneck=False, use_projection=False, scaling_factor=1.0,
                 is_inplace_of_activation=False, use_batchnorm=False):
        super().__init__()
        self.out_channels = in_channels // stride if out_channels is None else out_channels
        if padding is None:
            padding = 1 if is_inplace_of_activation else (kernel_size - 1) // 2
        self.padding = padding
        self.kernel_size = kernel_size
        if is_inplace_of_activation and isinstance(activation, nn.ReLU):
            activation = nn.ReLU(inplace=True)
        self.activation = activation
        self.scaling_factor = scaling_factor
        self.use_projection = use_projection or (stride != 1 or in_channels != out_channels)
        self.shortcut = conv(in_channels, out_channels, 1, stride, 0, weight_init, None, False) if use_projection \
            else nn.Sequential()
        if is_bottleneck:
            mid_in, mid_out = in_channels // 2, self.out_channels // 2
            stack = [conv(in_channels, mid_in, 1, 1, 0, weight_init, activation, use_batchnorm, False),
                     conv(mid_in, mid_out, kernel_size, stride, (kernel_size - 1) // 2, weight_init, activation,
                          use_batchnorm, False),
                     conv(mid_out, self.out_channels, 1, 1, 0, None, None, use_batchnorm, False)]
        else:
            stack = [conv(in_channels, in_channels, kernel_size, 1, padding, weight_init, activation, use_batchnorm, False),
                     conv(in_channels, self.out_channels, kernel_size, 1, padding, None, None, use_batchnorm, False)]
        self.layers = nn.Sequential(*stack)

    def forward(self, x):
        y = self.layers(x) + self.scaling_factor * self.shortcut(x)
        return y if self.activation is None else self.activation(y)
