"""ORACLE (test infrastructure): fp32 PyTorch restatement ("port") of the reference Generator / Discriminator forward,
driven by a reference-format state_dict.  It exists because /root/reference cannot travel to the GPU box; it is pinned
against the live (shimmed) reference in tests/test_oracle_cpu.py and against committed golden vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this file.

Follows (file:line in /root/reference):
  conv()/deconv()/ResidualBlock        ModificationLayer.py:54-123, 158-202, 233-302
  LocalPathway                         D_and_G_model.py:18-110
  LocalFuser                           D_and_G_model.py:112-159
  GlobalPathway                        D_and_G_model.py:161-329   (with fix F3: dim128 includes I128's 3 channels)
  FeaturePredict / Generator           D_and_G_model.py:331-407
  Discriminator                        D_and_G_model.py:409-435
Config defaults: use_batchnorm=False for G and D (config.py:63,68) - the only configuration restated.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LEAKY = 0.01  # nn.LeakyReLU(1e-2) and nn.LeakyReLU() both have negative_slope 0.01

# ---- optional TF32 operand emulation ---------------------------------------------------------------------------------
# The CUDA path multiplies TF32 operands (10-bit mantissa, round-to-nearest-away = PTX cvt.rna.tf32.f32) and accumulates
# in fp32.  With EMULATE_TF32 = True the port rounds exactly the same operands (weights, and every conv output where the
# kernels round at store), so the only remaining difference to the device is fp32 summation order.  This separates
# "arithmetic format" deviations from logic errors: LeakyReLU sign flips caused by TF32 noise otherwise dominate any
# gradient comparison against pure fp32.  Default False = the reference's fp32 arithmetic.
EMULATE_TF32 = False


def tf32_rna(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32: keep 10 mantissa bits, round to nearest, ties away from zero."""
    i = x.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF)
    r = torch.where((i & 0x7F800000) == 0x7F800000, i, r)  # inf / nan unchanged
    return r.view(torch.float32)


class _RoundST(torch.autograd.Function):
    """tf32 rounding in forward, and of the gradient in backward (the kernels round stored activations AND stored
    activation gradients)."""

    @staticmethod
    def forward(ctx, x):
        return tf32_rna(x)

    @staticmethod
    def backward(ctx, g):
        return tf32_rna(g)


def _q(x):
    return _RoundST.apply(x) if EMULATE_TF32 else x


class _RoundW(torch.autograd.Function):
    """tf32 rounding of a weight in forward; its gradient passes through unchanged."""

    @staticmethod
    def forward(ctx, w):
        return tf32_rna(w)

    @staticmethod
    def backward(ctx, g):
        return g


def _qw(w):
    return _RoundW.apply(w) if EMULATE_TF32 else w


# ---- optional bf16 operand emulation ---------------------------------------------------------------------------------
# The bf16 path (BASELINE configs[2]) keeps fp32 activations / gradients and hands the tensor cores bf16 COPIES of them: a
# convolution reads bf16(x) and bf16(w), accumulates in fp32, and its backward reads bf16(dY) (dY = the masked total gradient
# of its output) for both the input and the weight gradient; residual adds, gradient accumulation, masks, losses stay fp32.
# With EMULATE_BF16 = True the port rounds exactly those operands, so that the bf16 path can be checked at the precision of
# fp32 summation order instead of the format's 1e-2 (a stale or missing bf16 copy would hide inside 1e-2).  Rounding is
# written with differentiable ops (value + detached correction) and a gradient hook, so it composes with autograd.grad(...,
# create_graph=True); the gradient-penalty double backward is NOT claimed to be emulated exactly (the CUDA path evaluates it
# as a tangent forward with its own rounding points).
EMULATE_BF16 = False


def _bf(x):
    return x + (x.bfloat16().float() - x).detach()


def _xin(x):
    """A tensor-core operand (activation or weight)."""
    return _bf(x) if EMULATE_BF16 else x


def _gout(y):
    """A conv / linear pre-activation output: its incoming gradient is read through a bf16 copy by dgrad and wgrad."""
    if EMULATE_BF16 and y.requires_grad:
        y.register_hook(_bf)
    return y


# ---- optional activation-mask injection ---------------------------------------------------------------------------------
# (Leaky)ReLU makes gradients discontinuous in the forward values: an element whose pre-activation changes sign between
# two implementations changes its gradient by ~100 %, so a forward deviation eps turns into a gradient deviation of order
# sqrt(eps) that has nothing to do with the backward arithmetic being checked.  MASK_HOOK(name) may return the stored
# output of the CUDA path for the activation site `name`; the port then evaluates the activation's *backward* with the
# sign pattern of that tensor (forward values stay the port's own).  Gradient parity tests use it to verify dgrad / wgrad
# / plumbing at the precision of the arithmetic format.  Default None = plain autograd.
MASK_HOOK = None


class _ActWithMask(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, mask_src, slope):
        ctx.save_for_backward(mask_src > 0)
        ctx.slope = slope
        return F.leaky_relu(y, slope)

    @staticmethod
    def backward(ctx, g):
        (pos,) = ctx.saved_tensors
        return torch.where(pos, g, g * ctx.slope), None, None


def _act(y, slope, name):
    if slope is None:
        return y
    if MASK_HOOK is not None:
        m = MASK_HOOK(name)
        if m is not None:
            return _ActWithMask.apply(y, m.reshape(y.shape), slope)
    return F.leaky_relu(y, slope)


# The two selection sites of the generator - the maxout after fc1 and the LocalFuser's max over the four padded patches -
# are discontinuous in the same way: a near-tie that resolves differently routes a whole gradient element elsewhere.  With
# MASK_HOOK installed, MASK_HOOK("<fc1 key>#maxout") may return the CUDA path's stored fc1 output and
# MASK_HOOK("local_fuser#<site>") its stored arg-max map; the port's *backward* then follows those selections (forward
# values stay the port's own), exactly like the activation masks above.
class _MaxoutWithSel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pairs, sel_src):          # (B, F, 2)
        ctx.save_for_backward(sel_src[..., 0] >= sel_src[..., 1])   # first of an equal pair wins (maxout2_backward_kernel)
        return pairs.max(dim=2).values

    @staticmethod
    def backward(ctx, g):
        (first,) = ctx.saved_tensors
        z = torch.zeros_like(g)
        return torch.stack([torch.where(first, g, z), torch.where(first, z, g)], dim=2), None


class _FuseWithSel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stacked, idx):            # (4, B, C, 128, 128), (B, C, 128, 128) int64
        ctx.save_for_backward(idx)
        return stacked.max(dim=0).values

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        out = torch.zeros((4,) + tuple(g.shape), dtype=g.dtype)
        return out.scatter_(0, idx.unsqueeze(0), g.unsqueeze(0)), None


def _conv(sd, key, x, stride=1, pad=0, slope=LEAKY):
    """conv(): [ReflectionPad2d(4-list)] -> Conv2d(bias) -> activation (ModificationLayer.py:83-119)."""
    if isinstance(pad, (list, tuple)):
        x = F.pad(x, tuple(pad), mode="reflect")  # ReflectionPad2d((left, right, top, bottom)), :93
        idx, p = 1, 0
    else:
        idx, p = 0, pad
    y = _gout(F.conv2d(_xin(x), _xin(_qw(sd[f"{key}.{idx}.weight"])), sd[f"{key}.{idx}.bias"], stride=stride, padding=p))
    return _q(_act(y, slope, key))


def _deconv(sd, key, x, stride, pad, out_pad):
    """deconv(): ConvTranspose2d(bias) -> ReLU (ModificationLayer.py:189-198)."""
    return _q(_act(_gout(F.conv_transpose2d(_xin(x), _xin(_qw(sd[f"{key}.0.weight"])), sd[f"{key}.0.bias"], stride=stride,
                                            padding=pad, output_padding=out_pad)), 0.0, key))


def _res(sd, key, x, k=3, pad=None):
    """Non-bottleneck ResidualBlock with identity shortcut, scaling_factor 1.0 (ModificationLayer.py:292-302)."""
    pad = (k - 1) // 2 if pad is None else pad
    h = _conv(sd, f"{key}.layers.0", x, 1, pad, LEAKY)
    if isinstance(pad, (list, tuple)):
        hp, idx, p = F.pad(h, tuple(pad), mode="reflect"), 1, 0
    else:
        hp, idx, p = h, 0, pad
    y = _gout(F.conv2d(_xin(hp), _xin(_qw(sd[f"{key}.layers.1.{idx}.weight"])), sd[f"{key}.layers.1.{idx}.bias"], padding=p))
    return _q(_act(y + 1.0 * x, LEAKY, f"{key}.layers.1"))


def _conv_res(sd, key, x, k, stride, pad):
    """sequential(conv(...), ResidualBlock(...)) as used by every encoder stage."""
    return _res(sd, f"{key}.1", _conv(sd, f"{key}.0", x, stride, pad), k if key.endswith(("conv0", "conv1")) else 3)


def local_pathway(sd, pre, x):
    """D_and_G_model.py:84-110.  Returns (local_img, deconv2)."""
    c = lambda name, t, s: _res(sd, f"{pre}.{name}.1", _conv(sd, f"{pre}.{name}.0", t, s, 1), 3)
    conv0 = c("conv0", x, 1)
    conv1 = c("conv1", conv0, 2)
    conv2 = c("conv2", conv1, 2)
    conv3 = c("conv3", conv2, 2)
    deconv0 = _deconv(sd, f"{pre}.deconv0", conv3, 2, 1, 1)
    as0 = c("after_select0", torch.cat([deconv0, conv2], 1), 1)
    deconv1 = _deconv(sd, f"{pre}.deconv1", as0, 2, 1, 1)
    as1 = c("after_select1", torch.cat([deconv1, conv1], 1), 1)
    deconv2 = _deconv(sd, f"{pre}.deconv2", as1, 2, 1, 1)
    as2 = c("after_select2", torch.cat([deconv2, conv0], 1), 1)
    local_img = _conv(sd, f"{pre}.local_img", as2, 1, 0, None)
    return local_img, deconv2


# (left, top, width, height) of the fixed paste rectangles, D_and_G_model.py:148-157
FUSE_RECTS = ((39 - 20 - 1, 40 - 20 - 1, 40, 40), (86 - 20 - 1, 39 - 20 - 1, 40, 40), (64 - 20 - 1, 64 - 16 - 1, 40, 32),
              (65 - 24 - 1, 89 - 16 - 1, 48, 32))


def local_fuser(parts, return_index=False, site=None):
    """max over the four zero-padded maps (D_and_G_model.py:154-159).  site: name of the call for selection injection."""
    padded = []
    for t, (left, top, w, h) in zip(parts, FUSE_RECTS):
        assert t.shape[2] == h and t.shape[3] == w
        padded.append(F.pad(t, (left, 128 - left - w, top, 128 - top - h)))
    if MASK_HOOK is not None and site is not None and not return_index:
        sel = MASK_HOOK("local_fuser#" + site)
        if sel is not None:
            return _FuseWithSel.apply(torch.stack(padded, 0), sel)
    val, idx = torch.max(torch.stack(padded, 0), 0)
    return (val, idx) if return_index else val


def global_pathway(sd, pre, I128, local_fake, local_feat, z):
    """D_and_G_model.py:281-329."""
    p = lambda n: f"{pre}.{n}"
    conv0 = _res(sd, p("conv0.1"), _conv(sd, p("conv0.0"), I128, 1, 3), 7)
    conv1 = _res(sd, p("conv1.1"), _conv(sd, p("conv1.0"), conv0, 2, 2), 5)
    conv2 = _res(sd, p("conv2.1"), _conv(sd, p("conv2.0"), conv1, 2, 1), 3)
    conv3 = _res(sd, p("conv3.1"), _conv(sd, p("conv3.0"), conv2, 2, 1), 3)
    conv4 = _conv(sd, p("conv4.0"), conv3, 2, 1)
    for i in range(1, 5):
        conv4 = _res(sd, p(f"conv4.{i}"), conv4, 3)
    B = I128.shape[0]
    fc1 = _q(_gout(F.linear(_xin(conv4.reshape(B, -1)), _xin(_qw(sd[p("fc1.weight")])), sd[p("fc1.bias")])))
    sel = MASK_HOOK(p("fc1") + "#maxout") if MASK_HOOK is not None else None
    if sel is not None:
        fc2 = _MaxoutWithSel.apply(fc1.view(B, -1, 2), sel.reshape(B, -1, 2))
    else:
        fc2 = F.max_pool1d(fc1.view(B, -1, 2), 2, 2).view(B, -1)
    deconv_8 = _deconv(sd, p("deconv_8"), torch.cat([fc2, z], 1).view(B, -1, 1, 1), 1, 0, 0)
    deconv_32 = _deconv(sd, p("deconv_32"), deconv_8, 4, 0, 1)
    deconv_64 = _deconv(sd, p("deconv_64"), deconv_32, 2, 1, 1)
    deconv_128 = _deconv(sd, p("deconv_128"), deconv_64, 2, 1, 1)
    rp = [1, 0, 1, 0]
    f8 = _res(sd, p("add_conv_and_deconv_8"), torch.cat([deconv_8, conv4], 1), 2, rp)
    f8 = _res(sd, p("enhance_features_8.1"), _res(sd, p("enhance_features_8.0"), f8, 2, rp), 2, rp)
    up16 = _deconv(sd, p("upsample_16"), f8, 2, 1, 1)
    a16 = _res(sd, p("add_conv_and_deconv_16"), conv3, 3)
    f16 = torch.cat([up16, a16], 1)
    f16 = _res(sd, p("enhance_features_16.1"), _res(sd, p("enhance_features_16.0"), f16, 3), 3)
    up32 = _deconv(sd, p("upsample_32"), f16, 2, 1, 1)
    a32 = _res(sd, p("add_conv_and_deconv_32"), torch.cat([deconv_32, conv2], 1), 3)
    f32 = torch.cat([up32, a32], 1)
    f32 = _res(sd, p("enhance_features_32.1"), _res(sd, p("enhance_features_32.0"), f32, 3), 3)
    up64 = _deconv(sd, p("upsample_64"), f32, 2, 1, 1)
    a64 = _res(sd, p("add_conv_and_deconv_64"), torch.cat([deconv_64, conv1], 1), 5)
    f64 = torch.cat([up64, a64], 1)
    f64 = _res(sd, p("enhance_features_64.1"), _res(sd, p("enhance_features_64.0"), f64, 3), 3)
    up128 = _deconv(sd, p("upsample_128"), f64, 2, 1, 1)
    a128 = _res(sd, p("add_conv_and_deconv_128"), torch.cat([deconv_128, conv0, I128], 1), 7)
    f128 = _res(sd, p("enhance_features_128.0"), torch.cat([up128, a128, local_feat, local_fake], 1), 5)
    conv5 = _res(sd, p("conv5.1"), _conv(sd, p("conv5.0"), f128, 1, 2), 3)
    conv6 = _conv(sd, p("conv6"), conv5, 1, 1)
    img = _conv(sd, p("decoded_img128"), conv6, 1, 1, None)
    return img, fc2


def generator(sd, I128, left_eye, right_eye, nose, mouth, z, dropout_mask=None):
    """Generator.forward (D_and_G_model.py:374-407); returns the same 8-tuple.  dropout_mask (B,256), already scaled by
    1/(1-p), replaces nn.Dropout(0.3) when use_dropout is wanted; None = use_dropout False."""
    if EMULATE_TF32:  # the kernels round the staged inputs too
        I128, left_eye, right_eye, nose, mouth, z = (tf32_rna(t) for t in (I128, left_eye, right_eye, nose, mouth, z))
    parts_in = (left_eye, right_eye, nose, mouth)
    names = ("left_eye", "right_eye", "nose", "mouth")
    imgs, feats = [], []
    for n, t in zip(names, parts_in):
        im, ft = local_pathway(sd, f"local_pathway_{n}", t)
        imgs.append(im)
        feats.append(ft)
    fused_feat = local_fuser(feats, site="feature")
    fused_img = local_fuser(imgs, site="fake_image")
    fused_in = local_fuser(parts_in, site="origin")
    fake, enc = global_pathway(sd, "global_pathway", I128, fused_img, fused_feat, z)
    e = enc if dropout_mask is None else enc * dropout_mask
    logits = _q(_gout(F.linear(_xin(e), _xin(_qw(sd["feature_predict.fc.weight"])), sd["feature_predict.fc.bias"])))
    return fake, logits, fused_img, imgs[0], imgs[1], imgs[2], imgs[3], fused_in


def discriminator(sd, x):
    """Discriminator.forward (D_and_G_model.py:421-435): model.{0..3}=conv s2, .4=RB, .5=conv s2, .6=RB, .7=conv."""
    if EMULATE_TF32:
        x = _q(x)
    for i in range(4):
        x = _conv(sd, f"model.{i}", x, 2, 1)
    x = _res(sd, "model.4", x, 3)
    x = _conv(sd, "model.5", x, 2, 1)
    x = _res(sd, "model.6", x, 3)
    return _conv(sd, "model.7", x, 1, 1, None)
