"""Drop-in for the reference's model file (D_and_G_model.py:18-435 in PandaKenWei/TP-GAN).

Same classes, constructor signatures, sub-module attribute names, forward signatures / return tuples and state_dict
keys+shapes as the reference, so a training loop written against the reference's `Generator` / `Discriminator` runs
unchanged and checkpoints load either way.  What differs is how forward/backward execute: every module here *traces*
itself once per input geometry into a static launch plan (tpgan_b200.engine.Plan) over the C-ABI kernels of
libtpgan_b200.so - tcgen05 implicit-GEMM convolutions with fused bias/residual/(Leaky)ReLU epilogues, hand-written
dgrad/wgrad, grouped launches for the four local pathways, a fused LocalFuser stitch - and replays that plan.
torch.cat never executes: producers write into channel slices of pre-allocated NHWC concat buffers.  There is no ATen
convolution anywhere on the path and no CPU fallback: tensors must be CUDA tensors and the extension must be built.

Reference defects are resolved the way `forward()` behaves: F3 - `dim128` includes the 3 channels of
I128 that forward concatenates (D_and_G_model.py:268 vs :323), so enhance_features_128 has 206 channels.

Gradients: module outputs are connected to autograd through one torch.autograd.Function per call, whose backward replays
the traced backward plan and accumulates into the `.grad` of the reference-layout parameters.  Double backward (WGAN-GP
through autograd.grad(create_graph=True)) is not available through this generic interface; the fused training step
(tpgan_b200.train_step) implements the gradient penalty natively.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from .engine import LINEAR, ConvLayer, DeconvAsLinear, Plan, T
from .ModificationLayer import *  # noqa: F401,F403  (the reference does the same, D_and_G_model.py:15)
from .ModificationLayer import ResidualBlock, conv, deconv, sequential, _negative_slope
from .ops import Act
from .UtilityMethods import elementwise_multiply_and_cast_to_int as EMaC2I

EXACT_MODE = None   # tests only: force (True / False) the fp32-exact 3xTF32-split mode of newly traced plans; None = the module's own default

PATCH_HW = ((40, 40), (40, 40), (32, 40), (32, 48))  # (h, w) of left eye, right eye, nose, mouth
PART_NAMES = ("left_eye", "right_eye", "nose", "mouth")


# ---------------------------------------------------------------------------------------------------- tracing helpers
def _unpack_conv_seq(seq: nn.Sequential):
    """conv()/deconv() factory output -> (reflection pad or None, conv module, fused activation slope or None, BatchNorm2d
    or None).  With use_batchnorm=True the factories emit [pad] -> conv(bias=False) -> BatchNorm2d -> act
    (ModificationLayer.py:125-156)."""
    pad_mod, conv_mod, slope, bn_mod = None, None, None, None
    for m in seq:
        if isinstance(m, nn.ReflectionPad2d):
            pad_mod = m
        elif isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            conv_mod = m
        elif isinstance(m, nn.BatchNorm2d):
            assert conv_mod is not None, "pre_activation stacks are not used by the G/D models"
            bn_mod = m
        else:
            assert conv_mod is not None, "pre_activation stacks are not used by the G/D models"
            slope = _negative_slope(m)
    return pad_mod, conv_mod, slope, bn_mod


def _layer(conv_mod, name: str) -> ConvLayer:
    L = getattr(conv_mod, "_tc_layer", None)
    if L is None:
        tr = isinstance(conv_mod, nn.ConvTranspose2d)
        k, s, p = conv_mod.kernel_size[0], conv_mod.stride[0], conv_mod.padding[0]
        assert conv_mod.kernel_size[0] == conv_mod.kernel_size[1] and conv_mod.dilation == (1, 1) and conv_mod.groups == 1
        if tr:
            assert conv_mod.output_padding[0] == (1 if s > 1 else 0), "only the out = stride*in transposed convs are used"
        L = ConvLayer(conv_mod.weight, conv_mod.bias, tr, k, s, p, name)
        object.__setattr__(conv_mod, "_tc_layer", L)
    return L


def t_conv(plan: Plan, seqs: Sequence[nn.Sequential], xs: Sequence[T], outs=None, residuals=None, names=None,
           slope_override="same") -> List[T]:
    """Grouped conv()/deconv() stack: [reflect pad] -> conv (+bias, +residual) -> activation, one launch for all groups."""
    layers, xin, slopes, bns = [], [], [], []
    for i, (seq, x) in enumerate(zip(seqs, xs)):
        pad_mod, cm, slope, bn = _unpack_conv_seq(seq)
        if pad_mod is not None:
            l, r, t, b = pad_mod.padding
            assert r == 0 and b == 0, "only left/top reflection padding is used (D_and_G_model.py:235)"
            x = plan.reflect_pad(x, l, t)
        layers.append(_layer(cm, names[i] if names else ""))
        xin.append(x)
        slopes.append(slope)
        bns.append(bn)
    assert all(s == slopes[0] for s in slopes) and all((b is None) == (bns[0] is None) for b in bns)
    slope = slopes[0] if slope_override == "same" else slope_override
    if bns[0] is None:
        return plan.conv(layers, xin, slope, outs=outs, residuals=residuals)
    # use_batchnorm=True: one grouped conv launch (full fp32 out, no bias), then BatchNorm (+ residual) + activation per group,
    # written straight into the concat slice where there is one
    from .MobileNetV2 import BNLayer, _aux
    hs = plan.conv(layers, xin, None, round_out=False)
    res = list(residuals) if residuals is not None else [None] * len(hs)
    return [plan.batchnorm(_aux(bn, BNLayer, (names[i] if names else "") + ".bn"), h, res=res[i], slope=slope,
                           out=None if outs is None else outs[i])
            for i, (bn, h) in enumerate(zip(bns, hs))]


def t_res(plan: Plan, blocks: Sequence[ResidualBlock], xs: Sequence[T], outs=None, names=None) -> List[T]:
    """ResidualBlock (ModificationLayer.py:292-302): act(conv2(act(conv1(x))) + x), the add and the outer activation fused
    into conv2's epilogue."""
    for b in blocks:
        assert len(b.shortcut) == 0 and b.scaling_factor == 1.0 and len(b.layers) == 2, "identity-shortcut blocks only"
    nm = lambda j: [f"{n}.layers.{j}" for n in names] if names else None
    h = t_conv(plan, [b.layers[0] for b in blocks], xs, names=nm(0))
    slope = _negative_slope(blocks[0].activation)
    return t_conv(plan, [b.layers[1] for b in blocks], h, outs=outs, residuals=list(xs), names=nm(1), slope_override=slope)


def t_stage(plan: Plan, seqs: Sequence[nn.Sequential], xs: Sequence[T], outs=None, names=None) -> List[T]:
    """sequential(conv(...), ResidualBlock(...), ...) encoder stage."""
    nm = lambda j: [f"{n}.{j}" for n in names] if names else None
    cur = t_conv(plan, [s[0] for s in seqs], xs, names=nm(0))
    nblk = len(seqs[0]) - 1
    for j in range(1, nblk + 1):
        cur = t_res(plan, [s[j] for s in seqs], cur, outs=outs if j == nblk else None, names=nm(j))
    return cur


# ---------------------------------------------------------------------------------------------------- autograd bridge
class _TraceCache:
    def __init__(self):
        self.plans: Dict[tuple, "_Traced"] = {}


class _Traced:
    def __init__(self, plan: Plan, ins: List[T], outs: List[T]):
        self.plan, self.ins, self.outs = plan, ins, outs


class _EngineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, traced: _Traced, n_in: int, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        plan = traced.plan
        for t, x in zip(traced.ins, inputs):
            x4 = x.detach().float()
            t.act.from_nchw(x4.reshape(t.act.n, t.act.c, t.act.h, t.act.w), round_tf32=not plan.exact)
        for L in plan.layers + plan.aux:
            L.refresh()
        plan.run_forward()
        outs = []
        for t in traced.outs:
            o = t.act.to_nchw()
            if t.flat:
                o = o.reshape(o.shape[0], -1)
            outs.append(o)
        ctx.traced, ctx.module = traced, module
        ctx.n_in, ctx.n_par = n_in, len(params)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        traced: _Traced = ctx.traced
        plan = traced.plan
        for t, g in zip(traced.outs, grads):
            ga = plan.grad_act(t)
            if g is None:
                ops.fill(ga, 0.0)
            else:
                ga.from_nchw(g.detach().float().reshape(ga.n, ga.c, ga.h, ga.w))
        for L in plan.layers + plan.aux:
            L.zero_grad()
        plan.run_backward()
        for L in plan.layers + plan.aux:
            L.export_grad_autograd()
        gin = []
        for t in traced.ins:
            g = plan.grad_act(t).to_nchw() if t.requires_grad else None
            gin.append(g.reshape(g.shape[0], -1) if (g is not None and t.flat) else g)
        return (None, None, None) + tuple(gin) + (None,) * ctx.n_par


class TracedModule(nn.Module):
    """nn.Module whose forward replays a traced launch plan.  Subclasses implement
    `_trace(plan, *T_inputs, **static_kwargs) -> sequence of T outputs` and `_input_specs(*tensors)`."""

    def _cache(self) -> _TraceCache:
        c = self.__dict__.get("_trace_cache")
        if c is None:
            c = _TraceCache()
            self.__dict__["_trace_cache"] = c
        return c

    def _traced_call(self, tensors: Sequence[torch.Tensor], static: tuple = ()):
        for x in tensors:
            if not x.is_cuda:
                raise RuntimeError("tpgan_b200 modules run on CUDA tensors only (there is no CPU fallback); "
                                   "move the module and its inputs to a B200 device")
        key = (tuple((tuple(x.shape), bool(x.requires_grad)) for x in tensors), static, torch.is_grad_enabled(),
               bool(self.training))
        cache = self._cache()
        traced = cache.plans.get(key)
        if traced is None:
            traced = self._build(tensors, static, torch.is_grad_enabled())
            cache.plans[key] = traced
        if self.training:   # nn.BatchNorm bookkeeping in train mode (the statistics themselves are updated by the kernels)
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None:
                    m.num_batches_tracked += 1
        params = [p for p in self.parameters()]
        outs = _EngineFn.apply(self, traced, len(tensors), *tensors, *params)
        return outs

    def _build(self, tensors, static, grad_enabled) -> _Traced:
        dev = tensors[0].device
        exact = getattr(self, "_exact_default", False) if EXACT_MODE is None else bool(EXACT_MODE)
        plan = Plan(dev, training=grad_enabled, need_wgrad=grad_enabled, exact=exact)
        plan.bn_training = bool(self.training)     # BatchNorm follows module.train() / .eval(), not the autograd mode
        ins = []
        for x in tensors:
            if x.dim() == 2:
                t = plan.new(x.shape[0], 1, 1, x.shape[1], name="in", requires_grad=bool(x.requires_grad))
                t.flat = True
            else:
                n, c, h, w = x.shape
                t = plan.new(n, h, w, c, name="in", requires_grad=bool(x.requires_grad))
            ins.append(t)
        outs = list(self._trace(plan, *ins, static=static))
        if grad_enabled:
            for o in outs:
                plan.seed_grad(o)
            plan.trace_backward()
        return _Traced(plan, ins, outs)


# ---------------------------------------------------------------------------------------------------- LocalPathway
class LocalPathway(TracedModule):
    """Patch auto-encoder (reference D_and_G_model.py:18-110).  forward(x) -> (local_img, deconv2 feature)."""

    def __init__(self, use_batchnorm=True, feature_layer_dim=64, FM_multiplier=1.0):
        super().__init__()
        enc = EMaC2I([64, 128, 256, 512], FM_multiplier)
        dec = EMaC2I([256, 128], FM_multiplier)
        leaky = lambda: nn.LeakyReLU(1e-2)
        self.conv0 = sequential(conv(3, enc[0], 3, 1, 1, "kaiming", leaky(), use_batchnorm),
                                ResidualBlock(enc[0], activation=nn.LeakyReLU()))
        self.conv1 = sequential(conv(enc[0], enc[1], 3, 2, 1, "kaiming", leaky(), use_batchnorm),
                                ResidualBlock(enc[1], activation=nn.LeakyReLU()))
        self.conv2 = sequential(conv(enc[1], enc[2], 3, 2, 1, "kaiming", leaky(), use_batchnorm),
                                ResidualBlock(enc[2], activation=nn.LeakyReLU()))
        self.conv3 = sequential(conv(enc[2], enc[3], 3, 2, 1, "kaiming", leaky(), use_batchnorm),
                                ResidualBlock(enc[3], activation=nn.LeakyReLU()))
        self.deconv0 = deconv(enc[3], dec[0], 3, 2, 1, 1, "kaiming", nn.ReLU(), use_batchnorm)
        self.after_select0 = sequential(conv(dec[0] + self.conv2.out_channels, dec[0], 3, 1, 1, "kaiming", nn.LeakyReLU(),
                                             use_batchnorm), ResidualBlock(dec[0], activation=nn.LeakyReLU()))
        self.deconv1 = deconv(self.after_select0.out_channels, dec[1], 3, 2, 1, 1, "kaiming", nn.ReLU(), use_batchnorm)
        self.after_select1 = sequential(conv(dec[1] + self.conv1.out_channels, dec[1], 3, 1, 1, "kaiming", nn.LeakyReLU(),
                                             use_batchnorm), ResidualBlock(dec[1], activation=nn.LeakyReLU()))
        self.deconv2 = deconv(self.after_select1.out_channels, feature_layer_dim, 3, 2, 1, 1, "kaiming", nn.ReLU(),
                              use_batchnorm)
        self.after_select2 = sequential(conv(feature_layer_dim + self.conv0.out_channels, feature_layer_dim, 3, 1, 1,
                                             "kaiming", nn.LeakyReLU(), use_batchnorm),
                                        ResidualBlock(feature_layer_dim, activation=nn.LeakyReLU()))
        self.local_img = conv(feature_layer_dim, 3, 1, 1, 0, None, None, False)

    def forward(self, x):
        local_img, feat = self._traced_call([x])
        assert local_img.shape == x.shape, "{} {}".format(local_img.shape, x.shape)
        return local_img, feat

    def _trace(self, plan, x, static=()):
        imgs, feats = trace_local_pathways(plan, [self], [x], ["local_pathway"])
        return imgs[0], feats[0]


def trace_local_pathways(plan: Plan, paths: Sequence[LocalPathway], xs: Sequence[T], names: Sequence[str]):
    """The local pathways as grouped launches: layer i of every pathway runs in ONE kernel launch
    (reference D_and_G_model.py:84-110 run per pathway, :390-393)."""
    G = len(paths)
    nm = lambda a: [f"{n}.{a}" for n in names]
    cat0, cat1, cat2 = [], [], []
    for p, x in zip(paths, xs):
        n, h, w = x.act.n, x.act.h, x.act.w
        assert h % 8 == 0 and w % 8 == 0, "local patches must be multiples of 8 (three stride-2 stages)"
        c0, c1, c2 = p.conv0.out_channels, p.conv1.out_channels, p.conv2.out_channels
        cat2.append(plan.concat(n, h, w, [p.deconv2.out_channels, c0], name="ls2"))
        cat1.append(plan.concat(n, h // 2, w // 2, [p.deconv1.out_channels, c1], name="ls1"))
        cat0.append(plan.concat(n, h // 4, w // 4, [p.deconv0.out_channels, c2], name="ls0"))
    part = lambda cats, i: [c.parts[i] for c in cats]
    conv0 = t_stage(plan, [p.conv0 for p in paths], xs, outs=part(cat2, 1), names=nm("conv0"))
    conv1 = t_stage(plan, [p.conv1 for p in paths], conv0, outs=part(cat1, 1), names=nm("conv1"))
    conv2 = t_stage(plan, [p.conv2 for p in paths], conv1, outs=part(cat0, 1), names=nm("conv2"))
    conv3 = t_stage(plan, [p.conv3 for p in paths], conv2, names=nm("conv3"))
    t_conv(plan, [p.deconv0 for p in paths], conv3, outs=part(cat0, 0), names=nm("deconv0"))
    as0 = t_stage(plan, [p.after_select0 for p in paths], cat0, names=nm("after_select0"))
    t_conv(plan, [p.deconv1 for p in paths], as0, outs=part(cat1, 0), names=nm("deconv1"))
    as1 = t_stage(plan, [p.after_select1 for p in paths], cat1, names=nm("after_select1"))
    feats = t_conv(plan, [p.deconv2 for p in paths], as1, outs=part(cat2, 0), names=nm("deconv2"))
    as2 = t_stage(plan, [p.after_select2 for p in paths], cat2, names=nm("after_select2"))
    imgs = t_conv(plan, [p.local_img for p in paths], as2, names=nm("local_img"))
    return imgs, feats


# ---------------------------------------------------------------------------------------------------- LocalFuser
class LocalFuser(TracedModule):
    """max over the four zero-padded patches at the reference's fixed offsets (D_and_G_model.py:112-159)."""

    def __init__(self):
        super().__init__()

    def forward(self, f_left_eye, f_right_eye, f_nose, f_mouth):
        return self._traced_call([f_left_eye, f_right_eye, f_nose, f_mouth])[0]

    def _trace(self, plan, le, re, nose, mouth, static=()):
        out = plan.new(le.act.n, 128, 128, le.act.c, name="fused")
        plan.local_fuse([le, re, nose, mouth], out)
        return (out,)


# ---------------------------------------------------------------------------------------------------- GlobalPathway
class GlobalPathway(TracedModule):
    """Encoder - fc/maxout bottleneck - decoder with skip concatenations (reference D_and_G_model.py:161-329)."""

    def __init__(self, zdim, local_feature_layer_dim=64, use_batchnorm=True, use_residual_block=True, scaling_factor=1.0,
                 FM_multiplier=1.0):
        super().__init__()
        enc = EMaC2I([64, 64, 128, 256, 512], FM_multiplier)
        dec = EMaC2I([64, 32, 16, 8], FM_multiplier)
        enh = EMaC2I([512, 256, 128, 64], FM_multiplier)
        dconv = EMaC2I([64, 32], FM_multiplier)
        self.zdim = zdim
        self.use_residual_block = use_residual_block
        lk = lambda: nn.LeakyReLU(1e-2)
        sf = scaling_factor
        self.conv0 = sequential(conv(3, enc[0], 7, 1, 3, "kaiming", lk(), use_batchnorm),
                                ResidualBlock(64, 64, 7, 1, 3, "kaiming", lk(), scaling_factor=sf))
        self.conv1 = sequential(conv(enc[1], enc[1], 5, 2, 2, "kaiming", lk(), use_batchnorm),
                                ResidualBlock(64, 64, 5, 1, 2, "kaiming", lk(), scaling_factor=sf))
        self.conv2 = sequential(conv(enc[1], enc[2], 3, 2, 1, "kaiming", lk(), use_batchnorm),
                                ResidualBlock(128, 128, 3, 1, 1, "kaiming", lk(), scaling_factor=sf))
        self.conv3 = sequential(conv(enc[2], enc[3], 3, 2, 1, "kaiming", lk(), use_batchnorm),
                                ResidualBlock(256, 256, 3, 1, 1, "kaiming", lk(), is_bottleneck=False, scaling_factor=sf))
        self.conv4 = sequential(conv(enc[3], enc[4], 3, 2, 1, "kaiming", lk(), use_batchnorm),
                                *[ResidualBlock(512, 512, 3, 1, 1, "kaiming", lk(), is_bottleneck=False, scaling_factor=sf)
                                  for _ in range(4)])
        self.fc1 = nn.Linear(enc[4] * 8 * 8, 512)
        self.fc2 = nn.MaxPool1d(2, 2, 0)
        self.deconv_8 = deconv(256 + self.zdim, dec[0], 8, 1, 0, 0, "kaiming", nn.ReLU(), use_batchnorm)
        self.deconv_32 = deconv(dec[0], dec[1], 3, 4, 0, 1, "kaiming", nn.ReLU(), use_batchnorm)
        self.deconv_64 = deconv(dec[1], dec[2], 3, 2, 1, 1, "kaiming", nn.ReLU(), use_batchnorm)
        self.deconv_128 = deconv(dec[2], dec[3], 3, 2, 1, 1, "kaiming", nn.ReLU(), use_batchnorm)
        dim8 = self.deconv_8.out_channels + self.conv4.out_channels
        self.add_conv_and_deconv_8 = ResidualBlock(dim8, dim8, 2, 1, padding=[1, 0, 1, 0], activation=nn.LeakyReLU())
        self.enhance_features_8 = sequential(*[ResidualBlock(dim8, dim8, 2, 1, padding=[1, 0, 1, 0],
                                                             activation=nn.LeakyReLU()) for _ in range(2)])
        self.upsample_16 = deconv(self.enhance_features_8.out_channels, enh[0], 3, 2, 1, 1, "kaiming", nn.ReLU(),
                                  use_batchnorm)
        dim16 = self.conv3.out_channels
        self.add_conv_and_deconv_16 = ResidualBlock(dim16, activation=nn.LeakyReLU())
        self.enhance_features_16 = sequential(*[ResidualBlock(
            self.upsample_16.out_channels + self.add_conv_and_deconv_16.out_channels, activation=nn.LeakyReLU())
            for _ in range(2)])
        self.upsample_32 = deconv(self.enhance_features_16.out_channels, enh[1], 3, 2, 1, 1, "kaiming", nn.ReLU(),
                                  use_batchnorm)
        dim32 = self.conv2.out_channels + self.deconv_32.out_channels
        self.add_conv_and_deconv_32 = ResidualBlock(dim32, activation=nn.LeakyReLU())
        self.enhance_features_32 = sequential(*[ResidualBlock(
            self.upsample_32.out_channels + self.add_conv_and_deconv_32.out_channels, activation=nn.LeakyReLU())
            for _ in range(2)])
        self.upsample_64 = deconv(self.enhance_features_32.out_channels, enh[2], 3, 2, 1, 1, "kaiming", nn.ReLU(),
                                  use_batchnorm)
        dim64 = self.conv1.out_channels + self.deconv_64.out_channels
        self.add_conv_and_deconv_64 = ResidualBlock(dim64, kernel_size=5, activation=nn.LeakyReLU())
        self.enhance_features_64 = sequential(*[ResidualBlock(
            self.upsample_64.out_channels + self.add_conv_and_deconv_64.out_channels, activation=nn.LeakyReLU())
            for _ in range(2)])
        self.upsample_128 = deconv(self.enhance_features_64.out_channels, enh[3], 3, 2, 1, 1, "kaiming", nn.ReLU(),
                                   use_batchnorm)
        # F3: forward() concatenates [deconv_128, conv0, I128] (reference :323), so the 3 image channels belong here
        dim128 = self.conv0.out_channels + self.deconv_128.out_channels + 3
        self.add_conv_and_deconv_128 = ResidualBlock(dim128, kernel_size=7, activation=nn.LeakyReLU())
        self.enhance_features_128 = sequential(*[ResidualBlock(
            self.upsample_128.out_channels + self.add_conv_and_deconv_128.out_channels + local_feature_layer_dim + 3,
            kernel_size=5, activation=nn.LeakyReLU())])
        self.conv5 = sequential(conv(self.enhance_features_128.out_channels, dconv[0], 5, 1, 2, "kaiming", nn.LeakyReLU(),
                                     use_batchnorm), ResidualBlock(dconv[0], kernel_size=3, activation=nn.LeakyReLU()))
        self.conv6 = conv(dconv[0], dconv[1], 3, 1, 1, "kaiming", nn.LeakyReLU(), use_batchnorm)
        self.decoded_img128 = conv(dconv[1], 3, 3, 1, 1, None, activation=None)

    def forward(self, I128, local_fake_image, local_feature, z):
        img, fc2 = self._traced_call([I128, local_fake_image, local_feature, z])
        return img, fc2

    def _trace(self, plan, I128, local_fake_image, local_feature, z, static=()):
        bufs = self.alloc_concats(plan, I128.act.n)
        plan.copy(I128, bufs["a128"].parts[2])
        plan.copy(local_feature, bufs["f128"].parts[2])
        plan.copy(local_fake_image, bufs["f128"].parts[3])
        plan.copy(z, bufs["zin"].parts[1])
        return self.trace_body(plan, bufs)

    # -- shared with Generator (which writes I128 / z / fused maps straight into the concat buffers)
    def alloc_concats(self, plan: Plan, n: int) -> Dict[str, T]:
        c = lambda m: m.out_channels
        return dict(
            a128=plan.concat(n, 128, 128, [c(self.deconv_128), c(self.conv0), 3], name="a128_in"),
            f128=plan.concat(n, 128, 128, [c(self.upsample_128), c(self.add_conv_and_deconv_128), 64, 3], name="f128_in"),
            a64=plan.concat(n, 64, 64, [c(self.deconv_64), c(self.conv1)], name="a64_in"),
            f64=plan.concat(n, 64, 64, [c(self.upsample_64), c(self.add_conv_and_deconv_64)], name="f64_in"),
            a32=plan.concat(n, 32, 32, [c(self.deconv_32), c(self.conv2)], name="a32_in"),
            f32=plan.concat(n, 32, 32, [c(self.upsample_32), c(self.add_conv_and_deconv_32)], name="f32_in"),
            f16=plan.concat(n, 16, 16, [c(self.upsample_16), c(self.add_conv_and_deconv_16)], name="f16_in"),
            f8=plan.concat(n, 8, 8, [c(self.deconv_8), c(self.conv4)], name="f8_in"),
            zin=plan.concat(n, 1, 1, [256, self.zdim], name="z_in"),
        )

    def _special_layers(self):
        L = self.__dict__.get("_tc_special")
        if L is None:
            fc1 = ConvLayer(self.fc1.weight, self.fc1.bias, False, 8, 1, 0, "global_pathway.fc1",
                            w_shape=(self.fc1.out_features, self.fc1.in_features // 64, 8, 8))
            fc1.dgrad_gemm = True      # its input gradient is one GEMM over the dgrad packing (engine._emit_dgrad_gemm)
            d8 = self.deconv_8[0]
            assert d8.kernel_size == (8, 8) and d8.stride == (1, 1) and d8.padding == (0, 0)
            dec8 = DeconvAsLinear(d8.weight, d8.bias, 8, "global_pathway.deconv_8")
            L = (fc1, dec8)
            self.__dict__["_tc_special"] = L
        return L

    def trace_body(self, plan: Plan, B: Dict[str, T], pre: str = "global_pathway"):
        """reference forward, D_and_G_model.py:281-329.  B = alloc_concats(); I128 must already sit in a128.parts[2], the
        fused local feature / image in f128.parts[2:4] and z in zin.parts[1]."""
        nm = lambda a: [f"{pre}.{a}"]
        I128 = B["a128"].parts[2]
        n = I128.act.n
        conv0 = t_stage(plan, [self.conv0], [I128], outs=[B["a128"].parts[1]], names=nm("conv0"))
        conv1 = t_stage(plan, [self.conv1], conv0, outs=[B["a64"].parts[1]], names=nm("conv1"))
        conv2 = t_stage(plan, [self.conv2], conv1, outs=[B["a32"].parts[1]], names=nm("conv2"))
        conv3 = t_stage(plan, [self.conv3], conv2, names=nm("conv3"))
        conv4 = t_stage(plan, [self.conv4], conv3, outs=[B["f8"].parts[1]], names=nm("conv4"))
        fc1_layer, dec8_layer = self._special_layers()
        fc1 = plan.conv([fc1_layer], conv4, None)[0]                       # Linear(32768, 512) as an 8x8 valid conv
        fc2 = plan.maxout2(fc1, name="fc2")                               # MaxPool1d(2,2) "maxout"
        plan.copy(fc2, B["zin"].parts[0])
        bn8 = next((m for m in self.deconv_8 if isinstance(m, nn.BatchNorm2d)), None)
        if bn8 is None:
            d8_flat = plan.conv([dec8_layer], [B["zin"]], 0.0)[0]          # ConvTranspose2d k8 on 1x1 = GEMM, ReLU
            d8 = plan.alias(d8_flat, 8, 8, self.deconv_8.out_channels, name="deconv_8")
        else:                                                              # use_batchnorm=True: GEMM -> BatchNorm2d(8x8 map) -> ReLU
            from .MobileNetV2 import BNLayer, _aux
            d8_flat = plan.conv([dec8_layer], [B["zin"]], None, round_out=False)[0]
            d8 = plan.alias(d8_flat, 8, 8, self.deconv_8.out_channels, name="deconv_8.gemm")
            d8 = plan.batchnorm(_aux(bn8, BNLayer, f"{pre}.deconv_8.bn"), d8, slope=0.0, name="deconv_8")
        plan.copy(d8, B["f8"].parts[0])
        d32 = t_conv(plan, [self.deconv_32], [d8], outs=[B["a32"].parts[0]], names=nm("deconv_32"))
        d64 = t_conv(plan, [self.deconv_64], d32, outs=[B["a64"].parts[0]], names=nm("deconv_64"))
        t_conv(plan, [self.deconv_128], d64, outs=[B["a128"].parts[0]], names=nm("deconv_128"))
        f8 = t_res(plan, [self.add_conv_and_deconv_8], [B["f8"]], names=nm("add_conv_and_deconv_8"))
        f8 = t_res(plan, [self.enhance_features_8[0]], f8, names=nm("enhance_features_8.0"))
        f8 = t_res(plan, [self.enhance_features_8[1]], f8, names=nm("enhance_features_8.1"))
        assert f8[0].act.h == 8
        t_conv(plan, [self.upsample_16], f8, outs=[B["f16"].parts[0]], names=nm("upsample_16"))
        t_res(plan, [self.add_conv_and_deconv_16], conv3, outs=[B["f16"].parts[1]], names=nm("add_conv_and_deconv_16"))
        f16 = t_res(plan, [self.enhance_features_16[0]], [B["f16"]], names=nm("enhance_features_16.0"))
        f16 = t_res(plan, [self.enhance_features_16[1]], f16, names=nm("enhance_features_16.1"))
        assert f16[0].act.h == 16
        t_conv(plan, [self.upsample_32], f16, outs=[B["f32"].parts[0]], names=nm("upsample_32"))
        t_res(plan, [self.add_conv_and_deconv_32], [B["a32"]], outs=[B["f32"].parts[1]], names=nm("add_conv_and_deconv_32"))
        f32 = t_res(plan, [self.enhance_features_32[0]], [B["f32"]], names=nm("enhance_features_32.0"))
        f32 = t_res(plan, [self.enhance_features_32[1]], f32, names=nm("enhance_features_32.1"))
        t_conv(plan, [self.upsample_64], f32, outs=[B["f64"].parts[0]], names=nm("upsample_64"))
        t_res(plan, [self.add_conv_and_deconv_64], [B["a64"]], outs=[B["f64"].parts[1]], names=nm("add_conv_and_deconv_64"))
        f64 = t_res(plan, [self.enhance_features_64[0]], [B["f64"]], names=nm("enhance_features_64.0"))
        f64 = t_res(plan, [self.enhance_features_64[1]], f64, names=nm("enhance_features_64.1"))
        t_conv(plan, [self.upsample_128], f64, outs=[B["f128"].parts[0]], names=nm("upsample_128"))
        t_res(plan, [self.add_conv_and_deconv_128], [B["a128"]], outs=[B["f128"].parts[1]],
              names=nm("add_conv_and_deconv_128"))
        f128 = t_res(plan, [self.enhance_features_128[0]], [B["f128"]], names=nm("enhance_features_128.0"))
        conv5 = t_stage(plan, [self.conv5], f128, names=nm("conv5"))
        conv6 = t_conv(plan, [self.conv6], conv5, names=nm("conv6"))
        img = t_conv(plan, [self.decoded_img128], conv6, names=nm("decoded_img128"))[0]
        fc2.flat = True
        return img, fc2


# ---------------------------------------------------------------------------------------------------- FeaturePredict
class FeaturePredict(TracedModule):
    """[Dropout(0.3)] -> Linear(256, num_classes)  (reference D_and_G_model.py:331-348)."""

    def __init__(self, num_classes, global_feature_layer_dim=256, dropout=0.3):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.fc = nn.Linear(global_feature_layer_dim, num_classes)

    def _fc_layer(self) -> ConvLayer:
        L = self.__dict__.get("_tc_fc")
        if L is None:
            L = ConvLayer(self.fc.weight, self.fc.bias, False, 1, 1, 0, "feature_predict.fc",
                          w_shape=(self.fc.out_features, self.fc.in_features, 1, 1))
            self.__dict__["_tc_fc"] = L
        return L

    def draw_mask(self, n: int, device) -> torch.Tensor:
        """Bernoulli keep-mask scaled by 1/(1-p), as nn.Dropout applies in training mode."""
        p = self.dropout.p
        return (torch.rand((n, 1, 1, self.fc.in_features), device=device) >= p).float().div_(1.0 - p)

    def trace(self, plan: Plan, x: T, mask: Optional[Act]) -> T:
        if mask is not None:
            x = plan.mul_mask(x, mask)
        out = plan.conv([self._fc_layer()], [x], None)[0]
        out.flat = True
        return out

    def forward(self, x, use_dropout):
        if use_dropout and self.training:
            mask = self.draw_mask(x.shape[0], x.device)
            return self._traced_call([x, mask.view(x.shape[0], -1)], static=("dropout",))[0]
        return self._traced_call([x])[0]

    def _trace(self, plan, x, mask=None, static=()):
        return (self.trace(plan, x, None if mask is None else mask.act),)


# ---------------------------------------------------------------------------------------------------- Generator
class Generator(TracedModule):
    """Two-pathway generator (reference D_and_G_model.py:350-407).

    forward(I128, left_eye, right_eye, nose, mouth, z, use_dropout) -> (I128_fake, encoder_predict,
    fused_local_fake_image, left_eye_fake, right_eye_fake, nose_fake, mouth_fake, fused_local_origin_4_part)."""

    def __init__(self, zdim, num_classes, use_batchnorm=True, use_residual_block=True):
        super().__init__()
        self.local_pathway_left_eye = LocalPathway(use_batchnorm=use_batchnorm)
        self.local_pathway_right_eye = LocalPathway(use_batchnorm=use_batchnorm)
        self.local_pathway_nose = LocalPathway(use_batchnorm=use_batchnorm)
        self.local_pathway_mouth = LocalPathway(use_batchnorm=use_batchnorm)
        self.global_pathway = GlobalPathway(zdim, use_batchnorm=use_batchnorm, use_residual_block=use_residual_block)
        self.local_fuser = LocalFuser()
        self.feature_predict = FeaturePredict(num_classes)

    def local_pathways(self):
        return [self.local_pathway_left_eye, self.local_pathway_right_eye, self.local_pathway_nose,
                self.local_pathway_mouth]

    def forward(self, I128, left_eye, right_eye, nose, mouth, z, use_dropout):
        tensors = [I128, left_eye, right_eye, nose, mouth, z]
        if use_dropout and self.training:
            mask = self.feature_predict.draw_mask(I128.shape[0], I128.device).view(I128.shape[0], -1)
            return self._traced_call(tensors + [mask], static=("dropout",))
        return self._traced_call(tensors)

    def _trace(self, plan, I128, le, re, nose, mouth, z, mask=None, static=()):
        gp = self.global_pathway
        B = gp.alloc_concats(plan, I128.act.n)
        plan.copy(I128, B["a128"].parts[2])
        plan.copy(z, B["zin"].parts[1])
        return self.trace_body(plan, B, [le, re, nose, mouth], None if mask is None else mask.act)

    def trace_body(self, plan: Plan, B: Dict[str, T], patches: Sequence[T], mask: Optional[Act]):
        """I128 in B['a128'].parts[2], z in B['zin'].parts[1], the four input patches in `patches`."""
        n = patches[0].act.n
        names = [f"local_pathway_{p}" for p in PART_NAMES]
        imgs, feats = trace_local_pathways(plan, self.local_pathways(), patches, names)
        plan.local_fuse(feats, B["f128"].parts[2], name="feature")
        fused_img = plan.local_fuse(imgs, B["f128"].parts[3], name="fake_image")
        fused_in = plan.new(n, 128, 128, 3, name="fused_local_origin_4_part", requires_grad=False)
        plan.local_fuse(patches, fused_in, name="origin")
        fake, fc2 = self.global_pathway.trace_body(plan, B)
        logits = self.feature_predict.trace(plan, fc2, mask)
        self_outs = (fake, logits, fused_img, imgs[0], imgs[1], imgs[2], imgs[3], fused_in)
        return self_outs


# ---------------------------------------------------------------------------------------------------- Discriminator
class Discriminator(TracedModule):
    """PatchGAN-style critic (reference D_and_G_model.py:409-435): (B,3,128,128) -> (B,1,4,4) logits, no sigmoid."""

    def __init__(self, use_batchnorm=False, FM_multiplier=1.0):
        super().__init__()
        layers = []
        n_fmap = EMaC2I([3, 64, 128, 256, 512, 512], FM_multiplier)
        for i in range(len(n_fmap) - 1):
            layers.append(conv(n_fmap[i], n_fmap[i + 1], 3, 2, 1, "kaiming", nn.LeakyReLU(1e-2), use_batchnorm))
            if i >= 3:
                layers.append(ResidualBlock(n_fmap[i + 1], activation=nn.LeakyReLU()))
        layers.append(conv(n_fmap[-1], 1, kernel_size=3, stride=1, padding=1, init=None, activation=None))
        self.model = sequential(*layers)

    def forward(self, x):
        return self._traced_call([x])[0]

    def _trace(self, plan, x, static=()):
        cur = [x]
        for i, m in enumerate(self.model):
            if isinstance(m, ResidualBlock):
                cur = t_res(plan, [m], cur, names=[f"model.{i}"])
            else:
                cur = t_conv(plan, [m], cur, names=[f"model.{i}"])
        return (cur[0],)
