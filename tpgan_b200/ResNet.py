"""Identity / feature network: a ResNet18 for 128x128 inputs with the interface the reference INTENDS
(ResNet.py:5-119: conv1 7x7 s2 p3 -> 64 + BN + ReLU (:30-31), MaxPool2d(3,2,1) (:33), stages [64,128,256,512] x [2,2,2,2]
residual blocks (:28-29,38-42), AdaptiveAvgPool2d((1,1)) (:45), optional FC0 512 -> feature_layer_dim_before_FC (+BN1d)
(:48-49), Dropout (:52), FC -> classes (:55), forward(x, use_dropout) -> (out, out_FC0) (:80-119)).

The reference's own class cannot be constructed (conv() has no `bias` kwarg, the block factory is called
with stride in the kernel_size slot, only 3 of 4 stages are built and all with stride 1, `resnet18()` is mis-indented), so
this is a RESTATEMENT, not a drop-in for a working class: the canonical ResNet18 wiring (stage strides 1,2,2,2, projection
shortcut where shape changes) with the reference's constructor/forward signatures.  Parity for it is oracle-defined
(oracle/identity_port.py), stated in DESIGN.md.

Execution: in eval mode (the mode the TP-GAN step uses it in: a FROZEN identity network) BatchNorm is folded into the
preceding conv / linear and the whole forward (and the input gradient) runs as a traced plan on the tcgen05 kernels.
In training mode (pre-training the feature extractor, BASELINE config 5 "ResNet backbones") the same layers run unfolded:
conv -> training-mode BatchNorm (+ReLU, + the block's shortcut add) with the kernels of csrc/pretrain.cu, forward and
backward through the autograd bridge of TracedModule, parameter gradients in the reference layout.
"""
from __future__ import annotations

import copy
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from .D_and_G_model import TracedModule, _layer
from .engine import ConvLayer, Plan, T
from .ModificationLayer import conv, linear


class BasicBlock(nn.Module):
    """act( BN(conv3x3(act(BN(conv3x3_s(x))))) + shortcut(x) ); shortcut = 1x1 stride-s conv + BN where the shape changes."""

    def __init__(self, in_channels, out_channels, stride=1, activation=nn.ReLU(inplace=True), use_batchnorm=True):
        super().__init__()
        self.out_channels = out_channels
        self.conv_a = conv(in_channels, out_channels, 3, stride, 1, "kaiming", activation, use_batchnorm)
        self.conv_b = conv(out_channels, out_channels, 3, 1, 1, "kaiming", None, use_batchnorm)
        self.shortcut = conv(in_channels, out_channels, 1, stride, 0, "kaiming", None, use_batchnorm) \
            if (stride != 1 or in_channels != out_channels) else nn.Sequential()
        self.activation = activation

    def forward(self, x):
        return self.activation(self.conv_b(self.conv_a(x)) + self.shortcut(x))


def _fold(seq: nn.Sequential):
    """conv()/linear() stack [Conv2d|Linear, BatchNorm?, act?] -> (weight, bias, has_relu) with eval-mode BN folded in."""
    lin, bn, relu = None, None, False
    for m in seq:
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            lin = m
        elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
            bn = m
        elif isinstance(m, nn.ReLU):
            relu = True
        else:
            raise NotImplementedError(type(m).__name__)
    w = lin.weight.detach().float()
    b = lin.bias.detach().float() if lin.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if bn is not None:
        scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        w = w * scale.view(-1, *([1] * (w.dim() - 1)))
        b = (b - bn.running_mean) * scale + bn.bias.detach()
    return lin, w.contiguous(), b.contiguous(), relu


class ResNet18(TracedModule):
    def __init__(self, residualBlock=BasicBlock, num_of_output_classes=1000, use_batchnorm=True,
                 feature_layer_dim_before_FC=None, activation=nn.ReLU(inplace=True), dropout_rate=0.0):
        super().__init__()
        self.use_batchnorm = use_batchnorm
        self.activation = activation
        self.feature_layer_dim_before_FC = feature_layer_dim_before_FC
        num_features, num_sections, strides = [64, 128, 256, 512], [2, 2, 2, 2], [1, 2, 2, 2]
        self.conv1 = conv(3, num_features[0], 7, 2, 3, "kaiming", activation, use_batchnorm)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        sections, cin = [], num_features[0]
        for cout, nblk, st in zip(num_features, num_sections, strides):
            sections.append(self._build_blocks(residualBlock, cin, cout, st, nblk))
            cin = cout
        self.sections = nn.Sequential(*sections)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        if feature_layer_dim_before_FC is not None:
            self.FC0 = linear(num_features[-1], feature_layer_dim_before_FC, use_batchnorm=use_batchnorm)
        self.dropout = nn.Dropout(dropout_rate)
        fc_in = feature_layer_dim_before_FC if feature_layer_dim_before_FC is not None else num_features[-1]
        self.FC = linear(fc_in, num_of_output_classes, use_batchnorm=False)

    def _build_blocks(self, residualBlock, in_channels, out_channels, stride, num_of_residual_block):
        layers = []
        for i in range(num_of_residual_block):
            layers.append(residualBlock(in_channels if i == 0 else out_channels, out_channels, stride if i == 0 else 1,
                                        activation=copy.deepcopy(self.activation), use_batchnorm=self.use_batchnorm))
        return nn.Sequential(*layers)

    # ---- traced execution (eval mode, BN folded)
    def _fold_specs(self):
        """(name, conv()/linear() stack, kernel, stride, pad) of every folded layer, in forward order."""
        specs = [("conv1", self.conv1, 7, 2, 3)]
        for si, sec in enumerate(self.sections):
            for bi, blk in enumerate(sec):
                st = blk.conv_a[0].stride[0]
                specs.append((f"sections.{si}.{bi}.conv_a", blk.conv_a, 3, st, 1))
                specs.append((f"sections.{si}.{bi}.conv_b", blk.conv_b, 3, 1, 1))
                if len(blk.shortcut):
                    specs.append((f"sections.{si}.{bi}.shortcut", blk.shortcut, 1, st, 0))
        if hasattr(self, "FC0"):
            specs.append(("FC0", self.FC0, 1, 1, 0))
        specs.append(("FC", self.FC, 1, 1, 0))
        return specs

    def _fold_fingerprint(self):
        """Changes whenever a parameter / running statistic is written through torch (optimizer.step, load_state_dict,
        train-mode forward) or a fused trainer reports raw-pointer writes (invalidate_folded())."""
        ts = list(self.parameters()) + list(self.buffers())
        return (self.__dict__.get("_fold_epoch", 0),) + tuple((t._version, t.data_ptr()) for t in ts)

    def invalidate_folded(self):
        """The fused trainers (ClassifierTrainer) update parameters and running statistics through raw device pointers,
        which torch's version counters do not see: they call this after every step."""
        self.__dict__["_fold_epoch"] = self.__dict__.get("_fold_epoch", 0) + 1

    def _folded_layers(self, device):
        """The BatchNorm-folded ConvLayers of the eval-mode network.  Built once; when the weights or running statistics
        changed since (train -> eval cycles, load_state_dict) they are re-folded IN PLACE, so plans that hold the packed
        operand pointers (IdentityPlan) see the new weights."""
        fp = self._fold_fingerprint()
        L = self.__dict__.get("_traced_layers")
        if L is not None and self.__dict__.get("_fold_fp") == fp:
            return L
        fresh = L is None
        if fresh:
            L = {}
        for name, seq, k, s, p in self._fold_specs():
            lin, w, b, relu = _fold(seq)
            w4 = w if w.dim() == 4 else w.view(w.shape[0], w.shape[1], 1, 1)
            if fresh:
                L[name] = (ConvLayer(w4.to(device), b.to(device), False, k, s, p, name), relu)
            else:
                layer = L[name][0]
                layer.weight.copy_(w4)
                layer.bias.copy_(b)
                layer.repack()
        self.__dict__["_traced_layers"] = L
        self.__dict__["_fold_fp"] = fp
        return L

    def refresh_folded(self):
        """Re-fold if stale (cheap fingerprint check); trainers that hold an IdentityPlan call this before a step."""
        if self.__dict__.get("_traced_layers") is not None:
            dev = next(self.parameters()).device
            self._folded_layers(dev)

    def trace(self, plan: Plan, x: T, with_logits: bool = True):
        """x: (N,128,128,3) NHWC -> (pooled 512 feature T, FC0 feature T or None, logits T or None)."""
        assert not self.training, "the traced identity network is eval-only (BatchNorm folded with running statistics)"
        # one set of folded tensor-core layers per module: several plans (fake / gt batches) share the packed weights
        L = self._folded_layers(x.act.buf.device)

        def cv(name, t, residual=None, relu=None):
            layer, r = L[name]
            r = r if relu is None else relu
            return plan.conv([layer], [t], 0.0 if r else None, residuals=None if residual is None else [residual])[0]
        h = cv("conv1", x)
        h = plan.maxpool3s2(h)
        for si, sec in enumerate(self.sections):
            for bi, blk in enumerate(sec):
                pre = f"sections.{si}.{bi}"
                a = cv(pre + ".conv_a", h)
                sc = cv(pre + ".shortcut", h) if len(blk.shortcut) else h
                h = cv(pre + ".conv_b", a, residual=sc, relu=True)
        pooled = plan.avgpool(h)
        fc0 = cv("FC0", pooled) if hasattr(self, "FC0") else None
        logits = cv("FC", fc0 if fc0 is not None else pooled) if with_logits else None
        return pooled, fc0, logits

    # ---- traced execution (training mode: batch-statistics BatchNorm, weight gradients)
    @staticmethod
    def _parts(seq: nn.Sequential):
        lin = bn = None
        relu = False
        for m in seq:
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                lin = m
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                bn = m
            elif isinstance(m, nn.ReLU):
                relu = True
            else:
                raise NotImplementedError(type(m).__name__)
        return lin, bn, relu

    @staticmethod
    def _dense_layer(lin, name: str) -> ConvLayer:
        if isinstance(lin, nn.Conv2d):
            return _layer(lin, name)
        L = lin.__dict__.get("_tc_layer")
        if L is None:   # nn.Linear as a 1x1 convolution over the (B,1,1,C) row
            L = ConvLayer(lin.weight, lin.bias, False, 1, 1, 0, name, w_shape=(lin.out_features, lin.in_features, 1, 1))
            object.__setattr__(lin, "_tc_layer", L)
        return L

    def trace_train(self, plan: Plan, x: T):
        """x: (N,128,128,3) NHWC -> (logits T, FC0 feature T or None, pooled T), BatchNorm on batch statistics."""
        from .MobileNetV2 import BNLayer, _aux

        def cbr(name, seq, t, res=None, relu=None, feeds_conv=True):
            lin, bn, has_relu = self._parts(seq)
            act = has_relu if relu is None else relu
            L = self._dense_layer(lin, name + ".0")
            if bn is None:      # use_batchnorm=False: bias (+ residual) (+ ReLU) in the conv epilogue
                return plan.conv([L], [t], 0.0 if act else None, residuals=None if res is None else [res],
                                 round_out=feeds_conv)[0]
            h = plan.conv([L], [t], None, round_out=False)[0]
            return plan.batchnorm(_aux(bn, BNLayer, name + ".1"), h, res=res, slope=0.0 if act else None,
                                  round_out=feeds_conv)
        h = cbr("conv1", self.conv1, x)
        h = plan.maxpool3s2(h)
        for si, sec in enumerate(self.sections):
            for bi, blk in enumerate(sec):
                pre = f"sections.{si}.{bi}"
                a = cbr(pre + ".conv_a", blk.conv_a, h)
                sc = cbr(pre + ".shortcut", blk.shortcut, h, feeds_conv=False) if len(blk.shortcut) else h
                h = cbr(pre + ".conv_b", blk.conv_b, a, res=sc, relu=True)
        pooled = plan.avgpool(h)
        fc0 = cbr("FC0", self.FC0, pooled) if hasattr(self, "FC0") else None
        logits = cbr("FC", self.FC, fc0 if fc0 is not None else pooled, feeds_conv=False)
        return logits, fc0, pooled

    def _trace(self, plan, x, static=()):
        plan.training = plan.bn_training = True     # max-pool arg-max is recorded, BatchNorm uses batch statistics
        logits, fc0, _ = self.trace_train(plan, x)
        logits.flat = True
        if fc0 is not None:
            fc0.flat = True
        return [logits] if fc0 is None else [logits, fc0]

    def forward(self, x, use_dropout=False):
        if not x.is_cuda:
            raise RuntimeError("tpgan_b200 modules run on CUDA tensors only (there is no CPU fallback)")
        if self.training:
            if use_dropout and self.dropout.p > 0:
                raise NotImplementedError("dropout > 0 between FC0 and FC is not built (ResNet.py:52 default 0.0)")
            self.invalidate_folded()        # the folded eval-mode copies are stale once the weights train
            outs = self._traced_call([x], static=("train",))
            return (outs[0], outs[1]) if len(outs) > 1 else (outs[0], None)
        if use_dropout and self.dropout.p > 0:
            raise NotImplementedError("dropout is a training-time option")
        if x.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("ResNet18.eval()(x) returns plain tensors (no autograd graph): the identity loss and its "
                               "input gradient run through TPGANTrainer(identity_net=...) / IdentityPlan; call under "
                               "torch.no_grad() or detach the input")
        n = x.shape[0]
        key = (n, x.shape[2], x.shape[3], str(x.device))
        cache = self.__dict__.setdefault("_eval_plans", {})
        if key not in cache:     # one plan per input shape: activations are allocated once, not per call
            plan = Plan(x.device, training=False, need_wgrad=False)
            t = plan.new(n, x.shape[2], x.shape[3], 3, name="in", requires_grad=False)
            cache[key] = (plan, t) + tuple(self.trace(plan, t))
        plan, t, pooled, fc0, logits = cache[key]
        self.refresh_folded()
        t.act.from_nchw(x.detach().float(), round_tf32=True)
        plan.run_forward()
        out = logits.act.to_nchw().reshape(n, -1)
        out_fc0 = fc0.act.to_nchw().reshape(n, -1) if fc0 is not None else None
        return out, out_fc0
