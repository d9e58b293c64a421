"""Drop-in for the reference's MobileNetV2.py: `SSDHead`, `InvertedResidual`, `MobileNetV2`,
`MultiTaskLoss` with the reference's constructors, attribute names, state_dict keys (356 tensors), initialisation order
(same RNG consumption => seeded weights are bit-identical) and forward signatures/return values:

    MobileNetV2().forward(x, use_dropout=False) -> (locations (B, 394, 2), classifications (B, 394, 5))   MobileNetV2.py:180-217
    MultiTaskLoss(alpha, beta, distance_threshold_ratio).forward(locations_pred, classifications_pred, locations_true,
                                                                 image_size) -> scalar loss               MobileNetV2.py:432-534

Execution is a traced launch plan over the C ABI (include/tpgan_b200.h):
  * 1x1 expand / project convolutions, the 3x3 stem, the extra layers and the 12 SSD heads: tcgen05 implicit-GEMM kernels
    (tpgan_conv2d / tpgan_conv2d_wgrad), heads with bias (+ReLU for the locations, MobileNetV2.py:67) fused in the epilogue;
  * depthwise 3x3 convolutions: tpgan_dwconv3x3{,_dgrad,_wgrad} (HBM-bound CUDA-core kernels on the reference weight layout);
  * BatchNorm2d in training mode (batch statistics, running statistics updated) or eval mode, with ReLU6 and the
    InvertedResidual skip-add fused: tpgan_bn_forward / tpgan_bn_backward;
  * permute/view/cat of SSDHead.forward: NHWC is the permuted layout already; tpgan_rows_gather lays the heads end to end;
  * MultiTaskLoss: ONE launch for the whole batch (tpgan_multitask_loss), no host synchronisation - the reference's
    Python loops with .item() calls per point (MobileNetV2.py:393-430) are the host-bound part of Pretrain.py.
There is no CPU fallback: CPU tensors raise.

Differences that are deliberate: the loss accepts batch > 1 (mean over samples of the reference's per-sample
loss) and takes the background sub-sampling draw as explicit keys `u` (drawn on the device when omitted) instead of
torch.multinomial; it does not print.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from .D_and_G_model import TracedModule, _layer
from .engine import Plan, T
from .ops import Act


# ---------------------------------------------------------------------------------------------------- parameter holders
class _BNState:
    def __init__(self, c: int, device):
        self.sums = torch.zeros(2 * c + 1, dtype=torch.float64, device=device)     # + block ticket; kernels leave it zeroed
        self.dsums = torch.zeros(2 * c + 1, dtype=torch.float64, device=device)
        self.coef = torch.zeros(4 * c, dtype=torch.float32, device=device)


def _accumulate(dst: torch.Tensor, src: torch.Tensor):
    """dst += src through the library (no ATen arithmetic on the path)."""
    n = dst.numel()
    ops.view_copy(Act(src.view(1, 1, 1, n)), Act(dst.view(1, 1, 1, n)), True)


class _AuxLayer:
    """Gradient plumbing shared by BNLayer / DepthwiseLayer: kernels write into private buffers (autograd path: accumulated
    into param.grad afterwards) or straight into param.grad (plan.direct_grads: the trainer's flat buffers)."""

    def _params(self) -> List[torch.nn.Parameter]:
        raise NotImplementedError

    def _target(self, plan: Plan, p: torch.nn.Parameter, slot: str) -> torch.Tensor:
        if plan.direct_grads:
            assert p.grad is not None, "direct_grads needs pre-allocated .grad (FlatParams)"
            return p.grad
        buf = self.__dict__.get(slot)
        if buf is None:
            buf = torch.zeros_like(p.data)
            self.__dict__[slot] = buf
        return buf

    def zero_grad(self):
        for slot in ("_gw", "_gb"):
            b = self.__dict__.get(slot)
            if b is not None:
                b.zero_()

    def refresh(self):
        pass

    def export_grad_autograd(self):
        for p, slot in zip(self._params(), ("_gw", "_gb")):
            b = self.__dict__.get(slot)
            if b is None or not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            _accumulate(p.grad, b)


class BNLayer(_AuxLayer):
    """nn.BatchNorm2d parameters/buffers + per-plan scratch (batch sums, folded coefficients)."""

    def __init__(self, module: nn.BatchNorm2d, name: str = ""):
        self.module, self.name = module, name
        self._states = {}

    def _params(self):
        return [self.module.weight, self.module.bias]

    def state(self, plan: Plan, c: int) -> _BNState:
        st = self._states.get(id(plan))
        if st is None:
            st = _BNState(c, plan.device)
            self._states[id(plan)] = st
            plan.keep.append(st)
        return st

    def grad_targets(self, plan: Plan):
        return self._target(plan, self.module.weight, "_gw"), self._target(plan, self.module.bias, "_gb")


class DepthwiseLayer(_AuxLayer):
    def __init__(self, module: nn.Conv2d, name: str = ""):
        assert module.groups == module.in_channels == module.out_channels and module.kernel_size == (3, 3) and \
            module.padding == (1, 1) and module.bias is None and module.stride[0] in (1, 2)
        self.module, self.name = module, name
        self.weight = module.weight
        self.stride = module.stride[0]

    def _params(self):
        return [self.weight]

    def grad_target(self, plan: Plan):
        return self._target(plan, self.weight, "_gw")


def _layer_at(conv_mod: nn.Conv2d, name: str, x: T):
    """ConvLayer of a dense conv for input x.  A 3x3 / stride-2 / pad-1 conv applied to a 1x1 map (extra_layers 4 and 6,
    MobileNetV2.py:170-171) touches only its centre tap and yields a 1x1 map - the same arithmetic as stride 1, which is
    what the tensor-core kernel is given (its stride-2 parity planes need even extents)."""
    if conv_mod.stride[0] == 2 and (x.act.h == 1) != (x.act.w == 1):
        raise NotImplementedError(f"{name}: a stride-2 conv over a {x.act.h}x{x.act.w} map (extent 1 in one dimension only) is not "
                                  "built; square power-of-two inputs (64, 128, 256) keep both extents equal down to 1x1")
    if conv_mod.stride[0] == 2 and x.act.h == 1 and x.act.w == 1:
        L = conv_mod.__dict__.get("_tc_layer_1x1")
        if L is None:
            from .engine import ConvLayer
            L = ConvLayer(conv_mod.weight, conv_mod.bias, False, conv_mod.kernel_size[0], 1, conv_mod.padding[0], name)
            object.__setattr__(conv_mod, "_tc_layer_1x1", L)
        return L
    return _layer(conv_mod, name)


def _aux(module, cls, name):
    a = module.__dict__.get("_tc_aux")
    if a is None:
        a = cls(module, name)
        object.__setattr__(module, "_tc_aux", a)
    return a


# ---------------------------------------------------------------------------------------------------- modules
class SSDHead(nn.Module):
    """MobileNetV2.py:10-79 (same layer lists; executed by MobileNetV2's plan)."""

    def __init__(self, num_of_out_classes=4):
        super().__init__()
        self.num_of_out_classes = num_of_out_classes
        self.num_of_out_location = 2
        self.location_layer = nn.ModuleList()
        self.classification_layer = nn.ModuleList()
        for cin, anchors in zip((96, 1280, 512, 256, 256, 128), (4, 6, 6, 6, 6, 6)):      # MobileNetV2.py:28-44
            self.location_layer += [nn.Conv2d(cin, anchors * self.num_of_out_location, kernel_size=3, padding=1)]
            self.classification_layer += [nn.Conv2d(cin, anchors * self.num_of_out_classes, kernel_size=3, padding=1)]

    def trace(self, plan: Plan, features: List[T], pre: str = "ssd_head"):
        locs, clss = [], []
        for i, f in enumerate(features):
            # ReLU on the locations (MobileNetV2.py:67) fused into the conv epilogue
            locs.append(plan.conv([_layer(self.location_layer[i], f"{pre}.location_layer.{i}")], [f], 0.0,
                                  round_out=False)[0])
            clss.append(plan.conv([_layer(self.classification_layer[i], f"{pre}.classification_layer.{i}")], [f], None,
                                  round_out=False)[0])
        return plan.gather_rows(locs, "locations"), plan.gather_rows(clss, "classifications")

    def forward(self, features):
        raise RuntimeError("SSDHead runs inside MobileNetV2's traced plan; call the network")


class InvertedResidual(nn.Module):
    """MobileNetV2.py:81-120: 1x1 expand -> BN -> ReLU6 -> 3x3 depthwise -> BN -> ReLU6 -> 1x1 project -> BN (+x)."""

    def __init__(self, inp, oup, stride=1, expand_ratio=6):
        super().__init__()
        self.stride = stride
        self.use_res_connect = self.stride == 1 and inp == oup
        hid = inp * expand_ratio
        self.conv = nn.Sequential(
            nn.Conv2d(inp, hid, 1, 1, 0, bias=False), nn.BatchNorm2d(hid), nn.ReLU6(inplace=True),
            nn.Conv2d(hid, hid, 3, stride, 1, groups=hid, bias=False), nn.BatchNorm2d(hid), nn.ReLU6(inplace=True),
            nn.Conv2d(hid, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup))

    def trace(self, plan: Plan, x: T, pre: str) -> T:
        c = self.conv
        h = plan.conv([_layer(c[0], f"{pre}.conv.0")], [x], None, round_out=False)[0]       # BatchNorm reads full fp32
        h = plan.batchnorm(_aux(c[1], BNLayer, f"{pre}.conv.1"), h, relu6=True, round_out=False)      # feeds the depthwise
        h = plan.dwconv(_aux(c[3], DepthwiseLayer, f"{pre}.conv.3"), h)
        h = plan.batchnorm(_aux(c[4], BNLayer, f"{pre}.conv.4"), h, relu6=True, round_dx=False)
        h = plan.conv([_layer(c[6], f"{pre}.conv.6")], [h], None, round_out=False)[0]
        return plan.batchnorm(_aux(c[7], BNLayer, f"{pre}.conv.7"), h, res=x if self.use_res_connect else None)

    def forward(self, x):
        raise RuntimeError("InvertedResidual runs inside MobileNetV2's traced plan; call the network")


class MobileNetV2(TracedModule):
    """MobileNetV2.py:122-250.

    Arithmetic: the dense convolutions of this network run in the fp32-accurate 3xTF32 split mode by default
    (a*w = a_hi*w_hi + a_hi*w_lo + a_lo*w_hi, three tensor-core launches per product, nothing rounded at store).  At batch
    32 every tensor-core launch of this 0.27 GFLOP/image network sits at its ~12 us launch floor, so single-pass TF32 buys no
    throughput per FLOP, while training-mode BatchNorm at random initialisation amplifies operand rounding ~100x from the
    stem to the heads (3.7e-2 against the fp32 reference); the split mode matches the reference to 2e-5."""
    _exact_default = True

    def __init__(self):
        super().__init__()
        self.interverted_residual_setting = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1],
                                             [6, 160, 3, 2], [6, 320, 1, 1]]
        self.conv1 = nn.Sequential(nn.Conv2d(3, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU6(inplace=True))
        input_channel = 32
        self.bottlenecks = nn.ModuleList()
        for t, c, n, s in self.interverted_residual_setting:
            for idx in range(n):
                self.bottlenecks.append(InvertedResidual(input_channel, c, s if idx == 0 else 1, t))
                input_channel = c
        self.conv2 = nn.Sequential(nn.Conv2d(320, 1280, 1, 1, 0, bias=False), nn.BatchNorm2d(1280), nn.ReLU6(inplace=True))
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.ssd_head = SSDHead(4 + 1)
        self.extra_layers = nn.ModuleList([
            nn.Conv2d(1280, 512, kernel_size=1), nn.Conv2d(512, 512, kernel_size=3, stride=2, padding=1),
            nn.Conv2d(512, 256, kernel_size=1), nn.Conv2d(256, 256, kernel_size=3, stride=2, padding=1),
            nn.Conv2d(256, 256, kernel_size=3, stride=2, padding=1), nn.Conv2d(256, 128, kernel_size=1),
            nn.Conv2d(128, 128, kernel_size=3, stride=2, padding=1)])
        self._initialize_weights()

    def _initialize_weights(self):
        """MobileNetV2.py:219-250."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    # ---- traced execution
    def trace(self, plan: Plan, x: T):
        """x: (N,H,W,3) NHWC -> (locations flat T (N,1,1,n*2), classifications flat T (N,1,1,n*5))."""
        feats = []
        h = plan.conv([_layer(self.conv1[0], "conv1.0")], [x], None, round_out=False)[0]
        h = plan.batchnorm(_aux(self.conv1[1], BNLayer, "conv1.1"), h, relu6=True)
        for idx, b in enumerate(self.bottlenecks):
            h = b.trace(plan, h, f"bottlenecks.{idx}")
            if idx == 12:                                   # MobileNetV2.py:192-193
                feats.append(h)
        h = plan.conv([_layer(self.conv2[0], "conv2.0")], [h], None, round_out=False)[0]
        h = plan.batchnorm(_aux(self.conv2[1], BNLayer, "conv2.1"), h, relu6=True)
        feats.append(h)
        for idx, l in enumerate(self.extra_layers):          # plain convs with bias, no activation (:202-206)
            h = plan.conv([_layer_at(l, f"extra_layers.{idx}", h)], [h], None)[0]
            if idx in (1, 3, 4, 6):
                feats.append(h)
        return self.ssd_head.trace(plan, feats)

    def _trace(self, plan, x, static=()):
        plan.bn_training = bool(static[0]) if static else self.training   # batch statistics <=> module.train()
        return self.trace(plan, x)

    def forward(self, x, use_dropout=False):
        # use_dropout is accepted and ignored, as in the reference (MobileNetV2.py:180: the flag is never read)
        loc, cls = self._traced_call([x], static=(self.training,))     # (bumps num_batches_tracked in train mode)
        n = x.shape[0]
        return loc.reshape(n, -1, 2), cls.reshape(n, -1, self.ssd_head.num_of_out_classes)

    @staticmethod
    def num_points(h: int = 128, w: int = 128) -> int:
        s = lambda v, k: [v := (v + 1) // 2 for _ in range(k)][-1]
        f = [(s(h, 4), s(w, 4)), (s(h, 5), s(w, 5)), (s(h, 6), s(w, 6)), (s(h, 7), s(w, 7)), (s(h, 8), s(w, 8)),
             (s(h, 9), s(w, 9))]
        return sum(a * b * k for (a, b), k in zip(f, (4, 6, 6, 6, 6, 6)))


# ---------------------------------------------------------------------------------------------------- loss
class _MTLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc, cls, truth, u, hw, alpha, beta, ratio, ratio_nb, holder):
        B, n = loc.shape[0], loc.shape[1]
        locc, clsc = loc.detach().float().contiguous(), cls.detach().float().contiguous()
        dloc, dcls = torch.empty_like(locc), torch.empty_like(clsc)
        labels = torch.empty((B, n), dtype=torch.int32, device=loc.device)
        sums = torch.zeros(3, dtype=torch.float32, device=loc.device)
        ops.multitask_loss(locc, clsc, truth.detach().float().contiguous().view(B, 8), u, n, 2 * n, 5 * n,
                           int(ratio * n), float(hw[1]), float(hw[0]), alpha, beta, ratio_nb, 1.0 / B, dloc, dcls, labels, sums)
        ctx.save_for_backward(dloc, dcls)
        holder.labels, holder.parts = labels, sums
        return sums[0].clone()

    @staticmethod
    def backward(ctx, g):
        dloc, dcls = ctx.saved_tensors
        return dloc * g, dcls * g, None, None, None, None, None, None, None, None


class MultiTaskLoss(nn.Module):
    """MobileNetV2.py:342-534 on the device, batched (see the module docstring).  After a call `.labels` (B, n) int32 holds
    the assignment (-1 = background) and `.parts` = (total, location, classification) device scalars."""

    def __init__(self, alpha=30.0, beta=0.1, distance_threshold_ratio=0.1, ratio_non_background=5.0):
        super().__init__()   # defaults = config.py:25-27 (pretrain['loss'])
        self.alpha, self.beta = alpha, beta
        self.distance_threshold_ratio = distance_threshold_ratio
        self.ratio_non_background = ratio_non_background
        self.labels = self.parts = None

    def forward(self, locations_pred, classifications_pred, locations_true, image_size, u: Optional[torch.Tensor] = None):
        if not locations_pred.is_cuda:
            raise RuntimeError("tpgan_b200.MultiTaskLoss runs on CUDA tensors only (there is no CPU fallback)")
        B, n = locations_pred.shape[0], locations_pred.shape[1]
        if u is None:   # the reference's torch.multinomial draw (MobileNetV2.py:505), as per-point keys
            u = torch.rand((B, n), dtype=torch.float32, device=locations_pred.device)
        return _MTLFn.apply(locations_pred, classifications_pred, locations_true, u.float().contiguous(), tuple(image_size),
                            float(self.alpha), float(self.beta), float(self.distance_threshold_ratio),
                            float(self.ratio_non_background), self)


# ---------------------------------------------------------------------------------------------------- decoder
class MultiTaskDecoder(nn.Module):
    """MobileNetV2.py:536-649 on the device: ONE launch decodes the whole batch (softmax, confidence filter, greedy
    distance-NMS, top-k per class) instead of a Python loop with a sort and a host sync per kept point.

    forward(locations, classifications) returns the reference's structure - per sample a list of
    (class_idx, score tensor, point tensor (2,)) in class order, best first - built from the fixed-shape device outputs with a
    single device-to-host read.  decode(...) returns those device tensors without synchronising:
    count (B,K) int32, score (B,K,top_k), point (B,K,top_k,2) and, given the ground truth (B,8), accuracy (B,) =
    `_calculate_accuracy` of Pretrain.py:17-64 on the top-1 landmark detections."""

    def __init__(self, confidence_threshold=0.5, top_k=1, nms_distance_threshold=20):
        super().__init__()
        self.confidence_threshold = confidence_threshold
        self.top_k = top_k
        self.nms_distance_threshold = nms_distance_threshold

    def decode(self, locations, classifications, locations_true: Optional[torch.Tensor] = None):
        if not locations.is_cuda:
            raise RuntimeError("tpgan_b200.MultiTaskDecoder runs on CUDA tensors only (there is no CPU fallback)")
        B, n, K = classifications.shape
        loc, cls = locations.detach().float().contiguous(), classifications.detach().float().contiguous()
        dev = loc.device
        count = torch.zeros((B, K), dtype=torch.int32, device=dev)
        score = torch.empty((B, K, self.top_k), dtype=torch.float32, device=dev)
        point = torch.empty((B, K, self.top_k, 2), dtype=torch.float32, device=dev)
        truth = acc = None
        if locations_true is not None:
            truth = locations_true.detach().float().contiguous().view(B, 8)
            acc = torch.empty(B, dtype=torch.float32, device=dev)
        ops.ssd_decode(loc, cls, n, 2 * n, K * n, K, int(self.top_k), float(self.confidence_threshold),
                       float(self.nms_distance_threshold), count, score, point, truth, acc)
        return count, score, point, acc

    def forward(self, locations, classifications):
        count, score, point, _ = self.decode(locations, classifications)
        cnt = count.cpu()                                   # the one host read
        out = []
        for b in range(cnt.shape[0]):
            res = []
            for c in range(cnt.shape[1]):
                for t in range(int(cnt[b, c])):
                    res.append((c, score[b, c, t], point[b, c, t]))
            out.append(res)
        return out
