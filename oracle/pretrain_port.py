"""ORACLE (test infrastructure, never shipped on the product path): the Pretrain path.

CPU / plain-PyTorch fp32 restatement of
  * `MobileNetV2` + `SSDHead` + `InvertedResidual`          (MobileNetV2.py:10-250),
  * `MultiTaskLoss`                                         (MobileNetV2.py:342-534),
  * the Pretrain optimisation step                          (Pretrain.py:159-181, UtilityMethods.py:14-41,
                                                             config.py:29-35: SGD lr 5e-4, momentum 0.9, nesterov,
                                                             weight decay 5e-4).
It can travel to the GPU box (the reference tree cannot).  PINNED:
  * the network is bit-exact against the live reference with a shared state_dict
    (tests/test_pretrain_cpu.py::test_port_matches_live_reference) and against golden vectors recorded from it
    (tools/make_golden_pretrain.py -> tests/golden/pretrain_golden.pt),
  * the loss reproduces the reference's only known answer, Temp.py:8-29 -> 0.8939134478569031
    (test_multitask_loss_known_answer) and the live `MultiTaskLoss` on random inputs
    (test_multitask_loss_matches_live_reference).

Two things are ORACLE-DEFINED because the reference leaves them open (stated in DESIGN.md):
  1. batch > 1.  The reference loss is written for batch_size = 1 (config.py:12; `[0]` indexing at MobileNetV2.py:407,
     465,498).  The batched loss here is the reference's per-sample loss averaged over the batch.
  2. background sub-sampling.  The reference draws `torch.multinomial(background.float(), m)` (MobileNetV2.py:505):
     m background points, uniformly, without replacement.  Here the draw is an explicit input: per-point uniform keys
     `u` (B, n); the m background points with the smallest keys are kept (ties -> lower index).  For i.i.d. keys that is
     the same distribution, and it makes the step reproducible across implementations.

Distances: `torch.cdist` switches to a matmul formula above 25 points (ATen `_euclidean_dist`), whose rounding depends on
the BLAS in use.  The port computes sqrt(fl(fl(dx*dx) + fl(dy*dy))) in fp32 for every n - identical to ATen's direct
path (n <= 25, e.g. the Temp.py known answer) and within 2 ulp of the matmul path; an assignment can only differ from
the live reference at an exact near-tie of two distances (the live-reference test draws seeds where none occurs).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

# MobileNetV2.py:138-147 (t, c, n, s)
SETTING = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]
HEAD_IN = [96, 1280, 512, 256, 256, 128]      # MobileNetV2.py:28-44
HEAD_ANCHORS = [4, 6, 6, 6, 6, 6]
NUM_CLASSES = 5                                # 4 landmarks + background (MobileNetV2.py:177)
ALPHA, BETA, RATIO, RATIO_NON_BACKGROUND = 30.0, 0.1, 0.1, 5.0   # config.py:25-27, MobileNetV2.py:343
SGD = dict(lr=5e-4, momentum=0.9, nesterov=True, weight_decay=5e-4)  # config.py:31-35


class InvertedResidualPort(nn.Module):
    """MobileNetV2.py:81-120."""

    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        self.use_res_connect = stride == 1 and inp == oup
        hid = inp * expand_ratio
        self.conv = nn.Sequential(
            nn.Conv2d(inp, hid, 1, 1, 0, bias=False), nn.BatchNorm2d(hid), nn.ReLU6(inplace=True),
            nn.Conv2d(hid, hid, 3, stride, 1, groups=hid, bias=False), nn.BatchNorm2d(hid), nn.ReLU6(inplace=True),
            nn.Conv2d(hid, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup))

    def forward(self, x):
        return x + self.conv(x) if self.use_res_connect else self.conv(x)


class SSDHeadPort(nn.Module):
    """MobileNetV2.py:10-79."""

    def __init__(self, num_classes=NUM_CLASSES):
        super().__init__()
        self.num_of_out_classes, self.num_of_out_location = num_classes, 2
        self.location_layer = nn.ModuleList(nn.Conv2d(c, a * 2, 3, padding=1) for c, a in zip(HEAD_IN, HEAD_ANCHORS))
        self.classification_layer = nn.ModuleList(nn.Conv2d(c, a * num_classes, 3, padding=1)
                                                  for c, a in zip(HEAD_IN, HEAD_ANCHORS))

    def forward(self, features):
        locs, clss = [], []
        for i, x in enumerate(features):
            l = self.location_layer[i](x).permute(0, 2, 3, 1).contiguous()
            locs.append(torch.relu(l.view(l.size(0), -1, 2)))                      # :63-67
            c = self.classification_layer[i](x).permute(0, 2, 3, 1).contiguous()
            clss.append(c.view(c.size(0), -1, self.num_of_out_classes))            # :70-72
        return torch.cat(locs, 1), torch.cat(clss, 1)


class MobileNetV2Port(nn.Module):
    """MobileNetV2.py:122-250; same attribute names => same state_dict keys as the reference class."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(3, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU6(inplace=True))
        cin = 32
        self.bottlenecks = nn.ModuleList()
        for t, c, n, s in SETTING:
            for i in range(n):
                self.bottlenecks.append(InvertedResidualPort(cin, c, s if i == 0 else 1, t))
                cin = c
        self.conv2 = nn.Sequential(nn.Conv2d(320, 1280, 1, 1, 0, bias=False), nn.BatchNorm2d(1280), nn.ReLU6(inplace=True))
        self.avgpool = nn.AdaptiveAvgPool2d(1)   # constructed, never used by forward (MobileNetV2.py:174)
        self.ssd_head = SSDHeadPort(NUM_CLASSES)
        self.extra_layers = nn.ModuleList([
            nn.Conv2d(1280, 512, 1), nn.Conv2d(512, 512, 3, 2, 1), nn.Conv2d(512, 256, 1), nn.Conv2d(256, 256, 3, 2, 1),
            nn.Conv2d(256, 256, 3, 2, 1), nn.Conv2d(256, 128, 1), nn.Conv2d(128, 128, 3, 2, 1)])
        self._initialize_weights()

    def _initialize_weights(self):
        """MobileNetV2.py:219-250, same module order => same RNG consumption as the reference."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def forward(self, x, use_dropout=False):
        feats = []
        x = self.conv1(x)
        for i, b in enumerate(self.bottlenecks):
            x = b(x)
            if i == 12:
                feats.append(x)
        x = self.conv2(x)
        feats.append(x)
        for i, l in enumerate(self.extra_layers):
            x = l(x)
            if i in (1, 3, 4, 6):
                feats.append(x)
        return self.ssd_head(feats)


# ------------------------------------------------------------------------------------------------------ MultiTaskLoss
def point_distances(loc: torch.Tensor, true4: torch.Tensor) -> torch.Tensor:
    """(n,2),(4,2) -> (n,4) Euclidean distances, fp32, operation order fixed (see the module docstring)."""
    dx = loc[:, None, 0] - true4[None, :, 0]
    dy = loc[:, None, 1] - true4[None, :, 1]
    return torch.sqrt(dx * dx + dy * dy)


def assign_labels(dist: torch.Tensor, ratio: float = RATIO) -> torch.Tensor:
    """MobileNetV2.py:393-430 on one sample.  dist (n,4) -> int32 labels (n,), -1 = background.

    Per label: threshold = k-th smallest distance (k = int(ratio*n), :399-401), positives = dist <= threshold (:404); a
    point positive for several labels takes the one with the strictly smallest distance, first label on ties (:420-430)."""
    n = dist.shape[0]
    k = int(ratio * n)
    thr = dist.topk(k, dim=0, largest=False)[0].max(dim=0)[0]          # (4,)
    pos = dist <= thr[None, :]
    d = torch.where(pos, dist, torch.full_like(dist, float("inf")))
    best, lab = d.min(dim=1)    # torch.min returns the first index among equal minima == the reference's strict `<` scan
    # torch.min's tie rule on CPU is "first occurrence"; make it explicit so that it does not depend on the backend
    first = (d == best[:, None]).to(torch.int32).argmax(dim=1)
    return torch.where(torch.isinf(best), torch.full_like(first, -1), first).to(torch.int32)


def select_background(labels: torch.Tensor, u: Optional[torch.Tensor], ratio_nb: float = RATIO_NON_BACKGROUND) -> torch.Tensor:
    """MobileNetV2.py:494-507: bool mask of the background points that enter the loss.  `u` = per-point keys (see the
    module docstring); None is allowed only when no sub-sampling is needed."""
    bg = labels < 0
    n_pos = int((~bg).sum())
    m = int(n_pos * ratio_nb)
    if int(bg.sum()) <= m:
        return bg
    assert u is not None, "background sub-sampling needs the per-point keys u"
    key = torch.where(bg, u.to(torch.float32), torch.full_like(u, float("inf"), dtype=torch.float32))
    order = torch.sort(key, stable=True)[1][:m]
    sel = torch.zeros_like(bg)
    sel[order] = True
    return sel


def multitask_loss_sample(loc, cls, true8, image_size, u=None, alpha=ALPHA, beta=BETA, ratio=RATIO, labels_out=None):
    """MultiTaskLoss.forward (MobileNetV2.py:432-534) for ONE sample: loc (n,2), cls (n,5), true8 (8,), image_size (H, W)."""
    true4 = true8.view(4, 2)
    with torch.no_grad():
        labels = assign_labels(point_distances(loc.detach(), true4), ratio)
        bg_sel = select_background(labels, u)
    if labels_out is not None:
        labels_out.append(labels)
    h, w = image_size
    size = torch.tensor([w, h], dtype=loc.dtype, device=loc.device)       # :457 (width, height)
    lp = torch.clamp(loc / size, 0, 1)
    lt = torch.clamp(true4 / size, 0, 1)
    loc_loss = loc.new_zeros(())
    cls_loss = loc.new_zeros(())
    if bool(bg_sel.any()):                                                # :510-516
        idx = bg_sel.nonzero().flatten()
        cls_loss = cls_loss + F.cross_entropy(cls[idx], torch.full((idx.numel(),), 4, dtype=torch.long, device=cls.device))
    for j in range(4):
        idx = (labels == j).nonzero().flatten()
        if idx.numel():
            loc_loss = loc_loss + F.mse_loss(lp[idx], lt[j].expand(idx.numel(), 2))          # :471-481
            cls_loss = cls_loss + F.cross_entropy(cls[idx], torch.full((idx.numel(),), j, dtype=torch.long,
                                                                      device=cls.device))  # :519-528
    return alpha * loc_loss + beta * cls_loss, loc_loss, cls_loss


def multitask_loss(loc, cls, true, image_size, u=None, labels_out=None, **kw):
    """Batched loss = mean over the batch of the reference's per-sample loss (oracle-defined, see module docstring)."""
    B = loc.shape[0]
    tot = loc.new_zeros(())
    for b in range(B):
        t, _, _ = multitask_loss_sample(loc[b], cls[b], true[b].reshape(8), image_size, None if u is None else u[b],
                                        labels_out=labels_out, **kw)
        tot = tot + t
    return tot / B


# ------------------------------------------------------------------------------------------------------ training step
def sgd_nesterov_step(params: List[torch.Tensor], grads: List[torch.Tensor], bufs: List[Optional[torch.Tensor]],
                      lr=SGD["lr"], momentum=SGD["momentum"], weight_decay=SGD["weight_decay"]):
    """torch.optim.SGD(nesterov=True) update (UtilityMethods.py:30), written out: first step buf = g."""
    for i, (p, g) in enumerate(zip(params, grads)):
        g = g + weight_decay * p
        if bufs[i] is None:
            bufs[i] = g.clone()
        else:
            bufs[i].mul_(momentum).add_(g)
        p.sub_(lr * (g + momentum * bufs[i]))


def pretrain_step(model: nn.Module, images, true, u, optimizer: Optional[torch.optim.Optimizer] = None):
    """One Pretrain.py:159-181 iteration (model.train(): batch-statistics BatchNorm, running stats updated).  Returns
    (loss, labels list); gradients are left in .grad; steps `optimizer` if given."""
    model.train()
    loc, cls = model(images, use_dropout=True)
    labels: list = []
    loss = multitask_loss(loc, cls, true, (images.shape[2], images.shape[3]), u, labels_out=labels)
    for p in model.parameters():
        p.grad = None
    loss.backward()
    if optimizer is not None:
        optimizer.step()
    return loss.detach(), labels, loc.detach(), cls.detach()


def make_batch(B: int, seed: int = 1234, device="cpu"):
    """Synthetic Pretrain batch: faces ~ U(-1,1) (B,3,128,128); 4 ground-truth points = canonical landmark
    means (D_and_G_model.py:120-128: eyes, nose, mouth centre) + U(-3,3) px
    jitter; u = background sub-sampling keys."""
    g = torch.Generator().manual_seed(seed)
    images = torch.rand((B, 3, 128, 128), generator=g) * 2 - 1
    means = torch.tensor([[39.4799, 40.2799], [85.9613, 38.7062], [63.6415, 63.6473], [64.7803, 89.3250]])
    true = (means[None] + (torch.rand((B, 4, 2), generator=g) * 6 - 3)).reshape(B, 8)
    u = torch.rand((B, 394), generator=g)
    return images.to(device), true.to(device), u.to(device)


# ------------------------------------------------------------------------------------------------------ TF32 emulation
def forward_tf32_emulated(model: MobileNetV2Port, x: torch.Tensor):
    """Forward of the port with the PRODUCTION path's rounding points restated (tpgan_b200/MobileNetV2.py): operands of
    every dense (tensor-core) convolution - stored activations feeding one, and its weights - are rounded to tf32
    (cvt.rna), products are accumulated in fp32; depthwise convolutions, BatchNorm, ReLU6 and the residual add are fp32.
    Lets the TF32 path be checked sharply although the network itself amplifies any perturbation ~100x from the stem
    to the heads at random initialisation (measured in tests/test_pretrain_gpu.py).  Forward only."""
    from .model_port import tf32_rna as q

    def dense(conv, h):          # h is already rounded where the product rounds it
        return F.conv2d(h, q(conv.weight), conv.bias, conv.stride, conv.padding)

    def bn(m, h):
        return F.batch_norm(h, m.running_mean, m.running_var, m.weight, m.bias, m.training, m.momentum, m.eps)

    with torch.no_grad():
        h = q(x)
        h = q(F.relu6(bn(model.conv1[1], dense(model.conv1[0], h))))
        feats = []
        for i, b in enumerate(model.bottlenecks):
            c = b.conv
            t = F.relu6(bn(c[1], dense(c[0], h)))                      # feeds the depthwise conv: not rounded
            t = F.conv2d(t, c[3].weight, None, c[3].stride, c[3].padding, 1, c[3].groups)
            t = q(F.relu6(bn(c[4], t)))
            t = bn(c[7], dense(c[6], t))
            h = q(t + h) if b.use_res_connect else q(t)
            if i == 12:
                feats.append(h)
        h = q(F.relu6(bn(model.conv2[1], dense(model.conv2[0], h))))
        feats.append(h)
        for i, l in enumerate(model.extra_layers):
            h = q(dense(l, h))
            if i in (1, 3, 4, 6):
                feats.append(h)
        locs, clss = [], []
        for i, f in enumerate(feats):
            l = dense(model.ssd_head.location_layer[i], f).permute(0, 2, 3, 1).contiguous()
            locs.append(torch.relu(l.view(l.size(0), -1, 2)))
            c = dense(model.ssd_head.classification_layer[i], f).permute(0, 2, 3, 1).contiguous()
            clss.append(c.view(c.size(0), -1, NUM_CLASSES))
        return torch.cat(locs, 1), torch.cat(clss, 1)


# ------------------------------------------------------------------------------------------------------ decoder / accuracy
def decode_sample(loc: torch.Tensor, cls: torch.Tensor, confidence_threshold=0.5, top_k=1, nms_distance_threshold=20.0):
    """MultiTaskDecoder.forward + nms (MobileNetV2.py:552-649) for one sample, as fixed-shape tensors:
    count (K,), score (K, top_k), point (K, top_k, 2).  Greedy distance-NMS in descending score order followed by top_k
    == top_k rounds of { arg-max among live candidates; drop everything within the distance threshold }."""
    K = cls.shape[1]
    scores = torch.softmax(cls, dim=-1)
    count = torch.zeros(K, dtype=torch.int32)
    score = torch.zeros(K, top_k)
    point = torch.zeros(K, top_k, 2)
    for c in range(K):
        s = torch.where(scores[:, c] > confidence_threshold, scores[:, c], torch.full_like(scores[:, c], -1.0))
        for t in range(top_k):
            best = s.max()
            if float(best) <= 0:
                break
            bi = int((s == best).nonzero()[0])            # lowest index among equal scores
            score[c, t], point[c, t] = best, loc[bi]
            count[c] += 1
            d = torch.norm(loc - loc[bi], dim=1)
            s = torch.where(d > nms_distance_threshold, s, torch.full_like(s, -1.0))
            s[bi] = -1.0
    return count, score, point


def accuracy_sample(count, point, true8) -> float:
    """_calculate_accuracy (Pretrain.py:17-64) on the top-1 detections of the four landmark classes.  The reference indexes
    `predicts[:-1]` and therefore only works when every class (incl. background) has exactly one detection; for that case
    this is the same number.  ORACLE-DEFINED otherwise: a landmark class without detection contributes 0."""
    thresholds, weights = [5, 10, 18, 30, 45], [1.0, 0.9, 0.65, 0.35, 0.1]
    gt = true8.view(4, 2)
    acc = 0.0
    for c in range(4):
        if int(count[c]) == 0:
            continue
        d = float(torch.sqrt(((point[c, 0] - gt[c]) ** 2).sum()))
        prev = 0
        for th, w in zip(thresholds, weights):
            if prev < d <= th:
                acc += w
            prev = th
    return acc / 4.0
