"""ORACLE (test infrastructure): fp32 PyTorch restatement of the identity / feature network of the TP-GAN step - a
ResNet18 for 128x128 inputs in eval mode - and of the identity-preserving loss built on it.

PARITY UNPINNED BY THE REFERENCE: /root/reference/ResNet.py and FeatureExtract.py cannot be constructed (SURVEY.md 2.3:
`conv()` is called with a `bias` keyword it does not have (ResNet.py:31), the block factory receives the stride in the
kernel-size slot (:77), only three of the four stages are built and all with stride 1 (:38-40), `resnet18()` is
mis-indented (:121-126); FeatureExtract.py:31 reads `.in_features` of an nn.Sequential), and the reference ships no
training step.  What is restated here is therefore what those files DESCRIBE:

  ResNet.py:30-31   conv1 = 7x7, stride 2, pad 3, 3 -> 64, BatchNorm, ReLU
  ResNet.py:33      MaxPool2d(3, 2, 1)
  ResNet.py:28-29   stages [64, 128, 256, 512] x [2, 2, 2, 2] two-conv residual blocks (canonical ResNet18 wiring:
                    stage strides 1, 2, 2, 2; 1x1 stride-s projection shortcut + BN where the shape changes)
  ResNet.py:45      AdaptiveAvgPool2d((1, 1))  -> pooled feature (B, 512)
  ResNet.py:48-49   FC0 = Linear(512, feature_layer_dim_before_FC) + BatchNorm1d -> out_FC0
  ResNet.py:55      FC  = Linear(., num_of_output_classes)
  ResNet.py:80-119  forward(x) -> (out, out_FC0)
  config.py:79      loss['weight_identity_preserving'] = 30: L_ip = sum over the two last feature layers (pooled, FC0) of
                    mean |F_i(gt) - F_i(fake)|   (TP-GAN paper eq. 5, cited at D_and_G_model.py:2)

It is driven by a state_dict with the key names of tpgan_b200.ResNet.ResNet18 (reference-style factories: every conv /
linear is an nn.Sequential [Conv2d|Linear, BatchNorm?, ReLU?]).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import this file.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.BatchNorm default


def _bn(sd: Dict[str, torch.Tensor], key: str, x: torch.Tensor, training: bool = False) -> torch.Tensor:
    """BatchNorm if the layer has one: running statistics (eval) or batch statistics with the running ones updated in
    place in `sd` (training, momentum 0.1 - nn.BatchNorm's defaults)."""
    if key + ".weight" not in sd:
        return x
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"],
                        training=training, momentum=0.1, eps=EPS)


def _cbr(sd, key, x, stride, pad, relu, training=False):
    y = F.conv2d(x, sd[key + ".0.weight"], sd.get(key + ".0.bias"), stride=stride, padding=pad)
    y = _bn(sd, key + ".1", y, training)
    return F.relu(y) if relu else y


def _block(sd, key, x, stride, training=False):
    a = _cbr(sd, key + ".conv_a", x, stride, 1, True, training)
    b = _cbr(sd, key + ".conv_b", a, 1, 1, False, training)
    sc = _cbr(sd, key + ".shortcut", x, stride, 0, False, training) if (key + ".shortcut.0.weight") in sd else x
    return F.relu(b + sc)


def resnet18_128(sd: Dict[str, torch.Tensor], x: torch.Tensor,
                 training: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """x (B,3,128,128) -> (logits, FC0 feature or None, pooled 512 feature).  training=True: the pre-training forward
    (batch-statistics BatchNorm; `sd`'s running statistics are updated in place)."""
    h = _cbr(sd, "conv1", x, 2, 3, True, training)
    h = F.max_pool2d(h, 3, 2, 1)
    for s, stride in enumerate((1, 2, 2, 2)):
        for b in range(2):
            h = _block(sd, f"sections.{s}.{b}", h, stride if b == 0 else 1, training)
    pooled = F.adaptive_avg_pool2d(h, 1).flatten(1)
    fc0 = None
    f = pooled
    if "FC0.0.weight" in sd:
        f = F.linear(pooled, sd["FC0.0.weight"], sd.get("FC0.0.bias"))
        f = _bn(sd, "FC0.1", f, training)
        fc0 = f
    logits = F.linear(f, sd["FC.0.weight"], sd.get("FC.0.bias"))
    return logits, fc0, pooled


def identity_loss(sd: Dict[str, torch.Tensor], fake: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """L_ip = mean|pooled(gt) - pooled(fake)| + mean|FC0(gt) - FC0(fake)|; the network is frozen, gt carries no gradient."""
    _, f0, p0 = resnet18_128(sd, fake)
    with torch.no_grad():
        _, f1, p1 = resnet18_128(sd, gt)
    loss = (p0 - p1).abs().mean()
    if f0 is not None:
        loss = loss + (f0 - f1).abs().mean()
    return loss
