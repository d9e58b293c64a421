"""Synthetic training batches with the TrainDataset conventions (reference DataAndDataset.py:206-226): images
U(-1, 1), the five landmarks = the canonical means of the LocalFuser docstring (D_and_G_model.py:120-128) + U(-3, 3) px
jitter, the 64 / 32 targets = average pools of the frontal image, z ~ U(-1, 1), labels, WGAN-GP interpolation coefficients.
The landmark patches are NOT part of the batch: the trainer crops them on the device (tpgan_patch_crop).

Same generator and draw order as the oracle's make_batch (tests assert the shared keys are bit-identical), so benchmark
inputs and parity-test inputs are the same numbers; this module is product code and imports nothing from oracle/."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import config

# (x, y) of left eye, right eye, nose, mouth-left, mouth-right on the 128x128 face (D_and_G_model.py:120-128)
MEAN_LANDMARKS = torch.tensor([[39.4799, 40.2799], [85.9613, 38.7062], [63.6415, 63.6473], [45.6705, 89.9648],
                               [83.9000, 88.6898]], dtype=torch.float32)
KEYS = ("img", "img_frontal", "img64_frontal", "img32_frontal", "landmarks", "z", "label", "gp_alpha")


def make_batch(B: int, seed: int = 1234) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g) * 2 - 1
    img, frontal = u(B, 3, 128, 128), u(B, 3, 128, 128)
    lm = MEAN_LANDMARKS[None] + 3.0 * u(B, 5, 2)
    z = u(B, config.G["zdim"])
    label = torch.randint(0, config.G["num_classes"], (B,), generator=g)
    alpha = torch.rand(B, generator=g)
    return dict(img=img, img_frontal=frontal, img64_frontal=F.avg_pool2d(frontal, 2), img32_frontal=F.avg_pool2d(frontal, 4),
                landmarks=lm.float(), z=z, label=label, gp_alpha=alpha)
