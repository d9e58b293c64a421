"""ORACLE (test infrastructure): the G+D training step in plain fp32 PyTorch.

The reference ships NO training loop and NO loss code - only the loss weights (config.py:71-82), the optimiser factory
(UtilityMethods.py:14-41), set_requires_grad (:43-56), the TrainDataset batch keys (DataAndDataset.py:206-226) and the
paper pointer (D_and_G_model.py:2).  The step below is therefore ORACLE-DEFINED: "parity unpinned" by
any reference test; it is the specification the CUDA path (tpgan_b200/train_step.py) is checked against.

  pixel    = w128*L1(fake, gt128) + w64*L1(avgpool2(fake), gt64) + w32*L1(avgpool4(fake), gt32)
  symmetry = same three scales/weights of L1(t, hflip(t))
  tv       = mean|d/dy fake| + mean|d/dx fake|
  local    = sum over the 4 parts of L1(part_fake, part_gt)
  adv      : WGAN-GP critic.  L_D = mean D(fake.detach()) - mean D(real) + 10 * mean((||grad_xhat sum D(xhat)||_2 - 1)^2)
             L_G_adv = -mean D(fake)
  ce       = cross_entropy(encoder_predict, label)
  L_G = 1.0*pixel + 3.0*local + 0.3*symmetry + 1e-3*adv + 1e-3*tv + 10*ce            (identity term: not in this round)
  Optimiser: Adam(lr=train['learning_rate']=1e-4), torch defaults otherwise (the reference leaves the type unpinned).
"""
from __future__ import annotations

import math
from typing import Callable, Dict

import numpy as np
import torch
import torch.nn.functional as F

# config.py:71-82 (kept literally; tests assert they equal the reference's config when it is present)
LOSS_W = dict(weight_gradient_penalty=10, weight_128=1.0, weight_64=1.0, weight_32=1.5, weight_pixelwise=1.0,
              weight_pixelwise_local=3.0, weight_symmetry=3e-1, weight_adv_G=1e-3, weight_identity_preserving=3e1,
              weight_total_varation=1e-3, weight_cross_entropy=1e1)
LEARNING_RATE = 1e-4   # config.py:52
ZDIM, NUM_CLASSES = 64, 347  # config.py:61,64

# mean landmark table of the LocalFuser docstring, D_and_G_model.py:120-128: (x, y) of LE, RE, nose, mouth-left, mouth-right
MEAN_LANDMARKS = np.array([[39.4799, 40.2799], [85.9613, 38.7062], [63.6415, 63.6473], [45.6705, 89.9648],
                           [83.9000, 88.6898]], dtype=np.float32)
PATCH_WH = ((40, 40), (40, 40), (40, 32), (48, 32))  # DataAndDataset.py:35-40 (w, h)


def crop_boxes(landmarks: np.ndarray) -> np.ndarray:
    """process() index arithmetic, DataAndDataset.py:42-54.  landmarks (N,5,2) float32 -> int boxes (N,4,4)
    (left, upper, right, lower).  The mouth centre is the float32 mean of the two mouth corners."""
    lm = np.asarray(landmarks, dtype=np.float32).copy()
    lm[:, 3, 0] = (lm[:, 3, 0] + lm[:, 4, 0]) / np.float32(2.0)
    lm[:, 3, 1] = (lm[:, 3, 1] + lm[:, 4, 1]) / np.float32(2.0)
    boxes = np.zeros((lm.shape[0], 4, 4), dtype=np.int32)
    for i, (w, h) in enumerate(PATCH_WH):
        x = np.floor(lm[:, i, 0]).astype(np.int32)
        y = np.floor(lm[:, i, 1]).astype(np.int32)
        boxes[:, i] = np.stack([x - w // 2 + 1, y - h // 2 + 1, x + w // 2 + 1, y + h // 2 + 1], 1)
    return boxes


def crop_patches(img: torch.Tensor, landmarks: np.ndarray, fill: float = -1.0):
    """PIL.Image.crop semantics on a normalised tensor: out-of-image pixels are PIL's 0, i.e. -1 after *2-1
    (DataAndDataset.py:51-54, 254-255).  img (N,C,128,128) -> 4 tensors."""
    boxes = crop_boxes(landmarks)
    N, C, H, W = img.shape
    outs = []
    for i, (w, h) in enumerate(PATCH_WH):
        out = torch.full((N, C, h, w), fill, dtype=img.dtype)
        for n in range(N):
            l, u = int(boxes[n, i, 0]), int(boxes[n, i, 1])
            x0, x1, y0, y1 = max(l, 0), min(l + w, W), max(u, 0), min(u + h, H)
            if x1 > x0 and y1 > y0:
                out[n, :, y0 - u:y1 - u, x0 - l:x1 - l] = img[n, :, y0:y1, x0:x1]
        outs.append(out)
    return outs


def make_batch(B: int, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Synthetic batch with the TrainDataset conventions (DataAndDataset.py:206-226): images U(-1,1),
    landmarks = canonical means + U(-3,3) px, patches cropped from the images, 64/32 targets = average pools."""
    g = torch.Generator().manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g) * 2 - 1
    img, frontal = u(B, 3, 128, 128), u(B, 3, 128, 128)
    lm = torch.from_numpy(MEAN_LANDMARKS)[None] + 3.0 * u(B, 5, 2)
    z = u(B, ZDIM)
    label = torch.randint(0, NUM_CLASSES, (B,), generator=g)
    alpha = torch.rand(B, generator=g)
    lm_np = lm.numpy().astype(np.float32)
    parts = crop_patches(img, lm_np)
    parts_f = crop_patches(frontal, lm_np)
    names = ("left_eye", "right_eye", "nose", "mouth")
    b = dict(img=img, img_frontal=frontal, img64_frontal=F.avg_pool2d(frontal, 2), img32_frontal=F.avg_pool2d(frontal, 4),
             landmarks=lm.float(), z=z, label=label, gp_alpha=alpha)
    for n, p, pf in zip(names, parts, parts_f):
        b[n] = p
        b[n + "_frontal"] = pf
    return b


def image_terms(fake, b):
    w = LOSS_W
    f64, f32 = F.avg_pool2d(fake, 2), F.avg_pool2d(fake, 4)
    l1 = lambda a, c: (a - c).abs().mean()
    pixel = w["weight_128"] * l1(fake, b["img_frontal"]) + w["weight_64"] * l1(f64, b["img64_frontal"]) + \
        w["weight_32"] * l1(f32, b["img32_frontal"])
    sym = w["weight_128"] * l1(fake, fake.flip(3)) + w["weight_64"] * l1(f64, f64.flip(3)) + \
        w["weight_32"] * l1(f32, f32.flip(3))
    tv = (fake[:, :, 1:, :] - fake[:, :, :-1, :]).abs().mean() + (fake[:, :, :, 1:] - fake[:, :, :, :-1]).abs().mean()
    return pixel, sym, tv


def g_loss(g_out, d_fake, b, identity_sd=None):
    """g_out = the Generator 8-tuple; d_fake = D(fake) patch logits; identity_sd = state_dict of the frozen identity
    network (oracle/identity_port.py) or None (the reference's own FeatureExtract/ResNet cannot be built: SURVEY 2.3)."""
    w = LOSS_W
    fake, logits, _, le, re, nose, mouth, _ = g_out
    pixel, sym, tv = image_terms(fake, b)
    local = sum((p - b[n + "_frontal"]).abs().mean() for p, n in zip((le, re, nose, mouth),
                                                                     ("left_eye", "right_eye", "nose", "mouth")))
    adv = -d_fake.mean()
    ce = F.cross_entropy(logits, b["label"])
    total = w["weight_pixelwise"] * pixel + w["weight_pixelwise_local"] * local + w["weight_symmetry"] * sym + \
        w["weight_adv_G"] * adv + w["weight_total_varation"] * tv + w["weight_cross_entropy"] * ce
    m = dict(pixel=pixel, local=local, symmetry=sym, adv_g=adv, tv=tv, ce=ce)
    if identity_sd is not None:
        from . import identity_port
        ip = identity_port.identity_loss(identity_sd, fake, b["img_frontal"])
        total = total + w["weight_identity_preserving"] * ip
        m["ip"] = ip
    m["g_total"] = total
    return total, m


def d_loss(D: Callable, fake, b):
    real = b["img_frontal"]
    a = b["gp_alpha"].view(-1, 1, 1, 1)
    xhat = (a * real + (1 - a) * fake).detach().requires_grad_(True)
    grad = torch.autograd.grad(D(xhat).sum(), xhat, create_graph=True)[0]
    gp = ((grad.flatten(1).norm(dim=1) - 1) ** 2).mean()
    d_fake, d_real = D(fake).mean(), D(real).mean()
    total = d_fake - d_real + LOSS_W["weight_gradient_penalty"] * gp
    return total, dict(d_fake=d_fake, d_real=d_real, gp=gp, d_total=total)


def train_step(G: Callable, D: Callable, g_params, d_params, opt_g, opt_d, b, step_optim: bool = True, identity_sd=None):
    """One oracle step.  G(b) -> 8-tuple, D(x) -> logits.  Returns python-float metrics."""
    g_out = G(b)
    fake = g_out[0]
    # ---- D phase
    for p in d_params:
        p.requires_grad_(True)
    opt_d.zero_grad(set_to_none=True)
    ld, md = d_loss(D, fake.detach(), b)
    ld.backward()
    if step_optim:
        opt_d.step()
    # ---- G phase (D frozen: set_requires_grad(D.parameters(), False), UtilityMethods.py:43-56)
    for p in d_params:
        p.requires_grad_(False)
    opt_g.zero_grad(set_to_none=True)
    lg, mg = g_loss(g_out, D(fake), b, identity_sd)
    lg.backward()
    if step_optim:
        opt_g.step()
    for p in d_params:
        p.requires_grad_(True)
    out = {k: float(v) for k, v in {**md, **mg}.items()}
    return out


def port_callables(sd_g, sd_d, use_dropout_mask=None):
    """(G, D) callables over the fp32 port (oracle/model_port.py) and reference-format state dicts."""
    from . import model_port as mp

    def G(b):
        return mp.generator(sd_g, b["img"], b["left_eye"], b["right_eye"], b["nose"], b["mouth"], b["z"], use_dropout_mask)

    def D(x):
        return mp.discriminator(sd_d, x)

    return G, D


def flops_per_image() -> Dict[str, float]:
    """Algorithmic forward FLOPs (2*MACs over conv/deconv/linear) of G and D per image, from the layer list; must
    reproduce SURVEY's probe: 176.56 and 1.298 GFLOP."""
    from . import layer_table
    return layer_table.flops()
