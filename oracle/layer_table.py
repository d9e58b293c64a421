"""ORACLE (test infrastructure): algorithmic FLOP counter for the G and D forward passes (2*MACs over every conv,
transposed conv and linear layer, unpadded channel counts), from the reference-format parameter shapes and the fixed
spatial sizes of the 128x128 model.  Reproduces the survey probe (176.56 / 1.298 GFLOP per image)."""
from __future__ import annotations

from typing import Dict

PATCH_HW = {"left_eye": (40, 40), "right_eye": (40, 40), "nose": (32, 40), "mouth": (32, 48)}


def _local(shapes, pre, h, w):
    """LocalPathway (D_and_G_model.py:84-110): layer -> output spatial size."""
    f = 0
    size = {"conv0": (h, w), "conv1": (h // 2, w // 2), "conv2": (h // 4, w // 4), "conv3": (h // 8, w // 8),
            "after_select0": (h // 4, w // 4), "after_select1": (h // 2, w // 2), "after_select2": (h, w),
            "local_img": (h, w)}
    dec_in = {"deconv0": (h // 8, w // 8), "deconv1": (h // 4, w // 4), "deconv2": (h // 2, w // 2)}
    for k, s in shapes.items():
        if not k.startswith(pre + ".") or not k.endswith("weight"):
            continue
        name = k[len(pre) + 1:].split(".")[0]
        if name in dec_in:                       # ConvTranspose2d weight (Cin, Cout, k, k): MACs per INPUT pixel
            hh, ww = dec_in[name]
            f += 2 * hh * ww * s[0] * s[1] * s[2] * s[3]
        else:
            hh, ww = size[name]
            f += 2 * hh * ww * s[0] * s[1] * s[2] * s[3]
    return f


def _global(shapes, pre="global_pathway"):
    out_hw = {"conv0": 128, "conv1": 64, "conv2": 32, "conv3": 16, "conv4": 8, "add_conv_and_deconv_8": 8,
              "enhance_features_8": 8, "add_conv_and_deconv_16": 16, "enhance_features_16": 16,
              "add_conv_and_deconv_32": 32, "enhance_features_32": 32, "add_conv_and_deconv_64": 64,
              "enhance_features_64": 64, "add_conv_and_deconv_128": 128, "enhance_features_128": 128, "conv5": 128,
              "conv6": 128, "decoded_img128": 128}
    dec_in = {"deconv_8": 1, "deconv_32": 8, "deconv_64": 32, "deconv_128": 64, "upsample_16": 8, "upsample_32": 16,
              "upsample_64": 32, "upsample_128": 64}
    f = 0
    for k, s in shapes.items():
        if not k.startswith(pre + ".") or not k.endswith("weight"):
            continue
        name = k[len(pre) + 1:].split(".")[0]
        if name == "fc1":
            f += 2 * s[0] * s[1]
        elif name in dec_in:
            f += 2 * dec_in[name] ** 2 * s[0] * s[1] * s[2] * s[3]
        else:
            f += 2 * out_hw[name] ** 2 * s[0] * s[1] * s[2] * s[3]
    return f


def flops(shapes_g: Dict[str, tuple] = None, shapes_d: Dict[str, tuple] = None) -> Dict[str, float]:
    if shapes_g is None or shapes_d is None:
        from tpgan_b200 import D_and_G_model as M, config
        import torch
        with torch.device("meta"):
            G = M.Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"],
                            config.G["use_residual_block"])
            D = M.Discriminator(config.D["use_batchnorm"])
        shapes_g = {k: tuple(v.shape) for k, v in G.state_dict().items()}
        shapes_d = {k: tuple(v.shape) for k, v in D.state_dict().items()}
    g = sum(_local(shapes_g, f"local_pathway_{n}", h, w) for n, (h, w) in PATCH_HW.items())
    g += _global(shapes_g)
    s = shapes_g["feature_predict.fc.weight"]
    g += 2 * s[0] * s[1]
    d, hw = 0, 128
    for i in range(8):
        for sub in ("", ".layers.0", ".layers.1"):
            k = f"model.{i}{sub}.0.weight"
            if k in shapes_d:
                s = shapes_d[k]
                if sub == "" and i in (0, 1, 2, 3, 5):
                    hw //= 2
                d += 2 * hw * hw * s[0] * s[1] * s[2] * s[3]
    return {"G": float(g), "D": float(d)}
