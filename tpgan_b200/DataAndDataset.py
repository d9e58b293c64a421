"""Device-side counterparts of the per-sample host work in the reference's DataAndDataset.py / UtilityMethods.py
(SURVEY.md 8 row f2).  Nothing here touches files or PIL: the functions take batches that are already in device memory.

  process_batch(img, landmarks)          process() (DataAndDataset.py:10-56): the four landmark-centred patches of every image
                                         + the int32 crop boxes handed to PIL.Image.crop, bit-exact (tpgan_patch_crop)
  to_tensor_normalized(img_u8)           transforms.ToTensor() then *2.0 - 1.0 (DataAndDataset.py:214-220,251-255)
  get_5_landmarks_pixal_position(lm68)   UtilityMethods.py:146-164, batched, with TestDataset's 128/width, 128/height rescale
  pyramid(img128)                        the 64x64 / 32x32 targets as average pools (the oracle step's convention)
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops
from .ops import Act

# UtilityMethods.py:146, literally (the fifth range lies outside a 68-point list: the reference takes np.mean of an empty
# slice there and gets NaN; pass ranges=FIVE_PTS_IDX_DLIB for the conventional right mouth corner, index 54)
five_pts_idx = [[36, 41], [42, 47], [27, 35], [48, 48], [68, 68]]
FIVE_PTS_IDX_DLIB = [[36, 41], [42, 47], [27, 35], [48, 48], [54, 54]]
PATCH_HW = ((40, 40), (40, 40), (32, 40), (32, 48))     # DataAndDataset.py:35-40 (h, w) of left eye, right eye, nose, mouth
PART_NAMES = ("left_eye", "right_eye", "nose", "mouth")


def _cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("tpgan_b200 input pipeline runs on CUDA tensors only (there is no CPU fallback)")


def to_tensor_normalized(img_u8: torch.Tensor, round_tf32: bool = False) -> Act:
    """(B,H,W,C) uint8 -> NHWC fp32 Act in [-1,1]."""
    _cuda(img_u8)
    B, H, W, C = img_u8.shape
    out = Act.empty(B, H, W, C, img_u8.device)
    ops.u8_to_nhwc(img_u8.contiguous(), out, round_tf32)
    return out


def get_5_landmarks_pixal_position(x: torch.Tensor, image_size: Optional[Tuple[int, int]] = None, ranges=None) -> torch.Tensor:
    """(B,P,2) float32 -> (B,5,2).  image_size = (width, height) applies TestDataset's rescale to a 128x128 frame."""
    _cuda(x)
    r = torch.tensor(five_pts_idx if ranges is None else ranges, dtype=torch.int32, device=x.device)
    out = torch.empty((x.shape[0], r.shape[0], 2), dtype=torch.float32, device=x.device)
    # `lm[i][0] *= 128/img.width` on a float32 array: the Python float is cast to float32 first (NumPy 2 promotion)
    sx, sy = (1.0, 1.0) if image_size is None else (128 / image_size[0], 128 / image_size[1])
    ops.landmarks_reduce(x.float().contiguous(), r, out, float(torch.tensor(sx, dtype=torch.float32)),
                         float(torch.tensor(sy, dtype=torch.float32)))
    return out


def pyramid(img128: Act) -> Tuple[Act, Act]:
    half = Act.empty(img128.n, img128.h // 2, img128.w // 2, img128.c, img128.buf.device)
    quarter = Act.empty(img128.n, img128.h // 4, img128.w // 4, img128.c, img128.buf.device)
    ops.pyramid(img128, half, quarter)
    return half, quarter


def process_batch(img: Act, landmarks: torch.Tensor, fill: float = -1.0) -> Dict[str, object]:
    """process() for a batch: {'left_eye', 'right_eye', 'nose', 'mouth'} NHWC patches + 'boxes' (B,4,4) int32."""
    _cuda(landmarks)
    patches = [Act.empty(img.n, h, w, img.c, img.buf.device) for h, w in PATCH_HW]
    boxes = torch.zeros((img.n, 4, 4), dtype=torch.int32, device=img.buf.device)
    ops.patch_crop(img, landmarks.float().contiguous(), patches, boxes, fill)
    out: Dict[str, object] = dict(zip(PART_NAMES, patches))
    out["boxes"] = boxes
    return out
