"""Drop-in for the hot part of /root/reference/utils/datasets_utils.py: `create_paired_transform`
(:166-172) and `MedianPad` (:69-90).

`create_paired_transform(size=224)` returns a callable `np.uint8[h,w,3] | PIL.Image -> f32[3,224,224]`
(CPU tensor, like the reference) that runs the fused crop kernel (csrc/crop.cu) for ONE crop.  It
exists for API parity; the pipeline never calls it per crop -- it batches all boxes of all lines into
one `effocr_crop_resize` launch writing the encoder's input buffer directly.
Raises ValueError on an empty crop, like PIL does inside the reference (infer_effocr.py:294-297).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


class MedianPad:
    """Pad right/bottom to a square.  With `override` set (the only way the inference path uses it)
    the fill is that colour; the GPU transform implements override=(255,255,255) only."""

    def __init__(self, override=None):
        self.override = override

    def __call__(self, image):
        arr = np.asarray(image)
        if arr.ndim != 3 or arr.shape[0] == 0 or arr.shape[1] == 0:
            raise ValueError("empty image")
        h, w, c = arr.shape
        s = max(h, w)
        if self.override is None:
            border = np.concatenate([arr[:, w - 1, :], arr[:, 0, :], arr[0, :, :], arr[h - 1, :, :]], axis=0)
            fill = tuple(int(v) for v in np.median(border, axis=0))
        else:
            fill = self.override
        out = np.empty((s, s, c), dtype=arr.dtype)
        out[...] = np.asarray(fill, dtype=arr.dtype)
        out[:h, :w] = arr
        return out


class PairedTransform:

    def __init__(self, size=224):
        if size != 224:
            raise _lib.EffocrError("the fused crop kernel produces 224x224 crops only")
        self.size = size

    def __call__(self, image):
        arr = np.ascontiguousarray(np.asarray(image))
        if arr.ndim != 3 or arr.shape[2] != 3 or arr.shape[0] == 0 or arr.shape[1] == 0:
            raise ValueError("tile cannot extend outside image")  # PIL's message for an empty crop
        return self.batch([arr])[0].cpu()

    def batch(self, arrays, layout=ops.CROP_NCHW_F32):
        """All crops in one launch -> CUDA tensor [n,3,224,224] (or patch-major fp16)."""
        pixels, images, _ = ops.pack_images(arrays)
        boxes, n = ops.pack_boxes([(i, 0, 0, a.shape[1], a.shape[0]) for i, a in enumerate(arrays)])
        return ops.crop_resize(pixels, images, boxes, n, layout)


def create_paired_transform(size=224, lang=None):
    """`lang` is accepted and ignored: the reference's own callers pass it although the reference
    signature is (size=224) (infer_effocr_onnx_multi.py:489 -- a TypeError there)."""
    if isinstance(size, str):  # scripts/recognizer_onnx_export.py:104 passes lang positionally
        size = 224
    return PairedTransform(size)
