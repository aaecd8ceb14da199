"""`mmdet.apis.init_detector / inference_detector` over the B200 YOLOv5s localizer.

/root/reference/infer_effocr.py has no YOLO branch: its localizer is `init_detector(config, checkpoint, device=...,
cfg_options=...)` (:150) and every line goes through `inference_detector(self.localizer, path)` (:262), whose result
it reads as mmdetection's `(bbox_results, segm_results)` pair -- `bbox_results` one `ndarray[n, 5]`
(x0, y0, x1, y1, score; ORIGINAL image pixels) per class, character boxes first, word boxes second
(:272-276, :348, :415; `mmdet_output_format` :245-254 shows the shape).  The score thresholds are applied by the
caller (`x[4] > self.score_thresh`, :350,352,417).

`effocr_b200.dropin.install()` registers this module as `mmdet.apis`, so the unmodified `EffOCR` class localizes with
csrc/yolo.cu + csrc/nms.cu: the line is letterboxed exactly like the reference's ONNX driver does
(onnx_engines/localizer_engine.py:75-85), YOLOv5s + NMS run on the device, and the boxes are mapped back to the
original image -- `(x - pad) / r`, clipped to the image like ultralytics' `scale_boxes` -- because this caller slices
the ORIGINAL line with them.

`checkpoint`: the YOLOv5s weights (`best_bbox_mAP.pth` in the localizer directory, :516-518) as an ultralytics-keyed
state dict / `best.pt` pickle / .npz / exported .onnx.  `config` (an mmdetection .py file for the reference) is not
interpreted; `cfg_options` may carry `effocr_b200.conf_thresh`, `effocr_b200.iou_thresh`, `effocr_b200.input_shape`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lineio
from .localizer_engine import EffLocalizer, letterbox_geometry

# NMS runs with a LOW confidence floor and the caller thresholds afterwards: greedy NMS decides a box's fate from
# higher-scored boxes only, so (NMS at 0.05, then score > t) keeps exactly the boxes (NMS at t) keeps for any t >= 0.05.
DEFAULT_CONF_FLOOR = 0.05
DEFAULT_IOU = 0.01  # the reference's YOLO driver default (infer_effocr_onnx_multi.py:441-444)


class YoloDetector(torch.nn.Module):
    """What `init_detector` returns: holds the EffLocalizer; `named_parameters()` walks the detector's weights so that
    `count_parameters(ocr_engine.localizer)` (infer_effocr.py:539, utils/eval_utils.py:4-11) works."""

    def __init__(self, localizer: EffLocalizer, state_dict):
        super().__init__()
        object.__setattr__(self, "localizer", localizer)
        object.__setattr__(self, "_state", state_dict)
        self.nc = localizer._eng_net.nc

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        for k, v in self._state.items():
            if k.endswith(("running_mean", "running_var", "num_batches_tracked", "anchors")):
                continue
            yield prefix + k, torch.nn.Parameter(torch.as_tensor(v, dtype=torch.float32), requires_grad=True)

    def forward(self, *a, **k):
        raise NotImplementedError("use mmdet.apis.inference_detector(model, image)")


def init_detector(config=None, checkpoint=None, device="cuda:0", cfg_options=None):
    from .localizer_engine import _load_state

    opts = dict(cfg_options or {})
    sd = _load_state(checkpoint)
    loc = EffLocalizer(sd, iou_thresh=float(opts.get("effocr_b200.iou_thresh", DEFAULT_IOU)),
                       conf_thresh=float(opts.get("effocr_b200.conf_thresh", DEFAULT_CONF_FLOOR)),
                       input_shape=tuple(opts.get("effocr_b200.input_shape", (640, 640))), max_batch=16)
    return YoloDetector(loc, sd)


def unletterbox(det: np.ndarray, height: int, width: int, input_shape=(640, 640)) -> np.ndarray:
    """[n, >=4] letterbox-pixel xyxy (float32) -> original-image pixels: subtract the padding, divide by the resize
    gain, clip to the image (float32 arithmetic throughout; the gain is ultralytics' `min(H'/h, W'/w)`)."""
    _nw, _nh, left, top = letterbox_geometry(height, width, input_shape)
    gain = np.float32(min(input_shape[0] / height, input_shape[1] / width))
    out = np.array(det, dtype=np.float32, copy=True)
    out[:, [0, 2]] = (out[:, [0, 2]] - np.float32(left)) / gain
    out[:, [1, 3]] = (out[:, [1, 3]] - np.float32(top)) / gain
    out[:, [0, 2]] = np.clip(out[:, [0, 2]], 0, np.float32(width))
    out[:, [1, 3]] = np.clip(out[:, [1, 3]], 0, np.float32(height))
    return out


def format_result(det: np.ndarray, nc: int):
    """NMS rows (x0, y0, x1, y1, conf, cls) in image pixels -> mmdetection's (bbox_results, segm_results)."""
    det = np.asarray(det, dtype=np.float32).reshape(-1, 6)
    per_class = [np.ascontiguousarray(det[det[:, 5] == c][:, :5]) for c in range(nc)]
    return per_class, [[] for _ in range(nc)]


def inference_detector(model: YoloDetector, imgs):
    """One image path / RGB-or-BGR-agnostic path list -> mmdetection-style result(s)."""
    single = not isinstance(imgs, (list, tuple))
    paths = [imgs] if single else list(imgs)
    loc = model.localizer
    results = []
    for p in paths:
        if isinstance(p, np.ndarray):  # mmdet accepts a loaded BGR image too
            im_bgr = p
            x = EffLocalizer.preprocess_bgr(im_bgr, loc._input_shape)
            h, w = im_bgr.shape[:2]
        else:
            rgb = lineio.decode_rgb(p)
            h, w = rgb.shape[:2]
            x = EffLocalizer.preprocess_bgr(np.ascontiguousarray(rgb[:, :, ::-1]), loc._input_shape)
        det = loc.run([x])[0].numpy()
        results.append(format_result(unletterbox(det, h, w, loc._input_shape), model.nc))
    return results[0] if single else results
