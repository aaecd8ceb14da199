"""Drop-in for /root/reference/onnx_engines/localizer_engine.py: `EffLocalizer`.

Same constructor and call surface -- `EffLocalizer(model_path, iou_thresh=0.01, conf_thresh=0.30, vertical=False,
num_cores=None, providers=None, input_shape=(640, 640), model_backend='yolo')`, `run(imgs)` with `imgs` a list of
paths or of letterboxed `np.float32[1,3,H,W]` arrays, returning one `Tensor f32[n, 6]` (x1, y1, x2, y2, conf, cls;
letterbox pixels; confidence-sorted) per image, and the static helpers `letterbox`, `load_localizer_img`,
`xywh2xyxy`, `box_iou`, `non_max_suppression` -- but the YOLOv5s forward pass and the NMS run in
csrc/yolo.cu / csrc/nms.cu on the B200 instead of onnxruntime + torchvision on the CPU.

`model_path`: an ultralytics-keyed YOLOv5s state dict (`model.{i}...`; what `best.pt['model'].state_dict()` holds),
as a dict, a .pth/.pt/.npz path, an ultralytics `best.pt` pickle, or the exported `.onnx` (its initializers are read,
the graph is not executed: effocr_b200/weights_io.py).  Only model_backend='yolo' is implemented
(mmdetection / detectron2 back-ends are out of scope).
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _lib

YOLO_CONV_ORDER_C3 = lambda p, n: [p + "cv1.", p + "cv2.", p + "cv3."] + [p + f"m.{j}.cv{k}." for j in range(n) for k in (1, 2)]


def yolo_conv_prefixes():
    """Conv blocks in the order csrc/yolo.cu::yolo_specs() expects."""
    c3 = YOLO_CONV_ORDER_C3
    out = ["model.0.", "model.1."] + c3("model.2.", 1) + ["model.3."] + c3("model.4.", 2) + ["model.5."] + c3("model.6.", 3)
    out += ["model.7."] + c3("model.8.", 1) + ["model.9.cv1.", "model.9.cv2.", "model.10."] + c3("model.13.", 1)
    out += ["model.14."] + c3("model.17.", 1) + ["model.18."] + c3("model.20.", 1) + ["model.21."] + c3("model.23.", 1)
    return out


def yolo_weight_order():
    keys = []
    for p in yolo_conv_prefixes():
        keys += [p + "conv.weight", p + "bn.weight", p + "bn.bias", p + "bn.running_mean", p + "bn.running_var"]
    for l in range(3):
        keys += [f"model.24.m.{l}.weight", f"model.24.m.{l}.bias"]
    keys.append("model.24.anchors")
    return keys


def _load_state(model):
    from .weights_io import load_yolo_state

    return load_yolo_state(model)


LETTERBOX_PLAN_DTYPE = np.dtype([("new_width", "<i4"), ("new_height", "<i4"), ("left", "<i4"), ("top", "<i4"),
                                 ("xtap_offset", "<i4"), ("ytap_offset", "<i4")])
_tap_cache: dict = {}


def _resize_taps(src: int, dst: int, clamp_fraction: bool) -> np.ndarray:
    """Tap table of cv2.resize(INTER_LINEAR) on 8-bit images along one axis: int32 [dst, 4] =
    (source index 0, source index 1, coefficient 0, coefficient 1), coefficients in 1/2048 units.
    Follows OpenCV imgproc/resize.cpp: position `(float)((d + 0.5) * scale - 0.5)` with `scale = 1 / (dst / src)` in
    double, `cvFloor`, coefficients `saturate_cast<short>(f * 2048)` (round half to even).  Along x the fraction is
    zeroed where the 2-tap window leaves the image (`clamp_fraction`); along y OpenCV keeps the fraction and clips
    the two row indices instead -- the fixed-point results differ, so both rules are reproduced."""
    key = (src, dst, clamp_fraction)
    hit = _tap_cache.get(key)
    if hit is not None:
        return hit
    scale = 1.0 / (float(dst) / float(src))
    f = ((np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        lo, hi = s < 0, s >= src - 1
        f[lo | hi] = 0
        s[lo] = 0
        s[hi] = src - 1
    taps = np.empty((dst, 4), dtype=np.int32)
    taps[:, 0] = np.clip(s, 0, src - 1)
    taps[:, 1] = np.clip(s + 1, 0, src - 1)
    taps[:, 2] = np.clip(np.rint((np.float32(1.0) - f) * np.float32(2048.0)), -32768, 32767).astype(np.int32)
    taps[:, 3] = np.clip(np.rint(f * np.float32(2048.0)), -32768, 32767).astype(np.int32)
    _tap_cache[key] = taps
    return taps


def letterbox_geometry(height: int, width: int, new_shape=(640, 640)):
    """localizer_engine.py:107-138 with auto=False, scaleFill=False, scaleup=True: -> (new_width, new_height, left, top)."""
    r = min(new_shape[0] / height, new_shape[1] / width)
    nw, nh = int(round(width * r)), int(round(height * r))
    dw, dh = (new_shape[1] - nw) / 2, (new_shape[0] - nh) / 2
    return nw, nh, int(round(dw - 0.1)), int(round(dh - 0.1))


def letterbox_plan(shapes, new_shape=(640, 640)):
    """Per-image plans + the concatenated tap tables for effocr_letterbox_resize.  shapes: iterable of (height, width)."""
    plans = np.zeros(len(shapes), dtype=LETTERBOX_PLAN_DTYPE)
    tables, offsets, off = [], {}, 0
    for i, (h, w) in enumerate(shapes):
        nw, nh, left, top = letterbox_geometry(h, w, new_shape)
        if nw < 1 or nh < 1:
            raise _lib.EffocrError(f"letterbox: image {h}x{w} collapses to {nh}x{nw} at model shape {tuple(new_shape)}")
        for key in ((w, nw, True), (h, nh, False)):
            if key not in offsets:
                offsets[key] = off
                tables.append(_resize_taps(*key))
                off += key[1]
        plans[i] = (nw, nh, left, top, offsets[(w, nw, True)], offsets[(h, nh, False)])
    return plans, np.ascontiguousarray(np.concatenate(tables, 0))


YOLO_SPLIT, YOLO_FP16 = 0, 1  # include/effocr_b200.h: EFFOCR_YOLO_SPLIT / EFFOCR_YOLO_FP16


class YoloEngine:
    """Device-resident YOLOv5s (NHWC activations, folded BN) behind effocr_yolo_*.
    precision: "split" (default) -- (hi, lo) fp16 plane pairs and [Whi | Wlo] weights, fp32-accurate like the reference's
    onnxruntime session; "fp16" -- single fp16 plane, about twice as fast, boxes within ~1 px of fp32."""

    def __init__(self, state_dict, max_batch=16, max_shape=(640, 640), precision="split"):
        self._lib = _lib.load()
        _lib.require_device()
        sd = state_dict
        self.nc = int(sd["model.24.m.0.bias"].numel() // 3 - 5)
        self.no = self.nc + 5
        keys = yolo_weight_order()
        keep, ptrs = [], (C.c_void_p * len(keys))()
        for i, k in enumerate(keys):
            a = np.ascontiguousarray(sd[k].detach().to("cpu", torch.float32).numpy())
            keep.append(a)
            ptrs[i] = a.ctypes.data
        h = C.c_void_p()
        _lib.check(self._lib.effocr_yolo_create(self.nc, int(max_batch), int(max_shape[0]), int(max_shape[1]), ptrs, len(keys),
                                                C.byref(h)), "effocr_yolo_create")
        self._h = h
        self._lock = threading.Lock()
        self.max_batch = max_batch
        self.set_precision(precision)

    def set_precision(self, precision: str) -> None:
        if precision not in ("split", "fp16"):
            raise ValueError("precision must be 'split' or 'fp16'")
        _lib.check(self._lib.effocr_yolo_set_mode(self._h, YOLO_SPLIT if precision == "split" else YOLO_FP16), "effocr_yolo_set_mode")
        self.precision = precision

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.effocr_yolo_destroy(h)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: CUDA f32 [B,3,H,W] -> CUDA f32 [B, npred, 5+nc] decoded predictions."""
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
            raise _lib.EffocrError("YoloEngine.forward expects a CUDA float32 [B,3,H,W] tensor")
        x = x.contiguous()
        B, _, H, W = x.shape
        npred = self._lib.effocr_yolo_num_predictions(H, W)
        if npred < 0:
            raise _lib.EffocrError("input height/width must be multiples of 32")
        out = torch.empty((B, npred, self.no), device=x.device, dtype=torch.float32)
        with self._lock:
            _lib.check(self._lib.effocr_yolo_forward(self._h, x.data_ptr(), B, H, W, out.data_ptr(), _lib.stream_ptr()),
                       "effocr_yolo_forward")
        return out


def nms_device(pred: torch.Tensor, conf_thres: float, iou_thres: float, max_det: int = 1000):
    """pred CUDA f32 [B, npred, no] -> (out f32 [B, max_det, 6], count i32 [B]) on the device."""
    lib = _lib.load()
    if not pred.is_cuda or pred.dtype != torch.float32:
        raise _lib.EffocrError("nms_device expects a CUDA float32 tensor")
    pred = pred.contiguous()
    B, npred, no = pred.shape
    out = torch.empty((B, max_det, 6), device=pred.device, dtype=torch.float32)
    cnt = torch.empty((B,), device=pred.device, dtype=torch.int32)
    _lib.check(lib.effocr_nms(pred.data_ptr(), B, npred, no, float(conf_thres), float(iou_thres), int(max_det), out.data_ptr(),
                              cnt.data_ptr(), _lib.stream_ptr()), "effocr_nms")
    return out, cnt


class EffLocalizer:

    def __init__(self, model_path, iou_thresh=0.01, conf_thresh=0.30, vertical=False, num_cores=None, providers=None,
                 input_shape=(640, 640), model_backend="yolo", max_batch=16, precision="split"):
        if model_backend != "yolo":
            raise NotImplementedError("Backend {} is not implemented".format(model_backend))
        if input_shape is None:  # the reference falls back to the model's own (static) shape, localizer_engine.py:38-41
            input_shape = (640, 640)
        self._eng_net = YoloEngine(_load_state(model_path), max_batch=max_batch, max_shape=tuple(input_shape), precision=precision)
        self._iou_thresh = iou_thresh
        self._conf_thresh = conf_thresh
        self._vertical = vertical
        self._input_shape = tuple(input_shape)
        self._model_input_shape = [1, 3, "height", "width"]  # dynamic, like an ONNX export with dynamic axes
        self._model_backend = model_backend
        self._input_name = "images"

    def __call__(self, imgs):
        return self.run(imgs)

    def run(self, imgs):
        if isinstance(imgs, list):
            if isinstance(imgs[0], str):
                imgs = [EffLocalizer.load_localizer_img(img, self._input_shape, backend=self._model_backend) for img in imgs]
        # batch images of equal shape into one forward pass (the reference loops one image at a time)
        arrs = [np.asarray(i, dtype=np.float32).reshape((-1,) + tuple(np.asarray(i).shape[-3:])) for i in imgs]
        results = [None] * len(arrs)
        by_shape = {}
        for k, a in enumerate(arrs):
            by_shape.setdefault(a.shape[1:], []).append(k)
        for shape, idxs in by_shape.items():
            x = torch.from_numpy(np.concatenate([arrs[k] for k in idxs], 0)).cuda(non_blocking=True)
            pred = self._eng_net.forward(x)
            out, cnt = nms_device(pred, self._conf_thresh, self._iou_thresh, max_det=1000)
            out, cnt = out.cpu(), cnt.cpu().tolist()
            for j, k in enumerate(idxs):
                results[k] = out[j, :cnt[j]].clone()
        return results

    # ---- device-resident variant used by the pipeline: no host round trip of predictions
    def run_device(self, x: torch.Tensor):
        pred = self._eng_net.forward(x)
        return nms_device(pred, self._conf_thresh, self._iou_thresh, max_det=1000)

    # ---- static helpers with the reference's signatures (host pre-processing stays on the host: OpenCV)
    @staticmethod
    def get_onnx_input_name(model):
        return "images"

    @staticmethod
    def load_localizer_img(input_path, input_shape, backend="yolo"):
        """localizer_engine.py:75-85."""
        if backend != "yolo":
            raise NotImplementedError("Backend {} is not implemented".format(backend))
        import cv2

        im0 = cv2.imread(input_path)
        return EffLocalizer.preprocess_bgr(im0, input_shape)

    @staticmethod
    def preprocess_bgr(im0, input_shape):
        """BGR u8 [H, W, 3] -> the model input f32 [1, 3, H', W'] (RGB, 0..1), letterboxed (localizer_engine.py:79-85)."""
        boxed = EffLocalizer.letterbox(im0, input_shape, stride=32, auto=False)[0]
        chw_rgb = np.ascontiguousarray(boxed[:, :, ::-1].transpose(2, 0, 1))
        return (chw_rgb.astype(np.float32) / 255.0)[None] if chw_rgb.ndim == 3 else chw_rgb.astype(np.float32) / 255.0

    @staticmethod
    def letterbox(im, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32):
        """localizer_engine.py:107-138 (the ultralytics letterbox): aspect-preserving resize + constant border.
        -> (image, (ratio_w, ratio_h), (dw, dh)).  The default path (auto=False, scaleFill=False, scaleup=True) goes
        through `letterbox_geometry`, the same arithmetic the device letterbox kernel's plan uses."""
        import cv2

        h, w = im.shape[:2]
        if isinstance(new_shape, int):
            new_shape = (new_shape, new_shape)
        gain = min(new_shape[0] / h, new_shape[1] / w)
        if not scaleup:
            gain = min(gain, 1.0)
        ratio = (gain, gain)
        if scaleFill and not auto:
            out_w, out_h, pad_w, pad_h = new_shape[1], new_shape[0], 0.0, 0.0
            ratio = (new_shape[1] / w, new_shape[0] / h)
        else:
            out_w, out_h = int(round(w * gain)), int(round(h * gain))
            pad_w, pad_h = new_shape[1] - out_w, new_shape[0] - out_h
            if auto:  # minimum rectangle: pad only up to the next stride multiple
                pad_w, pad_h = np.mod(pad_w, stride), np.mod(pad_h, stride)
        pad_w, pad_h = pad_w / 2, pad_h / 2
        if (w, h) != (out_w, out_h):
            im = cv2.resize(im, (out_w, out_h), interpolation=cv2.INTER_LINEAR)
        edges = [int(round(v)) for v in (pad_h - 0.1, pad_h + 0.1, pad_w - 0.1, pad_w + 0.1)]  # top, bottom, left, right
        if not auto and not scaleFill and scaleup:
            assert (out_w, out_h, edges[2], edges[0]) == letterbox_geometry(h, w, new_shape)
        im = cv2.copyMakeBorder(im, *edges, cv2.BORDER_CONSTANT, value=color)
        return im, ratio, (pad_w, pad_h)

    @staticmethod
    def xywh2xyxy(x):
        """[n, 4+] centre/size boxes -> corner boxes (localizer_engine.py:140-148); other columns are kept."""
        y = x.clone() if isinstance(x, torch.Tensor) else np.copy(x)
        half_w, half_h = x[:, 2] / 2, x[:, 3] / 2
        y[:, 0], y[:, 2] = x[:, 0] - half_w, x[:, 0] + half_w
        y[:, 1], y[:, 3] = x[:, 1] - half_h, x[:, 1] + half_h
        return y

    @staticmethod
    def box_iou(box1, box2, eps=1e-7):
        """Pairwise IoU of corner boxes [n, 4] x [m, 4] -> [n, m] (localizer_engine.py:150-169)."""
        lo = torch.max(box1[:, None, :2], box2[None, :, :2])
        hi = torch.min(box1[:, None, 2:4], box2[None, :, 2:4])
        inter = (hi - lo).clamp(min=0).prod(dim=2)
        area1 = (box1[:, 2:4] - box1[:, :2]).prod(dim=1)
        area2 = (box2[:, 2:4] - box2[:, :2]).prod(dim=1)
        return inter / (area1[:, None] + area2[None, :] - inter + eps)

    @staticmethod
    def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                            labels=(), max_det=300, nm=0):
        """Reference signature; single-label, class-aware path (the only one the reference exercises) on the GPU."""
        if classes is not None or agnostic or multi_label or labels or nm:
            raise NotImplementedError("only the default single-label class-aware NMS path is implemented")
        if isinstance(prediction, (list, tuple)):
            prediction = prediction[0]
        assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
        assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
        dev = prediction.device
        out, cnt = nms_device(torch.as_tensor(prediction, dtype=torch.float32).cuda(), conf_thres, iou_thres, max_det)
        out, cnt = out.to(dev), cnt.cpu().tolist()
        return [out[b, :cnt[b]].clone() for b in range(out.shape[0])]
