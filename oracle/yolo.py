"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the localizer: YOLOv5s forward + letterbox + NMS.

The reference runs an exported YOLOv5 ONNX graph through onnxruntime
(/root/reference/onnx_engines/localizer_engine.py:25-29,49-55) and its torch-hub variant loads
`ultralytics/yolov5` `yolov5s` (onnx_engines/infer_ocr_yolo.py:272-278, classes = 2 for en / 1 for jp).
ultralytics/yolov5 is un-vendored and unpinned, so the published v6.x/7.0 `yolov5s.yaml` graph
(depth 0.33, width 0.50) is restated here functionally on an ultralytics-keyed
state dict (`model.{i}....`), with BatchNorm (eps 1e-3) applied explicitly.

PARITY UNPINNED against a second YOLOv5 implementation (none is installable here); pinned instead by
analytic invariants the published model must satisfy (tests/test_oracle.py::test_yolov5s_analytic_invariants): 7 235 389
parameters at nc = 80, 25 200 predictions at 640 x 640, fp64-vs-fp32 self-consistency.

`letterbox` / `non_max_suppression` restate localizer_engine.py:107-138 / :171-277 and ARE pinned
against the live reference functions and the fixtures in tests/golden/.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

ANCHORS = (((10, 13), (16, 30), (33, 23)), ((30, 61), (62, 45), (59, 119)), ((116, 90), (156, 198), (373, 326)))
STRIDES = (8, 16, 32)
BN_EPS = 1e-3

# (index, kind, args) of yolov5s.yaml after width/depth scaling; `f` = input layer(s)
#   Conv: (c1, c2, k, s)   C3: (c1, c2, n, shortcut)   SPPF: (c1, c2, k)
LAYERS = [
    (0, "Conv", (3, 32, 6, 2), -1), (1, "Conv", (32, 64, 3, 2), -1), (2, "C3", (64, 64, 1, True), -1),
    (3, "Conv", (64, 128, 3, 2), -1), (4, "C3", (128, 128, 2, True), -1), (5, "Conv", (128, 256, 3, 2), -1),
    (6, "C3", (256, 256, 3, True), -1), (7, "Conv", (256, 512, 3, 2), -1), (8, "C3", (512, 512, 1, True), -1),
    (9, "SPPF", (512, 512, 5), -1), (10, "Conv", (512, 256, 1, 1), -1), (11, "Up", (), -1), (12, "Cat", (), (-1, 6)),
    (13, "C3", (512, 256, 1, False), -1), (14, "Conv", (256, 128, 1, 1), -1), (15, "Up", (), -1), (16, "Cat", (), (-1, 4)),
    (17, "C3", (256, 128, 1, False), -1), (18, "Conv", (128, 128, 3, 2), -1), (19, "Cat", (), (-1, 14)),
    (20, "C3", (256, 256, 1, False), -1), (21, "Conv", (256, 256, 3, 2), -1), (22, "Cat", (), (-1, 10)),
    (23, "C3", (512, 512, 1, False), -1),
]
DETECT_FROM = (17, 20, 23)
DETECT_CH = (128, 256, 512)


def conv_names(prefix):
    return [prefix + "conv.weight", prefix + "bn.weight", prefix + "bn.bias", prefix + "bn.running_mean", prefix + "bn.running_var"]


def init_yolov5s_state_dict(nc: int = 2, seed: int = 0, obj_bias: float | None = None):
    """Kaiming-uniform convs, BN (gamma 1, beta 0, mean 0, var 1) with a small random perturbation so
    that folding is exercised, Detect bias a la ultralytics (obj <- log(8 / (640 / s)^2))."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def conv(prefix, c1, c2, k):
        fan_in = c1 * k * k
        bound = 1.0 / math.sqrt(fan_in) * math.sqrt(3.0)  # kaiming_uniform(a=sqrt(5)) -> sqrt(1/fan_in) * sqrt(3)... gain-adjusted
        sd[prefix + "conv.weight"] = (torch.rand(c2, c1, k, k, generator=g) * 2 - 1) * bound
        sd[prefix + "bn.weight"] = 1.0 + 0.1 * torch.randn(c2, generator=g)
        sd[prefix + "bn.bias"] = 0.1 * torch.randn(c2, generator=g)
        sd[prefix + "bn.running_mean"] = 0.1 * torch.randn(c2, generator=g)
        sd[prefix + "bn.running_var"] = 1.0 + 0.2 * torch.rand(c2, generator=g)

    for i, kind, args, _f in LAYERS:
        p = f"model.{i}."
        if kind == "Conv":
            c1, c2, k, _s = args
            conv(p, c1, c2, k)
        elif kind == "C3":
            c1, c2, n, _sc = args
            c_ = c2 // 2
            conv(p + "cv1.", c1, c_, 1)
            conv(p + "cv2.", c1, c_, 1)
            conv(p + "cv3.", 2 * c_, c2, 1)
            for j in range(n):
                conv(p + f"m.{j}.cv1.", c_, c_, 1)
                conv(p + f"m.{j}.cv2.", c_, c_, 3)
        elif kind == "SPPF":
            c1, c2, _k = args
            conv(p + "cv1.", c1, c1 // 2, 1)
            conv(p + "cv2.", c1 // 2 * 4, c2, 1)
    no = 5 + nc
    for l, (c, s) in enumerate(zip(DETECT_CH, STRIDES)):
        bound = 1.0 / math.sqrt(c)
        sd[f"model.24.m.{l}.weight"] = (torch.rand(3 * no, c, 1, 1, generator=g) * 2 - 1) * bound
        b = ((torch.rand(3 * no, generator=g) * 2 - 1) * bound).view(3, no)
        b[:, 4] += math.log(8 / (640 / s) ** 2) if obj_bias is None else obj_bias
        b[:, 5:] += math.log(0.6 / (nc - 0.99999))
        sd[f"model.24.m.{l}.bias"] = b.view(-1)
    sd["model.24.anchors"] = torch.tensor(ANCHORS, dtype=torch.float32) / torch.tensor(STRIDES, dtype=torch.float32).view(3, 1, 1)
    return sd


def count_parameters(sd) -> int:
    return sum(v.numel() for k, v in sd.items() if k.endswith(("conv.weight", "bn.weight", "bn.bias")) or ".m." in k and k.startswith("model.24"))


def _conv(sd, prefix, x, k, s, dtype):
    w = sd[prefix + "conv.weight"].to(dtype)
    y = F.conv2d(x, w, None, stride=s, padding=2 if k == 6 else k // 2)  # layer 0 is Conv(3, 32, k6, s2, p2)
    y = F.batch_norm(y, sd[prefix + "bn.running_mean"].to(dtype), sd[prefix + "bn.running_var"].to(dtype),
                     sd[prefix + "bn.weight"].to(dtype), sd[prefix + "bn.bias"].to(dtype), False, 0.0, BN_EPS)
    return F.silu(y)


def _c3(sd, p, x, n, shortcut, dtype):
    y1 = _conv(sd, p + "cv1.", x, 1, 1, dtype)
    for j in range(n):
        t = _conv(sd, p + f"m.{j}.cv2.", _conv(sd, p + f"m.{j}.cv1.", y1, 1, 1, dtype), 3, 1, dtype)
        y1 = y1 + t if shortcut else t
    y2 = _conv(sd, p + "cv2.", x, 1, 1, dtype)
    return _conv(sd, p + "cv3.", torch.cat((y1, y2), 1), 1, 1, dtype)


def _sppf(sd, p, x, k, dtype):
    x = _conv(sd, p + "cv1.", x, 1, 1, dtype)
    y1 = F.max_pool2d(x, k, 1, k // 2)
    y2 = F.max_pool2d(y1, k, 1, k // 2)
    y3 = F.max_pool2d(y2, k, 1, k // 2)
    return _conv(sd, p + "cv2.", torch.cat((x, y1, y2, y3), 1), 1, 1, dtype)


def yolov5s_forward(sd, x: torch.Tensor, dtype=torch.float32, return_raw: bool = False):
    """x f32 [B,3,H,W] (RGB, 0..1; H, W multiples of 32) -> [B, sum(3*ny*nx), 5+nc] decoded predictions
    (xywh in input pixels, obj, class probabilities) exactly as ultralytics' Detect in inference mode."""
    x = x.to(dtype)
    outs = []
    for i, kind, args, f in LAYERS:
        p = f"model.{i}."
        if kind == "Conv":
            _c1, _c2, k, s = args
            x = _conv(sd, p, x, k, s, dtype)
        elif kind == "C3":
            _c1, _c2, n, sc = args
            x = _c3(sd, p, x, n, sc, dtype)
        elif kind == "SPPF":
            x = _sppf(sd, p, x, args[2], dtype)
        elif kind == "Up":
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        elif kind == "Cat":
            x = torch.cat([x if j == -1 else outs[j] for j in f], 1)
        outs.append(x)
    nc = sd["model.24.m.0.bias"].numel() // 3 - 5
    no = 5 + nc
    z, raw = [], []
    for l, (src, stride) in enumerate(zip(DETECT_FROM, STRIDES)):
        t = F.conv2d(outs[src], sd[f"model.24.m.{l}.weight"].to(dtype), sd[f"model.24.m.{l}.bias"].to(dtype))
        bs, _, ny, nx = t.shape
        t = t.view(bs, 3, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        raw.append(t)
        yv, xv = torch.meshgrid(torch.arange(ny, dtype=dtype, device=t.device), torch.arange(nx, dtype=dtype, device=t.device), indexing="ij")
        grid = torch.stack((xv, yv), 2).expand(1, 3, ny, nx, 2) - 0.5
        anchor = (sd["model.24.anchors"][l].to(dtype) * stride).view(1, 3, 1, 1, 2).expand(1, 3, ny, nx, 2)
        y = t.sigmoid()
        xy = (y[..., 0:2] * 2 + grid) * stride
        wh = (y[..., 2:4] * 2) ** 2 * anchor
        z.append(torch.cat((xy, wh, y[..., 4:]), 4).view(bs, 3 * ny * nx, no))
    out = torch.cat(z, 1)
    return (out, raw) if return_raw else out


# ------------------------------------------------------------------ letterbox (localizer_engine.py:107-138)
def letterbox_geometry(h: int, w: int, new_shape=(640, 640)):
    """-> (r, new_unpad (w,h), top, bottom, left, right) with auto=False, scaleup=True, like the reference call
    `letterbox(im0, input_shape, stride=32, auto=False)` (localizer_engine.py:79)."""
    r = min(new_shape[0] / h, new_shape[1] / w)
    new_unpad = int(round(w * r)), int(round(h * r))
    dw, dh = (new_shape[1] - new_unpad[0]) / 2, (new_shape[0] - new_unpad[1]) / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return r, new_unpad, top, bottom, left, right


def load_localizer_img_from_array(im_bgr: np.ndarray, input_shape=(640, 640)) -> np.ndarray:
    """localizer_engine.py:75-85 minus cv2.imread: BGR u8 HWC -> f32 [1,3,H,W] RGB / 255, letterboxed (114 grey).
    Uses cv2.resize(INTER_LINEAR) like the reference (OpenCV is third-party to both)."""
    import cv2

    h, w = im_bgr.shape[:2]
    _r, new_unpad, top, bottom, left, right = letterbox_geometry(h, w, input_shape)
    im = im_bgr
    if (w, h) != new_unpad:
        im = cv2.resize(im, new_unpad, interpolation=cv2.INTER_LINEAR)
    im = cv2.copyMakeBorder(im, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))
    im = np.ascontiguousarray(im.transpose((2, 0, 1))[::-1]).astype(np.float32) / 255.0
    return im[None]


def cv2_resize_linear_u8(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR) for 8-bit images, restated in OpenCV's own
    fixed-point arithmetic (OpenCV 4.x imgproc/resize.cpp: resizeGeneric_ tap computation, HResizeLinear,
    VResizeLinear<uchar,int,short> with FixedPtCast<int,uchar,22>) -- the third-party call the reference's letterbox
    makes (localizer_engine.py:125).  Pinned against live cv2 in tests/test_oracle.py; this is what the device
    letterbox kernel (csrc/crop.cu letterbox_resize_kernel) must equal bit for bit."""
    h, w = img.shape[:2]

    def taps(src, dst, clamp_fraction):
        scale = 1.0 / (float(dst) / float(src))
        f = ((np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(np.float32)).astype(np.float32)
        if clamp_fraction:  # x: fraction zeroed where the window leaves the image; y: rows are clipped instead
            f[(s < 0) | (s >= src - 1)] = 0
        i0, i1 = np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1)
        if clamp_fraction:
            i0 = np.where(s >= src - 1, src - 1, i0)
        c0 = np.clip(np.rint((np.float32(1) - f) * np.float32(2048)), -32768, 32767).astype(np.int64)
        c1 = np.clip(np.rint(f * np.float32(2048)), -32768, 32767).astype(np.int64)
        return i0, i1, c0, c1

    x0, x1, a0, a1 = taps(w, new_w, True)
    y0, y1, b0, b1 = taps(h, new_h, False)
    src = img.astype(np.int64)
    rows0 = src[y0][:, x0] * a0[None, :, None] + src[y0][:, x1] * a1[None, :, None]
    rows1 = src[y1][:, x0] * a0[None, :, None] + src[y1][:, x1] * a1[None, :, None]
    out = (((b0[:, None, None] * (rows0 >> 4)) >> 16) + ((b1[:, None, None] * (rows1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def load_localizer_img_restated(im_bgr: np.ndarray, input_shape=(640, 640)) -> np.ndarray:
    """load_localizer_img_from_array with cv2.resize replaced by cv2_resize_linear_u8 (no OpenCV call at all)."""
    h, w = im_bgr.shape[:2]
    _r, new_unpad, top, _bottom, left, _right = letterbox_geometry(h, w, input_shape)
    im = im_bgr if (w, h) == new_unpad else cv2_resize_linear_u8(im_bgr, new_unpad[0], new_unpad[1])
    canvas = np.full((input_shape[0], input_shape[1], 3), 114, dtype=np.uint8)
    canvas[top:top + im.shape[0], left:left + im.shape[1]] = im
    return (np.ascontiguousarray(canvas.transpose((2, 0, 1))[::-1]).astype(np.float32) / 255.0)[None]


# ------------------------------------------------------------------ NMS (localizer_engine.py:171-277)
def nms_greedy(boxes: torch.Tensor, scores: torch.Tensor, iou_thres: float) -> torch.Tensor:
    """torchvision.ops.nms semantics (its CPU kernel, fp32 op for op): visit boxes by descending score (ties:
    lower index first), keep a box unless its IoU with an already kept box is strictly greater than iou_thres;
    IoU = inter / (area_i + area_j - inter) with inter = max(0, xx2 - xx1) * max(0, yy2 - yy1)."""
    order = torch.sort(scores, descending=True, stable=True).indices.numpy()
    b = boxes.numpy().astype(np.float32)[order]
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = ((x2 - x1) * (y2 - y1)).astype(np.float32)
    n = len(order)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float32(iou_thres)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(int(order[i]))
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(np.float32(0), (xx2 - xx1).astype(np.float32))
        h = np.maximum(np.float32(0), (yy2 - yy1).astype(np.float32))
        inter = (w * h).astype(np.float32)
        ovr = (inter / ((areas[i] + areas[i + 1:]).astype(np.float32) - inter).astype(np.float32)).astype(np.float32)
        suppressed[i + 1:] |= ovr > thr
    return torch.tensor(keep, dtype=torch.int64)


def non_max_suppression(prediction: torch.Tensor, conf_thres=0.25, iou_thres=0.45, max_det=300, max_wh=7680, max_nms=30000):
    """Restates the reference's single-label, class-aware path: obj > conf -> cls *= obj -> xywh -> xyxy ->
    best class, conf > thr -> sort by conf desc -> class-offset greedy NMS -> cap max_det.
    -> list of [n, 6] (x1, y1, x2, y2, conf, cls) per image."""
    out = []
    for x in prediction:
        x = x[x[:, 4] > conf_thres].clone()
        if not x.shape[0]:
            out.append(torch.zeros((0, 6)))
            continue
        x[:, 5:] *= x[:, 4:5]
        box = x[:, :4].clone()
        box[:, 0] = x[:, 0] - x[:, 2] / 2
        box[:, 1] = x[:, 1] - x[:, 3] / 2
        box[:, 2] = x[:, 0] + x[:, 2] / 2
        box[:, 3] = x[:, 1] + x[:, 3] / 2
        conf, j = x[:, 5:].max(1, keepdim=True)
        x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        n = x.shape[0]
        if not n:
            out.append(torch.zeros((0, 6)))
            continue
        x = x[torch.sort(x[:, 4], descending=True, stable=True).indices[:max_nms]]
        c = x[:, 5:6] * max_wh
        keep = nms_greedy(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        out.append(x[keep])
    return out
