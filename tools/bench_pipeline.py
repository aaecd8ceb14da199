"""Phase-by-phase timing of the whole hot path (config C3: YOLOv5s localizer + ViT-S recognizer + kNN) on synthetic lines."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops, synth
from effocr_b200.encoders import TimmViTParams
from effocr_b200.infer import EffOCRPipeline, run_effocr
from effocr_b200.localizer_engine import EffLocalizer, nms_device
from effocr_b200.pipeline import RecognizerPipeline

n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
net = TimmViTParams("vit_small_patch16_224")
vsd = {"net." + k: v.detach().clone() for k, v in net.state_dict().items()}
ysd = synth.background_suppressed_yolo_state(nc=2, seed=0)
g = torch.Generator().manual_seed(1)
index = torch.nn.functional.normalize(torch.randn(10000, 384, generator=g), dim=1)
loc = EffLocalizer(ysd, iou_thresh=0.01, conf_thresh=0.35, input_shape=(640, 640), max_batch=batch_lines)
rec = RecognizerPipeline(vsd, index, max_batch=4096)
pipe = EffOCRPipeline(loc, rec, [chr(33 + i % 94) for i in range(10000)], lang="en", knn=1)
t0 = time.perf_counter()
lines = [l[0] for l in synth.synthetic_lines(n_lines, seed=0)]
print(f"rendered {n_lines} lines in {time.perf_counter()-t0:.2f}s", flush=True)


def sync():
    torch.cuda.synchronize()


# ---- calibrate the confidence threshold so that ~32 boxes per line survive NMS
chunk = lines[:batch_lines]
px, im, _ = ops.pack_images(chunk)
x = ops.letterbox_resize(px, im, [c.shape[:2] for c in chunk], 640, 640)
pred = loc._eng_net.forward(x)
lo, hi = 0.001, 0.999
for _ in range(18):
    mid = 0.5 * (lo + hi)
    o_, cnt = nms_device(pred, mid, 0.01)
    live_ = torch.arange(o_.shape[1], device=o_.device)[None, :] < cnt[:, None]
    if float(((o_[:, :, 5] == 0) & live_).sum()) / len(chunk) > 32:
        lo = mid
    else:
        hi = mid
loc._conf_thresh = hi
_, cnt = nms_device(pred, hi, 0.01)
print(f"calibrated conf_thresh {hi:.4f}: {float(cnt.float().mean()):.1f} boxes/line after NMS", flush=True)

run_effocr(lines[:batch_lines], pipe, batch_lines=batch_lines)  # warm-up
sync()
t0 = time.perf_counter()
out = run_effocr(lines, pipe, batch_lines=batch_lines)
sync()
dt = time.perf_counter() - t0

# ---- the same phases one by one
T = dict(pack=0.0, letterbox=0.0, yolo=0.0, nms=0.0, d2h=0.0, host_boxes=0.0, recognize=0.0, decode=0.0)
ncrops = 0
for i0 in range(0, n_lines, batch_lines):
    chunk = lines[i0:i0 + batch_lines]
    sync(); a = time.perf_counter()
    packed = ops.pack_images(chunk); sync(); b = time.perf_counter(); T["pack"] += b - a
    x = ops.letterbox_resize(packed[0], packed[1], [c.shape[:2] for c in chunk], 640, 640); sync(); c = time.perf_counter(); T["letterbox"] += c - b
    pred = loc._eng_net.forward(x); sync(); d = time.perf_counter(); T["yolo"] += d - c
    o, cnt = nms_device(pred, loc._conf_thresh, 0.01); sync(); e = time.perf_counter(); T["nms"] += e - d
    o, cnt = o.cpu(), cnt.cpu().tolist(); f = time.perf_counter(); T["d2h"] += f - e
    dets = [o[i, :cnt[i]] for i in range(len(chunk))]
    rects = []
    for li, (img, det) in enumerate(zip(chunk, dets)):
        char_b, wei, r, hs, bs = pipe._boxes_for_line(det, img.shape[0], img.shape[1])
        rects += [(li,) + q for q in r]
    gg = time.perf_counter(); T["host_boxes"] += gg - f
    if rects:
        boxes, n = ops.pack_boxes(rects)
        dist, idx, _ = rec.recognize_device(packed[0], packed[1], boxes, n, 1)
        idx = idx.cpu().numpy()
        ncrops += n
    sync(); h = time.perf_counter(); T["recognize"] += h - gg
print(f"end-to-end run_effocr: {n_lines} lines, {ncrops} char crops in {dt*1e3:.1f} ms -> {n_lines/dt:.0f} lines/s, {ncrops/dt:.0f} crops/s")
print("phases (ms, serialised with syncs): " + ", ".join(f"{k} {v*1e3:.1f}" for k, v in T.items()))
yolo_flops = 15.7626368e9 * n_lines
print(f"yolo forward: {n_lines/T['yolo']:.0f} lines/s, {yolo_flops/T['yolo']/1e12:.1f} TFLOP/s; recognizer: {ncrops/T['recognize']:.0f} crops/s")
