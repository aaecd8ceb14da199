"""ncu / timing target: one YOLOv5s forward (64 letterboxed 640x640 lines) + NMS."""
import sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops, synth
from effocr_b200.localizer_engine import EffLocalizer, nms_device
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lines = [l[0] for l in synth.synthetic_lines(B, seed=0)]
prec = sys.argv[2] if len(sys.argv) > 2 else "split"
ysd = synth.random_yolov5s_state_dict(nc=2, seed=0, obj_bias=-0.5)
loc = EffLocalizer(ysd, iou_thresh=0.01, conf_thresh=0.35, input_shape=(640, 640), max_batch=B, precision=prec)
px, im, _ = ops.pack_images(lines)
x = ops.letterbox_resize(px, im, [c.shape[:2] for c in lines], 640, 640)
for _ in range(2):
    pred = loc._eng_net.forward(x)
    nms_device(pred, 0.3, 0.01)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    pred = loc._eng_net.forward(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"yolo forward B={B} precision={prec}: {ms:.2f} ms  ({B/ms*1e3:.0f} lines/s, {15.7626368e9*B/ms/1e9:.0f} TFLOP/s)")
