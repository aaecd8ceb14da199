"""Small end-to-end target for compute-sanitizer: ViT-S recognizer (fused proj+LN, fused MLP, one-pass attention,
class-token-only last block) + kNN on 12 crops, and the YOLOv5s localizer pipeline on 2 lines."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from effocr_b200 import synth
from effocr_b200.encoders import TimmViTParams
from effocr_b200.infer import EffOCRPipeline
from effocr_b200.localizer_engine import EffLocalizer
from effocr_b200.pipeline import RecognizerPipeline

torch.manual_seed(0)
net = TimmViTParams("vit_small_patch16_224")
vsd = {"net." + k: v.detach().clone() for k, v in net.state_dict().items()}
index = torch.nn.functional.normalize(torch.randn(300, 384), dim=1)
rec = RecognizerPipeline(vsd, index, max_batch=16)
crops, _ = synth.synthetic_crops(12, seed=0)
d, i = rec.recognize_crops(crops, k=5)
assert np.isfinite(d).all() and (i >= 0).all()
loc = EffLocalizer(synth.random_yolov5s_state_dict(nc=2, seed=0, obj_bias=2.0), iou_thresh=0.01, conf_thresh=0.3, input_shape=(640, 640), max_batch=2)
pipe = EffOCRPipeline(loc, rec, [chr(33 + k % 94) for k in range(300)], lang="en", knn=1)
lines = [l[0] for l in synth.synthetic_lines(2, seed=1)]
res = pipe.infer_lines(lines)
print("sanitizer target ok:", [len(r["char_boxes"]) for r in res])
