"""Where one 64-line batch of the whole path (config C3) spends its HOST time: wall-clock timers around the stages of
EffOCRPipeline.infer_batches, quick-fit weights, 640 lines.  Prints the median per batch in ms."""
import sys, time, statistics
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import driver_fixture as DF
from effocr_b200 import synth
from effocr_b200.infer import EffOCRPipeline
from effocr_b200.localizer_engine import EffLocalizer
from effocr_b200.pipeline import RecognizerPipeline

ysd, vsd = DF.load_npz_state(DF.YOLO_WEIGHTS), DF.load_npz_state(DF.VIT_WEIGHTS)
loc = EffLocalizer(ysd, iou_thresh=0.01, conf_thresh=0.35, input_shape=(640, 640), max_batch=64, precision="split")
protos = DF.prototype_crops()
rec = RecognizerPipeline(vsd, torch.zeros(1, 384), list(synth.ASCII_GLYPHS), max_batch=2048)
rec.train_knn(protos, list(synth.ASCII_GLYPHS))
pipe = EffOCRPipeline(loc, rec, list(synth.ASCII_GLYPHS), lang="en", knn=1)
lines = [l[0] for l in synth.synthetic_lines(640, seed=7, tracking=4.0)]
T = {}


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3); return r
    return w


pipe.localize = timed("stage1: localize (pack, H2D, YOLO+NMS wait, D2H)", pipe.localize)
pipe._boxes_for_line = timed("stage1: host box logic per line", pipe._boxes_for_line)
pipe.stage_localize = timed("stage1 total", pipe.stage_localize)
pipe.launch_recognize = timed("launch_recognize (enqueue)", pipe.launch_recognize)
pipe.finish_recognize = timed("finish_recognize (wait + decode)", pipe.finish_recognize)
for rep in range(2):
    T.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = sum(len(r) for r in pipe.infer_batches((lines[i:i + 64] for i in range(0, len(lines), 64)), overlap=True))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{n} lines in {dt*1e3:.1f} ms = {n/dt:.0f} lines/s, {dt*1e3/(len(lines)/64):.2f} ms per batch")
for k, v in T.items():
    per_batch = sum(v) / (len(lines) / 64)
    print(f"{k:55s} calls {len(v):5d}  per batch {per_batch:7.2f} ms   median call {statistics.median(v):.3f} ms")
