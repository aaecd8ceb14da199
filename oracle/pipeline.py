"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the whole hot path for one batch of line images, stage by
stage in the reference's order and arithmetic:

    letterbox            onnx_engines/localizer_engine.py:75-85,107-138   (oracle/yolo.py, OpenCV-exact restatement)
    YOLOv5s forward      :49-55 (onnxruntime)                              (oracle/yolo.py, fp32)
    NMS                  :171-277                                          (oracle/yolo.py)
    box ordering etc.    infer_effocr_onnx_multi.py:256-322 ("onnx" convention)
                         infer_effocr.py:261-302 behind an mmdet-style localizer ("torch" convention)
    per-crop transform   utils/datasets_utils.py:166-172                   (oracle/transform.py)
    encoder + L2 norm    models/encoders.py:58, infer_effocr.py:314-316    (oracle/vit.py, fp32)
    kNN                  faiss IndexFlatIP via FaissKNN                    (oracle/knn.py)
    decode + postprocess infer_effocr_onnx_multi.py:373-393 / infer_effocr.py:318-341

The string rules (`en_preprocess`, `en_postprocess`) are the product's host functions in effocr_b200/textproc.py,
which tests/test_oracle.py pins against the live reference functions; everything numeric is the oracle's.
Pinned end to end by tests/golden/driver_golden.json (the unmodified reference drivers' own output on the same job,
oracle/make_driver_golden.py; tests/test_oracle.py::test_oracle_pipeline_matches_reference_driver_golden).
"""
from __future__ import annotations

import numpy as np
import torch

from . import knn as OK, transform as OT, vit as OV, yolo as OY


def localize(lines_rgb, ysd, conf_thres=0.35, iou_thres=0.01, input_shape=(640, 640), batch=8):
    """-> list of [n, 6] float32 tensors (letterbox pixels, confidence-sorted), one per line."""
    out = []
    for i0 in range(0, len(lines_rgb), batch):
        x = np.concatenate([OY.load_localizer_img_restated(np.ascontiguousarray(im[:, :, ::-1]), input_shape)
                            for im in lines_rgb[i0:i0 + batch]], 0)
        with torch.no_grad():
            pred = OY.yolov5s_forward(ysd, torch.from_numpy(x))
        out += OY.non_max_suppression(pred, conf_thres=conf_thres, iou_thres=iou_thres, max_det=1000)
    return out


def unletterbox(det, h, w, input_shape=(640, 640)):
    r, _unpad, top, _b, left, _r = OY.letterbox_geometry(h, w, input_shape)
    det = np.array(det, dtype=np.float32, copy=True)
    gain = np.float32(r)
    det[:, [0, 2]] = np.clip((det[:, [0, 2]] - np.float32(left)) / gain, 0, np.float32(w))
    det[:, [1, 3]] = np.clip((det[:, [1, 3]] - np.float32(top)) / gain, 0, np.float32(h))
    return det


def boxes_for_line(det, h, w, convention="onnx", score_thresh=0.3, score_thresh_word=0.3):
    """-> (sorted char boxes, word_end_idx, crop rectangles, heights, bottoms) with the scalar reference loops."""
    from effocr_b200 import textproc
    det = np.asarray(det, dtype=np.float32).reshape(-1, 6)
    if convention == "torch":
        det = unletterbox(det, h, w)
        chars = [list(r[:5]) for r in det if r[5] == 0]
        words = [list(r[:5]) for r in det if r[5] == 1]
        sc, wei = textproc.en_preprocess(chars, words, score_thresh=score_thresh, score_thresh_word=score_thresh_word)
        rects = [OT.crop_rect_torch_path(b, h, w) for b in sc]
    else:
        chars = [r[:4] for r in det if r[5] == 0]
        words = [r[:4] for r in det if r[5] == 1]
        if not chars:
            return [], [], [], [], []
        sc, wei = textproc.en_preprocess(chars, words)
        rects = [OT.crop_rect_onnx_path(b, h, w) for b in sc]
    return sc, wei, rects, [b[3] - b[1] for b in sc], [b[3] for b in sc]


def run(lines_rgb, ysd, vsd, xb, chars, conf_thres=0.35, iou_thres=0.01, convention="onnx", k=1, score_thresh=0.3,
        score_thresh_word=0.3, enc_batch=256):
    """-> one dict per line: text, nns, rects, char_boxes, word_end_idx, margin (top-1 / top-2 gap per character)."""
    from effocr_b200 import textproc
    dets = localize(lines_rgb, ysd, conf_thres=conf_thres, iou_thres=iou_thres)
    per_line, crops = [], []
    for im, det in zip(lines_rgb, dets):
        h, w = im.shape[:2]
        sc, wei, rects, heights, bottoms = boxes_for_line(det.numpy(), h, w, convention, score_thresh, score_thresh_word)
        usable = []
        for r in rects:
            c = OT.numpy_slice(im, r)
            usable.append(c.size > 0)
            crops.append(c if c.size > 0 else None)
        per_line.append((sc, wei, rects, heights, bottoms, usable))
    # the ONNX driver feeds a zero image for a crop its transform rejects (infer_effocr_onnx_multi.py:151-152,200-204)
    x = np.stack([OT.paired_transform(c) if c is not None else np.zeros((3, 224, 224), np.float32) for c in crops]) if crops else None
    embs = []
    with torch.no_grad():
        for i0 in range(0, len(crops), enc_batch):
            embs.append(OV.l2_normalize(OV.vit_forward(vsd, torch.from_numpy(x[i0:i0 + enc_batch]))))
    emb = torch.cat(embs) if embs else torch.zeros((0, xb.shape[1]))
    _d, idx = OK.flat_ip_search(xb, emb, k)
    _s, margin = OK.margins(xb, emb, 1) if len(emb) else (None, torch.zeros(0))
    out, pos = [], 0
    for (sc, wei, rects, heights, bottoms, _usable) in per_line:
        n = len(rects)
        if n == 0:
            out.append({"text": None, "nns": [], "rects": [], "char_boxes": [], "word_end_idx": [], "margin": []})
            continue
        rows = idx[pos:pos + n].tolist()
        nearest = [[chars[j] for j in row if j >= 0] for row in rows]
        text = textproc.en_postprocess("".join(c[0] for c in nearest).strip(), wei, heights, bottoms)
        out.append({"text": text, "nns": ["".join(c).strip() for c in nearest], "rects": rects, "char_boxes": [list(map(float, b[:4])) for b in sc],
                    "word_end_idx": wei, "margin": margin[pos:pos + n].tolist()})
        pos += n
    return out
