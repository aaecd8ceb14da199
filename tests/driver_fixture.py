"""TEST INFRASTRUCTURE -- the on-disk job both reference drivers expect (image dir + COCO json + recognizer dir +
localizer dir), built deterministically from the committed quick-fit weights.  oracle/make_driver_golden.py runs the
UNMODIFIED reference scripts over it with oracle-backed back-ends (CPU) and stores their transcriptions in
tests/golden/driver_golden.json; the GPU tests run effocr_b200 over the same job and must reproduce them.

Layout (what infer_effocr.py:516-534 / infer_effocr_onnx_multi.py:470-505 read):
    images/line_000.png ...            synthetic 64 x 1024 lines (effocr_b200.synth, seeded)
    coco.json                          {"images": [{"file_name", "height", "width", "id", "text"}], ...}
    recognizer/enc_best.pth            timm-keyed `net.*` state dict (quick-fit ViT-S)
    recognizer/enc_best.onnx           placeholder: the ONNX driver asserts it exists; weights come from the sibling .pth
    recognizer/ref.index, ref.txt      faiss IndexFlatIP file + one character per line
    localizer/best_bbox_mAP.pth        ultralytics-keyed YOLOv5s state dict (quick-fit)
    localizer/best_bbox_mAP.onnx       placeholder (as above)
    localizer/config.py                placeholder mmdetection config (infer_effocr.py:519-521 globs for a *.py)
"""
from __future__ import annotations

import json
import os
import struct
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
VIT_WEIGHTS = GOLDEN / "quickfit_vit_small.npz"
YOLO_WEIGHTS = GOLDEN / "quickfit_yolov5s.npz"
INDEX_VECTORS = GOLDEN / "driver_ref_index.npy"
DRIVER_GOLDEN = GOLDEN / "driver_golden.json"
N_LINES, SEED, TRACKING = 24, 31337, 4.0  # letter-spacing: the reference NMS runs at IoU 0.01


def available() -> bool:
    return VIT_WEIGHTS.exists() and YOLO_WEIGHTS.exists()


def load_npz_state(path):
    return {k: torch.from_numpy(v.astype(np.float32)) for k, v in np.load(path).items()}


def prototype_crops():
    """One canonical render per printable-ASCII glyph (font size 40), cut like the pipeline cuts characters."""
    from effocr_b200 import synth
    out = []
    for ch in synth.ASCII_GLYPHS:
        im, cb, _wb, _chars = synth.render_line(ch, font_size=40, x0=6, width=128)
        out.append(np.ascontiguousarray(im[:, int(round(float(cb[0][0]))):int(round(float(cb[0][2]))), :]))
    return out


def write_flat_ip_index(path, xb):
    """faiss IndexFlatIP layout, spelled out independently of the product's writer (faiss index_write.cpp)."""
    xb = np.ascontiguousarray(xb, dtype="<f4")
    n, d = xb.shape
    with open(path, "wb") as f:
        f.write(b"IxFI" + struct.pack("<iqqqBiQ", d, n, 1 << 20, 1 << 20, 1, 0, n * d) + xb.tobytes())


def build(root, index_vectors=None, n_lines: int = N_LINES, seed: int = SEED):
    """-> dict(images=[paths], coco_json, recognizer_dir, localizer_dir, lines=[synth tuples])."""
    from PIL import Image

    from effocr_b200 import synth
    root = Path(root)
    (root / "images").mkdir(parents=True, exist_ok=True)
    (root / "recognizer").mkdir(exist_ok=True)
    (root / "localizer").mkdir(exist_ok=True)
    lines = synth.synthetic_lines(n_lines, seed=seed, tracking=TRACKING)
    coco = {"info": {"": ""}, "licenses": [{"": ""}], "images": [], "annotations": [], "categories": [{"id": 0, "name": "char"}]}
    paths = []
    for i, (img, _cb, _wb, chars) in enumerate(lines):
        name = f"line_{i:03d}.png"
        Image.fromarray(img).save(root / "images" / name)
        coco["images"].append({"file_name": name, "height": int(img.shape[0]), "width": int(img.shape[1]), "id": i, "text": "".join(chars)})
        paths.append(str(root / "images" / name))
    with open(root / "coco.json", "w") as f:
        json.dump(coco, f)
    torch.save(load_npz_state(VIT_WEIGHTS), root / "recognizer" / "enc_best.pth")
    (root / "recognizer" / "enc_best.onnx").write_bytes(b"placeholder: the weights are in enc_best.pth")
    if index_vectors is None:
        index_vectors = np.load(INDEX_VECTORS)
    write_flat_ip_index(root / "recognizer" / "ref.index", index_vectors)
    with open(root / "recognizer" / "ref.txt", "w") as f:
        f.write("\n".join(synth.ASCII_GLYPHS) + "\n")
    torch.save(load_npz_state(YOLO_WEIGHTS), root / "localizer" / "best_bbox_mAP.pth")
    (root / "localizer" / "best_bbox_mAP.onnx").write_bytes(b"placeholder: the weights are in best_bbox_mAP.pth")
    (root / "localizer" / "config.py") .write_text("# placeholder for the mmdetection config infer_effocr.py globs for\n")
    return {"images": paths, "coco_json": str(root / "coco.json"), "recognizer_dir": str(root / "recognizer"),
            "localizer_dir": str(root / "localizer"), "image_dir": str(root / "images"), "lines": lines}
