"""Caller side of the hot path: where the text-line images come from and where the
transcriptions go.

The reference drivers enumerate line images from a COCO JSON or an image directory, decode every line twice on the
main thread (cv2 for the localizer, PIL for the crops: infer_effocr_onnx_multi.py:181,307; infer_effocr.py:262,280) and
write `inference_results.json` / `inference_coco.json` (infer_effocr.py:542-578, infer_effocr_onnx_multi.py:516-526).
Here a small thread pool decodes each line ONCE, one batch ahead of the GPU (`LineDecoder`; PIL releases the GIL while
it inflates a PNG), `EffOCRPipeline.infer_batches` consumes the batches lazily, and the writers reproduce the reference's
files.  Nothing here touches the device: it is host plumbing around `effocr_b200.infer`.
"""
from __future__ import annotations

import copy
import glob
import json
import os
import shutil
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import textproc

# utils/coco_utils.py:1-7
COCO_JSON_SKELETON = {"info": {"": ""}, "licenses": [{"": ""}], "images": [], "annotations": [],
                      "categories": [{"id": 0, "name": "char"}]}


def create_coco_anno_entry(x, y, w, h, ann_id, image_id, cat_id=0, text=None):
    """One COCO annotation with the reference's fields and types (utils/coco_utils.py:10-26): integer box and
    rectangle polygon, `area` from the un-truncated w*h, score 1.0, `text` only when given."""
    xi, yi, wi, hi = int(x), int(y), int(w), int(h)
    entry = {"segmentation": [[xi, yi, xi + wi, yi, xi + wi, yi + hi, xi, yi + hi]], "area": w * h, "iscrowd": 0,
             "image_id": image_id, "bbox": [xi, yi, wi, hi], "category_id": cat_id, "id": ann_id, "score": 1.0}
    if text is not None:
        entry["text"] = text
    return entry


def create_coco_image_entry(path, height, width, image_id, text=None):
    """utils/coco_utils.py:29-44."""
    entry = {"file_name": path, "height": height, "width": width, "id": image_id}
    if text is not None:
        entry["text"] = text
    return entry


def list_images(image_dir: str, coco_json: str | None = None):
    """-> (paths, coco or None).  With a COCO file: its `images[*].file_name` joined to image_dir, in file order
    (infer_effocr.py:470-473); without: every *.png, then every *.jpg below image_dir, recursively
    (`--infer_over_img_dir`, infer_effocr_onnx_multi.py:465-467)."""
    if coco_json is not None:
        with open(coco_json) as f:
            coco = json.load(f)
        return [os.path.join(image_dir, im["file_name"]) for im in coco["images"]], coco
    paths = glob.glob(os.path.join(image_dir, "**/*.png"), recursive=True)
    paths += glob.glob(os.path.join(image_dir, "**/*.jpg"), recursive=True)
    return paths, None


def decode_rgb(path: str) -> np.ndarray:
    """u8 [H, W, 3] RGB, the pixels both reference decoders deliver for the lossless formats EffOCR line crops are
    stored in (PIL `convert("RGB")`, infer_effocr.py:280; cv2.imread + BGR->RGB, localizer_engine.py:75-85)."""
    from PIL import Image
    with Image.open(path) as im:
        return np.ascontiguousarray(np.asarray(im.convert("RGB")))


class LineDecoder:
    """Iterate over (paths_of_batch, [u8 RGB arrays]) in input order.  `workers` threads decode; at most `prefetch`
    batches beyond the one handed out are in flight, so memory stays bounded for 100k-line jobs (config C5).  A decode
    error surfaces at the batch that contains the file, as the reference's would at that line."""

    def __init__(self, paths, batch_lines: int = 64, workers: int | None = None, prefetch: int = 2, decode=decode_rgb):
        if batch_lines <= 0:
            raise ValueError("batch_lines must be positive")
        self.paths = list(paths)
        self.batch_lines = batch_lines
        self.workers = workers or min(8, os.cpu_count() or 1)
        self.prefetch = max(int(prefetch), 0)
        self.decode = decode

    def __len__(self):
        return (len(self.paths) + self.batch_lines - 1) // self.batch_lines

    def __iter__(self):
        starts = iter(range(0, len(self.paths), self.batch_lines))
        with ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="effocr-decode") as pool:
            pending = deque()

            def submit_next():
                i0 = next(starts, None)
                if i0 is None:
                    return False
                chunk = self.paths[i0:i0 + self.batch_lines]
                pending.append((chunk, [pool.submit(self.decode, p) for p in chunk]))
                return True

            for _ in range(self.prefetch + 1):
                if not submit_next():
                    break
            while pending:
                chunk, futures = pending.popleft()
                images = [f.result() for f in futures]
                submit_next()
                yield chunk, images


IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")  # torchvision ImageFolder


def render_dataset_files(root_dir: str, font_name: str = ""):
    """The files `create_render_dataset(root_dir, lang, font_name=...)` embeds for an ad-hoc index
    (effocr_datasets/recognizer_datasets.py:213-223): a torchvision ImageFolder walk of root_dir -- class directories in
    sorted order, inside each every image file below it in sorted walk order -- keeping the paths that contain
    `font_name` and whose basename does not start with "PAIRED" (those are scanned crops, not renders)."""
    classes = sorted(e.name for e in os.scandir(root_dir) if e.is_dir())
    files = []
    for cls in classes:
        for base, _dirs, names in sorted(os.walk(os.path.join(root_dir, cls), followlinks=True)):
            for name in sorted(names):
                if name.lower().endswith(IMG_EXTENSIONS):
                    files.append(os.path.join(base, name))
    return [p for p in files if font_name in p and not os.path.basename(p).startswith("PAIRED")]


def chars_from_render_files(files):
    """One candidate character per render file (infer_effocr.py:197-198): `0x<hex>_...` names give chr(hex), any other
    name gives its first character."""
    out = []
    for p in files:
        name = os.path.basename(p)
        out.append(chr(int(name.split("_")[0], base=16)) if name.startswith("0x") else name[0])
    return out


def build_ad_hoc_index(root_dir: str, recognizer, lang: str = "en", font_name: str | None = None, decode=decode_rgb,
                       workers: int | None = None):
    """`--ad_hoc_index_root_dir` (infer_effocr.py:190-201): embed the rendered glyphs of one font on the device and make
    them the recognizer's index; returns the candidate characters.  Default font as in the reference:
    NotoSerifCJKjp-Regular for jp, NotoSerif-Regular otherwise.  `recognizer` is a `pipeline.RecognizerPipeline`."""
    if font_name is None:
        font_name = "NotoSerifCJKjp-Regular" if lang == "jp" else "NotoSerif-Regular"
    files = render_dataset_files(root_dir, font_name)
    if not files:
        raise ValueError(f"no rendered glyphs of font {font_name!r} below {root_dir}")
    chars = chars_from_render_files(files)
    glyphs = [im for _, images in LineDecoder(files, batch_lines=256, workers=workers, decode=decode) for im in images]
    recognizer.train_knn(glyphs, candidate_chars=chars)
    return chars


def mmdet_output_format(result):
    """Detectron2-style predictions -> the mmdet-style nesting the reference's pre-processing expects
    (infer_effocr.py:245-254): [[char_boxes]] or [[char_boxes, word_boxes]], each box [x0, y0, x1, y1, score];
    class 0 = character, class 1 = word."""
    inst = result[0]["instances"]
    classes = inst.pred_classes.tolist()
    boxes = inst.pred_boxes.tensor.tolist()
    scores = inst.scores.tolist()
    chars = [list(b) + [s] for b, s, c in zip(boxes, scores, classes) if c == 0]
    words = [list(b) + [s] for b, s, c in zip(boxes, scores, classes) if c == 1]
    return [[chars, words]] if words else [[chars]]


def transcribe_paths(paths, pipeline, batch_lines: int = 64, workers: int | None = None, prefetch: int = 2,
                     overlap: bool = True, decode=decode_rgb):
    """Decode, localise, recognise: -> one record (path, height, width, per-line result dict) per input line, in input
    order.  The decoder runs one batch ahead of `pipeline.infer_batches`, which pulls its batches lazily."""
    decoder = LineDecoder(paths, batch_lines=batch_lines, workers=workers, prefetch=prefetch, decode=decode)
    chunks, shapes = [], []

    def batches():
        for chunk, images in decoder:
            chunks.append(chunk)
            shapes.append([im.shape[:2] for im in images])
            yield images

    if hasattr(pipeline, "infer_batches"):
        results = pipeline.infer_batches(batches(), overlap=overlap)
    else:  # any object with infer_lines() works as a pipeline
        results = map(pipeline.infer_lines, batches())
    records = []
    for b, res in enumerate(results):
        records += [(path, int(h), int(w), r) for path, (h, w), r in zip(chunks[b], shapes[b], res)]
        chunks[b] = shapes[b] = None  # keep the bookkeeping of finished batches small
    return records


def assemble_results(records, lang: str = "en"):
    """Records in input order -> (inference_results {path: text}, inference_coco), the pair the reference's `run_effocr`
    returns (infer_effocr_onnx_multi.py:397).  Lines without a detected character are left out of inference_results, as
    `if output is None: continue` does (infer_effocr.py:551-552).  For Japanese the COCO structure carries one image
    entry per line (with its transcription) and one annotation per character box whose `text` is the string of its k
    nearest glyphs (infer_effocr.py:554-559); image ids count the recognised lines."""
    inference_results = {}
    inference_coco = copy.deepcopy(COCO_JSON_SKELETON)
    image_id = anno_id = 0
    for path, h, w, r in records:
        if r["text"] is None:
            continue
        if lang == "jp":
            inference_coco["images"].append(create_coco_image_entry(os.path.basename(path), h, w, image_id, text=r["text"]))
            for nn_chars, box in zip(r.get("nns", []), r.get("char_boxes", [])):
                x0, y0, x1, y1 = (int(round(v)) for v in box[:4])
                inference_coco["annotations"].append(
                    create_coco_anno_entry(x0, y0, x1 - x0, y1 - y0, anno_id, image_id, cat_id=0, text=nn_chars))
        inference_results[path] = r["text"]
        image_id += 1
    return inference_results, inference_coco


def run_effocr_paths(paths, pipeline, batch_lines: int = 64, workers: int | None = None, prefetch: int = 2,
                     overlap: bool = True, lang: str | None = None, decode=decode_rgb):
    """One process, one GPU: `assemble_results(transcribe_paths(...))`."""
    lang = lang if lang is not None else getattr(pipeline, "lang", "en")
    return assemble_results(transcribe_paths(paths, pipeline, batch_lines, workers, prefetch, overlap, decode), lang)


def run_effocr_paths_sharded(paths, pipeline, batch_lines: int = 64, workers: int | None = None, prefetch: int = 2,
                             overlap: bool = True, lang: str | None = None, decode=decode_rgb, weights=None, dst: int = 0):
    """Data-parallel over lines: rank r decodes and transcribes the paths
    `dist.shard_indices` gives it (strided, or balanced by `weights`, e.g. file sizes), the per-line records travel to
    rank `dst` in one `gather_object`, and the result pair is assembled there in INPUT order -- identical to the
    single-process output whatever the world size.  Other ranks return (None, None).  No steady-state collective."""
    import torch.distributed as tdist

    from . import dist as D

    rank, world, _ = D.init_from_env()
    lang = lang if lang is not None else getattr(pipeline, "lang", "en")
    paths = list(paths)
    mine = D.shard_indices(len(paths), rank, world, weights=weights)
    local = transcribe_paths([paths[i] for i in mine], pipeline, batch_lines, workers, prefetch, overlap, decode)
    if world == 1:
        return assemble_results(local, lang)
    parts = [None] * world if rank == dst else None
    tdist.gather_object(list(zip(mine, local)), parts, dst=dst)
    if rank != dst:
        return None, None
    merged = sorted((item for part in parts for item in part), key=lambda t: t[0])
    return assemble_results([rec for _, rec in merged], lang)


def save_output(save_dir: str, paths, inference_results, inference_coco, copy_images: bool = True) -> None:
    """`--save_output` (infer_effocr.py:566-575): <dir>/images/<basename> for every input line, inference_results.json
    and inference_coco.json, both indent=2.  The reference re-encodes each image through PIL; the files are copied
    here (same pixels, no decode)."""
    os.makedirs(os.path.join(save_dir, "images"), exist_ok=True)
    if copy_images:
        for p in paths:
            shutil.copyfile(p, os.path.join(save_dir, "images", os.path.basename(p)))
    with open(os.path.join(save_dir, "inference_results.json"), "w") as f:
        json.dump(inference_results, f, indent=2)
    with open(os.path.join(save_dir, "inference_coco.json"), "w") as f:
        json.dump(inference_coco, f, indent=2)


def gt_collect(results, gts):
    """[(file_name, ground truth)] + {file_name: prediction} -> [(ground truth, prediction or "")]
    (infer_effocr.py:82-90)."""
    pairs = []
    for fn, gt in gts:
        pred = results.get(fn)
        pairs.append((gt, "" if pred is None else pred))
    return pairs


def evaluate_against_coco(coco, inference_results, no_spaces_in_eval=False, norm_edit_distance=False, uncased=False):
    """The reference's closing block (infer_effocr.py:577-598): predictions re-keyed by basename, ground truth from
    `images[*].text`, -> (textline accuracy in percent, CER)."""
    by_name = {os.path.basename(k): v for k, v in inference_results.items()}
    gts = [(im["file_name"], im["text"]) for im in coco["images"]]
    return textproc.textline_evaluation(gt_collect(by_name, gts), print_incorrect=False, no_spaces_in_eval=no_spaces_in_eval,
                                        norm_edit_distance=norm_edit_distance, uncased=uncased)
