"""Line-image -> transcription orchestration: the B200-first counterpart of the reference's two drivers

  * `run_effocr(...)`      /root/reference/infer_effocr_onnx_multi.py:227-397  (YOLO + ViT + kNN, k = 1)
  * `EffOCR.infer(...)`    /root/reference/infer_effocr.py:257-343            (per line, k = self.knn)

The reference runs three globally barriered thread-pool phases (localize, transform, recognize) with four
host round trips per line.  Here every batch of lines makes ONE pass:

    host: letterbox (OpenCV, as the reference)            -> H2D  f32 [B,3,H,W]  + u8 line pixels
    GPU : YOLOv5s forward -> NMS                          -> D2H  boxes [B, <=1000, 6] (a few KB)
    host: per-line box ordering / word segmentation / crop-rectangle arithmetic (exact reference semantics)
                                                          -> H2D  crop rectangles (20 B each)
    GPU : fused crop+resize+normalise -> ViT -> L2 norm -> exact kNN
                                                          -> D2H  k ids (+ distances) per character
    host: id -> char decode, en_postprocess

Pixels, crops, embeddings and scores never leave the device.  Results are keyed and ordered by input, so they
are independent of batch composition and of how lines are sharded over ranks.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, textproc
from .localizer_engine import EffLocalizer
from .pipeline import PackedCrops, RecognizerPipeline


def crop_rect_onnx_path(bbox, im_height: int, im_width: int, vertical: bool):
    """infer_effocr_onnx_multi.py:311-318: torch.round (half-to-even) on the letterbox-space box, rescale the
    text-direction axis by (image extent / 640) in Python floats, int(round(.)), full extent on the other axis."""
    x0, y0, x1, y1 = torch.round(torch.as_tensor(bbox[:4], dtype=torch.float32))
    if vertical:
        return 0, int(round(y0.item() * im_height / 640)), im_width, int(round(y1.item() * im_height / 640))
    return int(round(x0.item() * im_width / 640)), 0, int(round(x1.item() * im_width / 640)), im_height


def crop_rects_onnx_path(bboxes, im_height: int, im_width: int, vertical: bool):
    """crop_rect_onnx_path for all boxes of a line at once: np.rint on float32 == torch.round (half to even), the
    rescale in float64 in the reference's operation order ((x * extent) / 640), np.rint == Python round()."""
    if len(bboxes) == 0:
        return []
    if isinstance(bboxes, torch.Tensor):
        b = bboxes.numpy()
    elif isinstance(bboxes, np.ndarray) and bboxes.ndim == 2:
        b = bboxes
    else:
        b = np.asarray([np.asarray(v, dtype=np.float32) for v in bboxes])
    k = (1, 3) if vertical else (0, 2)
    extent = im_height if vertical else im_width
    r = np.rint(b[:, k].astype(np.float32, copy=False)).astype(np.float64)
    lo, hi = np.rint(r * extent / 640).astype(np.int64).T.tolist()
    if vertical:
        return [(0, a, im_width, c) for a, c in zip(lo, hi)]
    return [(a, 0, c, im_height) for a, c in zip(lo, hi)]


def crop_rect_torch_path(bbox, im_height: int, im_width: int, vertical: bool, double_clipped: bool = True):
    """infer_effocr.py:286-291: int(round(.)) (banker's) on all four coordinates, then double clipping."""
    x0, y0, x1, y1 = map(int, map(round, (float(v) for v in bbox[:4])))
    if double_clipped:
        if vertical:
            x0, x1 = 0, im_width
        else:
            y0, y1 = 0, im_height
    return x0, y0, x1, y1


class EffOCRPipeline:
    """localizer (EffLocalizer) + recognizer (RecognizerPipeline) + candidate chars."""

    def __init__(self, localizer: EffLocalizer, recognizer: RecognizerPipeline, candidate_chars, lang: str = "en",
                 vertical: bool = False, knn: int = 1, anchor_margin=None, blacklist=None, box_convention: str = "onnx",
                 score_thresh: float = 0.5, score_thresh_word: float = 0.5):
        """box_convention: which of the reference's two drivers the host stage between the GPU phases follows.
        "onnx"  -- infer_effocr_onnx_multi.py:256-322: letterbox-pixel boxes, no score filter after NMS, crop columns
                   rescaled by (width / 640), heights / bottoms in letterbox pixels;
        "torch" -- infer_effocr.py:261-302 behind the mmdet.apis shim (effocr_b200/mmdet_shim.py): boxes mapped back to
                   ORIGINAL image pixels, `score > score_thresh` / `score_thresh_word` applied by the caller (:350-352),
                   all four coordinates int(round(.)) and double-clipped; the localizer should then be built with the
                   shim's low confidence floor (mmdet_shim.DEFAULT_CONF_FLOOR)."""
        if box_convention not in ("onnx", "torch"):
            raise ValueError("box_convention must be 'onnx' or 'torch'")
        self.box_convention = box_convention
        self.score_thresh = score_thresh
        self.score_thresh_word = score_thresh_word
        self.localizer = localizer
        self.recognizer = recognizer
        self.candidate_chars = list(candidate_chars)
        self.lang = lang
        self.vertical = vertical
        self.knn = knn
        self.anchor_margin = anchor_margin
        if blacklist:  # infer_effocr.py:209-212
            ids = np.array([self.candidate_chars.index(c) for c in blacklist])
            self.recognizer.index.remove_ids(ids)
            self.candidate_chars = [c for c in self.candidate_chars if c not in blacklist]

    # -- phase 1: localize a batch of RGB u8 line images
    def localize(self, images_rgb, packed=None):
        """The u8 lines go to the device once (`packed` = ops.pack_images result, shared with the crop kernel); the
        reference's letterbox (cv2 resize + pad + /255) runs there bit-exactly, so neither the 4.9 MB float tensor
        per line nor the OpenCV call is on the host path."""
        shape = self.localizer._input_shape
        if packed is None:
            packed = ops.pack_images(images_rgb)
        buf = getattr(self, "_letterbox_buf", None)
        if buf is None or buf.shape[0] < len(images_rgb) or tuple(buf.shape[2:]) != tuple(shape):
            buf = self._letterbox_buf = torch.empty((len(images_rgb), 3, shape[0], shape[1]), device="cuda", dtype=torch.float32)
        x = ops.letterbox_resize(packed[0], packed[1], [im.shape[:2] for im in images_rgb], shape[0], shape[1], out=buf)
        out, cnt = self.localizer.run_device(x)
        cnt = cnt.cpu().tolist()  # (synchronises) survivors per line: a few dozen of the max_det = 1000 rows
        top = max(cnt) if cnt else 0
        out = out[:, :max(top, 1)].cpu()  # only the occupied rows cross PCIe: ~60 KB instead of 1.5 MB per 64 lines
        return [out[i, :cnt[i]] for i in range(len(images_rgb))]

    # -- host logic between the two GPU phases (reference semantics, ONNX path)
    def _boxes_for_line(self, result, im_h, im_w):
        # numpy float32 rows: the same float32 arithmetic / comparisons as the reference's torch rows, without a
        # ~5 us torch dispatch per scalar operation (the O(n w) word-end search is ~500 scalar ops per line)
        result = np.asarray(result, dtype=np.float32)
        if self.box_convention == "torch":
            return self._boxes_for_line_torch(result, im_h, im_w)
        bboxes, labels = result[:, :4], result[:, -1]
        word_end_idx = []
        if self.lang == "en":
            char_b, word_b = bboxes[labels == 0], bboxes[labels == 1]
            if len(char_b) != 0:
                char_b, word_end_idx = textproc.en_preprocess_np(char_b, word_b)  # == en_preprocess, vectorised
        else:
            char_b = bboxes[labels == 0]
            if len(char_b) != 0:  # jp_preprocess: stable sort along the text direction
                char_b = char_b[np.argsort(char_b[:, 1 if self.vertical else 0], kind="stable")]
        rects = crop_rects_onnx_path(char_b, im_h, im_w, self.vertical)
        heights = list(char_b[:, 3] - char_b[:, 1])  # float32 scalars, as b[3] - b[1] row by row
        bottoms = list(char_b[:, 3])
        return list(char_b), word_end_idx, rects, heights, bottoms

    def _boxes_for_line_torch(self, result, im_h, im_w):
        """infer_effocr.py:270-302 on the shim's un-letterboxed detections (see __init__)."""
        from .mmdet_shim import unletterbox

        det = unletterbox(result, im_h, im_w, self.localizer._input_shape)
        labels, score = det[:, 5], det[:, 4]
        char_b = det[(labels == 0) & (score > np.float32(self.score_thresh))][:, :4]
        word_end_idx = []
        if self.lang == "en":
            word_b = det[(labels == 1) & (score > np.float32(self.score_thresh_word))][:, :4]
            if len(char_b) != 0:  # filtering commutes with the reference's stable sort
                char_b, word_end_idx = textproc.en_preprocess_np(char_b, word_b, vertical=self.vertical)
        elif len(char_b) != 0:
            char_b = char_b[np.argsort(char_b[:, 1 if self.vertical else 0], kind="stable")]
        rects = [crop_rect_torch_path(b, im_h, im_w, self.vertical) for b in char_b]
        heights = list(char_b[:, 3] - char_b[:, 1])
        bottoms = list(char_b[:, 3])
        return list(char_b), word_end_idx, rects, heights, bottoms

    # -- the two GPU phases as separately schedulable stages (run_effocr overlaps stage 1 of batch i+1 with stage 2 of batch i)
    def stage_localize(self, images_rgb):
        """Stage 1 on the CURRENT stream: upload the u8 lines once, letterbox + YOLOv5s + NMS on the device, boxes to
        the host, reference box ordering / word ends / crop rectangles, rectangles back to the device."""
        pool = getattr(self, "_pin_pool", None)
        if pool is None:
            pool = self._pin_pool = ops.PinnedPool()
        packed_images = ops.pack_images(images_rgb, pool=pool)  # ONE upload of the u8 lines for both GPU phases
        dets = self.localize(images_rgb, packed_images)
        per_line, all_rects = [], []
        for li, (im, det) in enumerate(zip(images_rgb, dets)):
            h, w = im.shape[:2]
            char_b, wei, rects, heights, bottoms = self._boxes_for_line(det, h, w)
            per_line.append((char_b, wei, heights, bottoms, rects))
            all_rects += [(li,) + r for r in rects]
        boxes, n = ops.pack_boxes(all_rects) if all_rects else (None, 0)
        done = torch.cuda.current_stream().record_event()
        return {"packed": packed_images, "per_line": per_line, "boxes": boxes, "n": n, "event": done}

    def launch_recognize(self, st):
        """Stage 2a on the CURRENT stream, asynchronous: crop -> encoder -> kNN for every character of the batch, then the
        ids into a pinned host buffer with an asynchronous copy.  Returns (host ids, event) or None; nothing is
        synchronised, so the next batch's kernels can be enqueued behind this one before its ids are read."""
        if not st["n"]:
            return None
        torch.cuda.current_stream().wait_event(st["event"])
        _dist, idx, _ = self.recognizer.recognize_device(st["packed"][0], st["packed"][1], st["boxes"], st["n"], self.knn)
        ring = getattr(self, "_id_ring", None)
        if ring is None:
            ring = self._id_ring = {"bufs": [None] * 3, "next": 0}  # three batches can be between launch and finish
        slot = ring["next"]
        ring["next"] = (slot + 1) % len(ring["bufs"])
        buf = ring["bufs"][slot]
        if buf is None or buf.shape[0] < idx.shape[0] or buf.shape[1] != idx.shape[1] or buf.dtype != idx.dtype:
            # pinned allocations synchronise the device and take a driver-wide lock: all three at once, generously sized
            rows = max(2 * int(idx.shape[0]), 4096)
            ring["bufs"] = [torch.empty((rows, idx.shape[1]), dtype=idx.dtype, pin_memory=True) for _ in ring["bufs"]]
            buf = ring["bufs"][slot]
        host = buf[:idx.shape[0]]
        host.copy_(idx, non_blocking=True)
        return host, torch.cuda.current_stream().record_event()

    def finish_recognize(self, st, idx):
        """Stage 2b: wait for the batch's ids (its own event, not the stream), decode + en_postprocess."""
        results = []
        if idx is not None:
            host, event = idx
            event.synchronize()
            idx = host.numpy()
        pos = 0
        for (char_b, wei, heights, bottoms, rects) in st["per_line"]:
            n = len(rects)
            if n == 0:
                results.append({"text": None, "nns": [], "char_boxes": [], "word_end_idx": [], "rects": []})
                continue
            rows = idx[pos:pos + n]
            pos += n
            nearest = [[self.candidate_chars[j] for j in row if j >= 0] for row in rows.tolist()]
            nns = ["".join(c).strip() for c in nearest]
            text = "".join(x[0] for x in nearest).strip()
            if self.lang == "en":
                # the ONNX driver passes the un-stripped first-NN string's characters; lengths must agree (:94)
                first = [x[0] for x in nearest]
                text = textproc.en_postprocess(first, wei, heights, bottoms, anchor_margin=self.anchor_margin)
            results.append({"text": text, "nns": nns, "char_boxes": [b.tolist() for b in char_b], "word_end_idx": wei,
                            "rects": rects})
        return results

    def stage_recognize(self, st):
        return self.finish_recognize(st, self.launch_recognize(st))

    def infer_lines(self, images_rgb):
        """images_rgb: list of u8 [H, W, 3] arrays -> list of dicts (text, nns, char_boxes, word_end_idx)."""
        if len(images_rgb) == 0:
            return []
        return self.stage_recognize(self.stage_localize(images_rgb))

    def infer_batches(self, batches, overlap: bool = True):
        """Yield the per-line results of each batch in order.  With overlap the reference's three barriered phases
        become a software pipeline on ONE host thread: the recognizer kernels of batch i are enqueued without a
        host synchronisation, then stage 1 of batch i+1 (pinned-memory packing, upload, letterbox, YOLOv5s, NMS, host
        box logic) runs on a second CUDA stream while the GPU works through batch i, recognize(i+1) is enqueued behind it,
        and only then are the ids of batch i (copied to pinned memory asynchronously) decoded.  (A worker thread was tried first: the two Python threads fought over the GIL
        and the pipeline got slower.)  Results are identical to the sequential order: the stages share no mutable
        state and every kernel is deterministic."""
        # `batches` is consumed lazily, one batch ahead of the one being recognised, so that a decoding iterator
        # (effocr_b200.lineio.LineDecoder) keeps working while the GPU does
        it = iter(batches)
        first = next(it, None)
        if first is None:
            return
        second = next(it, None) if overlap else None
        if second is None:  # sequential order (also: a single batch)
            yield self.infer_lines(first)
            for chunk in it:
                yield self.infer_lines(chunk)
            return
        side = getattr(self, "_side_stream", None)  # persistent: the caching allocator keeps one block pool per stream
        if side is None:
            side = self._side_stream = torch.cuda.Stream()

        def stage1(chunk):
            if len(chunk) == 0:
                return None
            with torch.cuda.stream(side):
                return self.stage_localize(chunk)

        # order per iteration: enqueue recognize(i) -> decode batch i-1 (its ids arrived long ago) -> stage 1 of batch i+1 on
        # the side stream (the host blocks on ITS detections while the GPU runs recognize(i)) -> next iteration enqueues
        # recognize(i+1) behind recognize(i) before anything waits for batch i: the GPU never idles on a host round trip
        st, nxt = stage1(first), second
        prev = None
        while True:
            handle = self.launch_recognize(st) if st is not None else None
            if prev is not None:
                yield self.finish_recognize(*prev) if prev[0] is not None else []
            st_next = stage1(nxt) if nxt is not None else None
            prev = (st, handle)
            if nxt is None:
                break
            st, nxt = st_next, next(it, None)
        yield self.finish_recognize(*prev) if prev[0] is not None else []


def run_effocr_sharded(images_rgb, pipeline, keys=None, batch_lines: int = 64, weights=None):
    """Data-parallel driver: every rank transcribes its shard of the lines, rank 0 gets the
    merged {key: text}.  Results are keyed by input and per-line arithmetic does not depend on batch composition,
    so the output is independent of the world size."""
    from . import dist as D

    rank, world, _ = D.init_from_env()
    keys = list(range(len(images_rgb))) if keys is None else list(keys)
    mine = D.shard_indices(len(images_rgb), rank, world, weights=weights)
    local = run_effocr([images_rgb[i] for i in mine], pipeline, batch_lines=batch_lines, keys=[keys[i] for i in mine])
    return D.gather_results(local)


def run_effocr(images_rgb, pipeline: EffOCRPipeline, batch_lines: int = 64, keys=None, overlap: bool = True):
    """-> {key: text}; `keys` default to the line index.  Mirrors run_effocr's return (inference_results)."""
    keys = list(range(len(images_rgb))) if keys is None else list(keys)
    out = {}
    starts = list(range(0, len(images_rgb), batch_lines))
    batches = (images_rgb[i:i + batch_lines] for i in starts)
    # any object with infer_lines() works as a pipeline; the overlapped driver needs the two-stage interface
    results = pipeline.infer_batches(batches, overlap=overlap) if hasattr(pipeline, "infer_batches") else map(pipeline.infer_lines, batches)
    for i0, res in zip(starts, results):
        for k, r in zip(keys[i0:i0 + batch_lines], res):
            out[k] = r["text"]
    return out
