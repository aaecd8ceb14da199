"""GPU: the whole hot path (localize -> crop -> embed -> kNN -> decode) through EffOCRPipeline, checked stage by
stage against the oracle GIVEN the pipeline's own upstream outputs (random-init weights make end-to-end string
equality ill-defined: SURVEY.md section 0.5 -- top-1 margins ~1e-5 -- so ids are compared margin-aware)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(knn=1):
    from effocr_b200 import synth
    from effocr_b200.infer import EffOCRPipeline
    from effocr_b200.localizer_engine import EffLocalizer
    from effocr_b200.pipeline import RecognizerPipeline
    from oracle import transform as OT, vit as OV, yolo as OY
    ysd = OY.init_yolov5s_state_dict(nc=2, seed=0, obj_bias=-0.5)
    vsd = OV.randomize_affine(OV.init_vit_state_dict("vit_tiny_patch16_224", seed=0))
    glyphs = synth.glyph_images(94)
    with torch.no_grad():
        xb = OV.l2_normalize(OV.vit_forward(vsd, torch.from_numpy(np.stack([OT.paired_transform(g) for g in glyphs]))))
    loc = EffLocalizer(ysd, iou_thresh=0.01, conf_thresh=0.35, input_shape=(640, 640), max_batch=4)
    rec = RecognizerPipeline(vsd, xb, max_batch=256)
    return EffOCRPipeline(loc, rec, synth.ASCII_GLYPHS, lang="en", knn=knn), vsd, xb, ysd


def _build_trained(knn=1):
    """The same pipeline on the quick-fit YOLOv5s + ViT-S (tests/golden/): margins far above the embedding tolerance."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import driver_fixture as DF
    from effocr_b200 import synth
    from effocr_b200.infer import EffOCRPipeline
    from effocr_b200.localizer_engine import EffLocalizer
    from effocr_b200.pipeline import RecognizerPipeline
    from oracle import transform as OT, vit as OV
    if not DF.available():
        pytest.skip("quick-fit weights not generated")
    vsd, ysd = DF.load_npz_state(DF.VIT_WEIGHTS), DF.load_npz_state(DF.YOLO_WEIGHTS)
    with torch.no_grad():
        xb = OV.l2_normalize(OV.vit_forward(vsd, torch.from_numpy(np.stack([OT.paired_transform(c) for c in DF.prototype_crops()]))))
    loc = EffLocalizer(ysd, iou_thresh=0.01, conf_thresh=0.35, input_shape=(640, 640), max_batch=8)
    rec = RecognizerPipeline(vsd, xb, max_batch=512)
    return EffOCRPipeline(loc, rec, synth.ASCII_GLYPHS, lang="en", knn=knn), vsd, xb, ysd


def test_pipeline_stagewise_parity():
    """Every stage after the localizer, checked against the oracle GIVEN the pipeline's own detections: word ends,
    crop rectangles, top-1 characters and the assembled text, on trained weights (every row decidable)."""
    from effocr_b200 import synth, textproc
    from effocr_b200.infer import crop_rect_onnx_path
    from oracle import knn as OK, transform as OT, vit as OV
    pipe, vsd, xb, _ = _build_trained(knn=1)
    lines = [l[0] for l in synth.synthetic_lines(6, seed=5, tracking=4.0)]
    res = pipe.infer_lines(lines)
    assert len(res) == 6
    dets = pipe.localize(lines)
    n_checked = n_chars = 0
    for im, det, r in zip(lines, dets, res):
        labels = det[:, -1]
        char_b = det[:, :4][labels == 0]
        word_b = det[:, :4][labels == 1]
        assert len(char_b) >= 10, "the quick-fit localizer finds the characters of a synthetic line"
        sc, wei = textproc.en_preprocess(char_b, word_b)
        assert wei == r["word_end_idx"]
        crops = []
        for j, b in enumerate(sc):
            rect = OT.crop_rect_onnx_path(b, im.shape[0], im.shape[1])
            assert rect == crop_rect_onnx_path(b, im.shape[0], im.shape[1], False) == tuple(r["rects"][j])
            crops.append(OT.numpy_slice(im, rect))
        assert all(c.size for c in crops)
        with torch.no_grad():
            emb = OV.l2_normalize(OV.vit_forward(vsd, torch.from_numpy(np.stack([OT.paired_transform(c) for c in crops]))))
        _, ri = OK.flat_ip_search(xb, emb, 1)
        _, margin = OK.margins(xb, emb, 1)
        exp = [pipe.candidate_chars[int(i)] for i in ri[:, 0]]
        dec = (margin > 4e-3).tolist()  # 4 x the 1e-3 embedding tolerance (SURVEY.md section 8c rule 3)
        for g, e, d in zip(r["nns"], exp, dec):
            n_chars += 1
            if d:
                assert g == e
                n_checked += 1
        heights = [b[3] - b[1] for b in sc]
        bottoms = [b[3] for b in sc]
        if all(dec):  # the text is en_postprocess over the oracle's first neighbours
            assert r["text"] == textproc.en_postprocess(exp, wei, heights, bottoms)
    assert n_chars >= 60 and n_checked >= 0.9 * n_chars, (n_checked, n_chars)


def test_pipeline_batch_composition_independence():
    """Same line alone or inside a batch -> identical transcription and neighbours (fixed per-crop arithmetic)."""
    from effocr_b200 import synth
    pipe, *_ = _build(knn=3)
    lines = [l[0] for l in synth.synthetic_lines(4, seed=9)]
    all_res = pipe.infer_lines(lines)
    for i in (0, 2):
        single = pipe.infer_lines([lines[i]])[0]
        assert single["text"] == all_res[i]["text"] and single["nns"] == all_res[i]["nns"]


def test_run_effocr_returns_keyed_results():
    from effocr_b200 import synth
    from effocr_b200.infer import run_effocr
    pipe, *_ = _build()
    lines = [l[0] for l in synth.synthetic_lines(5, seed=2)]
    out = run_effocr(lines, pipe, batch_lines=2, keys=[f"l{i}.png" for i in range(5)])
    assert list(out) == [f"l{i}.png" for i in range(5)]


def test_overlapped_pipeline_equals_sequential():
    """The two-stage overlapped driver (stage 1 of batch i+1 on a side stream while stage 2 of batch i runs) returns
    exactly what the sequential per-batch loop returns."""
    from effocr_b200 import synth
    pipe, *_ = _build(knn=2)
    lines = [l[0] for l in synth.synthetic_lines(11, seed=21)]
    batches = [lines[i:i + 3] for i in range(0, len(lines), 3)]
    seq = [r for chunk in batches for r in pipe.infer_lines(chunk)]
    ovl = [r for res in pipe.infer_batches(batches, overlap=True) for r in res]
    assert len(seq) == len(ovl) == len(lines)
    for a, b in zip(seq, ovl):
        assert a["text"] == b["text"] and a["nns"] == b["nns"] and a["char_boxes"] == b["char_boxes"]


@pytest.mark.parametrize("vertical", [False, True])
def test_pipeline_japanese_mode_box_order_and_rects(vertical):
    """lang='jp': character boxes are ordered along the text direction (y for vertical lines) and every crop spans the
    full extent across it (infer_effocr_onnx_multi.py:134-140, 311-318); transcription = first neighbours joined."""
    from effocr_b200 import synth
    from effocr_b200.infer import EffOCRPipeline, crop_rect_onnx_path
    base, *_ = _build(knn=1)
    pipe = EffOCRPipeline(base.localizer, base.recognizer, base.candidate_chars, lang="jp", vertical=vertical, knn=1)
    lines = [l[0] for l in synth.synthetic_lines(2, seed=11)]
    if vertical:
        lines = [np.ascontiguousarray(np.transpose(l, (1, 0, 2))) for l in lines]  # 1024 x 64 "vertical" lines
    res = pipe.infer_lines(lines)
    dets = pipe.localize(lines)
    for im, det, r in zip(lines, dets, res):
        chars = det[det[:, -1] == 0][:, :4]
        assert len(r["char_boxes"]) == len(chars)
        if len(chars) == 0:
            assert r["text"] is None
            continue
        key = [b[1] if vertical else b[0] for b in r["char_boxes"]]
        assert key == sorted(key)
        h, w = im.shape[:2]
        for b in r["char_boxes"][:5]:
            x0, y0, x1, y1 = crop_rect_onnx_path(b, h, w, vertical)
            assert (x0, x1) == (0, w) if vertical else (y0, y1) == (0, h)
        assert r["text"] == "".join(n[:1] for n in r["nns"]).strip()
