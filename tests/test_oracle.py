"""CPU: the oracle restatements against (a) the committed golden fixtures generated from the live
reference (oracle/make_golden.py) and (b) the live reference itself when /root/reference is present."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import knn as OK, make_golden as MG, ref_harness as rh, transform as OT, vit as OV, yolo as OY

GOLDEN = Path(__file__).resolve().parent / "golden"
needs_ref = pytest.mark.skipif(not rh.available(), reason="/root/reference absent (GPU box)")


def test_transform_matches_golden():
    g = np.load(GOLDEN / "transform_golden.npz")
    for i, crop in enumerate(MG.transform_inputs()):
        out = OT.paired_transform(crop)
        assert np.abs(out[:, ::7, ::7] - g[f"sub_{i}"]).max() <= 2e-6
        assert abs(out.astype(np.float64).sum() - float(g[f"sum_{i}"])) <= 1e-2


def test_transform_empty_crop_raises():
    with pytest.raises(ValueError):
        OT.paired_transform(np.zeros((64, 0, 3), np.uint8))


def test_crop_rects():
    # banker's rounding + double clipping (infer_effocr.py:286-291)
    assert OT.crop_rect_torch_path([10.5, 3.2, 21.5, 40.0], 64, 1024) == (10, 0, 22, 64)
    assert OT.crop_rect_torch_path([10.5, 3.5, 21.5, 40.5], 64, 1024, vertical=True) == (0, 4, 1024, 40)
    # onnx path: torch.round, x scaled by W/640 in double, full height (infer_effocr_onnx_multi.py:311-318)
    assert OT.crop_rect_onnx_path([100.4, 1, 120.6, 60], 64, 1024) == (160, 0, 194, 64)


@needs_ref
def test_transform_matches_live_reference():
    t = rh.import_reference("utils.datasets_utils").create_paired_transform()
    rng = np.random.default_rng(5)
    for (h, w) in [(64, 31), (17, 90), (400, 260)]:
        crop = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.abs(t(crop).numpy() - OT.paired_transform(crop)).max() <= 2e-6


def test_vit_matches_golden_reference_hf_backend():
    sd, x = MG.vit_golden_inputs()
    with torch.no_grad():
        e = OV.vit_forward(sd, x).numpy()
    ref = np.load(GOLDEN / "vit_tiny_golden.npz")["emb"]
    assert np.linalg.norm(e - ref) / np.linalg.norm(ref) <= 5e-6


def test_vit_matches_torchvision():
    from torchvision.models.vision_transformer import VisionTransformer
    name = "vit_tiny_patch16_224"
    sd = OV.randomize_affine(OV.init_vit_state_dict(name, seed=2))
    d, h, depth, mlp = OV.VIT_CONFIGS[name]
    tv = VisionTransformer(224, 16, depth, h, d, mlp)
    tv.heads = torch.nn.Identity()
    tv.load_state_dict(OV.timm_to_torchvision(sd), strict=True)
    tv.eval()
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        a, b = OV.vit_forward(sd, x), tv(x)
    assert ((a - b).norm() / b.norm()).item() <= 5e-6


def test_convnext_matches_torchvision():
    """The ConvNeXt-Tiny restatement (timm keys, num_classes=0) against torchvision's convnext_tiny with the classifier's
    Linear removed (avgpool -> LayerNorm2d -> flatten: the same pooled pre-logits), through oracle.convnext.timm_to_torchvision."""
    from torchvision.models import convnext_tiny
    from oracle import convnext as OC
    sd = OC.init_convnext_tiny_state_dict(seed=4)
    tv = convnext_tiny(weights=None)
    tv.classifier[2] = torch.nn.Identity()
    tv.load_state_dict(OC.timm_to_torchvision(sd), strict=True)
    tv.eval()  # stochastic depth off
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        a, b = OC.convnext_forward(sd, x), tv(x)
    assert a.shape == (2, 768)
    assert ((a - b).norm() / b.norm()).item() <= 5e-6


def test_vit_key_maps_roundtrip():
    sd = OV.init_vit_state_dict("vit_tiny_patch16_224", seed=0)
    back = OV.hf_to_timm(OV.timm_to_hf(sd))
    assert list(back) == list(sd) and all(torch.equal(back[k], sd[k]) for k in sd)


def test_knn_oracle_ties_and_padding():
    xb = torch.eye(4)[[0, 1, 1, 2]]  # ids 1 and 2 identical
    q = torch.tensor([[0.0, 1.0, 0.0, 0.0]])
    d, i = OK.flat_ip_search(xb, q, 6)
    assert i[0].tolist() == [1, 2, 0, 3, -1, -1]
    assert d[0, 4].item() < -3e38


def test_index_file_roundtrip(tmp_path):
    xb = np.random.default_rng(0).normal(size=(7, 5)).astype(np.float32)
    OK.write_index_flat_ip(tmp_path / "ref.index", xb)
    assert np.array_equal(OK.read_index_flat_ip(tmp_path / "ref.index"), xb)
    # the product's reader/writer agree with the oracle's statement of the layout
    from effocr_b200 import knn as PK
    idx = PK.read_index(tmp_path / "ref.index")
    assert idx.ntotal == 7 and idx.d == 5 and np.array_equal(idx.reconstruct_n(), xb)
    PK.write_index(idx, tmp_path / "ref2.index")
    assert (tmp_path / "ref2.index").read_bytes() == (tmp_path / "ref.index").read_bytes()


def test_nms_matches_golden():
    g = np.load(GOLDEN / "nms_golden.npz")
    for i, (pred, conf, iou) in enumerate(MG.nms_inputs()):
        out = OY.non_max_suppression(pred.clone(), conf_thres=conf, iou_thres=iou, max_det=1000)[0].numpy()
        assert out.shape == g[f"out_{i}"].shape and np.array_equal(out, g[f"out_{i}"])


def test_letterbox_matches_golden():
    g = np.load(GOLDEN / "letterbox_golden.npz")
    for i, im in enumerate(MG.letterbox_inputs()):
        x = OY.load_localizer_img_from_array(im)  # [1,3,640,640] RGB/255
        u8 = np.rint(x[0][::-1].transpose(1, 2, 0) * 255).astype(np.uint8)  # back to BGR HWC u8
        assert np.array_equal(u8[::9, ::9], g[f"sub_{i}"])
        assert int(u8.astype(np.int64).sum()) == int(g[f"sum_{i}"])


LETTERBOX_SHAPES = [(64, 1024), (64, 1000), (30, 500), (700, 90), (100, 100), (48, 333), (1, 10), (640, 640), (123, 457),
                    (800, 1200), (64, 64), (33, 977), (2, 3)]


def test_cv2_resize_restatement_matches_live_cv2_and_golden():
    """The fixed-point restatement of cv2.resize(INTER_LINEAR, u8) equals live OpenCV bit for bit (down- and
    up-scaling, both edge rules), and the OpenCV-free letterbox equals the reference-generated golden fixtures."""
    import cv2
    rng = np.random.default_rng(0)
    for (h, w) in LETTERBOX_SHAPES:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for (nw, nh) in [(max(1, w * 5 // 8), max(1, h * 5 // 8)), (w * 2 + 1, h * 3), (640, max(1, round(h * 640 / w)))]:
            assert np.array_equal(OY.cv2_resize_linear_u8(img, nw, nh), cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR))
    g = np.load(GOLDEN / "letterbox_golden.npz")
    for i, im in enumerate(MG.letterbox_inputs()):
        x = OY.load_localizer_img_restated(im)
        assert np.array_equal(x, OY.load_localizer_img_from_array(im))
        u8 = np.rint(x[0][::-1].transpose(1, 2, 0) * 255).astype(np.uint8)
        assert np.array_equal(u8[::9, ::9], g[f"sub_{i}"]) and int(u8.astype(np.int64).sum()) == int(g[f"sum_{i}"])


def test_letterbox_plan_tables_match_oracle():
    """Product host logic (localizer_engine.letterbox_plan: geometry + tap tables the device kernel consumes),
    emulated in numpy, equals the oracle letterbox for every shape."""
    from effocr_b200.localizer_engine import letterbox_plan
    rng = np.random.default_rng(1)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in LETTERBOX_SHAPES]
    plans, taps = letterbox_plan([im.shape[:2] for im in imgs], (640, 640))
    for im, pl in zip(imgs, plans):
        tx = taps[pl["xtap_offset"]:pl["xtap_offset"] + pl["new_width"]].astype(np.int64)
        ty = taps[pl["ytap_offset"]:pl["ytap_offset"] + pl["new_height"]].astype(np.int64)
        src = im.astype(np.int64)
        s0 = src[ty[:, 0]][:, tx[:, 0]] * tx[None, :, 2, None] + src[ty[:, 0]][:, tx[:, 1]] * tx[None, :, 3, None]
        s1 = src[ty[:, 1]][:, tx[:, 0]] * tx[None, :, 2, None] + src[ty[:, 1]][:, tx[:, 1]] * tx[None, :, 3, None]
        o = (((ty[:, 2, None, None] * (s0 >> 4)) >> 16) + ((ty[:, 3, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
        canvas = np.full((640, 640, 3), 114, np.uint8)
        canvas[pl["top"]:pl["top"] + pl["new_height"], pl["left"]:pl["left"] + pl["new_width"]] = np.clip(o, 0, 255)
        got = (canvas.transpose(2, 0, 1).astype(np.float32) / 255.0)[None]
        assert np.array_equal(got, OY.load_localizer_img_from_array(np.ascontiguousarray(im[:, :, ::-1])))


def test_yolov5s_analytic_invariants():
    assert OY.count_parameters(OY.init_yolov5s_state_dict(nc=80)) == 7_235_389  # ultralytics yolov5s
    sd = OY.init_yolov5s_state_dict(nc=2)
    x = torch.rand(1, 3, 64, 1024, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        o32 = OY.yolov5s_forward(sd, x)
        o64 = OY.yolov5s_forward(sd, x, dtype=torch.float64)
    assert o32.shape == (1, 3 * (8 * 128 + 4 * 64 + 2 * 32), 7)
    assert ((o32 - o64).abs().max() / o64.abs().max()).item() < 1e-5
    assert 3 * (80 * 80 + 40 * 40 + 20 * 20) == 25200  # prediction count at 640 x 640


def test_textproc_matches_golden():
    from effocr_b200 import textproc as tp
    recs = json.loads((GOLDEN / "textproc_golden.json").read_text())
    for c, r in zip(MG.textproc_cases(), recs):
        chars, words = torch.tensor(c["chars"]), torch.tensor(c["words"]).reshape(-1, 4)
        sc, wei = tp.en_preprocess(chars, words)
        assert wei == r["word_end_idx"] and [float(b[0]) for b in sc] == r["order_x0"]
        assert [float(b[1]) for b in tp.jp_preprocess(chars, vertical=True)] == r["jp_order_y0"]
        for am in (None, 0.1, 0.3):
            assert tp.en_postprocess(c["text"], wei, c["heights"], c["bottoms"], anchor_margin=am) == r["post"][str(am)]


def test_edit_distance_and_cer():
    from effocr_b200 import textproc as tp
    assert tp.edit_distance("kitten", "sitting") == 3 and tp.edit_distance("", "abc") == 3
    acc, cer = tp.textline_evaluation([("hello world", "hello world"), ("abc", "abd")])
    assert acc == 50.0 and abs(cer - 1 / 14) < 1e-12


def test_pack_images_layout_and_pool_reuse():
    """Host staging: pixels and the descriptor table travel in ONE buffer (descriptors 256-byte aligned behind the
    pixels); a PinnedPool hands the same staging buffer out again once its copy has completed."""
    from effocr_b200 import ops
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in [(64, 1024), (40, 333), (7, 5)]]
    pool = ops.PinnedPool()
    for _ in range(3):
        pixels, images, descs = ops.pack_images(imgs, device="cpu", pool=pool)
        table = np.frombuffer(images.numpy().tobytes(), dtype=ops.IMAGE_DESC_DTYPE)
        assert np.array_equal(table, descs)
        for im, d in zip(imgs, descs):
            o = int(d["offset"])
            assert o % 256 == 0 and (d["height"], d["width"], d["pitch"]) == (im.shape[0], im.shape[1], im.shape[1] * 3)
            assert np.array_equal(pixels.numpy()[o:o + im.size].reshape(im.shape), im)
    assert len(pool._bufs) <= 1 or not torch.cuda.is_available()


def test_en_preprocess_np_equals_scalar_version():
    """The vectorised word-end search the pipeline uses must return exactly what the (golden-pinned) scalar en_preprocess
    returns: golden cases, random float32 boxes, exact ties, words left of / right of every character (index carry-over),
    no words, no characters."""
    from effocr_b200 import textproc as tp
    cases = [(np.asarray(c["chars"], np.float32).reshape(-1, 4), np.asarray(c["words"], np.float32).reshape(-1, 4)) for c in MG.textproc_cases()]
    rng = np.random.default_rng(11)
    for _ in range(300):
        n, w = int(rng.integers(0, 40)), int(rng.integers(0, 9))
        x0 = rng.uniform(0, 640, n).astype(np.float32)
        ch = np.stack([x0, rng.uniform(0, 60, n), x0 + rng.uniform(1, 40, n), rng.uniform(60, 64, n)], 1).astype(np.float32) if n else np.zeros((0, 4), np.float32)
        wx = rng.uniform(-50, 700, w).astype(np.float32)
        wd = np.stack([wx, np.zeros(w), wx + 80, np.full(w, 64.0)], 1).astype(np.float32) if w else np.zeros((0, 4), np.float32)
        if n > 3 and w > 1:  # exact ties: duplicated boxes, a word edge exactly on / halfway between character edges
            ch[1] = ch[0]
            ch[3, 2] = ch[2, 2]
            wd[0, 0] = ch[2, 2]
            wd[1, 0] = np.float32((ch[0, 2] + ch[2, 2]) / 2)
        cases.append((ch, wd))
    for vertical in (False, True):
        for ch, wd in cases:
            sc, wei = tp.en_preprocess(list(ch), list(wd), vertical=vertical)
            sc2, wei2 = tp.en_preprocess_np(ch, wd, vertical=vertical)
            assert wei2 == wei
            assert len(sc2) == len(sc) and all(np.array_equal(a, b) for a, b in zip(sc, sc2))


def test_vectorised_crop_rects_equal_scalar_reference_semantics():
    """infer.crop_rects_onnx_path (all boxes of a line at once; array, tensor and list input) == the scalar restatement
    of infer_effocr_onnx_multi.py:311-318 box by box, half-way cases included, both text directions."""
    from effocr_b200 import infer
    rng = np.random.default_rng(3)
    for (h, w) in [(64, 1024), (37, 640), (900, 55)]:
        b = rng.uniform(-5, 660, (200, 4)).astype(np.float32)
        b[::7] = np.round(b[::7]) + 0.5  # ties for round-half-to-even
        b[::11, 0] = np.float32(0.8 * 3)  # products that land on x.5 after the rescale
        for vertical in (False, True):
            exp = [infer.crop_rect_onnx_path(row, h, w, vertical) for row in b]
            assert [OT.crop_rect_onnx_path(list(map(float, row)), h, w, vertical=vertical) for row in b[:50]] == exp[:50]
            assert infer.crop_rects_onnx_path(b, h, w, vertical) == exp
            assert infer.crop_rects_onnx_path(torch.from_numpy(b), h, w, vertical) == exp
            assert infer.crop_rects_onnx_path(list(b), h, w, vertical) == exp
            assert all(type(v) is int for r in infer.crop_rects_onnx_path(b, h, w, vertical) for v in r)
    assert infer.crop_rects_onnx_path(np.zeros((0, 4), np.float32), 64, 1024, False) == []


@pytest.mark.parametrize("convention,key", [("onnx", "infer_effocr_onnx_multi"), ("torch", "infer_effocr")])
def test_oracle_pipeline_matches_reference_driver_golden(convention, key):
    """oracle/pipeline.py (the whole hot path restated) reproduces, string for string, what the UNMODIFIED reference
    drivers wrote for the same job when run over oracle back-ends (tests/golden/driver_golden.json, generated by
    oracle/make_driver_golden.py): the orchestration restatement -- box ordering, word ends, crop rectangles, zero-padded
    batches, decode, en_postprocess -- is pinned by the reference's own scripts."""
    import json
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import driver_fixture as DF
    from effocr_b200 import synth
    from oracle import pipeline as OP
    if not (DF.available() and DF.DRIVER_GOLDEN.exists()):
        pytest.skip("quick-fit weights / driver golden not generated")
    golden = json.loads(DF.DRIVER_GOLDEN.read_text())
    n = 8  # a third of the job keeps the CPU suite short; the GPU tests cover all 24 lines
    lines = [l[0] for l in synth.synthetic_lines(DF.N_LINES, seed=DF.SEED, tracking=DF.TRACKING)][:n]
    vsd, ysd = DF.load_npz_state(DF.VIT_WEIGHTS), DF.load_npz_state(DF.YOLO_WEIGHTS)
    xb = torch.from_numpy(np.load(DF.INDEX_VECTORS))
    conf, k = (0.35, 1) if convention == "onnx" else (0.05, 10)
    res = OP.run(lines, ysd, vsd, xb, synth.ASCII_GLYPHS, conf_thres=conf, iou_thres=0.01, convention=convention, k=k)
    for i, r in enumerate(res):
        assert r["text"] == golden[key].get(f"line_{i:03d}.png"), (i, r["text"], golden[key].get(f"line_{i:03d}.png"))
