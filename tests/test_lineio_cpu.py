"""Caller-side plumbing: image enumeration, the prefetching decoder, the COCO / JSON writers and
the lazy batch consumption of the pipeline driver.  CPU only; the reference's own helpers are used as the checker where
they import here (utils/coco_utils.py, EffOCR.mmdet_output_format through oracle.ref_harness)."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

from effocr_b200 import lineio

REF = "/root/reference"


def _write_lines(tmp_path, n, ext="png"):
    from PIL import Image
    rng = np.random.default_rng(7)
    paths, arrays = [], []
    for i in range(n):
        a = rng.integers(0, 256, size=(8 + i % 3, 20 + i, 3), dtype=np.uint8)
        p = tmp_path / f"line_{i:03d}.{ext}"
        Image.fromarray(a).save(p)
        paths.append(str(p))
        arrays.append(a)
    return paths, arrays


def test_list_images_coco_order_and_directory_glob(tmp_path):
    paths, _ = _write_lines(tmp_path, 3)
    sub = tmp_path / "sub"
    sub.mkdir()
    from PIL import Image
    Image.fromarray(np.zeros((4, 4, 3), np.uint8)).save(sub / "deep.jpg")
    coco = {"images": [{"file_name": "line_002.png", "text": "b"}, {"file_name": "line_000.png", "text": "a"}]}
    cj = tmp_path / "coco.json"
    cj.write_text(json.dumps(coco))
    got, loaded = lineio.list_images(str(tmp_path), str(cj))
    assert got == [os.path.join(str(tmp_path), "line_002.png"), os.path.join(str(tmp_path), "line_000.png")] and loaded == coco
    got, loaded = lineio.list_images(str(tmp_path))
    assert loaded is None and sorted(got[:3]) == sorted(paths) and got[3].endswith("deep.jpg")  # all png first, then jpg


@pytest.mark.parametrize("batch_lines,prefetch,workers", [(1, 0, 1), (3, 2, 4), (4, 1, 2), (64, 2, 8)])
def test_line_decoder_order_content_and_raggedness(tmp_path, batch_lines, prefetch, workers):
    paths, arrays = _write_lines(tmp_path, 10)
    dec = lineio.LineDecoder(paths, batch_lines=batch_lines, workers=workers, prefetch=prefetch)
    seen_paths, seen = [], []
    nb = 0
    for chunk, images in dec:
        assert len(chunk) == len(images) <= batch_lines
        seen_paths += chunk
        seen += images
        nb += 1
    assert nb == len(dec) == (10 + batch_lines - 1) // batch_lines
    assert seen_paths == paths
    assert all(a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"] and np.array_equal(a, b) for a, b in zip(seen, arrays))
    assert list(lineio.LineDecoder([], batch_lines=4)) == []


def test_line_decoder_bounded_lookahead_and_error_at_its_batch(tmp_path):
    paths, _ = _write_lines(tmp_path, 12)
    started = []

    def decode(p):
        started.append(p)
        return lineio.decode_rgb(p)

    it = iter(lineio.LineDecoder(paths, batch_lines=2, workers=1, prefetch=1, decode=decode))
    next(it)
    # batch 0 handed out: batches 0, 1 were submitted up front and batch 2 when batch 0 was collected -- nothing further
    assert len(started) <= 6
    it.close()
    bad = paths[:5] + [str(tmp_path / "missing.png")] + paths[5:]
    got = []
    with pytest.raises(FileNotFoundError):
        for chunk, _ in lineio.LineDecoder(bad, batch_lines=2, workers=2):
            got.append(chunk)
    assert got == [bad[0:2], bad[2:4]]  # the error surfaces at the batch that holds the missing file


def test_coco_entries_equal_reference_helpers():
    cases = [(3, 4, 10, 7, 5, 2, 0, None), (3.7, 4.2, 10.5, 7.9, 1, 0, 0, "ab"), (0, 0, 1, 1, 9, 9, 1, "")]
    expect0 = {"segmentation": [[3, 4, 13, 4, 13, 11, 3, 11]], "area": 70, "iscrowd": 0, "image_id": 2, "bbox": [3, 4, 10, 7],
               "category_id": 0, "id": 5, "score": 1.0}
    assert lineio.create_coco_anno_entry(*cases[0]) == expect0
    assert lineio.create_coco_image_entry("a.png", 64, 1024, 3) == {"file_name": "a.png", "height": 64, "width": 1024, "id": 3}
    assert lineio.create_coco_image_entry("a.png", 64, 1024, 3, text="x")["text"] == "x"
    if os.path.isdir(REF):  # live reference (not present on the GPU box; this suite runs without a GPU anyway)
        sys.path.insert(0, REF)
        try:
            from utils import coco_utils as R
        finally:
            sys.path.remove(REF)
        assert lineio.COCO_JSON_SKELETON == R.COCO_JSON_SKELETON
        for c in cases:
            assert lineio.create_coco_anno_entry(*c) == R.create_coco_anno_entry(*c)
            assert json.dumps(lineio.create_coco_anno_entry(*c)) == json.dumps(R.create_coco_anno_entry(*c))  # key order too
        for t in (None, "some text"):
            assert json.dumps(lineio.create_coco_image_entry("p.png", 5, 6, 7, text=t)) == json.dumps(R.create_coco_image_entry("p.png", 5, 6, 7, text=t))


def test_mmdet_output_format_adapter():
    inst = types.SimpleNamespace(pred_classes=torch.tensor([0, 1, 0]),
                                 pred_boxes=types.SimpleNamespace(tensor=torch.tensor([[1., 2., 3., 4.], [0., 0., 9., 9.], [5., 6., 7., 8.]])),
                                 scores=torch.tensor([0.9, 0.8, 0.7]))
    out = lineio.mmdet_output_format([{"instances": inst}])
    assert len(out) == 1 and len(out[0]) == 2
    assert out[0][0] == [[1.0, 2.0, 3.0, 4.0, pytest.approx(0.9)], [5.0, 6.0, 7.0, 8.0, pytest.approx(0.7)]]
    assert out[0][1] == [[0.0, 0.0, 9.0, 9.0, pytest.approx(0.8)]]
    inst.pred_classes = torch.tensor([0, 0, 0])
    assert len(lineio.mmdet_output_format([{"instances": inst}])[0]) == 1  # no word boxes: [[char_boxes]]


class _FakePipeline:
    """Two-stage contract of EffOCRPipeline without a device: text = mean pixel of the line."""
    lang = "jp"

    def __init__(self):
        self.calls = []

    def infer_lines(self, imgs):
        self.calls.append(len(imgs))
        out = []
        for im in imgs:
            if im.shape[0] == 8:  # "no character found" on every third line (heights cycle 8, 9, 10)
                out.append({"text": None, "nns": [], "char_boxes": [], "word_end_idx": []})
            else:
                out.append({"text": f"m{int(im.mean())}", "nns": ["ab", "cd"],
                            "char_boxes": [[0.4, 1.5, 2.5, 3.6], [4.0, 0.0, 6.0, 2.0]], "word_end_idx": []})
        return out


def test_run_effocr_paths_results_coco_and_save(tmp_path):
    paths, arrays = _write_lines(tmp_path, 7)
    pipe = _FakePipeline()
    res, coco = lineio.run_effocr_paths(paths, pipe, batch_lines=3, workers=2)
    assert pipe.calls == [3, 3, 1]
    kept = [i for i in range(7) if arrays[i].shape[0] != 8]
    assert res == {paths[i]: f"m{int(arrays[i].mean())}" for i in kept}
    assert [im["file_name"] for im in coco["images"]] == [os.path.basename(paths[i]) for i in kept]
    assert [im["id"] for im in coco["images"]] == list(range(len(kept)))
    assert coco["images"][0]["height"] == arrays[kept[0]].shape[0] and coco["images"][0]["width"] == arrays[kept[0]].shape[1]
    assert len(coco["annotations"]) == 2 * len(kept)
    a0 = coco["annotations"][0]
    assert a0["bbox"] == [0, 2, 2, 2] and a0["text"] == "ab" and a0["image_id"] == 0  # round(0.4, 1.5, 2.5, 3.6) = 0, 2, 2, 4
    assert lineio.COCO_JSON_SKELETON["images"] == []  # the skeleton is copied, not filled
    pipe.lang = "en"
    res_en, coco_en = lineio.run_effocr_paths(paths, pipe, batch_lines=64)
    assert res_en == res and coco_en["images"] == [] and coco_en["annotations"] == []
    out = tmp_path / "out"
    lineio.save_output(str(out), paths, res, coco)
    assert json.loads((out / "inference_results.json").read_text()) == res
    assert json.loads((out / "inference_coco.json").read_text()) == coco
    assert sorted(os.listdir(out / "images")) == sorted(os.path.basename(p) for p in paths)
    assert (out / "inference_results.json").read_text().startswith("{\n  ")  # indent=2 like the reference


def test_gt_collect_and_evaluation_against_coco():
    coco = {"images": [{"file_name": "a.png", "text": "hello"}, {"file_name": "b.png", "text": "world"}, {"file_name": "c.png", "text": "x"}]}
    res = {"/some/dir/a.png": "hello", "/other/b.png": "w0rld"}
    assert lineio.gt_collect({"a.png": "hello"}, [("a.png", "hello"), ("zz.png", "q")]) == [("hello", "hello"), ("q", "")]
    acc, cer = lineio.evaluate_against_coco(coco, res)
    assert acc == pytest.approx(100 / 3)  # percent, like utils/eval_utils.py
    assert cer == pytest.approx((0 + 1 + 1) / (5 + 5 + 1))


def test_infer_batches_consumes_lazily_and_matches_sequential(monkeypatch):
    """The overlapped driver pulls its batches ahead of the batch being recognised (so a decoder keeps working) and
    yields what the sequential order yields; stage functions mocked, CUDA stream objects stubbed."""
    import contextlib
    from effocr_b200.infer import EffOCRPipeline

    class P(EffOCRPipeline):
        def __init__(self):  # no device objects
            self.log = []

        def stage_localize(self, chunk):
            self.log.append(("loc", chunk[0]))
            return {"chunk": chunk}

        def launch_recognize(self, st):
            self.log.append(("launch", st["chunk"][0]))
            return len(st["chunk"])

        def finish_recognize(self, st, idx):
            self.log.append(("finish", st["chunk"][0]))
            return [f"r{v}" for v in st["chunk"]]

    monkeypatch.setattr(torch.cuda, "Stream", lambda: object())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    pulled = []

    def gen():
        for chunk in ([0, 1], [2, 3], [], [4], [5, 6]):
            pulled.append(chunk[0] if chunk else None)
            yield chunk

    p = P()
    out = []
    for res in p.infer_batches(gen(), overlap=True):
        out.append(res)
        assert len(pulled) <= len(out) + 3  # never more than three batches ahead of the results handed out
    assert out == [["r0", "r1"], ["r2", "r3"], [], ["r4"], ["r5", "r6"]]
    # batch i+1 is localised AND enqueued before anything waits for batch i: the device never idles on a host round trip
    # (the third batch is empty: nothing is localised or launched for it)
    assert p.log[:7] == [("loc", 0), ("launch", 0), ("loc", 2), ("launch", 2), ("finish", 0), ("finish", 2), ("loc", 4)]
    q = P()
    seq = list(q.infer_batches(gen(), overlap=False))
    assert seq == out
    assert list(P().infer_batches(iter([]), overlap=True)) == []
    assert list(P().infer_batches(iter([[7]]), overlap=True)) == [["r7"]]


def _sharded_worker(rank, world, port, q, paths):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from effocr_b200 import dist as D
    from effocr_b200 import lineio as L
    D.init_from_env("gloo")
    pipe = _FakePipeline()
    sizes = [os.path.getsize(p) for p in paths]
    out = L.run_effocr_paths_sharded(paths, pipe, batch_lines=2, workers=2, weights=sizes if rank >= 0 else None)
    q.put((rank, out, sum(pipe.calls)))
    torch.distributed.destroy_process_group()


def test_run_effocr_paths_sharded_two_ranks_equals_single_process(tmp_path):
    """world_size 2 over gloo: every rank decodes and transcribes its shard only; rank 0 assembles results and the COCO
    structure in input order -- equal to the single-process pair (ids included)."""
    import socket
    import torch.multiprocessing as mp
    paths, _ = _write_lines(tmp_path, 9)
    single = lineio.run_effocr_paths(paths, _FakePipeline(), batch_lines=4)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q, paths)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    (_, out0, n0), (_, out1, n1) = res
    assert out1 == (None, None)
    assert n0 + n1 == len(paths) and 0 < n0 < len(paths)  # the lines were split, none transcribed twice
    assert out0[0] == single[0] and list(out0[0]) == list(single[0])
    assert out0[1] == single[1]


def test_render_dataset_files_match_torchvision_imagefolder_and_chars(tmp_path):
    """File selection and order of the ad-hoc index = the reference's FontImageFolder (a torchvision ImageFolder walk)
    filtered by font name / PAIRED prefix; characters parsed from the file names as infer_effocr.py:197-198 does."""
    from PIL import Image
    from torchvision.datasets import ImageFolder
    layout = {
        "0x41": ["0x41_NotoSerif-Regular.png", "0x41_Other-Font.png", "PAIRED_0x41_NotoSerif-Regular_3.png"],
        "0x3042": ["0x3042_NotoSerifCJKjp-Regular.png", "0x3042_NotoSerif-Regular.jpg"],
        "b": ["b_NotoSerif-Regular.png", "notes.txt"],
        "0x42": ["sub/0x42_NotoSerif-Regular.png"],
    }
    for cls, names in layout.items():
        for n in names:
            p = tmp_path / cls / n
            p.parent.mkdir(parents=True, exist_ok=True)
            if n.endswith(".txt"):
                p.write_text("x")
            else:
                Image.fromarray(np.zeros((5, 5, 3), np.uint8)).save(p)
    font = "NotoSerif-Regular"
    got = lineio.render_dataset_files(str(tmp_path), font)
    ref = [p for p, _ in ImageFolder(str(tmp_path)).samples if font in p and not os.path.basename(p).startswith("PAIRED")]
    assert got == ref and len(got) == 4
    assert lineio.chars_from_render_files(got) == ["あ", "A", "B", "b"]
    assert lineio.render_dataset_files(str(tmp_path), "NotoSerifCJKjp-Regular") == [str(tmp_path / "0x3042" / "0x3042_NotoSerifCJKjp-Regular.png")]

    class Rec:  # stands in for RecognizerPipeline: records what build_ad_hoc_index hands over
        def train_knn(self, glyphs, candidate_chars=None):
            self.shapes, self.chars = [g.shape for g in glyphs], candidate_chars

    rec = Rec()
    chars = lineio.build_ad_hoc_index(str(tmp_path), rec, lang="en")
    assert chars == rec.chars == ["あ", "A", "B", "b"] and rec.shapes == [(5, 5, 3)] * 4
    with pytest.raises(ValueError):
        lineio.build_ad_hoc_index(str(tmp_path), rec, lang="en", font_name="NoSuchFont")
