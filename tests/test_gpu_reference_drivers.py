"""GPU: end-to-end transcription identity against the reference's OWN drivers.

tests/golden/driver_golden.json holds what the unmodified reference scripts print for a 24-line job when their
third-party back-ends are the CPU oracle (oracle/make_driver_golden.py: `infer_effocr_onnx_multi.py` through its own
EffLocalizer / EffRecognizer classes, `infer_effocr.py` through its own encoder factory + an mmdet-style localizer).
Here the same job (tests/driver_fixture.py) runs on the B200:

  * through effocr_b200's own drivers (EffOCRPipeline + lineio.run_effocr_paths) in both box conventions, and
  * when a copy of the reference tree is reachable (EFFOCR_REFERENCE_ROOT, /root/reference, baseline/_ref/effocr --
    `tools/stage_reference.sh` stages one for a gpurun call; it is never committed), through the UNMODIFIED reference
    scripts themselves, run as `__main__` after `effocr_b200.dropin.install()`.

Every transcription must equal the golden one: CER vs the reference = 0 (BASELINE.json metric).
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import driver_fixture as DF  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (DF.available() and DF.DRIVER_GOLDEN.exists()),
                                                  reason="quick-fit weights / driver golden not generated")]
ROOT = Path(__file__).resolve().parent.parent
REF = next((p for p in (os.environ.get("EFFOCR_REFERENCE_ROOT"), "/root/reference", str(ROOT / "baseline" / "_ref" / "effocr"))
            if p and os.path.isfile(os.path.join(p, "infer_effocr_onnx_multi.py"))), None)


def _golden():
    with open(DF.DRIVER_GOLDEN) as f:
        return json.load(f)


def _pipeline(job, convention):
    from effocr_b200 import knn, mmdet_shim
    from effocr_b200.infer import EffOCRPipeline
    from effocr_b200.localizer_engine import EffLocalizer
    from effocr_b200.pipeline import RecognizerPipeline
    from effocr_b200.weights_io import load_encoder_state
    with open(os.path.join(job["recognizer_dir"], "ref.txt")) as f:
        chars = f.read().split()
    index = knn.read_index(os.path.join(job["recognizer_dir"], "ref.index"))
    rec = RecognizerPipeline(load_encoder_state(os.path.join(job["recognizer_dir"], "enc_best.onnx")), index, chars, max_batch=1024)
    if convention == "onnx":  # infer_effocr_onnx_multi.py defaults: conf 0.35, iou 0.01, k = 1
        loc = EffLocalizer(os.path.join(job["localizer_dir"], "best_bbox_mAP.onnx"), iou_thresh=0.01, conf_thresh=0.35,
                           input_shape=None, max_batch=32)
        return EffOCRPipeline(loc, rec, chars, lang="en", knn=1)
    loc = EffLocalizer(os.path.join(job["localizer_dir"], "best_bbox_mAP.pth"), iou_thresh=mmdet_shim.DEFAULT_IOU,
                       conf_thresh=mmdet_shim.DEFAULT_CONF_FLOOR, max_batch=32)
    return EffOCRPipeline(loc, rec, chars, lang="en", knn=10, box_convention="torch", score_thresh=0.3, score_thresh_word=0.3)


def _compare(results, golden, skip_none=False):
    want = {k: v for k, v in golden.items() if not (skip_none and v is None)}
    got = {os.path.basename(k): v for k, v in results.items() if not (skip_none and v is None)}
    diff = {k: (want.get(k), got.get(k)) for k in set(want) | set(got) if want.get(k) != got.get(k)}
    assert not diff, f"{len(diff)} of {len(want)} transcriptions differ from the reference driver's: {list(diff.items())[:3]}"
    assert len(want) >= 20


@pytest.mark.parametrize("convention,key", [("onnx", "infer_effocr_onnx_multi"), ("torch", "infer_effocr")])
def test_own_drivers_reproduce_the_reference_drivers_transcriptions(tmp_path, convention, key):
    from effocr_b200 import lineio, textproc
    job = DF.build(tmp_path)
    golden = _golden()
    pipe = _pipeline(job, convention)
    results, _coco = lineio.run_effocr_paths(job["images"], pipe, batch_lines=8)
    # the torch driver leaves lines without detections out of inference_results (infer_effocr.py:555-556)
    _compare(results, golden[key], skip_none=(convention == "torch"))
    # and the transcriptions are real: the quick-fit models read the rendered text
    pairs = [(golden["ground_truth"][os.path.basename(k)], (v or "").replace(" ", "")) for k, v in results.items()]
    _acc, cer = textproc.textline_evaluation(pairs, no_spaces_in_eval=True)
    assert cer < 0.2, cer


def _run_reference_script(script, argv, tmp_path):
    code = f"""
import runpy, sys
sys.path.insert(0, {str(ROOT)!r})
import effocr_b200.dropin as d
d.install(reference_root={REF!r})
sys.argv = [{script!r}] + {argv!r}
try:
    runpy.run_path({os.path.join(REF, script)!r}, run_name="__main__")
except SystemExit as e:
    assert e.code in (0, None), e.code
from effocr_b200 import _lib
print("LAUNCHES", _lib.load().effocr_launch_count())
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    launches = int(r.stdout.strip().rsplit("LAUNCHES", 1)[1])
    assert launches > 0  # the script ran on the sm_100a kernels, not on a fallback
    return r


@pytest.mark.skipif(REF is None, reason="no copy of the reference tree on this machine (tools/stage_reference.sh)")
def test_unmodified_infer_effocr_onnx_multi_script_on_b200(tmp_path):
    job = DF.build(tmp_path)
    out = str(tmp_path / "out")
    _run_reference_script("infer_effocr_onnx_multi.py",
                          ["--image_dir", job["image_dir"], "--coco_json", job["coco_json"], "--recognizer_dir", job["recognizer_dir"],
                           "--lang", "en", "--localizer_dir", job["localizer_dir"], "--num_threads", "4", "--save_output", out], tmp_path)
    with open(os.path.join(out, "inference_results.json")) as f:
        _compare(json.load(f), _golden()["infer_effocr_onnx_multi"])


@pytest.mark.skipif(REF is None, reason="no copy of the reference tree on this machine (tools/stage_reference.sh)")
def test_unmodified_infer_effocr_script_on_b200(tmp_path):
    job = DF.build(tmp_path)
    out = str(tmp_path / "out")
    _run_reference_script("infer_effocr.py",
                          ["--image_dir", job["image_dir"], "--coco_json", job["coco_json"], "--recognizer_dir", job["recognizer_dir"],
                           "--lang", "en", "--localizer_dir", job["localizer_dir"], "--device", "cuda",
                           "--auto_model_timm", "vit_small_patch16_224", "--save_output", out], tmp_path)
    with open(os.path.join(out, "inference_results.json")) as f:
        _compare(json.load(f), _golden()["infer_effocr"], skip_none=True)
