"""CPU: `effocr_b200.dropin.install()` lets the UNMODIFIED reference drivers import (input_shape=None,
`.onnx` paths, the `models` package, mmcv / deepsparse / mmdet / detectron2 at module level).  The reference tree is
only present in the build container; these tests skip elsewhere."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = next((p for p in (os.environ.get("EFFOCR_REFERENCE_ROOT"), "/root/reference", str(ROOT / "baseline" / "_ref" / "effocr"))
            if p and os.path.isfile(os.path.join(p, "infer_effocr_onnx_multi.py"))), None)


def _run(code: str):
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=600)


@pytest.mark.skipif(REF is None, reason="reference tree not present")
def test_unmodified_reference_drivers_import_over_the_dropins():
    # a fresh interpreter: sys.modules must not carry stubs from other tests
    r = _run(f"""
import effocr_b200.dropin as d
d.install(reference_root={REF!r})
import infer_effocr_onnx_multi as onnx_driver, infer_effocr as torch_driver
import effocr_b200.localizer_engine as L, effocr_b200.recognizer_engine as R, effocr_b200.knn as K, effocr_b200.mmdet_shim as M
import effocr_b200.encoders as E
assert onnx_driver.EffLocalizer is L.EffLocalizer and onnx_driver.EffRecognizer is R.EffRecognizer
assert onnx_driver.FaissKNN is K.FaissKNN and onnx_driver.faiss.IndexFlatIP is K.IndexFlatIP
assert torch_driver.init_detector is M.init_detector and torch_driver.inference_detector is M.inference_detector
assert torch_driver.InferenceModel is K.InferenceModel and torch_driver.AutoEncoderFactory is E.AutoEncoderFactory
assert torch_driver.AutoClassifierFactory.__module__ == "models.classifiers"      # the reference's own file, via __path__
assert onnx_driver.run_effocr.__code__.co_filename.startswith({REF!r})            # the unmodified function bodies
assert torch_driver.EffOCR.infer.__code__.co_filename.startswith({REF!r})
assert torch_driver.create_worddict() is not None                                  # infer_effocr.py:471 runs it eagerly
print("ok")
""")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_install_without_a_reference_tree_registers_stub_classifiers():
    r = _run("""
import sys
import effocr_b200.dropin as d
names = d.install(reference_root=None)
assert "models.classifiers" in names or d.find_reference_root() is not None
import faiss, mmdet.apis, onnx_engines.localizer_engine, pytorch_metric_learning.utils.inference as pml
assert faiss.IndexFlatIP(4).d == 4 and callable(mmdet.apis.inference_detector) and pml.FaissKNN
print("ok")
""")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_localizer_accepts_the_drivers_defaults():
    """infer_effocr_onnx_multi.py:457,485 passes input_shape=None by default; construction must not choke on it before
    reaching the device check."""
    import inspect

    from effocr_b200.localizer_engine import EffLocalizer
    src = inspect.getsource(EffLocalizer.__init__)
    assert "input_shape is None" in src
    sig = inspect.signature(EffLocalizer.__init__)
    assert list(sig.parameters)[1:9] == ["model_path", "iou_thresh", "conf_thresh", "vertical", "num_cores", "providers",
                                          "input_shape", "model_backend"]


def test_unletterbox_and_mmdet_result_format():
    import numpy as np

    from effocr_b200 import mmdet_shim
    det = np.array([[10.0, 305.0, 20.0, 335.0, 0.9, 0.0], [0.0, 300.0, 640.0, 340.0, 0.8, 1.0], [630.0, 290.0, 650.0, 350.0, 0.5, 0.0]],
                   dtype=np.float32)
    out = mmdet_shim.unletterbox(det, 64, 1024)
    assert np.allclose(out[0, :4], [16.0, 8.0, 32.0, 56.0]) and np.allclose(out[1, :4], [0, 0, 1024, 64])
    assert np.allclose(out[2, :4], [1008.0, 0.0, 1024.0, 64.0])  # clipped to the image
    bbox, segm = mmdet_shim.format_result(out, nc=2)
    assert [b.shape for b in bbox] == [(2, 5), (1, 5)] and len(segm) == 2
    result = (bbox, segm)
    char, word = result if isinstance(result[0], np.ndarray) else result[0]  # infer_effocr.py:348
    assert char.shape == (2, 5) and word.shape == (1, 5) and result[0][0] is bbox[0]  # jp_preprocess reads result[0][0]
