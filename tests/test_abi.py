"""CPU: the C-ABI library loads and exports every symbol include/effocr_b200.h declares, the ctypes
table covers the header, and the product never imports the oracle (no compute calls here)."""
import ctypes
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "effocr_b200.h").read_text()
    return sorted(set(re.findall(r"EFFOCR_API[^;(]*?\b(effocr_\w+)\s*\(", text)))


def test_header_declares_symbols():
    names = _declared()
    assert "effocr_gemm_f16" in names and "effocr_vit_forward" in names and "effocr_knn_search" in names
    assert len(names) >= 20


def test_library_exports_every_declared_symbol(lib):
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.effocr_abi_version() == 1


def test_ctypes_table_matches_header():
    from effocr_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_exports_are_only_the_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "effocr_b200" / "libeffocr_b200.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "effocr_" in l)
    assert exported == _declared()


def test_no_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        return
    assert lib.effocr_device_ok() != 0
    assert lib.effocr_last_error()


def test_product_never_imports_oracle():
    for py in (ROOT / "effocr_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{py} imports the oracle"
    assert "oracle" not in (ROOT / "effocr_b200" / "_lib.py").read_text()


def test_product_fails_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from effocr_b200 import _lib
    from effocr_b200.encoders import AutoEncoderFactory
    enc = AutoEncoderFactory("timm", "vit_tiny_patch16_224")(device="cpu")
    with pytest.raises(_lib.EffocrError):
        enc(torch.zeros(1, 3, 224, 224))


def test_dropin_registers_reference_module_names():
    import sys
    saved = {k: sys.modules.get(k) for k in ("faiss", "models.encoders", "onnx_engines.localizer_engine")}
    try:
        from effocr_b200 import dropin
        names = dropin.install(force=True)
        assert "faiss" in names
        import faiss
        from models.encoders import AutoEncoderFactory
        from onnx_engines.localizer_engine import EffLocalizer
        from pytorch_metric_learning.utils.inference import FaissKNN, InferenceModel
        assert callable(faiss.IndexFlatIP) and callable(AutoEncoderFactory) and hasattr(EffLocalizer, "non_max_suppression")
        assert FaissKNN(reset_before=False, reset_after=False).index is None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in ("models", "onnx_engines", "onnx_engines.recognizer_engine", "pytorch_metric_learning",
                  "pytorch_metric_learning.utils", "pytorch_metric_learning.utils.inference", "nltk", "nltk.metrics",
                  "nltk.metrics.distance"):
            sys.modules.pop(k, None)
