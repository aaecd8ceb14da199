"""TEST INFRASTRUCTURE -- live import of the UNMODIFIED reference modules in the build container.

The reference (dell-research-harvard/effocr @ 0793b03d) is pure Python but its modules import
packages that are not installed here (timm, faiss, onnxruntime, mmcv, ...).  This harness injects
empty stub modules for exactly those names so that the reference's own functions that do not touch
them -- `create_paired_transform` / `MedianPad`, `EffLocalizer.letterbox / non_max_suppression /
xywh2xyxy / box_iou`, `en_preprocess` / `en_postprocess` / `jp_preprocess` / `create_batches`,
`AutoEncoderFactory("hf", ...)` -- execute unmodified and can pin the restatements in oracle/.

It is used ONLY by oracle/make_golden.py (fixture generation) and by CPU tests that skip when
/root/reference is absent.  Nothing in effocr_b200/, bench.py's GPU arm or smoke() imports it, and
/root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import importlib
import sys
import types
from importlib.machinery import ModuleSpec
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "infer_effocr_onnx_multi.py").exists()


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_installed = False


def install_stubs() -> None:
    """Must run after `import torch, torchvision` (torch._dynamo probes find_spec on onnx)."""
    global _installed
    if _installed:
        return
    import torch  # noqa: F401
    import torchvision  # noqa: F401

    class _Missing:
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class: not available in this container")

    def _try(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not _try("timm"):
        _stub("timm")
        _stub("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    for name in ("albumentations", "kornia", "mmcv", "onnx", "onnxruntime", "faiss"):
        if not _try(name):
            _stub(name)
    if not _try("deepsparse"):
        _stub("deepsparse", compile_model=_Missing)
        _stub("deepsparse.pipelines")
        _stub("deepsparse.pipelines.custom_pipeline", CustomTaskPipeline=_Missing)
    if not _try("pytorch_metric_learning"):
        _stub("pytorch_metric_learning")
        _stub("pytorch_metric_learning.utils")
        _stub("pytorch_metric_learning.utils.inference", FaissKNN=_Missing, InferenceModel=_Missing)
    if not _try("symspellpy"):
        _stub("symspellpy", SymSpell=_Missing, Verbosity=_Missing)
    if not _try("nltk"):
        _stub("nltk")
        _stub("nltk.metrics")
        _stub("nltk.metrics.distance", edit_distance=_Missing)
    _installed = True


def import_reference(module: str):
    """Import `module` (e.g. 'utils.datasets_utils') from /root/reference, unmodified."""
    if not available():
        raise RuntimeError("/root/reference is not present (expected on the GPU box)")
    install_stubs()
    root = str(REFERENCE_ROOT)
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module(module)
