"""`effocr_b200.dropin.install()` makes the UNMODIFIED reference scripts import this package's engines:
it registers modules under the names `infer_effocr.py` / `infer_effocr_onnx_multi.py` import
(SURVEY.md App. C lists them -- they are exactly the places where the reference reaches third-party code).

    import effocr_b200.dropin as d; d.install(reference_root="/path/to/effocr")
    import infer_effocr_onnx_multi            # EffLocalizer / EffRecognizer / FaissKNN / faiss resolve here
    import infer_effocr                       # AutoEncoderFactory / InferenceModel / mmdet.apis resolve here

Three kinds of names are registered:
  1. the hot path (engine-backed): `faiss`, `pytorch_metric_learning.utils.inference`, `models.encoders`,
     `onnx_engines.localizer_engine`, `onnx_engines.recognizer_engine`, `mmdet.apis` (the YOLOv5s localizer behind
     `init_detector` / `inference_detector`, effocr_b200/mmdet_shim.py), `nltk.metrics.distance.edit_distance`;
  2. the reference's own packages `models` / `onnx_engines`: their `__path__` points INTO the reference tree (when
     `reference_root` is given or found on sys.path), so everything else in them -- `models.classifiers`,
     `onnx_engines.infer_ocr_yolo` -- still resolves to the reference's files; without a reference tree,
     `models.classifiers` is a stub whose factory raises (the softmax-classifier ablation is out of scope);
  3. third-party packages the drivers import at module level but the YOLO + kNN path never calls (`mmcv`,
     `deepsparse`, `detectron2.*`, `timm`, `albumentations`, `kornia`, `symspellpy`, `onnx`, `onnxruntime`): empty
     stand-ins, registered only when the real package is not importable.

Modules that already exist (a real `faiss`, a real `timm`) are left alone unless force=True.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from importlib.machinery import ModuleSpec


def _module(name: str, path=None, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = ModuleSpec(name, None, is_package=True)
    m.__path__ = list(path or [])
    m.__spec__.submodule_search_locations = m.__path__
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _importable(name: str) -> bool:
    if name in sys.modules:
        return True
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def find_reference_root(reference_root=None):
    """The directory that holds the reference's `models/encoders.py` and `onnx_engines/`: the argument,
    $EFFOCR_REFERENCE_ROOT, or the first sys.path entry that looks like it."""
    cands = [reference_root, os.environ.get("EFFOCR_REFERENCE_ROOT")] + list(sys.path)
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "models", "encoders.py")) and os.path.isdir(os.path.join(c, "onnx_engines")):
            return os.path.abspath(c)
    return None


class _Unavailable:
    """Stand-in for a third-party class the hot path never constructs."""

    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__}: this third-party back-end is not part of effocr_b200 "
                                  "(the B200 path serves the YOLO localizer + timm/hf encoder + kNN route)")


def _unavailable(name):
    return type(name, (_Unavailable,), {})


class _SymSpellStub:
    """symspellpy.SymSpell for `create_worddict()`, which infer_effocr.py:471 calls even with --spell_check off:
    an empty dictionary (spell checking itself is out of scope)."""

    def __init__(self, *a, **k):
        self.words = {}

    def load_dictionary(self, *a, **k):
        return False


def install(force: bool = False, reference_root=None):
    from . import datasets_utils, encoders, knn, localizer_engine, mmdet_shim, recognizer_engine, textproc

    ref = find_reference_root(reference_root)
    if ref is not None and ref not in sys.path:
        sys.path.insert(0, ref)

    def ref_dir(name):
        return [os.path.join(ref, name)] if ref is not None else []

    def classifier_factory(*a, **k):
        raise NotImplementedError("AutoClassifierFactory (--N_classes softmax head) is outside the effocr_b200 hot path")

    hot = {
        "faiss": _module("faiss", IndexFlatIP=knn.IndexFlatIP, read_index=knn.read_index, write_index=knn.write_index,
                         METRIC_INNER_PRODUCT=knn.METRIC_INNER_PRODUCT),
        "pytorch_metric_learning": _module("pytorch_metric_learning"),
        "pytorch_metric_learning.utils": _module("pytorch_metric_learning.utils"),
        "pytorch_metric_learning.utils.inference": _module("pytorch_metric_learning.utils.inference", FaissKNN=knn.FaissKNN,
                                                           InferenceModel=knn.InferenceModel),
        "models": _module("models", path=ref_dir("models")),
        "models.encoders": _module("models.encoders", AutoEncoderFactory=encoders.AutoEncoderFactory),
        "onnx_engines": _module("onnx_engines", path=ref_dir("onnx_engines")),
        "onnx_engines.localizer_engine": _module("onnx_engines.localizer_engine", EffLocalizer=localizer_engine.EffLocalizer),
        "onnx_engines.recognizer_engine": _module("onnx_engines.recognizer_engine", EffRecognizer=recognizer_engine.EffRecognizer),
        "mmdet": _module("mmdet"),
        "mmdet.apis": _module("mmdet.apis", init_detector=mmdet_shim.init_detector,
                              inference_detector=mmdet_shim.inference_detector),
        "nltk": _module("nltk"),
        "nltk.metrics": _module("nltk.metrics"),
        "nltk.metrics.distance": _module("nltk.metrics.distance", edit_distance=textproc.edit_distance),
    }
    if ref is None:
        hot["models.classifiers"] = _module("models.classifiers", AutoClassifierFactory=classifier_factory)
    # third-party names the drivers import at module level; never called on the YOLO + kNN path
    cold = {
        "mmcv": lambda: {"mmcv": _module("mmcv")},
        "deepsparse": lambda: {"deepsparse": _module("deepsparse", compile_model=_unavailable("compile_model")),
                               "deepsparse.pipelines": _module("deepsparse.pipelines"),
                               "deepsparse.pipelines.custom_pipeline": _module("deepsparse.pipelines.custom_pipeline",
                                                                               CustomTaskPipeline=_unavailable("CustomTaskPipeline"))},
        "detectron2": lambda: {"detectron2": _module("detectron2"),
                               "detectron2.checkpoint": _module("detectron2.checkpoint",
                                                                DetectionCheckpointer=_unavailable("DetectionCheckpointer")),
                               "detectron2.config": _module("detectron2.config", LazyConfig=_unavailable("LazyConfig"),
                                                            instantiate=_unavailable("instantiate")),
                               "detectron2.engine": _module("detectron2.engine"),
                               "detectron2.engine.defaults": _module("detectron2.engine.defaults",
                                                                     create_ddp_model=_unavailable("create_ddp_model"))},
        "timm": lambda: {"timm": _module("timm", create_model=_unavailable("create_model")),
                         "timm.data": _module("timm.data", IMAGENET_DEFAULT_MEAN=datasets_utils.IMAGENET_DEFAULT_MEAN,
                                              IMAGENET_DEFAULT_STD=datasets_utils.IMAGENET_DEFAULT_STD)},
        "albumentations": lambda: {"albumentations": _module("albumentations")},
        "kornia": lambda: {"kornia": _module("kornia")},
        "symspellpy": lambda: {"symspellpy": _module("symspellpy", SymSpell=_SymSpellStub, Verbosity=_unavailable("Verbosity"))},
        "onnx": lambda: {"onnx": _module("onnx")},
        "onnxruntime": lambda: {"onnxruntime": _module("onnxruntime", InferenceSession=_unavailable("InferenceSession"))},
    }
    import torch  # noqa: F401  (torch._dynamo probes find_spec("onnx"): import torch before a stub `onnx` exists)
    import torchvision  # noqa: F401

    installed = []
    for name, mod in hot.items():
        if force or name not in sys.modules:
            sys.modules[name] = mod
            installed.append(name)
    for top, make in cold.items():
        if not _importable(top):
            for name, mod in make().items():
                sys.modules[name] = mod
                installed.append(name)
    # bind sub-modules as attributes of their parents (`import a.b` then `a.b.c`)
    for name in installed:
        parent, _, child = name.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, sys.modules[name])
    # expose the hot transform under the reference's module path without shadowing its training augmentations
    du = sys.modules.get("utils.datasets_utils")
    if du is None and ref is not None:
        # the drivers bind the name with `from utils.datasets_utils import *` at import time: the reference module has
        # to exist (and be patched) BEFORE they are imported
        import importlib

        try:
            du = importlib.import_module("utils.datasets_utils")
        except Exception:
            du = None
    if du is not None:
        du.create_paired_transform = datasets_utils.create_paired_transform
        du.MedianPad = datasets_utils.MedianPad
    return installed
