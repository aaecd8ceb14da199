"""`effocr_b200.dropin.install()` makes the UNMODIFIED reference scripts import this package's engines:
it registers modules under the names `infer_effocr.py` / `infer_effocr_onnx_multi.py` import
(SURVEY.md App. C lists them -- they are exactly the places where the reference reaches third-party code).

    import effocr_b200.dropin as d; d.install()
    import infer_effocr_onnx_multi            # now resolves EffLocalizer / EffRecognizer / FaissKNN / faiss here

Modules that already exist (a real `faiss`, a real `timm`) are left alone unless force=True.
"""
from __future__ import annotations

import sys
import types
from importlib.machinery import ModuleSpec


def _module(name: str, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install(force: bool = False):
    from . import datasets_utils, encoders, knn, localizer_engine, recognizer_engine, textproc

    mods = {
        "faiss": _module("faiss", IndexFlatIP=knn.IndexFlatIP, read_index=knn.read_index, write_index=knn.write_index,
                         METRIC_INNER_PRODUCT=knn.METRIC_INNER_PRODUCT),
        "pytorch_metric_learning": _module("pytorch_metric_learning"),
        "pytorch_metric_learning.utils": _module("pytorch_metric_learning.utils"),
        "pytorch_metric_learning.utils.inference": _module("pytorch_metric_learning.utils.inference", FaissKNN=knn.FaissKNN,
                                                           InferenceModel=knn.InferenceModel),
        "models": _module("models"),
        "models.encoders": _module("models.encoders", AutoEncoderFactory=encoders.AutoEncoderFactory),
        "onnx_engines": _module("onnx_engines"),
        "onnx_engines.localizer_engine": _module("onnx_engines.localizer_engine", EffLocalizer=localizer_engine.EffLocalizer),
        "onnx_engines.recognizer_engine": _module("onnx_engines.recognizer_engine", EffRecognizer=recognizer_engine.EffRecognizer),
        "nltk": _module("nltk"),
        "nltk.metrics": _module("nltk.metrics"),
        "nltk.metrics.distance": _module("nltk.metrics.distance", edit_distance=textproc.edit_distance),
    }
    installed = []
    for name, mod in mods.items():
        if force or name not in sys.modules:
            sys.modules[name] = mod
            installed.append(name)
    # expose the hot transform under the reference's module path without shadowing its training augmentations
    du = sys.modules.get("utils.datasets_utils")
    if du is not None:
        du.create_paired_transform = datasets_utils.create_paired_transform
        du.MedianPad = datasets_utils.MedianPad
    return installed
