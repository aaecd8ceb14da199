"""Weight ingestion for the two engines: what the reference loads from disk.

  * recognizer: `enc_best.pth` -- a timm-keyed `net.*` state dict (train_effocr_recognizer.py:65-72,
    models/encoders.py:60,69) -- or its ONNX export `enc_best.onnx` (scripts/recognizer_onnx_export.py:63-69; what
    onnx_engines/recognizer_engine.py:14-21 opens with onnxruntime);
  * localizer: an ultralytics `best.pt` pickle `{'model': <DetectionModel>}` (onnx_engines/infer_ocr_yolo.py:274-276), a
    plain ultralytics-keyed state dict, or the exported graph `best_bbox_mAP.onnx` (onnx_engines/localizer_engine.py:25-31).

ONNX files are NOT executed: only their initializers (the weights) are read, with a ~100-line protobuf wire-format
reader (the `onnx` package is not part of this stack), and mapped back to state-dict names:
  - initializers that kept their parameter name (`net.blocks.0.norm1.weight`, `model.0.conv.weight`, ...) map directly;
  - a `Linear` exported as MatMul(x, W^T) + Add(bias) has an anonymous transposed weight (`onnx::MatMul_123`): it is
    named after the bias its result is added to, and transposed back; `Gemm` nodes likewise (honouring `transB`);
  - an ultralytics export fuses BatchNorm into the convolution (`model.N.conv.weight` + `model.N.conv.bias`): the
    missing BatchNorm is synthesised as the identity (gamma 1, beta = conv bias, mean 0, var 1 - eps);
  - `model.24.anchors` (folded into constants by the exporter) defaults to the published yolov5s anchors.
A path ending in `.onnx` may also simply have its weights next to it (`enc_best.pth`, `best_bbox_mAP.pt`, `.npz`):
the reference's ONNX driver insists on the `.onnx` file names (infer_effocr_onnx_multi.py:470-476), so a sibling
checkpoint is preferred when one exists.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from . import _lib

YOLOV5_ANCHORS = (((10, 13), (16, 30), (33, 23)), ((30, 61), (62, 45), (59, 119)), ((116, 90), (156, 198), (373, 326)))
YOLOV5_STRIDES = (8, 16, 32)
_BN_EPS = 1e-3


# ------------------------------------------------------------------------------------------ protobuf wire format
def _varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field number, wire type, value) over one protobuf message; length-delimited values are memoryviews."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = bytes(buf[pos:pos + 4]), pos + 4
        else:
            raise _lib.EffocrError(f"unsupported protobuf wire type {wt}")
        yield num, wt, val


_ONNX_DTYPES = {1: np.dtype("<f4"), 10: np.dtype("<f2"), 11: np.dtype("<f8"), 6: np.dtype("<i4"), 7: np.dtype("<i8")}


def _tensor(buf):
    """TensorProto -> (name, ndarray or None)."""
    dims, dtype, name, raw, floats, ints, external = [], 1, "", None, [], [], False
    for num, wt, val in _fields(buf):
        if num == 1:  # dims (packed or not)
            if wt == 0:
                dims.append(val)
            else:
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    dims.append(d)
        elif num == 2:
            dtype = val
        elif num == 8:
            name = bytes(val).decode()
        elif num == 9:
            raw = bytes(val)
        elif num == 4:  # float_data
            floats.append(np.frombuffer(bytes(val), "<f4") if wt == 2 else np.frombuffer(val, "<f4"))
        elif num in (5, 7):  # int32_data / int64_data (varints)
            if wt == 0:
                ints.append(val)
            else:
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    ints.append(d)
        elif num == 14 and val == 1:
            external = True
    if external:
        raise _lib.EffocrError(f"ONNX initializer '{name}' uses external data; export with the weights embedded")
    np_dtype = _ONNX_DTYPES.get(dtype)
    if np_dtype is None:
        return name, None
    if raw is not None:
        arr = np.frombuffer(raw, np_dtype)
    elif floats:
        arr = np.concatenate(floats)
    elif ints:
        arr = np.asarray(ints, dtype=np.int64).astype(np_dtype)
    else:
        arr = np.zeros(0, np_dtype)
    return name, arr.reshape(dims) if dims else arr.reshape(())


def _node(buf):
    inputs, outputs, op, attrs = [], [], "", {}
    for num, _wt, val in _fields(buf):
        if num == 1:
            inputs.append(bytes(val).decode())
        elif num == 2:
            outputs.append(bytes(val).decode())
        elif num == 4:
            op = bytes(val).decode()
        elif num == 5:  # AttributeProto: only integer attributes matter here (transB)
            aname, ival = "", None
            for n2, _w2, v2 in _fields(val):
                if n2 == 1:
                    aname = bytes(v2).decode()
                elif n2 == 3:
                    ival = v2
            attrs[aname] = ival
    return op, inputs, outputs, attrs


def read_onnx_initializers(path):
    """-> ({initializer name: ndarray}, [(op_type, inputs, outputs, attrs)]) of the model's main graph."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    graph = None
    for num, wt, val in _fields(buf):
        if num == 7 and wt == 2:
            graph = val
    if graph is None:
        raise _lib.EffocrError(f"{path}: no GraphProto in this file (not an ONNX model?)")
    inits, nodes = {}, []
    for num, wt, val in _fields(graph):
        if num == 5 and wt == 2:
            name, arr = _tensor(val)
            if arr is not None:
                inits[name] = arr
        elif num == 1 and wt == 2:
            nodes.append(_node(val))
    return inits, nodes


def onnx_state_dict(path):
    """Initializers of an exported torch module, under their state-dict names (see the module docstring)."""
    inits, nodes = read_onnx_initializers(path)
    named = {k: v for k, v in inits.items() if not k.startswith("onnx::") and "/" not in k}
    produced_by = {}
    for op, ins, outs, attrs in nodes:
        for o in outs:
            produced_by[o] = (op, ins, attrs)
    for op, ins, outs, attrs in nodes:
        if op == "Gemm" and len(ins) >= 3 and ins[1] in inits and ins[2] in named and ins[2].endswith(".bias"):
            w = inits[ins[1]]
            named.setdefault(ins[2][:-4] + "weight", w if attrs.get("transB") else w.T)
        elif op == "Add":
            bias = [i for i in ins if i in named and i.endswith(".bias")]
            other = [i for i in ins if i not in named]
            if len(bias) == 1 and len(other) == 1 and other[0] in produced_by:
                pop, pins, _ = produced_by[other[0]]
                if pop == "MatMul" and len(pins) == 2 and pins[1] in inits:
                    named.setdefault(bias[0][:-4] + "weight", inits[pins[1]].T)
    return {k: torch.from_numpy(np.array(v, copy=True)) for k, v in named.items()}


# ------------------------------------------------------------------------------------------ checkpoints
def _sibling(path: str):
    stem = os.path.splitext(path)[0]
    for ext in (".pth", ".pt", ".npz"):
        if os.path.exists(stem + ext):
            return stem + ext
    return None


def _load_file(path: str):
    if path.endswith(".npz"):
        return {k: torch.from_numpy(v) for k, v in np.load(path).items()}
    if path.endswith(".onnx"):
        sib = _sibling(path)
        return _load_file(sib) if sib else onnx_state_dict(path)
    try:
        return torch.load(path, map_location="cpu", weights_only=False)
    except (ModuleNotFoundError, AttributeError):
        # an ultralytics best.pt pickles whole modules (models.yolo.DetectionModel, models.common.Conv, ...); without
        # the yolov5 sources on the path the classes are rebuilt as empty nn.Module subclasses -- parameters, buffers
        # and the module tree unpickle into them unchanged, which is all state_dict() needs
        return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_StubPickle)


class _StubUnpickler(pickle.Unpickler):
    _made: dict = {}

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            key = (module, name)
            if key not in self._made:
                self._made[key] = type(name, (torch.nn.Module,), {"__module__": module, "forward": _no_forward})
            return self._made[key]


def _no_forward(self, *a, **k):
    raise _lib.EffocrError("placeholder for a class that is not importable here; only its weights are used")


class _StubPickle:
    """The `pickle_module` interface torch.load expects."""
    __name__ = "effocr_b200_stub_pickle"
    Unpickler = _StubUnpickler
    load = staticmethod(lambda f, **kw: _StubUnpickler(f, **kw).load())


def load_encoder_state(model):
    """-> timm-keyed state dict (with or without the `net.` prefix) from a dict / .pth / .npz / .onnx."""
    sd = model if isinstance(model, dict) else _load_file(str(model))
    if not isinstance(sd, dict):
        sd = sd.state_dict()
    if not any(k.endswith("cls_token") or k.endswith("stem.0.weight") for k in sd):
        raise _lib.EffocrError("recognizer checkpoint holds neither a timm VisionTransformer nor a ConvNeXt state dict")
    return sd


def complete_yolo_state(sd):
    """ultralytics-keyed YOLOv5s weights -> the engine's key set: fused convolutions (`conv.weight` + `conv.bias`, what
    `model.fuse()` / the ONNX export leave) get an identity BatchNorm; missing anchors default to yolov5s.yaml's."""
    out = dict(sd)
    for k in list(sd):
        if k.endswith("conv.bias"):
            p = k[:-len("conv.bias")]
            if p + "bn.weight" not in sd:
                b = torch.as_tensor(sd[k], dtype=torch.float32)
                out[p + "bn.weight"] = torch.ones_like(b)
                out[p + "bn.bias"] = b.clone()
                out[p + "bn.running_mean"] = torch.zeros_like(b)
                out[p + "bn.running_var"] = torch.full_like(b, 1.0 - _BN_EPS)  # sqrt(var + eps) == 1
    if "model.24.anchors" not in out:
        out["model.24.anchors"] = (torch.tensor(YOLOV5_ANCHORS, dtype=torch.float32)
                                   / torch.tensor(YOLOV5_STRIDES, dtype=torch.float32).view(3, 1, 1))
    return out


def load_yolo_state(model):
    """-> ultralytics-keyed (`model.{i}...`) state dict from a dict / best.pt pickle / .pth / .npz / .onnx."""
    sd = model if isinstance(model, dict) else _load_file(str(model))
    if isinstance(sd, dict) and "model" in sd and hasattr(sd["model"], "state_dict"):
        sd = sd["model"].float().state_dict()  # ultralytics best.pt: {'model': DetectionModel, 'epoch': ..., ...}
    elif not isinstance(sd, dict) and hasattr(sd, "state_dict"):
        sd = sd.state_dict()
    if "model.24.m.0.weight" not in sd:
        raise _lib.EffocrError("localizer checkpoint is not an ultralytics-keyed YOLOv5 state dict (no model.24.m.0.weight)")
    return complete_yolo_state(sd)
