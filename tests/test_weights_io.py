"""CPU: on-disk artefacts the reference loads -- a faiss IndexFlatIP file, an ultralytics
best.pt pickle and ONNX exports -- read by effocr_b200 from fixtures built INDEPENDENTLY of the repo's own writers:
the index bytes are spelled out from faiss' index_write.cpp layout, the ONNX files are emitted by a protobuf
wire-format writer that lives in this test, and the best.pt is pickled from classes the loader cannot import."""
import struct
import sys
import types

import numpy as np
import pytest
import torch


# ------------------------------------------------------------------ faiss IndexFlatIP ("IxFI")
def test_read_index_against_hand_built_faiss_bytes(tmp_path):
    """faiss/impl/index_write.cpp: write_index_header (fourcc, d, ntotal, 2 dummies, is_trained, metric_type) then the
    vector as `size_t n_floats` + raw little-endian floats."""
    from effocr_b200 import knn
    x = np.array([[1.0, 0.0, -2.5], [0.25, 4.0, 8.0]], dtype="<f4")
    blob = b"".join([
        b"IxFI",
        (3).to_bytes(4, "little", signed=True),        # d
        (2).to_bytes(8, "little", signed=True),        # ntotal
        (1 << 20).to_bytes(8, "little", signed=True),  # dummy
        (1 << 20).to_bytes(8, "little", signed=True),  # dummy
        b"\x01",                                       # is_trained
        (0).to_bytes(4, "little", signed=True),        # METRIC_INNER_PRODUCT
        (6).to_bytes(8, "little", signed=False),       # number of floats
        bytes.fromhex("0000803f" "00000000" "000020c0" "0000803e" "00008040" "00000041"),
    ])
    assert len(blob) == 4 + 4 + 8 + 16 + 1 + 4 + 8 + 24
    p = tmp_path / "ref.index"
    p.write_bytes(blob)
    index = knn.read_index(str(p))
    assert index.d == 3 and index.ntotal == 2
    assert np.array_equal(index.reconstruct_n(), x)
    # and the writer emits exactly these bytes
    q = tmp_path / "out.index"
    knn.write_index(index, str(q))
    assert q.read_bytes() == blob
    # truncated / foreign files fail loudly
    (tmp_path / "bad.index").write_bytes(blob[:-4])
    with pytest.raises(Exception):
        knn.read_index(str(tmp_path / "bad.index"))
    (tmp_path / "ivf.index").write_bytes(b"IwFl" + blob[4:])
    with pytest.raises(Exception):
        knn.read_index(str(tmp_path / "ivf.index"))


# ------------------------------------------------------------------ a protobuf writer independent of the product reader
def _vi(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _ld(field, payload):
    return _vi(field << 3 | 2) + _vi(len(payload)) + payload


def _tensor_proto(name, arr, raw=True, packed_dims=False):
    arr = np.ascontiguousarray(arr)
    dt = {np.dtype("float32"): 1, np.dtype("float16"): 10, np.dtype("int64"): 7}[arr.dtype]
    out = b""
    if packed_dims:
        out += _ld(1, b"".join(_vi(d) for d in arr.shape))
    else:
        out += b"".join(_vi(1 << 3 | 0) + _vi(d) for d in arr.shape)
    out += _vi(2 << 3 | 0) + _vi(dt)
    out += _ld(8, name.encode())
    if raw:
        out += _ld(9, arr.tobytes())
    else:
        assert arr.dtype == np.float32
        out += _ld(4, arr.tobytes())  # packed float_data
    return out


def _node_proto(op, inputs, outputs, int_attrs=None):
    out = b"".join(_ld(1, i.encode()) for i in inputs) + b"".join(_ld(2, o.encode()) for o in outputs)
    out += _ld(3, (op + "_0").encode()) + _ld(4, op.encode())
    for k, v in (int_attrs or {}).items():
        out += _ld(5, _ld(1, k.encode()) + _vi(3 << 3 | 0) + _vi(v) + _vi(20 << 3 | 0) + _vi(2))
    return out


def _model_proto(nodes, tensors):
    graph = b"".join(_ld(1, n) for n in nodes) + _ld(2, b"torch_jit") + b"".join(_ld(5, t) for t in tensors)
    return _vi(1 << 3 | 0) + _vi(8) + _ld(2, b"pytorch") + _ld(7, graph) + _ld(8, _ld(1, b"") + _vi(2 << 3 | 0) + _vi(11))


def test_onnx_initializers_round_trip_to_state_dict(tmp_path):
    """A Linear exported as MatMul(x, W^T) + Add(bias) keeps only the bias' name; Gemm keeps orientation via transB."""
    from effocr_b200 import weights_io
    rng = np.random.default_rng(0)
    qkv_w = rng.standard_normal((12, 4)).astype(np.float32)   # [out, in]
    qkv_b = rng.standard_normal(12).astype(np.float32)
    fc_w = rng.standard_normal((6, 4)).astype(np.float32)
    fc_b = rng.standard_normal(6).astype(np.float32)
    ln_w = rng.standard_normal(4).astype(np.float16)
    pos = rng.standard_normal((1, 5, 4)).astype(np.float32)
    tensors = [
        _tensor_proto("net.pos_embed", pos, raw=False),
        _tensor_proto("net.blocks.0.norm1.weight", ln_w, packed_dims=True),
        _tensor_proto("net.blocks.0.attn.qkv.bias", qkv_b),
        _tensor_proto("onnx::MatMul_711", np.ascontiguousarray(qkv_w.T)),
        _tensor_proto("net.blocks.0.mlp.fc1.bias", fc_b),
        _tensor_proto("onnx::Gemm_9", fc_w),
        _tensor_proto("/Constant_3_output_0", np.array([1, 5], dtype=np.int64)),
    ]
    nodes = [
        _node_proto("MatMul", ["/blocks.0/norm1/Add_1_output_0", "onnx::MatMul_711"], ["/blocks.0/attn/qkv/MatMul_output_0"]),
        _node_proto("Add", ["net.blocks.0.attn.qkv.bias", "/blocks.0/attn/qkv/MatMul_output_0"], ["/blocks.0/attn/qkv/Add_output_0"]),
        _node_proto("Gemm", ["/x", "onnx::Gemm_9", "net.blocks.0.mlp.fc1.bias"], ["/y"], {"transB": 1}),
    ]
    p = tmp_path / "enc_best.onnx"
    p.write_bytes(_model_proto(nodes, tensors))
    sd = weights_io.onnx_state_dict(str(p))
    assert set(sd) == {"net.pos_embed", "net.blocks.0.norm1.weight", "net.blocks.0.attn.qkv.bias", "net.blocks.0.attn.qkv.weight",
                       "net.blocks.0.mlp.fc1.bias", "net.blocks.0.mlp.fc1.weight"}
    assert np.array_equal(sd["net.blocks.0.attn.qkv.weight"].numpy(), qkv_w)
    assert np.array_equal(sd["net.blocks.0.mlp.fc1.weight"].numpy(), fc_w)
    assert np.array_equal(sd["net.pos_embed"].numpy(), pos)
    assert sd["net.blocks.0.norm1.weight"].dtype == torch.float16 and np.array_equal(sd["net.blocks.0.norm1.weight"].numpy(), ln_w)


def test_full_vit_onnx_export_shape_recovers_every_weight(tmp_path):
    """Every tensor of a (1-block) timm-keyed ViT survives the export naming: parameters by name, Linear weights
    through their MatMul/Add pairing."""
    from effocr_b200 import weights_io
    from effocr_b200.engine import vit_weight_order
    rng = np.random.default_rng(1)
    d, mlp = 64, 256
    shapes = {"patch_embed.proj.weight": (d, 3, 16, 16), "patch_embed.proj.bias": (d,), "cls_token": (1, 1, d), "pos_embed": (1, 197, d),
              "norm.weight": (d,), "norm.bias": (d,)}
    for k, shp in (("norm1.weight", (d,)), ("norm1.bias", (d,)), ("attn.qkv.weight", (3 * d, d)), ("attn.qkv.bias", (3 * d,)),
                   ("attn.proj.weight", (d, d)), ("attn.proj.bias", (d,)), ("norm2.weight", (d,)), ("norm2.bias", (d,)),
                   ("mlp.fc1.weight", (mlp, d)), ("mlp.fc1.bias", (mlp,)), ("mlp.fc2.weight", (d, mlp)), ("mlp.fc2.bias", (d,))):
        shapes["blocks.0." + k] = shp
    ref = {"net." + k: rng.standard_normal(s).astype(np.float32) for k, s in shapes.items()}
    tensors, nodes, anon = [], [], 0
    for k, v in ref.items():
        if v.ndim == 2 and k.endswith(".weight"):  # Linear: anonymous transposed initializer + MatMul/Add
            anon += 1
            tensors.append(_tensor_proto(f"onnx::MatMul_{900 + anon}", np.ascontiguousarray(v.T)))
            nodes.append(_node_proto("MatMul", [f"/in_{anon}", f"onnx::MatMul_{900 + anon}"], [f"/mm_{anon}"]))
            nodes.append(_node_proto("Add", [k[:-6] + "bias", f"/mm_{anon}"], [f"/out_{anon}"]))
        else:
            tensors.append(_tensor_proto(k, v))
    p = tmp_path / "enc.onnx"
    p.write_bytes(_model_proto(nodes, tensors))
    sd = weights_io.load_encoder_state(str(p))
    assert set("net." + k for k in vit_weight_order(1)) <= set(sd)
    for k, v in ref.items():
        assert np.array_equal(sd[k].numpy(), v), k


def test_onnx_path_prefers_sibling_checkpoint(tmp_path):
    """infer_effocr_onnx_multi.py:470-476 insists on `enc_best.onnx` / `best_bbox_mAP.onnx`; a .pth next to it wins."""
    from effocr_b200 import weights_io
    sd = {"net.cls_token": torch.zeros(1, 1, 8), "net.x": torch.arange(4.0)}
    torch.save(sd, tmp_path / "enc_best.pth")
    (tmp_path / "enc_best.onnx").write_bytes(b"not a protobuf")
    got = weights_io.load_encoder_state(str(tmp_path / "enc_best.onnx"))
    assert torch.equal(got["net.x"], sd["net.x"])


# ------------------------------------------------------------------ YOLOv5 artefacts
def _tiny_yolo_sd(fused: bool):
    sd = {"model.24.m.0.weight": torch.randn(21, 8, 1, 1), "model.24.m.0.bias": torch.randn(21)}
    if fused:
        sd["model.0.conv.weight"], sd["model.0.conv.bias"] = torch.randn(4, 3, 6, 6), torch.randn(4)
    else:
        sd["model.0.conv.weight"] = torch.randn(4, 3, 6, 6)
        for k, v in (("weight", torch.rand(4) + 0.5), ("bias", torch.randn(4)), ("running_mean", torch.randn(4)), ("running_var", torch.rand(4) + 0.5)):
            sd["model.0.bn." + k] = v
    return sd


def test_fused_yolo_export_gets_identity_batchnorm(tmp_path):
    from effocr_b200 import weights_io
    sd = _tiny_yolo_sd(fused=True)
    tensors = [_tensor_proto(k, v.numpy()) for k, v in sd.items()]
    p = tmp_path / "best_bbox_mAP.onnx"
    p.write_bytes(_model_proto([], tensors))
    got = weights_io.load_yolo_state(str(p))
    b = sd["model.0.conv.bias"]
    # y = (conv(x) - mean) / sqrt(var + 1e-3) * gamma + beta == conv(x) + bias
    scale = got["model.0.bn.weight"] / torch.sqrt(got["model.0.bn.running_var"] + 1e-3)
    assert torch.allclose(scale, torch.ones(4), atol=1e-7) and torch.equal(got["model.0.bn.bias"], b)
    assert torch.equal(got["model.0.bn.running_mean"], torch.zeros(4))
    anchors = got["model.24.anchors"]
    assert anchors.shape == (3, 3, 2) and torch.equal(anchors[0, 0], torch.tensor([10 / 8, 13 / 8]))


def test_ultralytics_best_pt_pickle_without_yolov5_sources(tmp_path):
    """best.pt = {'model': DetectionModel(...), ...} pickled with classes from `models.yolo` / `models.common`
    (onnx_engines/infer_ocr_yolo.py:274-276).  The loader must rebuild the state dict without those modules."""
    from effocr_b200 import weights_io
    want = _tiny_yolo_sd(fused=False)
    saved = {k: sys.modules.get(k) for k in ("models", "models.yolo", "models.common")}
    try:
        pkg, yolo, common = types.ModuleType("models"), types.ModuleType("models.yolo"), types.ModuleType("models.common")
        pkg.__path__ = []

        class Conv(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.conv = torch.nn.Conv2d(3, 4, 6, bias=False)
                self.bn = torch.nn.BatchNorm2d(4)

        class Detect(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.m = torch.nn.ModuleList([torch.nn.Conv2d(8, 21, 1)])
                self.register_buffer("anchors", torch.ones(3, 3, 2))

        class DetectionModel(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.model = torch.nn.Sequential()
                self.model.add_module("0", Conv())
                self.model.add_module("24", Detect())

        for cls, mod in ((Conv, common), (Detect, yolo), (DetectionModel, yolo)):
            cls.__module__, cls.__qualname__ = mod.__name__, cls.__name__
            setattr(mod, cls.__name__, cls)
        sys.modules.update({"models": pkg, "models.yolo": yolo, "models.common": common})
        net = DetectionModel()
        sdn = net.state_dict()
        for k, v in want.items():
            sdn[k].copy_(v)
        torch.save({"epoch": -1, "model": net.half(), "optimizer": None}, tmp_path / "best.pt")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert "models.yolo" not in sys.modules or saved["models.yolo"] is not None
    got = weights_io.load_yolo_state(str(tmp_path / "best.pt"))
    for k, v in want.items():
        assert torch.allclose(got[k].float(), v.half().float()), k
    assert got["model.24.anchors"].shape == (3, 3, 2)
