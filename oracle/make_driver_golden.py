"""TEST INFRASTRUCTURE -- runs the UNMODIFIED reference drivers, as scripts, over oracle-backed back-ends (CPU) and
stores their transcriptions in tests/golden/driver_golden.json.

    python -m oracle.make_driver_golden            (build container only: needs /root/reference)

Both scripts run through `runpy.run_path(..., run_name="__main__")` with their own argparse command lines:

  * /root/reference/infer_effocr_onnx_multi.py  (YOLO + ONNX route): the reference's OWN `EffLocalizer` /
    `EffRecognizer` classes run unmodified; the third-party session underneath them (`onnxruntime.InferenceSession`)
    is an oracle session -- oracle/yolo.py or oracle/vit.py on the weights next to the `.onnx` placeholder -- and
    `faiss` / `FaissKNN` are oracle/knn.py;
  * /root/reference/infer_effocr.py  (torch route): the reference's own `models/encoders.py` factory runs unmodified
    over a stand-in `timm.create_model` (oracle ViT with timm parameter names); `mmdet.apis` is the oracle localizer
    (reference letterbox + oracle YOLOv5s + reference NMS + un-letterboxing).

Nothing under /root/reference is modified.  Two reference bugs have to be side-stepped from the outside to get the
ONNX script past its start-up: `create_paired_transform(lang=...)` is a TypeError against the
reference's own signature (the harness wraps the function to drop `lang`), and `num_streams` defaults to None
(`--num_threads` is given).

The job (images, weights, index) is tests/driver_fixture.py's; the index vectors -- oracle embeddings of the canonical
glyph renders -- are written to tests/golden/driver_ref_index.npy first.
"""
from __future__ import annotations

import json
import runpy
import sys
import tempfile
import types
from importlib.machinery import ModuleSpec
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import driver_fixture as DF  # noqa: E402
from oracle import knn as OK, ref_harness as rh, transform as OT, vit as OV, yolo as OY  # noqa: E402


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = ModuleSpec(name, None, is_package=True)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _sibling_state(path):
    stem = str(path).rsplit(".", 1)[0]
    return torch.load(stem + ".pth", map_location="cpu")


# ------------------------------------------------------------------ oracle "onnxruntime" + "onnx"
class _IO:
    def __init__(self, name, shape):
        self.name, self.shape = name, shape


class OracleSession:
    """onnxruntime.InferenceSession over the oracle forward passes."""

    def __init__(self, model_path, sess_options=None, providers=None):
        self.sd = _sibling_state(model_path)
        self.is_yolo = "model.24.m.0.weight" in self.sd

    def get_inputs(self):
        # a YOLOv5 export is static 640 x 640 by default; the driver's own default (--localizer_input_shape None) only
        # works with such a model (localizer_engine.py:38-41)
        return [_IO("images", [1, 3, 640, 640])] if self.is_yolo else [_IO("imgs", ["batch", 3, 224, 224])]

    def run(self, _outputs, feeds):
        (x,) = feeds.values()
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        with torch.no_grad():
            y = OY.yolov5s_forward(self.sd, x) if self.is_yolo else OV.vit_forward(self.sd, x)
        return [y.numpy()]


class _SessionOptions:
    intra_op_num_threads = 0


class _OnnxModel:
    class graph:  # noqa: N801
        input = [_IO("images", None)]
        initializer = []


# ------------------------------------------------------------------ oracle faiss / pytorch_metric_learning
class OracleIndex:
    def __init__(self, d):
        self.d, self.xb = d, np.zeros((0, d), np.float32)

    ntotal = property(lambda self: len(self.xb))

    def add(self, x):
        self.xb = np.concatenate([self.xb, np.asarray(x, np.float32).reshape(-1, self.d)])

    def remove_ids(self, ids):
        keep = np.ones(len(self.xb), bool)
        keep[np.asarray(ids)] = False
        self.xb = self.xb[keep]

    def search(self, q, k):
        d, i = OK.flat_ip_search(self.xb, np.asarray(q, np.float32), k)
        return d.numpy(), i.numpy()


def _read_index(path):
    xb = OK.read_index_flat_ip(path)
    ix = OracleIndex(xb.shape[1])
    ix.add(xb)
    return ix


class OracleFaissKNN:
    def __init__(self, reset_before=True, reset_after=True, index_init_fn=None, gpus=None):
        self.index = None

    def load(self, path):
        self.index = _read_index(path)

    def __call__(self, query, k, reference=None, ref_includes_query=False):
        d, i = OK.flat_ip_search(self.index.xb, query.detach().float().cpu(), k)
        return d.to(query.device), i.to(query.device)


class OracleInferenceModel:
    def __init__(self, trunk, embedder=None, match_finder=None, normalize_embeddings=True, knn_func=None, **_):
        self.trunk, self.knn_func = trunk, knn_func

    def load_knn_func(self, path):
        self.knn_func.load(path)


# ------------------------------------------------------------------ oracle "timm"
class OracleTimmViT(torch.nn.Module):
    """What `timm.create_model(name, num_classes=0)` returns, as far as models/encoders.py needs it: an nn.Module whose
    parameters carry timm's names and whose forward is the pooled pre-logits (oracle/vit.py)."""

    def __init__(self, name):
        super().__init__()
        for key, val in OV.init_vit_state_dict(name, seed=0, prefix="").items():
            mod, parts = self, key.split(".")
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, torch.nn.Module())
                mod = mod._modules[p]
            mod.register_parameter(parts[-1], torch.nn.Parameter(val.clone()))

    def to(self, *a, **k):  # encoders.py:56-59 hard-codes device='cuda'; this oracle lives on the CPU
        return self

    def forward(self, x):
        return OV.vit_forward({"net." + k: v for k, v in self.state_dict().items()}, x.float())


# ------------------------------------------------------------------ oracle "mmdet.apis" 
class OracleDetector(torch.nn.Module):
    def __init__(self, sd):
        super().__init__()
        self.sd = sd
        self.nc = sd["model.24.m.0.bias"].numel() // 3 - 5


def oracle_init_detector(config=None, checkpoint=None, device="cpu", cfg_options=None):
    return OracleDetector(torch.load(checkpoint, map_location="cpu"))


def oracle_inference_detector(model, path):
    RefLoc = rh.import_reference("onnx_engines.localizer_engine").EffLocalizer
    import cv2
    im0 = cv2.imread(path)
    h, w = im0.shape[:2]
    x = RefLoc.load_localizer_img(path, (640, 640), backend="yolo")  # the reference's own letterbox
    with torch.no_grad():
        pred = OY.yolov5s_forward(model.sd, torch.from_numpy(x))
    det = RefLoc.non_max_suppression(pred, conf_thres=0.05, iou_thres=0.01, max_det=1000)[0].numpy().astype(np.float32)
    # un-letterbox: subtract the padding, divide by the gain, clip to the image (float32)
    r, _unpad, top, _b, left, _r = OY.letterbox_geometry(h, w, (640, 640))
    gain = np.float32(r)
    det[:, [0, 2]] = np.clip((det[:, [0, 2]] - np.float32(left)) / gain, 0, np.float32(w))
    det[:, [1, 3]] = np.clip((det[:, [1, 3]] - np.float32(top)) / gain, 0, np.float32(h))
    per_class = [np.ascontiguousarray(det[det[:, 5] == c][:, :5]) for c in range(model.nc)]
    return per_class, [[] for _ in range(model.nc)]


class _SymSpell:
    def __init__(self, *a, **k):
        self.words = {}

    def load_dictionary(self, *a, **k):
        return False


def _levenshtein(a, b, **_):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def install_oracle_backends():
    import torchvision  # noqa: F401
    _mod("onnxruntime", InferenceSession=OracleSession, SessionOptions=_SessionOptions,
         get_available_providers=lambda: ["CPUExecutionProvider"])
    _mod("onnx", load=lambda path: _OnnxModel)
    _mod("faiss", IndexFlatIP=OracleIndex, read_index=_read_index)
    _mod("pytorch_metric_learning")
    _mod("pytorch_metric_learning.utils")
    _mod("pytorch_metric_learning.utils.inference", FaissKNN=OracleFaissKNN, InferenceModel=OracleInferenceModel)
    _mod("timm", create_model=lambda name, num_classes=0, pretrained=False, **k: OracleTimmViT(name))
    _mod("timm.data", IMAGENET_DEFAULT_MEAN=OT.IMAGENET_MEAN, IMAGENET_DEFAULT_STD=OT.IMAGENET_STD)
    _mod("mmdet")
    _mod("mmdet.apis", init_detector=oracle_init_detector, inference_detector=oracle_inference_detector)
    _mod("detectron2")
    _mod("detectron2.checkpoint", DetectionCheckpointer=None)
    _mod("detectron2.config", LazyConfig=None, instantiate=None)
    _mod("detectron2.engine")
    _mod("detectron2.engine.defaults", create_ddp_model=None)
    _mod("symspellpy", SymSpell=_SymSpell)
    _mod("nltk")
    _mod("nltk.metrics")
    _mod("nltk.metrics.distance", edit_distance=_levenshtein)
    rh.install_stubs()  # mmcv, deepsparse, albumentations, kornia: empty, never called on this route
    du = rh.import_reference("utils.datasets_utils")
    original = du.create_paired_transform
    if not getattr(original, "_lang_tolerant", False):
        def create_paired_transform(size=224, lang=None):  # App. B: the ONNX driver passes lang=
            return original(size)
        create_paired_transform._lang_tolerant = True
        du.create_paired_transform = create_paired_transform


def run_script(script, argv):
    old = sys.argv
    sys.argv = [script] + argv
    try:
        runpy.run_path(str(rh.REFERENCE_ROOT / script), run_name="__main__")
    except SystemExit as e:
        assert e.code in (0, None), f"{script} exited with {e.code}"
    finally:
        sys.argv = old


def main():
    assert rh.available(), "needs /root/reference"
    assert DF.available(), "needs tests/golden/quickfit_vit_small.npz and quickfit_yolov5s.npz"
    torch.set_num_threads(8)
    vsd = DF.load_npz_state(DF.VIT_WEIGHTS)
    with torch.no_grad():
        x = torch.from_numpy(np.stack([OT.paired_transform(c) for c in DF.prototype_crops()]))
        xb = OV.l2_normalize(OV.vit_forward(vsd, x)).numpy()
    np.save(DF.INDEX_VECTORS, xb)
    install_oracle_backends()
    out = {"n_lines": DF.N_LINES, "seed": DF.SEED}
    with tempfile.TemporaryDirectory() as tmp:
        job = DF.build(tmp, index_vectors=xb)
        common = ["--image_dir", job["image_dir"], "--coco_json", job["coco_json"], "--recognizer_dir", job["recognizer_dir"],
                  "--lang", "en", "--localizer_dir", job["localizer_dir"]]
        run_script("infer_effocr_onnx_multi.py", common + ["--num_threads", "4", "--save_output", tmp + "/out_onnx"])
        run_script("infer_effocr.py", common + ["--device", "cpu", "--auto_model_timm", "vit_small_patch16_224",
                                                "--save_output", tmp + "/out_torch"])
        import os
        for key, d in (("infer_effocr_onnx_multi", "out_onnx"), ("infer_effocr", "out_torch")):
            with open(os.path.join(tmp, d, "inference_results.json")) as f:
                res = json.load(f)
            out[key] = {os.path.basename(k): v for k, v in res.items()}
        out["ground_truth"] = {f"line_{i:03d}.png": "".join(l[3]) for i, l in enumerate(job["lines"])}
    with open(DF.DRIVER_GOLDEN, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for key in ("infer_effocr_onnx_multi", "infer_effocr"):
        hit = sum((out[key].get(k) or "").replace(" ", "") == v for k, v in out["ground_truth"].items())
        print(f"{key}: {len(out[key])} lines transcribed, {hit}/{len(out['ground_truth'])} equal to the rendered text (spaces ignored)")


if __name__ == "__main__":
    main()
