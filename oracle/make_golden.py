"""TEST INFRASTRUCTURE -- regenerates tests/golden/* by running the UNMODIFIED reference functions
(imported live from /root/reference through oracle/ref_harness.py) on seeded inputs.

Run in the build container only:  python -m oracle.make_golden
The fixtures pin the oracle (and the product's host logic) on machines where /root/reference is absent
(the GPU box).  Inputs are regenerated from seeds by the tests, only the reference OUTPUTS are stored
(sub-sampled where large).
"""
from __future__ import annotations

import json
import random
import tempfile
from pathlib import Path

import numpy as np
import torch

from . import ref_harness as rh
from . import vit as V

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"

TRANSFORM_SHAPES = [(64, 23), (64, 64), (64, 100), (30, 64), (224, 50), (300, 40), (1, 1), (2, 7), (64, 10), (48, 225), (64, 400)]


def transform_inputs():
    rng = np.random.default_rng(1234)
    return [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in TRANSFORM_SHAPES]


def nms_inputs():
    g = torch.Generator().manual_seed(7)
    cases = []
    for n, nc, conf, iou in [(3000, 2, 0.35, 0.01), (3000, 2, 0.1, 0.45), (500, 1, 0.3, 0.01), (25200, 2, 0.5, 0.01), (50, 2, 0.99, 0.5)]:
        pred = torch.rand(1, n, 5 + nc, generator=g)
        pred[..., 0:2] *= 640
        pred[..., 2:4] = pred[..., 2:4] * 60 + 5
        pred[..., 4] = pred[..., 4] ** 3
        cases.append((pred, conf, iou))
    return cases


def letterbox_inputs():
    out = []
    for (h, w) in [(64, 1024), (100, 300), (640, 640), (700, 50), (33, 977), (64, 640)]:
        out.append(np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 3), dtype=np.uint8))
    return out


def textproc_cases():
    random.seed(0)
    g = torch.Generator().manual_seed(0)
    cases = []
    for _ in range(40):
        n = random.randint(1, 30)
        w = random.randint(0, 6)
        chars = torch.rand(n, 4, generator=g) * 600
        chars[:, 2] = chars[:, 0] + torch.rand(n, generator=g) * 30
        chars[:, 3] = chars[:, 1] + torch.rand(n, generator=g) * 40
        words = torch.rand(w, 4, generator=g) * 600
        if w:
            words[:, 2] = words[:, 0] + 100
        text = "".join(random.choice("abcdeNnrXzw-uO.,") for _ in range(n))
        hs = [float(x) for x in torch.rand(n, generator=g) * 40 + 10]
        bs = [float(x) for x in torch.rand(n, generator=g) * 10 + 50]
        cases.append({"chars": chars.tolist(), "words": words.tolist(), "text": text, "heights": hs, "bottoms": bs})
    return cases


def vit_golden_inputs():
    sd = V.randomize_affine(V.init_vit_state_dict("vit_tiny_patch16_224", seed=0))
    x = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    return sd, x


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    # ---- per-crop transform (utils/datasets_utils.py:166-172), live
    du = rh.import_reference("utils.datasets_utils")
    t = du.create_paired_transform()
    outs = [t(c).numpy() for c in transform_inputs()]
    np.savez_compressed(GOLDEN / "transform_golden.npz",
                        **{f"sub_{i}": o[:, ::7, ::7].astype(np.float32) for i, o in enumerate(outs)},
                        **{f"sum_{i}": np.float64(o.astype(np.float64).sum()) for i, o in enumerate(outs)})
    # ---- NMS + letterbox (onnx_engines/localizer_engine.py), live
    le = rh.import_reference("onnx_engines.localizer_engine").EffLocalizer
    nms = {}
    for i, (pred, conf, iou) in enumerate(nms_inputs()):
        nms[f"out_{i}"] = le.non_max_suppression(pred.clone(), conf_thres=conf, iou_thres=iou, max_det=1000)[0].numpy()
    np.savez_compressed(GOLDEN / "nms_golden.npz", **nms)
    lb = {}
    for i, im in enumerate(letterbox_inputs()):
        ref, ratio, (dw, dh) = le.letterbox(im, (640, 640), stride=32, auto=False)
        lb[f"sub_{i}"] = ref[::9, ::9].copy()
        lb[f"sum_{i}"] = np.int64(ref.astype(np.int64).sum())
        lb[f"geom_{i}"] = np.array([ratio[0], dw, dh], dtype=np.float64)
    np.savez_compressed(GOLDEN / "letterbox_golden.npz", **lb)
    # ---- host text processing (infer_effocr_onnx_multi.py:70-140), live
    m = rh.import_reference("infer_effocr_onnx_multi")
    recs = []
    for c in textproc_cases():
        chars, words = torch.tensor(c["chars"]), torch.tensor(c["words"]).reshape(-1, 4)
        sc, wei = m.en_preprocess(chars, words)
        rec = {"word_end_idx": wei, "order_x0": [float(b[0]) for b in sc],
               "jp_order_y0": [float(b[1]) for b in m.jp_preprocess(chars, vertical=True)], "post": {}}
        for am in (None, 0.1, 0.3):
            rec["post"][str(am)] = m.en_postprocess(c["text"], wei, c["heights"], c["bottoms"], anchor_margin=am)
        recs.append(rec)
    (GOLDEN / "textproc_golden.json").write_text(json.dumps(recs))
    # ---- encoder: the reference's own "hf" back-end (models/encoders.py:72-91), live
    from transformers import ViTConfig, ViTModel

    sd, x = vit_golden_inputs()
    d, h, depth, mlp = V.VIT_CONFIGS["vit_tiny_patch16_224"]
    cfg = ViTConfig(hidden_size=d, num_hidden_layers=depth, num_attention_heads=h, intermediate_size=mlp,
                    layer_norm_eps=1e-6, image_size=224, patch_size=16)
    hf = ViTModel(cfg, add_pooling_layer=False)
    hf.load_state_dict(V.timm_to_hf(sd), strict=True)
    tmp = tempfile.mkdtemp()
    hf.save_pretrained(tmp)
    enc = rh.import_reference("models.encoders").AutoEncoderFactory("hf", tmp)(device="cpu").eval()
    with torch.no_grad():
        e = enc(x).numpy()
    np.savez_compressed(GOLDEN / "vit_tiny_golden.npz", emb=e.astype(np.float32))
    print("golden fixtures written to", GOLDEN)


if __name__ == "__main__":
    main()
