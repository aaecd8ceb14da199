"""Drop-in for /root/reference/onnx_engines/recognizer_engine.py: `EffRecognizer`.

Same surface -- `EffRecognizer(model, num_cores=None, providers=None)`, `__call__(imgs)` ==
`run(imgs: np.float32[B,3,224,224]) -> [np.float32[B,D]]` (an ORT-style list; the caller indexes
`embedding[0][0]`, infer_effocr_onnx_multi.py:161-163,371) -- but `run` executes csrc/vit.cu on the
B200 instead of an onnxruntime CPU session.  `run` may be called concurrently from several threads
on one object (infer_effocr_onnx_multi.py:357-364): calls are serialised on the handle's workspace.

`model` is the recognizer checkpoint: a timm-keyed state dict (`net.*`, what the reference's
training script saves as enc_best.pth) given as a path or a dict, or its ONNX export `enc_best.onnx`
(scripts/recognizer_onnx_export.py) -- the initializers are read back into the state dict, the graph itself is
not executed (effocr_b200/weights_io.py).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import VitEngine


def _load_state(model):
    from .weights_io import load_encoder_state

    return load_encoder_state(model)


class EffRecognizer:

    def __init__(self, model, num_cores=None, providers=None, max_batch=1024):
        sd = _load_state(model)
        prefix = "net." if any(k.startswith("net.") for k in sd) else ""
        self._eng_net = VitEngine(sd, prefix=prefix, max_batch=max_batch)
        self.embed_dim = self._eng_net.embed_dim

    def __call__(self, imgs):
        return self.run(imgs)

    def run(self, imgs):
        if isinstance(imgs, torch.Tensor):
            x = imgs.to("cuda", torch.float32, non_blocking=True)
        else:
            x = torch.from_numpy(np.ascontiguousarray(imgs, dtype=np.float32)).cuda(non_blocking=True)
        emb = self._eng_net.forward(x)
        return [emb.cpu().numpy()]

    def run_device(self, x: torch.Tensor) -> torch.Tensor:
        """Device-resident variant: CUDA f32 [B,3,224,224] or f16 patch-major [B*196,768] -> CUDA f32 [B,D]."""
        return self._eng_net.forward(x)
