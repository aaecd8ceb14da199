"""Drop-ins for the nearest-neighbour surface the reference uses from third-party packages:

  * `faiss.IndexFlatIP`, `faiss.read_index`, `faiss.write_index`            -> FlatIPIndex / read_index / write_index
  * `pytorch_metric_learning.utils.inference.FaissKNN`, `InferenceModel`      -> FaissKNN / InferenceModel

Call sites: /root/reference/infer_effocr.py:183-212,317; infer_effocr_onnx_multi.py:496-510,372;
train_effocr_recognizer.py:47-62.  The search runs in csrc/knn.cu (tcgen05 split-fp16 GEMM with an
in-epilogue top-k and an fp32 re-rank); nothing here falls back to torch.matmul or the CPU.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from . import _lib
from .engine import FlatIPIndex

METRIC_INNER_PRODUCT = 0


def IndexFlatIP(d: int) -> FlatIPIndex:  # noqa: N802  (faiss spelling)
    return FlatIPIndex(d)


def write_index(index: FlatIPIndex, path) -> None:
    """faiss IndexFlatIP file layout (faiss/impl/index_write.cpp), little-endian:
    "IxFI" | d i32 | ntotal i64 | dummy i64 x2 | is_trained u8 | metric_type i32 | size u64 | f32 data."""
    xb = np.ascontiguousarray(index.reconstruct_n(), dtype="<f4")
    n, d = xb.shape if xb.size else (0, index.d)
    with open(path, "wb") as f:
        f.write(b"IxFI")
        f.write(struct.pack("<i", d))
        f.write(struct.pack("<q", n))
        f.write(struct.pack("<qq", 1 << 20, 1 << 20))
        f.write(struct.pack("<B", 1))
        f.write(struct.pack("<i", METRIC_INNER_PRODUCT))
        f.write(struct.pack("<Q", n * d))
        f.write(xb.tobytes())


def read_index(path) -> FlatIPIndex:
    with open(path, "rb") as f:
        fourcc = f.read(4)
        if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
            raise _lib.EffocrError(f"{path}: unsupported faiss index type {fourcc!r} (only flat indexes)")
        (d,) = struct.unpack("<i", f.read(4))
        (n,) = struct.unpack("<q", f.read(8))
        f.read(16)  # two dummy i64
        f.read(1)   # is_trained
        (metric,) = struct.unpack("<i", f.read(4))
        if metric > 1:
            f.read(4)  # metric_arg
        if fourcc == b"IxFI" and metric != METRIC_INNER_PRODUCT:
            raise _lib.EffocrError(f"{path}: metric {metric} is not inner product")
        (size,) = struct.unpack("<Q", f.read(8))
        if size != n * d:
            raise _lib.EffocrError(f"{path}: corrupt flat index (size {size} != {n} x {d})")
        data = np.frombuffer(f.read(4 * size), dtype="<f4")
        if data.size != size:
            raise _lib.EffocrError(f"{path}: truncated flat index")
    index = FlatIPIndex(d)
    if n:
        index.add(data.reshape(n, d).copy())
    return index


class FaissKNN:
    """pytorch_metric_learning.utils.inference.FaissKNN with the reference's settings
    (`index_init_fn=faiss.IndexFlatIP, reset_before=False, reset_after=False`)."""

    def __init__(self, reset_before=True, reset_after=True, index_init_fn=None, gpus=None):
        self.reset_before = reset_before
        self.reset_after = reset_after
        self.index_init_fn = IndexFlatIP if index_init_fn is None else index_init_fn
        self.index = None
        self.gpus = gpus

    def __call__(self, query, k, reference=None, ref_includes_query=False):
        if ref_includes_query:
            k = k + 1
        device = query.device if isinstance(query, torch.Tensor) else torch.device("cpu")
        q = torch.as_tensor(query).detach()
        if self.reset_before:
            self.reset()
            if reference is None:
                raise ValueError("reference embeddings are required when reset_before=True")
            self.train(torch.as_tensor(reference))
        elif reference is not None and self.index is None:
            self.train(torch.as_tensor(reference))
        if self.index is None:
            raise ValueError("The index must be trained (or loaded) before it is searched")
        dist, idx = self.index.search_device(q.to("cuda", torch.float32), k)
        if self.reset_after:
            self.reset()
        if ref_includes_query:
            dist, idx = dist[:, 1:], idx[:, 1:]
        return dist.to(device), idx.to(device)

    def train(self, embeddings):
        emb = torch.as_tensor(embeddings).detach().to("cpu", torch.float32)
        self.index = self.index_init_fn(int(emb.shape[1]))
        self.add(emb)

    def add(self, embeddings):
        self.index.add(torch.as_tensor(embeddings).detach().to("cpu", torch.float32))

    def save(self, filename):
        write_index(self.index, filename)

    def load(self, filename):
        self.index = read_index(filename)

    def reset(self):
        self.index = None


class InferenceModel:
    """pytorch_metric_learning.utils.inference.InferenceModel (subset the reference uses):
    `trunk`, `knn_func`, `get_embeddings`, `train_knn`, `add_to_knn`, `get_nearest_neighbors`,
    `save_knn_func`, `load_knn_func`; embeddings are L2-normalised by default."""

    def __init__(self, trunk, embedder=None, match_finder=None, normalize_embeddings=True, knn_func=None,
                 data_device=None, dtype=None):
        self.trunk = trunk
        self.embedder = embedder
        self.normalize_embeddings = normalize_embeddings
        self.knn_func = FaissKNN(reset_before=False, reset_after=False) if knn_func is None else knn_func
        self.data_device = torch.device("cuda" if torch.cuda.is_available() else "cpu") if data_device is None else data_device
        self.dtype = dtype

    def get_embeddings(self, x):
        if isinstance(x, list):
            x = torch.stack([torch.as_tensor(t) for t in x])
        if isinstance(self.trunk, torch.nn.Module):
            self.trunk.eval()
        with torch.no_grad():
            e = self.trunk(torch.as_tensor(x).to(self.data_device))
            if self.embedder is not None:
                e = self.embedder(e)
        if self.normalize_embeddings:
            from . import ops

            e = ops.l2_normalize(e.to("cuda", torch.float32))
        return e

    def _embed_dataset(self, inputs, batch_size=64):
        if isinstance(inputs, torch.Tensor):
            chunks = [inputs[i:i + batch_size] for i in range(0, len(inputs), batch_size)]
        else:
            loader = torch.utils.data.DataLoader(inputs, batch_size=batch_size, shuffle=False)
            chunks = (b[0] if isinstance(b, (list, tuple)) else b for b in loader)
        return torch.cat([self.get_embeddings(c).cpu() for c in chunks], 0)

    def train_knn(self, inputs, batch_size=64):
        self.knn_func.train(self._embed_dataset(inputs, batch_size))

    def add_to_knn(self, inputs, batch_size=64):
        self.knn_func.add(self._embed_dataset(inputs, batch_size))

    def get_nearest_neighbors(self, query, k):
        return self.knn_func(self.get_embeddings(query), k)

    def save_knn_func(self, filename):
        self.knn_func.save(filename)

    def load_knn_func(self, filename):
        self.knn_func.load(filename)
