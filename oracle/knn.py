"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the nearest-neighbour match.

The reference searches a `faiss.IndexFlatIP` through pytorch_metric_learning's FaissKNN
(/root/reference/infer_effocr.py:184-187,317 with k=10; infer_effocr_onnx_multi.py:496-500,372 with
k=1).  Both packages are un-vendored and unpinned (README.md:26, requirements.txt:17) and absent
here, so IndexFlatIP's published behaviour is restated: exact fp32 inner products of every query
against every stored vector, results ordered by descending score.  Tie rule (faiss leaves it
unspecified): lowest id first.  Parity of the encoder is margin-aware; the
search itself must return identical ids wherever the oracle's own fp32 rounding noise
(~1e-7 * sqrt(D)) cannot reorder two candidates.

Also restates the on-disk IndexFlatIP layout (faiss index_write.cpp):
"IxFI" | d i32 | ntotal i64 | dummy i64 x2 | is_trained u8 | metric_type i32 | size u64 | f32 data.
PARITY UNPINNED for the file layout: no real faiss file is available in this container.
"""
from __future__ import annotations

import struct

import numpy as np
import torch


def flat_ip_search(xb, q, k: int, dtype=torch.float32):
    """(distances [nq,k] f32, ids [nq,k] i64); missing results (-3.4028235e38, -1) like faiss."""
    xb = torch.as_tensor(xb, dtype=dtype)
    q = torch.as_tensor(q, dtype=dtype)
    n = xb.shape[0]
    nq = q.shape[0]
    dist = torch.full((nq, k), -3.4028234663852886e38, dtype=torch.float32)
    idx = torch.full((nq, k), -1, dtype=torch.int64)
    if n == 0 or nq == 0:
        return dist, idx
    s = q @ xb.t()
    kk = min(k, n)
    # stable descending sort == (score desc, id asc)
    order = torch.sort(s, dim=1, descending=True, stable=True)
    dist[:, :kk] = order.values[:, :kk].float()
    idx[:, :kk] = order.indices[:, :kk]
    return dist, idx


def margins(xb, q, k: int):
    """fp64 scores and, per query, the smallest gap between consecutive scores among the top (k+1)."""
    s = torch.as_tensor(q, dtype=torch.float64) @ torch.as_tensor(xb, dtype=torch.float64).t()
    top = torch.sort(s, dim=1, descending=True).values[:, :min(k + 1, s.shape[1])]
    if top.shape[1] < 2:
        return s, torch.full((s.shape[0],), float("inf"), dtype=torch.float64)
    return s, (top[:, :-1] - top[:, 1:]).min(dim=1).values


def write_index_flat_ip(path, xb: np.ndarray) -> None:
    xb = np.ascontiguousarray(xb, dtype="<f4")
    n, d = xb.shape
    with open(path, "wb") as f:
        f.write(b"IxFI")
        f.write(struct.pack("<i", d))
        f.write(struct.pack("<q", n))
        f.write(struct.pack("<qq", 1 << 20, 1 << 20))
        f.write(struct.pack("<B", 1))
        f.write(struct.pack("<i", 0))  # METRIC_INNER_PRODUCT
        f.write(struct.pack("<Q", n * d))
        f.write(xb.tobytes())


def read_index_flat_ip(path) -> np.ndarray:
    with open(path, "rb") as f:
        if f.read(4) != b"IxFI":
            raise ValueError("not an IndexFlatIP file")
        (d,) = struct.unpack("<i", f.read(4))
        (n,) = struct.unpack("<q", f.read(8))
        f.read(16)
        f.read(1)
        (metric,) = struct.unpack("<i", f.read(4))
        (size,) = struct.unpack("<Q", f.read(8))
        assert metric == 0 and size == n * d
        return np.frombuffer(f.read(4 * size), dtype="<f4").reshape(n, d).copy()
