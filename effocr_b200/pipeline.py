"""Device-resident recognizer pipeline: boxes -> fused crop -> ViT -> L2 norm -> exact kNN.

This is the B200-first replacement for the per-line loop of /root/reference/infer_effocr.py:280-338
and phases 2-3 of infer_effocr_onnx_multi.py:307-386: all crops of all lines go through ONE crop
launch that writes the encoder's patch buffer in place, embeddings and scores never leave the
device, and only the k ids (and distances) per character come back to the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from .engine import ConvNextEngine, FlatIPIndex, VitEngine, make_encoder_engine


class PackedCrops:
    """Pinned host staging for one batch: pixels + image descriptors + crop boxes."""

    def __init__(self, arrays, boxes=None, pool: "ops.PinnedPool | None" = None):
        """`pool`: take the pixel staging buffer from a PinnedPool instead of page-locking a fresh one (a per-batch
        cudaHostAlloc costs milliseconds); the buffer returns to the pool once `to_device()`'s copy has completed."""
        descs = np.zeros(len(arrays), dtype=ops.IMAGE_DESC_DTYPE)
        off = 0
        for i, a in enumerate(arrays):
            h, w, _ = a.shape
            descs[i] = (off, h, w, w * 3, 0)
            off += (h * w * 3 + 255) // 256 * 256
        pin = torch.cuda.is_available()
        self._pool_entry = None
        if pool is not None and pin:
            self._pool_entry = pool.get(max(off, 1))
            self.pixels = self._pool_entry[0][:max(off, 1)]
        else:
            self.pixels = torch.empty(max(off, 1), dtype=torch.uint8, pin_memory=pin)
        hv = self.pixels.numpy()
        for i, a in enumerate(arrays):
            o = int(descs[i]["offset"])
            hv[o:o + a.size] = np.ascontiguousarray(a).reshape(-1)
        if boxes is None:
            boxes = [(i, 0, 0, a.shape[1], a.shape[0]) for i, a in enumerate(arrays)]
        arr = np.asarray(list(boxes), dtype=np.int32).reshape(-1, 5)
        rec = np.zeros(len(arr), dtype=ops.CROP_BOX_DTYPE)
        for k, name in enumerate(("image", "x0", "y0", "x1", "y1")):
            rec[name] = arr[:, k]
        self.images = torch.from_numpy(descs.view(np.uint8).copy())
        self.boxes = torch.from_numpy(rec.view(np.uint8).copy())
        if pin:
            self.images = self.images.pin_memory()
            self.boxes = self.boxes.pin_memory()
        self.n = len(arr)

    @property
    def h2d_bytes(self) -> int:
        return self.pixels.numel() + self.images.numel() + self.boxes.numel()

    def to_device(self):
        out = (self.pixels.cuda(non_blocking=True), self.images.cuda(non_blocking=True),
               self.boxes.cuda(non_blocking=True), self.n)
        if self._pool_entry is not None:
            self._pool_entry[1] = torch.cuda.current_stream().record_event()
        return out


class RecognizerPipeline:

    def __init__(self, encoder_state, index, candidate_chars=None, max_batch: int = 1024, prefix: str | None = None):
        if prefix is None:
            prefix = "net." if any(k.startswith("net.") for k in encoder_state) else ""
        self.encoder = make_encoder_engine(encoder_state, prefix=prefix, max_batch=max_batch)  # ViT or ConvNeXt
        max_batch = self.encoder.max_batch
        self._crop_layout = ops.CROP_PATCH4_F16 if isinstance(self.encoder, ConvNextEngine) else ops.CROP_PATCH_F16
        if not isinstance(index, FlatIPIndex):
            vec = torch.as_tensor(index, dtype=torch.float32)
            index = FlatIPIndex(vec.shape[1])
            index.add(vec)
        self.index = index
        self.candidate_chars = candidate_chars
        self.max_batch = max_batch

    # ---- device-resident stage-wise API
    def embed_boxes(self, pixels, images, boxes, n: int) -> torch.Tensor:
        """-> L2-normalised embeddings, CUDA f32 [n, D]."""
        out = torch.empty((n, self.encoder.embed_dim), device=pixels.device, dtype=torch.float32)
        itemsize = ops.CROP_BOX_DTYPE.itemsize
        for b0 in range(0, n, self.max_batch):
            b = min(self.max_batch, n - b0)
            ops.crop_resize(pixels, images, boxes[b0 * itemsize:], b, self._crop_layout, out=self.encoder.patch_buffer(b))
            self.encoder.forward(None, batch=b, out=out[b0:b0 + b])
        return ops.l2_normalize(out)

    def embed_crops(self, crops) -> torch.Tensor:
        """crops: list of u8 [h, w, 3] arrays (each taken whole) -> L2-normalised embeddings, CUDA f32 [n, D]; chunks of
        max_batch through the same crop kernel and encoder as recognition (the composition bench.py builds its
        10 000-glyph index with)."""
        out = torch.empty((len(crops), self.encoder.embed_dim), device="cuda", dtype=torch.float32)
        for i0 in range(0, len(crops), self.max_batch):
            pixels, images, boxes, n = PackedCrops(crops[i0:i0 + self.max_batch]).to_device()
            out[i0:i0 + n] = self.embed_boxes(pixels, images, boxes, n)
        return out

    def train_knn(self, glyph_crops, candidate_chars=None) -> None:
        """Index construction on the device: what `InferenceModel.train_knn(render_dataset)`
        does in the reference (infer_effocr.py:190-201: paired transform on the host per glyph, trunk forward in batches
        of 64, `IndexFlatIP.add`) -- here the rendered glyph images go through the fused crop kernel and the encoder in
        batches of max_batch and the normalised prototypes REPLACE the index; `candidate_chars` (one per glyph) replaces
        the label list when given."""
        vectors = self.embed_crops(list(glyph_crops))
        index = FlatIPIndex(self.encoder.embed_dim)
        index.add(vectors)
        self.index = index
        if candidate_chars is not None:
            candidate_chars = list(candidate_chars)
            if len(candidate_chars) != index.ntotal:
                raise ValueError(f"{len(candidate_chars)} candidate chars for {index.ntotal} glyph renders")
            self.candidate_chars = candidate_chars

    def recognize_device(self, pixels, images, boxes, n: int, k: int = 10):
        emb = self.embed_boxes(pixels, images, boxes, n)
        dist, idx = self.index.search_device(emb, k)
        return dist, idx, emb

    # ---- host-facing API (what a caller holding numpy crops uses)
    def recognize_packed(self, packed: PackedCrops, k: int = 10):
        pixels, images, boxes, n = packed.to_device()
        dist, idx, _ = self.recognize_device(pixels, images, boxes, n, k)
        return dist.cpu().numpy(), idx.cpu().numpy()

    def recognize_stream(self, packed_batches, k: int = 10):
        """Generator over (distances, ids) numpy pairs, one per PackedCrops batch, with ONE batch in flight: the upload
        and the kernels of batch i+1 are enqueued before the host waits for the results of batch i, whose ids and
        distances come back through pinned buffers with an asynchronous copy.  Same results as recognize_packed per
        batch; the GPU never idles on the host round trip (the reference does four host round trips per line,
        infer_effocr.py:280-338)."""
        copy_stream = getattr(self, "_copy_stream", None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()

        def upload(packed):  # H2D on the copy stream: overlaps the kernels of the batch in flight
            with torch.cuda.stream(copy_stream):
                bufs = packed.to_device()
                return bufs, copy_stream.record_event()

        it = iter(packed_batches)
        nxt = next(it, None)
        up = upload(nxt) if nxt is not None else None
        pending = None
        while up is not None:
            (pixels, images, boxes, n), ready = up
            nxt = next(it, None)
            up = upload(nxt) if nxt is not None else None  # next batch's upload is enqueued before this batch's kernels
            main.wait_event(ready)
            dist, idx, _ = self.recognize_device(pixels, images, boxes, n, k)
            # results come back through a small ring of persistent pinned buffers: a fresh pinned allocation per batch
            # goes through the caching host allocator, which falls back to cudaHostAlloc (milliseconds, device-wide
            # synchronisation) whenever its event for a recycled block has not completed yet
            ring = getattr(self, "_out_ring", None)
            if ring is None or ring["k"] != k or ring["rows"] < dist.shape[0]:
                rows = max(2 * int(dist.shape[0]), 1024)
                ring = self._out_ring = {"k": k, "rows": rows, "next": 0,
                                         "bufs": [(torch.empty((rows, k), dtype=dist.dtype, pin_memory=True),
                                                   torch.empty((rows, k), dtype=idx.dtype, pin_memory=True)) for _ in range(3)]}
            slot = ring["next"]
            ring["next"] = (slot + 1) % 3
            h_dist, h_idx = (t[:dist.shape[0]] for t in ring["bufs"][slot])
            h_dist.copy_(dist, non_blocking=True)
            h_idx.copy_(idx, non_blocking=True)
            ev = main.record_event()
            for t in (pixels, images, boxes):
                t.record_stream(main)  # allocated on the copy stream, consumed on the compute stream
            cur = (h_dist, h_idx, ev, (pixels, images, boxes, dist, idx))  # keep device buffers alive until consumed
            if pending is not None:
                pending[2].synchronize()
                yield pending[0].numpy().copy(), pending[1].numpy().copy()  # the ring slot is reused three batches later
            pending = cur
        if pending is not None:
            pending[2].synchronize()
            yield pending[0].numpy().copy(), pending[1].numpy().copy()

    def recognize_crops(self, crops, k: int = 10):
        """crops: list of u8 [h, w, 3] arrays -> (distances [n,k], ids [n,k]) numpy."""
        if len(crops) == 0:
            return np.zeros((0, k), np.float32), np.zeros((0, k), np.int64)
        return self.recognize_packed(PackedCrops(crops), k)

    def decode(self, idx: np.ndarray):
        """ids -> nearest-neighbour strings per char and the line text (infer_effocr.py:318-338)."""
        nearest = [[self.candidate_chars[j] for j in row if j >= 0] for row in idx.tolist()]
        return nearest, "".join(x[0] for x in nearest).strip()
