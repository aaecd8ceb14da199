"""Handle-owning Python objects over the C ABI: the ViT recognizer encoder and the flat
inner-product index.  These are the objects the reference-facing classes (encoders.py,
recognizer_engine.py, knn.py) delegate to.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _lib, ops

VIT_CONFIGS = {
    # timm name: (embed dim, heads, depth, mlp dim)   -- models/encoders.py:58 `timm.create_model(name)`
    "vit_tiny_patch16_224": (192, 3, 12, 768),
    "vit_small_patch16_224": (384, 6, 12, 1536),
    "vit_base_patch16_224": (768, 12, 12, 3072),
}

_BLOCK_KEYS = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
               "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")


def vit_weight_order(depth: int):
    keys = ["patch_embed.proj.weight", "patch_embed.proj.bias", "cls_token", "pos_embed"]
    for i in range(depth):
        keys += [f"blocks.{i}.{k}" for k in _BLOCK_KEYS]
    keys += ["norm.weight", "norm.bias"]
    return keys


class VitEngine:
    """Device-resident ViT encoder (fp16 weights, fp32 residual stream) behind effocr_vit_*."""

    def __init__(self, state_dict, prefix: str = "net.", max_batch: int = 1024, device=None, ln_eps: float = 1e-6):
        self._lib = _lib.load()
        if device is not None:
            torch.cuda.set_device(device)
        _lib.require_device()
        sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        if "cls_token" not in sd:
            raise _lib.EffocrError(f"state dict has no '{prefix}cls_token': not a timm VisionTransformer checkpoint")
        self.embed_dim = int(sd["cls_token"].shape[-1])
        self.depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
        self.mlp_dim = int(sd["blocks.0.mlp.fc1.weight"].shape[0])
        self.num_heads = self.embed_dim // 64
        if tuple(sd["patch_embed.proj.weight"].shape[1:]) != (3, 16, 16) or tuple(sd["pos_embed"].shape[-2:]) != (197, self.embed_dim):
            raise _lib.EffocrError("only patch16 / 224x224 ViT encoders are supported")
        self.max_batch = int(max_batch)
        keep = []  # keep host arrays alive during the call
        ptrs = (C.c_void_p * (4 + 12 * self.depth + 2))()
        for i, k in enumerate(vit_weight_order(self.depth)):
            a = np.ascontiguousarray(sd[k].detach().to("cpu", torch.float32).numpy())
            keep.append(a)
            ptrs[i] = a.ctypes.data
        h = C.c_void_p()
        _lib.check(self._lib.effocr_vit_create(self.embed_dim, self.num_heads, self.depth, self.mlp_dim, self.max_batch,
                                               float(ln_eps), ptrs, len(ptrs), C.byref(h)), "effocr_vit_create")
        self._h = h
        self._lock = threading.Lock()  # one workspace per handle: serialise concurrent run() callers
        self.device = torch.device("cuda", torch.cuda.current_device())

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.effocr_vit_destroy(h)

    def patch_buffer(self, n_crops: int) -> torch.Tensor:
        """A view of the handle's internal patch-major input buffer ([n*196, 768] fp16) so the crop
        kernel can write the encoder input in place (no copy)."""
        if n_crops > self.max_batch:
            raise _lib.EffocrError("patch_buffer: n_crops exceeds max_batch")
        p = self._lib.effocr_vit_patch_buffer(self._h)
        return _tensor_from_ptr(p, (n_crops * 196, 768), torch.float16, self.device)

    def forward(self, x: torch.Tensor | None, batch: int | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        """x: CUDA f32 [B,3,224,224] | CUDA f16 [B*196,768] | None (input already in patch_buffer)."""
        if x is None:
            kind, B, p = ops.INPUT_PATCH_BUFFER, int(batch), 0
        elif x.dtype == torch.float32 and x.dim() == 4:
            x = ops._cuda(x, torch.float32, "x").contiguous()
            if tuple(x.shape[1:]) != (3, 224, 224):
                raise _lib.EffocrError(f"expected [B,3,224,224], got {tuple(x.shape)}")
            kind, B, p = ops.INPUT_NCHW_F32, x.shape[0], x.data_ptr()
        elif x.dtype == torch.float16 and x.dim() == 2 and x.shape[1] == 768:
            x = ops._cuda(x, torch.float16, "x").contiguous()
            kind, B, p = ops.INPUT_PATCH_F16, x.shape[0] // 196, x.data_ptr()
        else:
            raise _lib.EffocrError("unsupported encoder input")
        if out is None:
            out = torch.empty((B, self.embed_dim), device=self.device, dtype=torch.float32)
        with self._lock:
            _lib.check(self._lib.effocr_vit_forward(self._h, p, kind, B, out.data_ptr(), _lib.stream_ptr()),
                       "effocr_vit_forward")
        return out

    __call__ = forward


CONVNEXT_DEPTHS, CONVNEXT_DIMS = (3, 3, 9, 3), (96, 192, 384, 768)


def convnext_weight_order():
    keys = ["stem.0.weight", "stem.0.bias", "stem.1.weight", "stem.1.bias"]
    for i, depth in enumerate(CONVNEXT_DEPTHS):
        if i > 0:
            keys += [f"stages.{i}.downsample.{n}.{w}" for n in (0, 1) for w in ("weight", "bias")]
        for j in range(depth):
            p = f"stages.{i}.blocks.{j}."
            keys += [p + k for k in ("conv_dw.weight", "conv_dw.bias", "norm.weight", "norm.bias", "mlp.fc1.weight",
                                     "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias", "gamma")]
    return keys + ["head.norm.weight", "head.norm.bias"]


class ConvNextEngine:
    """Device-resident ConvNeXt-Tiny encoder behind effocr_convnext_* (same surface as VitEngine)."""

    embed_dim = 768

    def __init__(self, state_dict, prefix: str = "net.", max_batch: int = 256, device=None):
        self._lib = _lib.load()
        if device is not None:
            torch.cuda.set_device(device)
        _lib.require_device()
        sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        keys = convnext_weight_order()
        missing = [k for k in keys if k not in sd]
        if missing:
            raise _lib.EffocrError(f"not a timm convnext_tiny checkpoint (missing {missing[:3]} ...)")
        keep, ptrs = [], (C.c_void_p * len(keys))()
        for i, k in enumerate(keys):
            a = np.ascontiguousarray(sd[k].detach().to("cpu", torch.float32).numpy())
            keep.append(a)
            ptrs[i] = a.ctypes.data
        h = C.c_void_p()
        _lib.check(self._lib.effocr_convnext_create(int(max_batch), ptrs, len(keys), C.byref(h)), "effocr_convnext_create")
        self._h = h
        self.max_batch = int(max_batch)
        self._lock = threading.Lock()
        self.device = torch.device("cuda", torch.cuda.current_device())

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.effocr_convnext_destroy(h)

    def patch_buffer(self, n_crops: int) -> torch.Tensor:
        if n_crops > self.max_batch:
            raise _lib.EffocrError("patch_buffer: n_crops exceeds max_batch")
        p = self._lib.effocr_convnext_patch_buffer(self._h)
        return _tensor_from_ptr(p, (n_crops * 3136, 48), torch.float16, self.device)

    def forward(self, x: torch.Tensor | None, batch: int | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        if x is None:
            kind, B, p = ops.INPUT_PATCH_BUFFER, int(batch), 0
        else:
            x = ops._cuda(x, torch.float32, "x").contiguous()
            if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
                raise _lib.EffocrError(f"expected [B,3,224,224], got {tuple(x.shape)}")
            kind, B, p = ops.INPUT_NCHW_F32, x.shape[0], x.data_ptr()
        if out is None:
            out = torch.empty((B, 768), device=self.device, dtype=torch.float32)
        with self._lock:
            _lib.check(self._lib.effocr_convnext_forward(self._h, p, kind, B, out.data_ptr(), _lib.stream_ptr()),
                       "effocr_convnext_forward")
        return out

    __call__ = forward


def make_encoder_engine(state_dict, prefix: str = "net.", max_batch: int = 1024):
    """ViT or ConvNeXt engine, chosen from the checkpoint's keys."""
    if prefix + "cls_token" in state_dict:
        return VitEngine(state_dict, prefix=prefix, max_batch=max_batch)
    if prefix + "stem.0.weight" in state_dict:
        return ConvNextEngine(state_dict, prefix=prefix, max_batch=min(max_batch, 4096))
    raise _lib.EffocrError("unrecognised encoder checkpoint (expected timm ViT or ConvNeXt keys)")


def _tensor_from_ptr(ptr: int, shape, dtype, device) -> torch.Tensor:
    """Wrap library-owned device memory as a torch tensor (no ownership transfer)."""
    n = int(np.prod(shape))
    itemsize = torch.empty((), dtype=dtype).element_size()

    class _Holder:
        pass

    holder = _Holder()
    holder.__cuda_array_interface__ = {
        "shape": (n * itemsize,), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}
    t = torch.as_tensor(holder, device=device)
    return t.view(dtype).view(*shape)


class FlatIPIndex:
    """Exact inner-product index (faiss.IndexFlatIP semantics) resident on the GPU."""

    def __init__(self, d: int):
        self._lib = _lib.load()
        self.d = int(d)
        self._h = None
        self._vectors = torch.empty((0, self.d), dtype=torch.float32)  # host master copy (fp32, row-major)
        self._lock = threading.Lock()
        self.is_trained = True

    @property
    def ntotal(self) -> int:
        return int(self._vectors.shape[0])

    def __del__(self):
        self._drop()

    def _drop(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.effocr_knn_destroy(h)

    def reset(self):
        self._drop()
        self._vectors = torch.empty((0, self.d), dtype=torch.float32)

    def add(self, x) -> None:
        x = torch.as_tensor(x).detach().to("cpu", torch.float32).reshape(-1, self.d)
        self._vectors = torch.cat([self._vectors, x], 0).contiguous()
        self._drop()

    def remove_ids(self, ids) -> int:
        """faiss flat-index semantics: rows are compacted, later ids shift down."""
        ids = np.unique(np.asarray(ids, dtype=np.int64).reshape(-1))
        ids = ids[(ids >= 0) & (ids < self.ntotal)]
        keep = np.ones(self.ntotal, dtype=bool)
        keep[ids] = False
        self._vectors = self._vectors[torch.from_numpy(keep)].contiguous()
        self._drop()
        return int(len(ids))

    def reconstruct_n(self, i0: int = 0, n: int | None = None) -> np.ndarray:
        n = self.ntotal - i0 if n is None else n
        return self._vectors[i0:i0 + n].numpy().copy()

    def _handle(self):
        if self._h is None:
            _lib.require_device()
            dev = self._vectors.cuda()
            h = C.c_void_p()
            _lib.check(self._lib.effocr_knn_create(dev.data_ptr(), self.ntotal, self.d, C.byref(h)), "effocr_knn_create")
            torch.cuda.synchronize()
            self._h = h
        return self._h

    def search_device(self, q: torch.Tensor, k: int):
        """q: CUDA f32 [nq, d] -> (distances f32 [nq,k], ids i64 [nq,k]) on the device."""
        q = ops._cuda(q, torch.float32, "queries").contiguous()
        if q.shape[1] != self.d:
            raise _lib.EffocrError(f"query dim {q.shape[1]} != index dim {self.d}")
        nq = q.shape[0]
        dist = torch.empty((nq, k), device=q.device, dtype=torch.float32)
        idx = torch.empty((nq, k), device=q.device, dtype=torch.int64)
        with self._lock:
            _lib.check(self._lib.effocr_knn_search(self._handle(), q.data_ptr(), nq, k, dist.data_ptr(), idx.data_ptr(),
                                                   _lib.stream_ptr()), "effocr_knn_search")
        return dist, idx

    def search(self, x, k: int):
        """faiss-style: numpy in, numpy out."""
        q = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        d, i = self.search_device(q, k)
        return d.cpu().numpy(), i.cpu().numpy()
