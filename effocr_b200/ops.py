"""Thin torch-tensor wrappers over the C ABI (include/effocr_b200.h).

Every function launches the hand-written sm_100a kernels on the current torch CUDA stream; none
has a PyTorch or CPU fallback.  Tensors must be CUDA tensors on the current device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_SILU = 0, 1, 2
CROP_NCHW_F16, CROP_NCHW_F32, CROP_PATCH_F16, CROP_PATCH4_F16 = 0, 1, 2, 3
INPUT_NCHW_F32, INPUT_PATCH_F16, INPUT_PATCH_BUFFER = 0, 1, 2

IMAGE_DESC_DTYPE = np.dtype([("offset", "<i8"), ("height", "<i4"), ("width", "<i4"), ("pitch", "<i4"),
                             ("reserved", "<i4")])
CROP_BOX_DTYPE = np.dtype([("image", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"), ("y1", "<i4")])


def _cuda(t: torch.Tensor, dtype=None, name="tensor") -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.EffocrError(f"{name} must be a CUDA tensor (effocr_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.EffocrError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def gemm(a: torch.Tensor, w: torch.Tensor, bias=None, act: int = ACT_NONE, out_dtype=torch.float16, resid=None,
         gamma=None, out=None, block_n: int = 0, direct_epilogue: bool = False) -> torch.Tensor:
    """out = act(a @ w.T + bias) * gamma + resid, fp16 operands, fp32 accumulate (tcgen05)."""
    lib = _lib.load()
    a = _cuda(a, torch.float16, "a")
    w = _cuda(w, torch.float16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        ld = (N + 7) // 8 * 8
        out = torch.empty((M, ld), device=a.device, dtype=out_dtype)[:, :N]
    f32 = 1 if out.dtype == torch.float32 else 0
    _lib.check(lib.effocr_gemm_f16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, _lib.ptr(bias),
                                   _lib.ptr(gamma), _lib.ptr(resid), resid.stride(0) if resid is not None else 0,
                                   out.data_ptr(), out.stride(0), act, f32, block_n | (0x10000 if direct_epilogue else 0),
                                   _lib.stream_ptr()), "effocr_gemm_f16")
    return out


def ln_gemm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, w: torch.Tensor, bias=None, eps: float = 1e-6,
            out=None) -> torch.Tensor:
    """(LayerNorm(x) * gamma + beta) @ w.T + bias -> fp16, one fused tcgen05 kernel (x fp32 [M, 384], w fp16 [N, 384])."""
    lib = _lib.load()
    x = _cuda(x, torch.float32, "x")
    w = _cuda(w, torch.float16, "w")
    M, D = x.shape
    N = w.shape[0]
    assert w.shape[1] == D and x.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float16)
    _lib.check(lib.effocr_ln_gemm_f16(x.data_ptr(), x.stride(0), _cuda(gamma, torch.float32, "gamma").data_ptr(),
                                      _cuda(beta, torch.float32, "beta").data_ptr(), float(eps), w.data_ptr(), w.stride(0),
                                      _lib.ptr(bias), out.data_ptr(), out.stride(0), M, N, D, _lib.stream_ptr()), "effocr_ln_gemm_f16")
    return out


def mlp_fused(x: torch.Tensor, h: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
              b2: torch.Tensor) -> torch.Tensor:
    """x += GELU(h @ w1.T + b1) @ w2.T + b2 in place (fp32 residual x, fp16 h / weights); one fused tcgen05 kernel."""
    lib = _lib.load()
    x = _cuda(x, torch.float32, "x")
    h = _cuda(h, torch.float16, "h")
    w1 = _cuda(w1, torch.float16, "w1")
    w2 = _cuda(w2, torch.float16, "w2")
    M, D = x.shape
    HID = w1.shape[0]
    assert h.shape == (M, D) and w1.shape == (HID, D) and w2.shape == (D, HID) and w1.is_contiguous() and w2.is_contiguous()
    assert x.stride(1) == 1 and h.stride(1) == 1
    _lib.check(lib.effocr_mlp_fused_f16(h.data_ptr(), h.stride(0), w1.data_ptr(), _cuda(b1, torch.float32, "b1").data_ptr(),
                                        w2.data_ptr(), _cuda(b2, torch.float32, "b2").data_ptr(), x.data_ptr(), x.stride(0),
                                        M, D, HID, _lib.stream_ptr()), "effocr_mlp_fused_f16")
    return x


def proj_ln(x: torch.Tensor, att: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
            eps: float = 1e-6, out=None) -> torch.Tensor:
    """x += att @ w.T + bias (in place, fp32);  returns LayerNorm(x) * gamma + beta as fp16 -- one fused tcgen05 kernel."""
    lib = _lib.load()
    x = _cuda(x, torch.float32, "x")
    att = _cuda(att, torch.float16, "att")
    w = _cuda(w, torch.float16, "w")
    M, D = x.shape
    assert att.shape == (M, D) and w.shape == (D, D) and w.is_contiguous() and x.stride(1) == 1 and att.stride(1) == 1
    if out is None:
        out = torch.empty((M, D), device=x.device, dtype=torch.float16)
    _lib.check(lib.effocr_proj_ln_f16(att.data_ptr(), att.stride(0), w.data_ptr(), _cuda(bias, torch.float32, "bias").data_ptr(),
                                      x.data_ptr(), x.stride(0), _cuda(gamma, torch.float32, "gamma").data_ptr(),
                                      _cuda(beta, torch.float32, "beta").data_ptr(), float(eps), out.data_ptr(), out.stride(0),
                                      M, D, _lib.stream_ptr()), "effocr_proj_ln_f16")
    return out


def block_tail(x: torch.Tensor, att: torch.Tensor, wp: torch.Tensor, bp: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
               w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """x += att @ wp.T + bp;  x += GELU(LayerNorm(x) @ w1.T + b1) @ w2.T + b2 -- in place, one fused tcgen05 kernel."""
    lib = _lib.load()
    x = _cuda(x, torch.float32, "x")
    att = _cuda(att, torch.float16, "att")
    wp, w1, w2 = _cuda(wp, torch.float16, "wp"), _cuda(w1, torch.float16, "w1"), _cuda(w2, torch.float16, "w2")
    M, D = x.shape
    HID = w1.shape[0]
    assert att.shape == (M, D) and wp.shape == (D, D) and w1.shape == (HID, D) and w2.shape == (D, HID)
    assert wp.is_contiguous() and w1.is_contiguous() and w2.is_contiguous() and x.stride(1) == 1 and att.stride(1) == 1
    f32 = lambda t, n: _cuda(t, torch.float32, n).data_ptr()  # noqa: E731
    _lib.check(lib.effocr_block_tail_f16(att.data_ptr(), att.stride(0), wp.data_ptr(), f32(bp, "bp"), f32(gamma, "gamma"),
                                         f32(beta, "beta"), float(eps), w1.data_ptr(), f32(b1, "b1"), w2.data_ptr(), f32(b2, "b2"),
                                         x.data_ptr(), x.stride(0), M, D, HID, _lib.stream_ptr()), "effocr_block_tail_f16")
    return x


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6,
              out_dtype=torch.float16) -> torch.Tensor:
    lib = _lib.load()
    x = _cuda(x, torch.float32, "x")
    rows, dim = x.shape
    out = torch.empty((rows, dim), device=x.device, dtype=out_dtype)
    _lib.check(lib.effocr_layernorm(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), dim,
                                    rows, dim, eps, 1 if out_dtype == torch.float32 else 0, _lib.stream_ptr()),
               "effocr_layernorm")
    return out


def attention(qkv: torch.Tensor, batch: int, heads: int, impl: int = 0) -> torch.Tensor:
    lib = _lib.load()
    qkv = _cuda(qkv, torch.float16, "qkv")
    tokens = qkv.shape[0] // batch
    out = torch.empty((qkv.shape[0], heads * 64), device=qkv.device, dtype=torch.float16)
    _lib.check(lib.effocr_attention_f16(qkv.data_ptr(), out.data_ptr(), batch, tokens, heads, impl, _lib.stream_ptr()),
               "effocr_attention_f16")
    return out


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    lib = _lib.load()
    x = _cuda(x, torch.float32, "x").contiguous()
    out = torch.empty_like(x)
    _lib.check(lib.effocr_l2_normalize(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], eps, _lib.stream_ptr()),
               "effocr_l2_normalize")
    return out


def crop_resize(pixels: torch.Tensor, images: torch.Tensor, boxes: torch.Tensor, n_boxes: int, layout: int,
                out: torch.Tensor | None = None) -> torch.Tensor:
    """pixels: u8 device buffer; images/boxes: device byte tensors viewed from IMAGE_DESC_DTYPE /
    CROP_BOX_DTYPE arrays (see pack_images / pack_boxes)."""
    lib = _lib.load()
    pixels = _cuda(pixels, torch.uint8, "pixels")
    if out is None:
        if layout == CROP_PATCH_F16:
            out = torch.empty((n_boxes * 196, 768), device=pixels.device, dtype=torch.float16)
        elif layout == CROP_PATCH4_F16:
            out = torch.empty((n_boxes * 3136, 48), device=pixels.device, dtype=torch.float16)
        else:
            out = torch.empty((n_boxes, 3, 224, 224), device=pixels.device,
                              dtype=torch.float32 if layout == CROP_NCHW_F32 else torch.float16)
    _lib.check(lib.effocr_crop_resize(pixels.data_ptr(), images.data_ptr(), boxes.data_ptr(), n_boxes, layout,
                                      out.data_ptr(), _lib.stream_ptr()), "effocr_crop_resize")
    return out


def letterbox_pad(pixels: torch.Tensor, images: torch.Tensor, n_images: int, height: int, width: int) -> torch.Tensor:
    """u8 RGB line images (no larger than height x width) -> letterboxed f32 [n, 3, height, width] on the device."""
    lib = _lib.load()
    pixels = _cuda(pixels, torch.uint8, "pixels")
    out = torch.empty((n_images, 3, height, width), device=pixels.device, dtype=torch.float32)
    _lib.check(lib.effocr_letterbox_pad(pixels.data_ptr(), images.data_ptr(), n_images, height, width, out.data_ptr(),
                                        _lib.stream_ptr()), "effocr_letterbox_pad")
    return out


def letterbox_resize(pixels: torch.Tensor, images: torch.Tensor, shapes, height: int, width: int,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """u8 RGB line images of any size -> the reference's letterboxed model input, f32 [n, 3, height, width], on the
    device: cv2.resize(INTER_LINEAR) restated bit-exactly + grey padding + / 255 (EffLocalizer.load_localizer_img).
    shapes: [(h, w)] of the packed images (host side, for the tap tables)."""
    from .localizer_engine import letterbox_plan

    lib = _lib.load()
    pixels = _cuda(pixels, torch.uint8, "pixels")
    plans, taps = letterbox_plan(list(shapes), (height, width))
    d_plans = torch.from_numpy(plans.view(np.uint8).copy()).to(pixels.device, non_blocking=True)
    d_taps = torch.from_numpy(taps).to(pixels.device, non_blocking=True)
    if out is None:
        out = torch.empty((len(plans), 3, height, width), device=pixels.device, dtype=torch.float32)
    else:  # caller-owned buffer (the pipeline keeps one per model shape: no allocator traffic per batch)
        out = _cuda(out, torch.float32, "out")[:len(plans)]
        assert out.shape[1:] == (3, height, width) and out.is_contiguous()
    _lib.check(lib.effocr_letterbox_resize(pixels.data_ptr(), images.data_ptr(), d_plans.data_ptr(), d_taps.data_ptr(),
                                           len(plans), height, width, out.data_ptr(), _lib.stream_ptr()),
               "effocr_letterbox_resize")
    return out


class PinnedPool:
    """A few reusable pinned staging buffers.  Allocating pinned memory per batch (cudaHostAlloc: an ioctl that page-locks
    the range under a driver-wide lock) costs milliseconds and serialises ACROSS processes -- with several ranks per box the
    per-batch allocations were the whole-pipeline scaling bottleneck.  A buffer is handed out again once the copy that
    last read it has completed (CUDA event)."""

    def __init__(self):
        self._bufs = []  # [tensor, event or None]

    def get(self, nbytes: int):
        for ent in self._bufs:
            if ent[0].numel() >= nbytes and (ent[1] is None or ent[1].query()):
                ent[1] = None
                return ent
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        ent = [t, None]
        self._bufs.append(ent)
        if len(self._bufs) > 8:  # unbounded growth guard: drop the oldest idle buffer
            for i, e in enumerate(self._bufs):
                if e is not ent and (e[1] is None or e[1].query()):
                    del self._bufs[i]
                    break
        return ent


def pack_images(arrays, device="cuda", pinned: bool = True, pool: PinnedPool | None = None):
    """Concatenate u8 HWC RGB images into one device buffer + descriptor table (one H2D copy each).
    `pool`: reuse pinned staging buffers across calls (see PinnedPool)."""
    descs = np.zeros(len(arrays), dtype=IMAGE_DESC_DTYPE)
    off = 0
    for i, a in enumerate(arrays):
        assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3
        h, w, _ = a.shape
        descs[i] = (off, h, w, w * 3, 0)
        off += (h * w * 3 + 255) // 256 * 256
    ent = None
    if pool is not None and torch.cuda.is_available():
        ent = pool.get(off + descs.nbytes + 256)
        host = ent[0]
    else:
        host = torch.empty(max(off, 1) + descs.nbytes + 256, dtype=torch.uint8, pin_memory=pinned and torch.cuda.is_available())
    hv = host.numpy()
    for i, a in enumerate(arrays):
        o = int(descs[i]["offset"])
        hv[o:o + a.size] = np.ascontiguousarray(a).reshape(-1)
    doff = (off + 255) // 256 * 256  # descriptors ride in the same pinned buffer: one upload for pixels + table
    hv[doff:doff + descs.nbytes] = descs.view(np.uint8).reshape(-1)
    dev = host[:doff + descs.nbytes].to(device, non_blocking=True)
    if ent is not None:
        ent[1] = torch.cuda.current_stream().record_event()
    pixels = dev[:max(off, 1)]
    images = dev[doff:doff + descs.nbytes]
    return pixels, images, descs


def pack_boxes(boxes, device="cuda"):
    """boxes: iterable of (image, x0, y0, x1, y1) integer Python-slice rectangles."""
    arr = np.asarray(list(boxes), dtype=np.int32).reshape(-1, 5)
    rec = np.zeros(len(arr), dtype=CROP_BOX_DTYPE)
    for k, name in enumerate(("image", "x0", "y0", "x1", "y1")):
        rec[name] = arr[:, k]
    return torch.from_numpy(rec.view(np.uint8).copy()).to(device, non_blocking=True), len(arr)
