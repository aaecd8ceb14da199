"""ctypes binding of libeffocr_b200.so (the C ABI declared in include/effocr_b200.h).

There is NO fallback: if the shared library is missing or the device is not sm_100 every product
entry point raises.  PyTorch is used only for device memory, streams and torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libeffocr_b200.so"

_lib = None
_lock = threading.Lock()

c_void_p = C.c_void_p
c_int = C.c_int
c_ll = C.c_longlong
c_float = C.c_float
c_float_p = C.c_void_p  # device pointers are passed as integers


class EffocrError(RuntimeError):
    pass


# name -> (restype, argtypes); every symbol include/effocr_b200.h declares must be listed here
# (tests/test_abi.py cross-checks this table against the header and the .so export table).
SIGNATURES = {
    "effocr_abi_version": (c_int, []),
    "effocr_build_flags": (c_int, []),
    "effocr_last_error": (C.c_char_p, []),
    "effocr_device_ok": (c_int, []),
    "effocr_launch_count": (c_ll, []),
    "effocr_profile_enable": (None, [c_int]),
    "effocr_profile_reset": (None, []),
    "effocr_profile_num_tags": (c_int, []),
    "effocr_profile_tag_name": (C.c_char_p, [c_int]),
    "effocr_profile_read": (c_int, [c_int, c_void_p, c_void_p]),
    "effocr_gemm_f16": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "effocr_ln_gemm_f16": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_float, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int,
                                   c_int, c_void_p]),
    "effocr_mlp_fused_f16": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int,
                                     c_int, c_void_p]),
    "effocr_proj_ln_f16": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_float, c_void_p,
                                   c_ll, c_int, c_int, c_void_p]),
    "effocr_block_tail_f16": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "effocr_crop_resize": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "effocr_letterbox_pad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "effocr_letterbox_resize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "effocr_vit_create": (c_int, [c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p]),
    "effocr_vit_destroy": (None, [c_void_p]),
    "effocr_vit_embed_dim": (c_int, [c_void_p]),
    "effocr_vit_max_batch": (c_int, [c_void_p]),
    "effocr_vit_patch_buffer": (c_void_p, [c_void_p]),
    "effocr_vit_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "effocr_convnext_create": (c_int, [c_int, c_void_p, c_int, c_void_p]),
    "effocr_convnext_destroy": (None, [c_void_p]),
    "effocr_convnext_patch_buffer": (c_void_p, [c_void_p]),
    "effocr_convnext_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "effocr_layernorm": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_float, c_int,
                                 c_void_p]),
    "effocr_attention_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "effocr_yolo_create": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "effocr_yolo_destroy": (None, [c_void_p]),
    "effocr_yolo_set_mode": (c_int, [c_void_p, c_int]),
    "effocr_yolo_num_predictions": (c_int, [c_int, c_int]),
    "effocr_yolo_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "effocr_nms": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "effocr_l2_normalize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "effocr_knn_create": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "effocr_knn_destroy": (None, [c_void_p]),
    "effocr_knn_ntotal": (c_int, [c_void_p]),
    "effocr_knn_dim": (c_int, [c_void_p]),
    "effocr_knn_vectors": (c_void_p, [c_void_p]),
    "effocr_knn_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}


def load(build_if_missing: bool = False):
    """Return the loaded CDLL; raises EffocrError when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            if build_if_missing or os.environ.get("EFFOCR_B200_AUTOBUILD") == "1":
                from . import build as _build

                _build.build()
            else:
                raise EffocrError(
                    f"{LIB_PATH} is missing: run `python -m effocr_b200.build` (needs nvcc). "
                    "effocr_b200 has no CPU or PyTorch fallback for its CUDA kernels.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().effocr_last_error()
        raise EffocrError(f"{what or 'effocr_b200'} failed with status {status}: {msg.decode() if msg else ''}")


def stream_ptr(stream=None) -> int:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


def ptr(t) -> int:
    """Device (or host) address of a torch tensor / None."""
    return 0 if t is None else int(t.data_ptr())


def require_device() -> None:
    """Fail loudly unless a CUDA sm_100 device is current and the extension is loadable."""
    check(load().effocr_device_ok(), "effocr_device_ok")
