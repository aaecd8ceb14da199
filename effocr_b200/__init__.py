"""effocr_b200 -- B200-native (sm_100a) EffOCR inference hot path.

Public surface mirrors the reference's engine layer; the arithmetic runs in
hand-written CUDA kernels behind the C ABI in include/effocr_b200.h.
"""
__version__ = "0.1.0"
