/* effocr_b200 -- C ABI of the B200-native EffOCR inference hot path.
 *
 * The reference (dell-research-harvard/effocr) is pure Python and reaches its arithmetic only
 * through third-party libraries (timm / onnxruntime / faiss / torchvision / OpenCV).  This
 * header is the boundary a maintainer would bind (ctypes stub in INTEGRATION.md) to replace
 * those library calls on the hot path  localize -> crop -> embed -> kNN  (SURVEY.md section 8a).
 * Each entry point names the reference call site it replaces.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory on the current CUDA device (sm_100a required);
 *     h_* is host memory; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - all calls are stream-ordered and never synchronise unless documented;
 *   - return value 0 = success, otherwise an EFFOCR_ERR_* code; effocr_last_error() returns a
 *     thread-local message; nothing throws, the caller owns every buffer it passes;
 *   - handles are immutable after create and may be used from several host threads as long as
 *     each thread passes its own workspace-owning handle or serialises calls on one handle
 *     (the Python shim holds a lock per handle, mirroring ORT's thread-safe session.run()).
 */
#ifndef EFFOCR_B200_H_
#define EFFOCR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFFOCR_B200_ABI_VERSION 1

#if defined(EFFOCR_BUILDING)
#define EFFOCR_API __attribute__((visibility("default")))
#else
#define EFFOCR_API
#endif

#define EFFOCR_OK 0
#define EFFOCR_ERR_INVALID 1
#define EFFOCR_ERR_CUDA 2
#define EFFOCR_ERR_NO_DEVICE 3
#define EFFOCR_ERR_NOMEM 4

/* ---- library ---------------------------------------------------------------------------- */
EFFOCR_API int effocr_abi_version(void);
EFFOCR_API const char* effocr_last_error(void);
/* 0 when the current CUDA device is an sm_100 part, EFFOCR_ERR_NO_DEVICE / _CUDA otherwise. */
EFFOCR_API int effocr_device_ok(void);

/* ---- tcgen05 GEMM building block ----------------------------------------------------------
 * out[M,N] = act(A[M,K] . W[N,K]^T + bias) * gamma + resid      fp16 operands, fp32 accumulate.
 * Replaces the ATen/oneDNN/ORT matmul + elementwise calls inside timm Block.forward and yolov5
 * Conv.forward that the reference reaches through models/encoders.py:62-64 and
 * onnx_engines/recognizer_engine.py:27 / localizer_engine.py:54.
 * act: 0 none, 1 GELU(erf), 2 SiLU.  out_f32: element type of out and resid (0 fp16, 1 fp32).
 * block_n: 0 = auto, else 64/128/192/256.  bias/gamma fp32 [N] or NULL; resid may alias out. */
EFFOCR_API int effocr_gemm_f16(const void* d_A, long long lda, const void* d_W, long long ldw, int M, int N, int K,
                    const float* d_bias, const float* d_gamma, const void* d_resid, long long ldr, void* d_out,
                    long long ldo, int act, int out_f32, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EFFOCR_B200_H_ */
