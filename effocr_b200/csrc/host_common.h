// Host-side plumbing shared by the C-ABI translation units: error reporting, TMA tensor-map
// encoding (driver entry point fetched through the runtime, so the library links only cudart
// and still dlopen()s on a machine without a GPU), device properties.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/effocr_b200.h"

namespace effocr {


void set_last_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define EFFOCR_CUDA(expr)                                          \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return ::effocr::cuda_fail(_e, #expr);  \
  } while (0)

#define EFFOCR_TRY(expr)          \
  do {                            \
    int _s = (expr);              \
    if (_s != 0) return _s;       \
  } while (0)

// ---- launch accounting + optional per-kernel CUDA-event timing (bench.py's roofline block) ----
enum ProfTag : int {
  PROF_CROP = 0, PROF_IM2PATCH, PROF_GEMM_PATCH, PROF_LAYERNORM, PROF_GEMM_QKV, PROF_ATTENTION, PROF_GEMM_PROJ,
  PROF_GEMM_FC1, PROF_GEMM_FC2, PROF_FINAL_LN, PROF_L2NORM, PROF_KNN_SPLIT, PROF_KNN_GEMM, PROF_KNN_MERGE, PROF_MISC,
  PROF_GEMM_OTHER, PROF_CONV_IM2COL, PROF_YOLO_MISC, PROF_NMS, PROF_DWCONV, PROF_MLP_FUSED, PROF_PROJ_LN, PROF_BLOCK_TAIL, PROF_NUM_TAGS
};
// Construct right before a kernel launch, destroy right after: counts the launch and, when
// profiling is enabled, brackets it with events on `stream`.
class KernelScope {
 public:
  KernelScope(int tag, cudaStream_t stream);
  ~KernelScope();
 private:
  int slot_;
  cudaStream_t stream_;
};

// Number of SMs persistent kernels may occupy: the device's SM count (148 on B200) unless a limit is in force.
int sm_count();
int sm_count_physical();
// Per-host-thread cap on sm_count() (0 = none): lets two independent launch sequences share the device side by side.
void set_sm_limit(int n);
// Fails loudly unless the current device is compute capability 10.x.
int require_sm100();

// 2-D row-major fp16 tensor [rows, cols] with leading dimension ld (elements), tiled in
// boxes of box_rows x 64 columns, 128-byte swizzle, out-of-bounds reads return zero.
int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows);

// General 2-D row-major tensor map: element size esz (2 = fp16, 4 = fp32), box = box_rows x box_cols,
// swizzle_bytes in {0, 32, 64, 128} (must equal box_cols * esz when non-zero).
int make_tmap_2d(CUtensorMap* out, const void* base, int esz, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, int swizzle_bytes);

// 4-D tiled tensor map (innermost dimension first), e.g. NHWC activations as {C, W, H, N}; elem_strides are the TMA
// traversal strides (a box of box[i] positions loads ceil(box[i] / elem_strides[i]) elements along dimension i).
int make_tmap_4d(CUtensorMap* out, const void* base, int esz, const uint64_t dims[4], const uint64_t strides_bytes[3],
                 const uint32_t box[4], const uint32_t elem_strides[4], int swizzle_bytes);

}  // namespace effocr
