// LayerNorm arithmetic shared by the stand-alone kernel (vit_kernels.cuh) and the LayerNorm warps of the fused norm1 + QKV
// kernel (lnqkv_sm100.cuh): every rounding step is spelled out so that both produce the same bits whatever the compiler
// would otherwise contract into FMAs.
#pragma once

namespace effocr {

__device__ __forceinline__ float ln_rstd(float sum_sq_dev, float inv_d, float eps) { return rsqrtf(__fmaf_rn(sum_sq_dev, inv_d, eps)); }
// (v - mean) * rstd * gamma + beta
__device__ __forceinline__ float ln_affine(float v, float mean, float rstd, float g, float b) {
  return __fmaf_rn(__fmul_rn(__fsub_rn(v, mean), rstd), g, b);
}

}  // namespace effocr
