// Internal C++ interface to the tcgen05 GEMM (gemm_sm100.cuh); the model translation units
// (vit.cu, yolo.cu, convnext.cu) and the C-ABI test entry effocr_gemm_f16 all go through it.
#pragma once
#include "host_common.h"

namespace effocr {

struct GemmArgs {
  const __half* A = nullptr;  // [M, K] row-major, leading dimension lda
  long long lda = 0;
  const __half* W = nullptr;  // [N, K] row-major (torch Linear / flattened conv weight), ld = ldw
  long long ldw = 0;
  int M = 0, N = 0, K = 0;
  int act = 0;          // 0 none, 1 GELU(erf), 2 SiLU
  int out_f32 = 0;      // output (and residual) element type: 0 fp16, 1 fp32
  void* out = nullptr;  // [M, N], leading dimension ldo
  long long ldo = 0;
  const float* bias = nullptr;   // [N] fp32 or null
  const float* gamma = nullptr;  // [N] fp32 layer-scale or null
  const void* resid = nullptr;   // same type as out, may alias out; null = no residual
  long long ldr = 0;
  // patch-embed epilogue (ViT): when pos != null the output row is remapped (see EpiPatchEmbed)
  const float* pos = nullptr;
  int patches = 0;
  int block_n = 0;  // 0 = choose
  int prof_tag = PROF_GEMM_OTHER;
  int epilogue = 0;  // 0 = TMA-store epilogue when applicable, 1 = force the direct-store epilogue
};

int gemm_f16(const GemmArgs& a, cudaStream_t stream);

struct LnGemmArgs {
  const float* x = nullptr;  // [M, D] fp32 residual stream, row pitch ldx
  long long ldx = 0;
  const float *gamma = nullptr, *beta = nullptr;  // LayerNorm affine [D]
  float eps = 1e-6f;
  const __half* W = nullptr;  // [N, D] (torch Linear layout), row pitch ldw
  long long ldw = 0;
  const float* bias = nullptr;  // [N] or null
  __half* out = nullptr;        // [M, N] fp16, row pitch ldo
  long long ldo = 0;
  int M = 0, N = 0, D = 0;
  int prof_tag = PROF_GEMM_QKV;
  int reverse = 0;  // walk the row blocks from the last to the first (L2-friendly kernel order, vit.cu)
};
bool ln_gemm_supported(int D, int N);
// out = (LayerNorm(x) * gamma + beta) . W^T + bias in one kernel: the normalised operand never reaches HBM
int ln_gemm_f16(const LnGemmArgs& a, cudaStream_t stream);
int choose_block_n(int N);

}  // namespace effocr
