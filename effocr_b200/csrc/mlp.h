// Internal C++ interface to the fused MLP block kernel (mlp_sm100.cuh).
#pragma once
#include "host_common.h"

namespace effocr {

struct MlpArgs {
  const __half* h = nullptr;  // [M, D] LayerNorm output, leading dimension ldh
  long long ldh = 0;
  const __half* w1 = nullptr;  // [HID, D] fc1 weight (torch Linear layout)
  const float* b1 = nullptr;   // [HID]
  const __half* w2 = nullptr;  // [D, HID] fc2 weight
  const float* b2 = nullptr;   // [D]
  float* x = nullptr;          // [M, D] fp32 residual stream, updated in place
  long long ldx = 0;
  int M = 0, D = 0, HID = 0;
  long long* dbg = nullptr;  // optional: 512 x int64 device buffer receiving clock64() stamps (profiling aid)
};

bool mlp_fused_supported(int D, int HID);
// x += GELU(h . w1^T + b1) . w2^T + b2
int mlp_fused_f16(const MlpArgs& a, cudaStream_t stream);

struct ProjLnArgs {
  const __half* att = nullptr;  // [M, D] attention output, leading dimension lda
  long long lda = 0;
  const __half* w = nullptr;    // [D, D] projection weight (torch Linear layout)
  const float* bias = nullptr;  // [D]
  float* x = nullptr;           // [M, D] fp32 residual stream, updated in place
  long long ldx = 0;
  const float *gamma = nullptr, *beta = nullptr;  // LayerNorm affine [D]
  float eps = 1e-6f;
  __half* h = nullptr;          // [M, D] LayerNorm(x) out
  long long ldh = 0;
  int M = 0, D = 0;
};

bool proj_ln_supported(int D);
// x += att . w^T + bias;  h = LayerNorm(x) * gamma + beta   (one kernel, full-row tiles)
int proj_ln_f16(const ProjLnArgs& a, cudaStream_t stream);

struct BlockTailArgs {
  const __half* att = nullptr;  // [M, D] attention output, leading dimension lda
  long long lda = 0;
  const __half* wp = nullptr;   // [D, D] projection weight
  const float* bp = nullptr;    // [D]
  const float *gamma = nullptr, *beta = nullptr;  // norm2 affine [D]
  float eps = 1e-6f;
  const __half* w1 = nullptr;   // [HID, D]
  const float* b1 = nullptr;
  const __half* w2 = nullptr;   // [D, HID]
  const float* b2 = nullptr;
  float* x = nullptr;           // [M, D] fp32 residual stream, updated in place
  long long ldx = 0;
  int M = 0, D = 0, HID = 0;
  int reverse = 0;  // walk the tiles from the last to the first (L2-friendly kernel order, vit.cu)
};

bool block_tail_supported(int D, int HID);
// x += att . wp^T + bp;  x += GELU(LayerNorm(x) . w1^T + b1) . w2^T + b2   (one kernel: blocktail_sm100.cuh)
int block_tail_f16(const BlockTailArgs& a, cudaStream_t stream);

}  // namespace effocr
