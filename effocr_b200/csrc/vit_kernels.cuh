// Non-GEMM kernels of the ViT recognizer forward pass (timm VisionTransformer, num_classes=0;
// SURVEY.md App. A.1): LayerNorm (fp32 residual stream -> fp16 GEMM operand), fused multi-head
// attention for the fixed 197-token sequence, NCHW -> patch-major gather, CLS/pos initialisation
// and the final LayerNorm on the CLS row.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ln_math.cuh"

namespace effocr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ LayerNorm rows
// One warp per row; D = 32 * PER_LANE, element e of a lane = (j * 32 + lane) * VEC + v.
// HBM-bound: reads 4 D bytes, writes 2 D (fp16) or 4 D (fp32) bytes per row.
template <int D, typename OutT>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ x, long long ldx,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, OutT* __restrict__ out,
                                                             long long ldo, int rows, float eps) {
  static_assert(D % 128 == 0 || D % 32 == 0, "D must be a multiple of 32");
  constexpr int VEC = (D % 128 == 0) ? 4 : 1;
  constexpr int ITERS = D / (32 * VEC);
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * ldx;
  float v[ITERS * VEC];
#pragma unroll
  for (int j = 0; j < ITERS; ++j) {
    if (VEC == 4) {
      const float4 t = *reinterpret_cast<const float4*>(xr + (j * 32 + lane) * 4);
      v[j * 4] = t.x; v[j * 4 + 1] = t.y; v[j * 4 + 2] = t.z; v[j * 4 + 3] = t.w;
    } else {
      v[j] = xr[j * 32 + lane];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS * VEC; ++i) s += v[i];
  const float mean = __fmul_rn(warp_sum(s), 1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS * VEC; ++i) { const float d = __fsub_rn(v[i], mean); q = __fmaf_rn(d, d, q); }
  const float rstd = ln_rstd(warp_sum(q), 1.0f / D, eps);
  OutT* orow = out + static_cast<long long>(row) * ldo;
#pragma unroll
  for (int j = 0; j < ITERS; ++j) {
    if (VEC == 4) {
      const int c = (j * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float y0 = ln_affine(v[j * 4], mean, rstd, g.x, b.x);
      const float y1 = ln_affine(v[j * 4 + 1], mean, rstd, g.y, b.y);
      const float y2 = ln_affine(v[j * 4 + 2], mean, rstd, g.z, b.z);
      const float y3 = ln_affine(v[j * 4 + 3], mean, rstd, g.w, b.w);
      if constexpr (sizeof(OutT) == 2) {
        uint2 pk;
        *reinterpret_cast<__half2*>(&pk.x) = __floats2half2_rn(y0, y1);
        *reinterpret_cast<__half2*>(&pk.y) = __floats2half2_rn(y2, y3);
        *reinterpret_cast<uint2*>(orow + c) = pk;
      } else {
        *reinterpret_cast<float4*>(orow + c) = make_float4(y0, y1, y2, y3);
      }
    } else {
      const int c = j * 32 + lane;
      const float y = ln_affine(v[j], mean, rstd, __ldg(gamma + c), __ldg(beta + c));
      if constexpr (sizeof(OutT) == 2) orow[c] = __float2half_rn(y);
      else orow[c] = y;
    }
  }
}

// ------------------------------------------------------------------ NCHW f32 image -> patch-major f16
// out[(b * 196 + py * 14 + px), c * 256 + iy * 16 + ix] = img[b, c, py * 16 + iy, px * 16 + ix]
// (the im2col of timm PatchEmbed's Conv2d(3, D, 16, 16): weight [D, 3, 16, 16] flattens to [D, 768]).
__global__ void __launch_bounds__(256) im2patch_kernel(const float* __restrict__ img, __half* __restrict__ out,
                                                       int batch) {
  // one thread per 8 consecutive ix: 2 float4 loads -> one 16-byte store
  const long long total = static_cast<long long>(batch) * 196 * 96;  // 768 / 8 = 96 vectors per patch row
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(i % 96);
    const long long prow = i / 96;
    const int p = static_cast<int>(prow % 196);
    const int b = static_cast<int>(prow / 196);
    const int c = vec / 32, iy = (vec % 32) / 2, ix0 = (vec % 2) * 8;
    const int py = p / 14, px = p % 14;
    const float* src = img + ((static_cast<long long>(b) * 3 + c) * 224 + py * 16 + iy) * 224 + px * 16 + ix0;
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 d = *reinterpret_cast<const float4*>(src + 4);
    // white-centred patch rows (see crop.cu): the encoder's patch-embedding bias carries W . white
    const float wl = c == 0 ? (1.0f - 0.485f) / 0.229f : (c == 1 ? (1.0f - 0.456f) / 0.224f : (1.0f - 0.406f) / 0.225f);
    uint4 pk;
    *reinterpret_cast<__half2*>(&pk.x) = __floats2half2_rn(a.x - wl, a.y - wl);
    *reinterpret_cast<__half2*>(&pk.y) = __floats2half2_rn(a.z - wl, a.w - wl);
    *reinterpret_cast<__half2*>(&pk.z) = __floats2half2_rn(d.x - wl, d.y - wl);
    *reinterpret_cast<__half2*>(&pk.w) = __floats2half2_rn(d.z - wl, d.w - wl);
    *reinterpret_cast<uint4*>(out + prow * 768 + vec * 8) = pk;
  }
}

// Attention for the CLS query only (last encoder block).  timm's ViT with num_classes = 0 pools the class token
// (`x[:, 0]` after the final norm; models/encoders.py:58 -> timm VisionTransformer.forward_head), so nothing the last
// block computes for the 196 patch tokens can reach the embedding: their queries, their attention rows, their
// projection and their MLP are dead.  K and V of ALL tokens are still needed.  One warp per (image, head): lane = key
// for the scores (fp32, exact softmax -- P is not rounded to fp16 here), lane = output dim pair for P.V.
__global__ void __launch_bounds__(128) cls_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int batch,
                                                            int T, int H, float scale) {
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (bh >= batch * H) return;
  const int b = bh / H, h = bh % H;
  const int D = H * 64;
  const __half* base = qkv + static_cast<long long>(b) * T * 3 * D;
  float q[64];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(base + h * 64);  // token 0
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = __ldg(qp + i);
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(hp[t]);
        q[8 * i + 2 * t] = f.x;
        q[8 * i + 2 * t + 1] = f.y;
      }
    }
  }
  float sc[7];
  float mx = -INFINITY;
#pragma unroll
  for (int it = 0; it < 7; ++it) {
    const int t = it * 32 + lane;
    float a = -INFINITY;
    if (t < T) {
      const uint4* kp = reinterpret_cast<const uint4*>(base + static_cast<long long>(t) * 3 * D + D + h * 64);
      a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 u = __ldg(kp + i);
        const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
          const float2 f = __half22float2(hp[tt]);
          a = fmaf(q[8 * i + 2 * tt], f.x, a);
          a = fmaf(q[8 * i + 2 * tt + 1], f.y, a);
        }
      }
      a *= scale;
    }
    sc[it] = a;
    mx = fmaxf(mx, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int it = 0; it < 7; ++it) {
    sc[it] = (it * 32 + lane < T) ? __expf(sc[it] - mx) : 0.f;
    sum += sc[it];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  float o0 = 0.f, o1 = 0.f;  // output dims 2*lane, 2*lane + 1
  const __half* vbase = base + 2 * D + h * 64 + 2 * lane;
#pragma unroll
  for (int it = 0; it < 7; ++it) {
    const int nt = (T - it * 32) < 32 ? (T - it * 32) : 32;
    for (int j = 0; j < nt; ++j) {
      const float p = __shfl_sync(0xffffffffu, sc[it], j);
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(vbase + static_cast<long long>(it * 32 + j) * 3 * D));
      o0 = fmaf(p, f.x, o0);
      o1 = fmaf(p, f.y, o1);
    }
  }
  *reinterpret_cast<__half2*>(out + static_cast<long long>(b) * D + h * 64 + 2 * lane) = __floats2half2_rn(o0 * inv, o1 * inv);
}

// x[b * T + 0, :] = cls + pos[0, :]
__global__ void cls_pos_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                               int batch, int T, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * D) return;
  const int b = i / D, c = i % D;
  x[static_cast<long long>(b) * T * D + c] = cls[c] + pos[c];
}

#ifdef EFFOCR_AB  // first-generation mma.sync attention kernel: A/B only
// ------------------------------------------------------------------ fused attention, 197 tokens, head dim 64
// One CTA per (image, head).  Q, K, V ([T, 64] fp16 each) are staged in shared memory; each
// warp owns 16-row stripes of the score matrix, keeps the whole 16 x 208 stripe in registers
// (fp32), applies the softmax there and feeds P (fp16) straight back into the P.V MMAs.
// Tensor-core path here is mma.sync m16n8k16 (HMMA): 8 % of the encoder FLOPs.
constexpr int kAttnTP = 208;     // 197 tokens padded to a multiple of 16
constexpr int kAttnLd = 72;      // smem row pitch in halves (144 B: conflict-free ldmatrix)
constexpr int kAttnThreads = 128;
constexpr int kAttnSmemBytes = 3 * kAttnTP * kAttnLd * 2;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void hmma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_197x64_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int T, int H, float scale_log2e) {
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __half* Qs = reinterpret_cast<__half*>(attn_smem);
  __half* Ks = Qs + kAttnTP * kAttnLd;
  __half* Vs = Ks + kAttnTP * kAttnLd;
  const int D = H * 64;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // stage Q, K, V: 16-byte cp.async chunks; rows T..207 are zero
  const __half* base = qkv + static_cast<long long>(b) * T * 3 * D + h * 64;
  for (int i = tid; i < 3 * kAttnTP * 8; i += kAttnThreads) {
    const int which = i / (kAttnTP * 8);
    const int r = (i / 8) % kAttnTP;
    const int c = i % 8;
    __half* dst = Qs + which * kAttnTP * kAttnLd + r * kAttnLd + c * 8;
    if (r < T) {
      const __half* src = base + static_cast<long long>(r) * 3 * D + which * D + c * 8;
      const uint32_t d32 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32), "l"(src) : "memory");
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  constexpr int NT = kAttnTP / 8;  // 26 key tiles of 8
  const int g = lane >> 2, t4 = lane & 3;
  for (int stripe = warp; stripe < kAttnTP / 16; stripe += kAttnThreads / 32) {
    if (stripe * 16 >= T) break;
    // Q fragments for the 4 k-steps of the head dimension
    uint32_t qf[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      ldmatrix_x4(qf[k], Qs + (stripe * 16 + (lane & 15)) * kAttnLd + k * 16 + (lane >> 4) * 8);
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      uint32_t kf0[4], kf1[4];
      const __half* kp = Ks + (j * 8 + (lane & 7)) * kAttnLd + (lane >> 3) * 8;
      ldmatrix_x4(kf0, kp);       // dims 0..31  -> (b0, b1) of k-steps 0, 1
      ldmatrix_x4(kf1, kp + 32);  // dims 32..63 -> (b0, b1) of k-steps 2, 3
      hmma_16816(s[j], qf[0], kf0[0], kf0[1]);
      hmma_16816(s[j], qf[1], kf0[2], kf0[3]);
      hmma_16816(s[j], qf[2], kf1[0], kf1[1]);
      hmma_16816(s[j], qf[3], kf1[2], kf1[3]);
    }
    // softmax over keys (rows g and g + 8 of the stripe); scores scaled by 1/sqrt(64) in log2 domain
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int c = j * 8 + t4 * 2;
      if (c >= T) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (c + 1 >= T) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
      m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float o0 = m0 * scale_log2e, o1 = m1 * scale_log2e;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = exp2f(fmaf(s[j][0], scale_log2e, -o0));
      s[j][1] = exp2f(fmaf(s[j][1], scale_log2e, -o0));
      s[j][2] = exp2f(fmaf(s[j][2], scale_log2e, -o1));
      s[j][3] = exp2f(fmaf(s[j][3], scale_log2e, -o1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // O = P . V
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < kAttnTP / 16; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, Vs + (kk * 16 + (lane & 15)) * kAttnLd + n2 * 16 + (lane >> 4) * 8);
        hmma_16816(o[2 * n2], pf, vf[0], vf[1]);
        hmma_16816(o[2 * n2 + 1], pf, vf[2], vf[3]);
      }
    }
    const float r0 = 1.0f / l0, r1 = 1.0f / l1;
    const int row0 = stripe * 16 + g, row1 = row0 + 8;
    __half* ob = out + static_cast<long long>(b) * T * D + h * 64 + t4 * 2;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (row0 < T)
        *reinterpret_cast<__half2*>(ob + static_cast<long long>(row0) * D + n * 8) = __floats2half2_rn(o[n][0] * r0, o[n][1] * r0);
      if (row1 < T)
        *reinterpret_cast<__half2*>(ob + static_cast<long long>(row1) * D + n * 8) = __floats2half2_rn(o[n][2] * r1, o[n][3] * r1);
    }
  }
}

#endif  // EFFOCR_AB

}  // namespace effocr
