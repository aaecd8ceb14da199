// ConvNeXt-Tiny recognizer encoder (timm convnext_tiny, num_classes=0) as a C-ABI handle.
// Replaces `timm.create_model("convnext_tiny", num_classes=0)(x)` reached from
// /root/reference/models/encoders.py:58,62-64 (BASELINE config 4).
//
// Layout: NHWC.  The residual stream x is fp32 [B*H*W, C]; every pointwise layer is the tcgen05 GEMM:
//   stem / downsample   patchify gather (4x4 s4 / 2x2 s2) -> GEMM (+bias) -> fp32
//   block               dwconv7x7 + bias + LayerNorm fused (fp32 in, fp16 out)  ->  GEMM fc1 + bias + GELU (fp16)
//                       ->  GEMM fc2 + bias, * gamma, TMA reduce-add into x (fp32, in place)
//   head                global average pool + LayerNorm -> fp32 [B, 768]
#include <stdlib.h>

#include <vector>

#include "../../include/effocr_b200.h"
#include "gemm.h"
#include "mlp.h"

namespace effocr {

int layernorm_f16(const float* x, long long ldx, const float* g, const float* b, __half* out, long long ldo, int rows,
                  int D, float eps, cudaStream_t s, int tag);
int layernorm_f32(const float* x, long long ldx, const float* g, const float* b, float* out, long long ldo, int rows,
                  int D, float eps, cudaStream_t s, int tag);

__device__ __forceinline__ float cnx_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NCHW f32 image -> 4x4 patch rows: out[(b*56 + py)*56 + px, c*16 + iy*4 + ix] (torch conv weight order), fp16
__global__ void __launch_bounds__(256) cnx_im2patch4_kernel(const float* __restrict__ img, __half* __restrict__ out,
                                                            int batch) {
  const long long total = static_cast<long long>(batch) * 3136 * 12;  // 12 = 3 channels x 4 rows of 4 px
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int piece = static_cast<int>(i % 12);
    const long long prow = i / 12;
    const int p = static_cast<int>(prow % 3136);
    const int b = static_cast<int>(prow / 3136);
    const int c = piece / 4, iy = piece % 4;
    const int py = p / 56, px = p % 56;
    const float4 a = *reinterpret_cast<const float4*>(img + ((static_cast<long long>(b) * 3 + c) * 224 + py * 4 + iy) * 224 + px * 4);
    // white-centred patch rows (see crop.cu): the stem bias carries W . white
    const float wl = c == 0 ? (1.0f - 0.485f) / 0.229f : (c == 1 ? (1.0f - 0.456f) / 0.224f : (1.0f - 0.406f) / 0.225f);
    uint2 pk;
    *reinterpret_cast<__half2*>(&pk.x) = __floats2half2_rn(a.x - wl, a.y - wl);
    *reinterpret_cast<__half2*>(&pk.y) = __floats2half2_rn(a.z - wl, a.w - wl);
    *reinterpret_cast<uint2*>(out + prow * 48 + piece * 4) = pk;
  }
}

// LayerNorm over C of fp32 NHWC rows, written as fp16 into the 2x2/s2 patchify layout of the next stage:
// out[(b, y/2, x/2), ((y&1)*2 + (x&1))*C + c]   (downsample = LayerNorm2d + Conv2d(k2, s2))
template <int C>
__global__ void __launch_bounds__(256) cnx_ln_patch2_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, __half* __restrict__ out,
                                                            int batch, int H, int W, float eps) {
  constexpr int PER = C / 32;
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(batch) * H * W) return;
  const int xx = static_cast<int>(row % W), yy = static_cast<int>((row / W) % H), b = static_cast<int>(row / (static_cast<long long>(W) * H));
  const float* xr = x + row * C;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PER; ++j) { v[j] = xr[j * 32 + lane]; s += v[j]; }
  const float mean = cnx_warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < PER; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(cnx_warp_sum(q) * (1.0f / C) + eps);
  const long long orow = (static_cast<long long>(b) * (H / 2) + yy / 2) * (W / 2) + xx / 2;
  __half* o = out + orow * (4 * C) + ((yy & 1) * 2 + (xx & 1)) * C;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int c = j * 32 + lane;
    o[c] = __float2half_rn((v[j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c));
  }
}

// Fused depthwise 7x7 (pad 3) + bias + LayerNorm over channels: x fp32 NHWC -> h fp16 [B*H*W, C].
// One warp computes XT = 4 adjacent output pixels of one row for ALL channels (lane l owns channels l, l+32, ...):
// every loaded input value feeds up to four outputs (2.8x fewer L1 reads than one pixel per warp), and the
// LayerNorm statistics are plain warp reductions.  w is [49][C] (tap-major) so weight loads are coalesced.
template <int C>
__global__ void __launch_bounds__(128) cnx_dwconv_ln_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, __half* __restrict__ out,
                                                            int batch, int H, int W, float eps) {
  constexpr int PER = C / 32;
  constexpr int XT = 4;
  const int lane = threadIdx.x & 31;
  const int tiles_x = (W + XT - 1) / XT;
  const long long wid = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= static_cast<long long>(batch) * H * tiles_x) return;
  const int tx = static_cast<int>(wid % tiles_x);
  const int y = static_cast<int>((wid / tiles_x) % H);
  const int b = static_cast<int>(wid / (static_cast<long long>(tiles_x) * H));
  const int x0 = tx * XT;
  float acc[XT][PER];
#pragma unroll
  for (int t = 0; t < XT; ++t)
#pragma unroll
    for (int j = 0; j < PER; ++j) acc[t][j] = __ldg(bias + j * 32 + lane);
  const float* xb = x + static_cast<long long>(b) * H * W * C;
  for (int ky = 0; ky < 7; ++ky) {
    const int iy = y + ky - 3;
    if (iy < 0 || iy >= H) continue;
    float wk[7][PER];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx)
#pragma unroll
      for (int j = 0; j < PER; ++j) wk[kx][j] = __ldg(w + (ky * 7 + kx) * C + j * 32 + lane);
#pragma unroll
    for (int dx = 0; dx < XT + 6; ++dx) {
      const int ix = x0 + dx - 3;
      if (ix < 0 || ix >= W) continue;
      const float* p = xb + (static_cast<long long>(iy) * W + ix) * C + lane;
      float in[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) in[j] = p[j * 32];
#pragma unroll
      for (int t = 0; t < XT; ++t) {
        const int kx = dx - t;  // input column ix contributes to output x0 + t through tap kx = ix - (x0 + t) + 3
        if (kx >= 0 && kx < 7) {
#pragma unroll
          for (int j = 0; j < PER; ++j) acc[t][j] = fmaf(in[j], wk[kx][j], acc[t][j]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < XT; ++t) {
    const int ox = x0 + t;
    if (ox >= W) break;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) s += acc[t][j];
    const float mean = cnx_warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) { const float d = acc[t][j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(cnx_warp_sum(q) * (1.0f / C) + eps);
    __half* o = out + ((static_cast<long long>(b) * H + y) * W + ox) * C;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int c = j * 32 + lane;
      o[c] = __float2half_rn((acc[t][j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c));
    }
  }
}

// head: global average pool over the 7x7 map + LayerNorm(768) -> fp32 [B, 768]; one CTA (256 threads) per image
__global__ void __launch_bounds__(256) cnx_head_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ out, int HW,
                                                       float eps) {
  constexpr int C = 768;
  __shared__ float red[16];
  const int b = blockIdx.x, tid = threadIdx.x;
  float v[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = j * 256 + tid;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += x[(static_cast<long long>(b) * HW + p) * C + c];
    v[j] = s / static_cast<float>(HW);
  }
  float s = v[0] + v[1] + v[2];
  s = cnx_warp_sum(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float mean = tot * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
  q = cnx_warp_sum(q);
  if ((tid & 31) == 0) red[8 + (tid >> 5)] = q;
  __syncthreads();
  float qt = 0.f;
  for (int i = 0; i < 8; ++i) qt += red[8 + i];
  const float rstd = rsqrtf(qt * (1.0f / C) + eps);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = j * 256 + tid;
    out[static_cast<long long>(b) * C + c] = (v[j] - mean) * rstd * gamma[c] + beta[c];
  }
}

struct CnxBlock {
  float *dw_w, *dw_b, *ln_w, *ln_b, *b1, *b2, *gamma;
  __half *w1, *w2;
  // stages of width 192 / 384 run the pointwise pair through the fused MLP kernel (mlp_sm100.cuh: fc1 + GELU + fc2 +
  // residual in one kernel, the [M, 4C] hidden activations never reach HBM); the layer scale is folded into fc2:
  // x += gamma * (P . W2^T + b2) == P . (diag(gamma) W2)^T + gamma * b2
  __half* w2g = nullptr;
  float* b2g = nullptr;
};
struct CnxStage {
  float *ds_ln_w = nullptr, *ds_ln_b = nullptr, *ds_b = nullptr;
  __half* ds_w = nullptr;
  std::vector<CnxBlock> blocks;
};

struct CnxHandle {
  int max_batch = 0;
  __half* stem_w = nullptr;
  float *stem_b = nullptr, *stem_ln_w = nullptr, *stem_ln_b = nullptr, *head_ln_w = nullptr, *head_ln_b = nullptr;
  CnxStage stages[4];
  // workspace
  __half *patches = nullptr, *h16 = nullptr, *mid = nullptr;
  float *xa = nullptr, *xb = nullptr;
  std::vector<void*> allocs;
  ~CnxHandle() {
    for (void* p : allocs) cudaFree(p);
  }
  template <typename T>
  int alloc(T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return EFFOCR_OK;
  }
  int up32(float** d, const float* s, size_t n) {
    EFFOCR_TRY(alloc(d, n));
    EFFOCR_CUDA(cudaMemcpy(*d, s, n * 4, cudaMemcpyHostToDevice));
    return EFFOCR_OK;
  }
  int up16(__half** d, const std::vector<__half>& v) {
    EFFOCR_TRY(alloc(d, v.size()));
    EFFOCR_CUDA(cudaMemcpy(*d, v.data(), v.size() * 2, cudaMemcpyHostToDevice));
    return EFFOCR_OK;
  }
};

static const int kCnxDepths[4] = {3, 3, 9, 3};
static const int kCnxDims[4] = {96, 192, 384, 768};

static inline int cnx_grid(long long total, int per_block) {
  long long g = (total + per_block - 1) / per_block;
  return static_cast<int>(g > 0 ? g : 1);
}

template <int C>
static void launch_dwconv(const float* x, const CnxBlock& k, __half* out, int B, int H, int W, cudaStream_t s) {
  const long long warps = static_cast<long long>(B) * H * ((W + 3) / 4);
  KernelScope ks(PROF_DWCONV, s);
  cnx_dwconv_ln_kernel<C><<<cnx_grid(warps, 4), 128, 0, s>>>(x, k.dw_w, k.dw_b, k.ln_w, k.ln_b, out, B, H, W, 1e-6f);
}
template <int C>
static void launch_ln_patch2(const float* x, const float* g, const float* b, __half* out, int B, int H, int W, cudaStream_t s) {
  const long long rows = static_cast<long long>(B) * H * W;
  KernelScope ks(PROF_LAYERNORM, s);
  cnx_ln_patch2_kernel<C><<<cnx_grid(rows, 8), 256, 0, s>>>(x, g, b, out, B, H, W, 1e-6f);
}

static int cnx_forward_chunk(CnxHandle* h, int B, float* emb, cudaStream_t s) {
  GemmArgs g;
  // stem: [B*3136, 48] x [96, 48]^T + bias -> fp32, then LayerNorm2d in place
  long long M = static_cast<long long>(B) * 3136;
  g = GemmArgs();
  g.A = h->patches; g.lda = 48; g.W = h->stem_w; g.ldw = 48; g.M = static_cast<int>(M); g.N = 96; g.K = 48;
  g.out = h->xa; g.ldo = 96; g.out_f32 = 1; g.bias = h->stem_b;
  EFFOCR_TRY(gemm_f16(g, s));
  EFFOCR_TRY(layernorm_f32(h->xa, 96, h->stem_ln_w, h->stem_ln_b, h->xa, 96, static_cast<int>(M), 96, 1e-6f, s, PROF_LAYERNORM));
  float* x = h->xa;
  float* xalt = h->xb;
  int H = 56, W = 56;
  static const bool fused_mlp = [] {
    const char* e = getenv("EFFOCR_MLP_FUSED");  // "0" = separate fc1 / fc2 GEMMs everywhere (A/B runs)
    return !(e && e[0] == '0');
  }();
  for (int st = 0; st < 4; ++st) {
    const int C = kCnxDims[st];
    if (st > 0) {
      const int Cp = kCnxDims[st - 1];
      const CnxStage& S = h->stages[st];
      switch (Cp) {
        case 96: launch_ln_patch2<96>(x, S.ds_ln_w, S.ds_ln_b, h->h16, B, H, W, s); break;
        case 192: launch_ln_patch2<192>(x, S.ds_ln_w, S.ds_ln_b, h->h16, B, H, W, s); break;
        default: launch_ln_patch2<384>(x, S.ds_ln_w, S.ds_ln_b, h->h16, B, H, W, s); break;
      }
      EFFOCR_CUDA(cudaGetLastError());
      H /= 2; W /= 2;
      M = static_cast<long long>(B) * H * W;
      g = GemmArgs();
      g.A = h->h16; g.lda = 4 * Cp; g.W = S.ds_w; g.ldw = 4 * Cp; g.M = static_cast<int>(M); g.N = C; g.K = 4 * Cp;
      g.out = xalt; g.ldo = C; g.out_f32 = 1; g.bias = S.ds_b;
      EFFOCR_TRY(gemm_f16(g, s));
      float* t = x; x = xalt; xalt = t;
    }
    for (const CnxBlock& k : h->stages[st].blocks) {
      switch (C) {
        case 96: launch_dwconv<96>(x, k, h->h16, B, H, W, s); break;
        case 192: launch_dwconv<192>(x, k, h->h16, B, H, W, s); break;
        case 384: launch_dwconv<384>(x, k, h->h16, B, H, W, s); break;
        default: launch_dwconv<768>(x, k, h->h16, B, H, W, s); break;
      }
      EFFOCR_CUDA(cudaGetLastError());
      if (k.w2g && fused_mlp) {
        MlpArgs m;
        m.h = h->h16; m.ldh = C; m.w1 = k.w1; m.b1 = k.b1; m.w2 = k.w2g; m.b2 = k.b2g;
        m.x = x; m.ldx = C; m.M = static_cast<int>(M); m.D = C; m.HID = 4 * C;
        EFFOCR_TRY(mlp_fused_f16(m, s));
        continue;
      }
      g = GemmArgs();
      g.A = h->h16; g.lda = C; g.W = k.w1; g.ldw = C; g.M = static_cast<int>(M); g.N = 4 * C; g.K = C;
      g.out = h->mid; g.ldo = 4 * C; g.bias = k.b1; g.act = 1; g.prof_tag = PROF_GEMM_FC1;
      EFFOCR_TRY(gemm_f16(g, s));
      g = GemmArgs();
      g.A = h->mid; g.lda = 4 * C; g.W = k.w2; g.ldw = 4 * C; g.M = static_cast<int>(M); g.N = C; g.K = 4 * C;
      g.out = x; g.ldo = C; g.out_f32 = 1; g.bias = k.b2; g.gamma = k.gamma; g.resid = x; g.ldr = C; g.prof_tag = PROF_GEMM_FC2;
      EFFOCR_TRY(gemm_f16(g, s));
    }
  }
  {
    KernelScope ks(PROF_FINAL_LN, s);
    cnx_head_kernel<<<B, 256, 0, s>>>(x, h->head_ln_w, h->head_ln_b, emb, H * W, 1e-6f);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

}  // namespace effocr

using namespace effocr;

extern "C" int effocr_convnext_create(int max_batch, const float* const* hw, int n_weights, effocr_convnext_t* out) {
  if (!out) return fail(EFFOCR_ERR_INVALID, "convnext_create: null out");
  *out = nullptr;
  EFFOCR_TRY(require_sm100());
  if (n_weights != 4 + 3 * 4 + 18 * 9 + 2) return fail(EFFOCR_ERR_INVALID, "convnext_create: expected 180 weight tensors");
  if (max_batch <= 0) return fail(EFFOCR_ERR_INVALID, "convnext_create: bad max_batch");
  CnxHandle* h = new CnxHandle();
  h->max_batch = max_batch;
  int st = EFFOCR_OK, i = 0;
  auto to16 = [](const float* src, size_t n) {
    std::vector<__half> v(n);
    for (size_t k = 0; k < n; ++k) v[k] = __float2half_rn(src[k]);
    return v;
  };
  do {
    if ((st = h->up16(&h->stem_w, to16(hw[i], 96 * 48)))) break; ++i;
    {
      // the 4x4-patch rows are white-centred (crop.cu): bias' = bias + sum_k fp16(W[n, k]) * white(channel of k), k = c*16 + ...
      const float white[3] = {(1.0f - 0.485f) / 0.229f, (1.0f - 0.456f) / 0.224f, (1.0f - 0.406f) / 0.225f};
      std::vector<float> b(96);
      for (int n = 0; n < 96; ++n) {
        double acc = hw[i][n];
        for (int k = 0; k < 48; ++k)
          acc += static_cast<double>(__half2float(__float2half_rn(hw[i - 1][n * 48 + k]))) * static_cast<double>(white[k / 16]);
        b[n] = static_cast<float>(acc);
      }
      if ((st = h->up32(&h->stem_b, b.data(), 96))) break; ++i;
    }
    if ((st = h->up32(&h->stem_ln_w, hw[i], 96))) break; ++i;
    if ((st = h->up32(&h->stem_ln_b, hw[i], 96))) break; ++i;
    for (int sidx = 0; sidx < 4 && !st; ++sidx) {
      const int C = kCnxDims[sidx];
      CnxStage& S = h->stages[sidx];
      if (sidx > 0) {
        const int Cp = kCnxDims[sidx - 1];
        if ((st = h->up32(&S.ds_ln_w, hw[i], Cp))) break; ++i;
        if ((st = h->up32(&S.ds_ln_b, hw[i], Cp))) break; ++i;
        // conv weight [C, Cp, 2, 2] -> [C, (ky, kx, cp)] to match the patch2 gather order
        std::vector<__half> w(static_cast<size_t>(C) * 4 * Cp);
        for (int o = 0; o < C; ++o)
          for (int c = 0; c < Cp; ++c)
            for (int t = 0; t < 4; ++t)
              w[(static_cast<size_t>(o) * 4 + t) * Cp + c] = __float2half_rn(hw[i][(static_cast<size_t>(o) * Cp + c) * 4 + t]);
        if ((st = h->up16(&S.ds_w, w))) break; ++i;
        if ((st = h->up32(&S.ds_b, hw[i], C))) break; ++i;
      }
      for (int j = 0; j < kCnxDepths[sidx] && !st; ++j) {
        CnxBlock k;
        std::vector<float> dw(static_cast<size_t>(49) * C);  // [C,1,7,7] -> [49][C]
        for (int c = 0; c < C; ++c)
          for (int t = 0; t < 49; ++t) dw[static_cast<size_t>(t) * C + c] = hw[i][static_cast<size_t>(c) * 49 + t];
        if ((st = h->up32(&k.dw_w, dw.data(), dw.size()))) break; ++i;
        if ((st = h->up32(&k.dw_b, hw[i], C))) break; ++i;
        if ((st = h->up32(&k.ln_w, hw[i], C))) break; ++i;
        if ((st = h->up32(&k.ln_b, hw[i], C))) break; ++i;
        if ((st = h->up16(&k.w1, to16(hw[i], static_cast<size_t>(4) * C * C)))) break; ++i;
        if ((st = h->up32(&k.b1, hw[i], 4 * C))) break; ++i;
        if ((st = h->up16(&k.w2, to16(hw[i], static_cast<size_t>(4) * C * C)))) break; ++i;
        if ((st = h->up32(&k.b2, hw[i], C))) break; ++i;
        if ((st = h->up32(&k.gamma, hw[i], C))) break; ++i;
        if (mlp_fused_supported(C, 4 * C)) {
          const float *w2 = hw[i - 3], *b2 = hw[i - 2], *gm = hw[i - 1];
          std::vector<__half> wg(static_cast<size_t>(4) * C * C);
          std::vector<float> bg(C);
          for (int n = 0; n < C; ++n) {
            bg[n] = gm[n] * b2[n];
            for (int c = 0; c < 4 * C; ++c) wg[static_cast<size_t>(n) * 4 * C + c] = __float2half_rn(gm[n] * w2[static_cast<size_t>(n) * 4 * C + c]);
          }
          if ((st = h->up16(&k.w2g, wg))) break;
          if ((st = h->up32(&k.b2g, bg.data(), C))) break;
        }
        S.blocks.push_back(k);
      }
    }
    if (st) break;
    if ((st = h->up32(&h->head_ln_w, hw[i], 768))) break; ++i;
    if ((st = h->up32(&h->head_ln_b, hw[i], 768))) break; ++i;
    const size_t M0 = static_cast<size_t>(max_batch) * 3136;
    if ((st = h->alloc(&h->patches, M0 * 48))) break;
    if ((st = h->alloc(&h->xa, M0 * 96))) break;
    if ((st = h->alloc(&h->xb, M0 / 4 * 192))) break;
    if ((st = h->alloc(&h->h16, M0 * 96))) break;
    if ((st = h->alloc(&h->mid, M0 * 384))) break;
  } while (0);
  if (st) { delete h; return st; }
  *out = reinterpret_cast<effocr_convnext_t>(h);
  return EFFOCR_OK;
}

extern "C" void effocr_convnext_destroy(effocr_convnext_t h) { delete reinterpret_cast<CnxHandle*>(h); }
extern "C" void* effocr_convnext_patch_buffer(effocr_convnext_t h) { return h ? reinterpret_cast<CnxHandle*>(h)->patches : nullptr; }

extern "C" int effocr_convnext_forward(effocr_convnext_t handle, const void* d_input, int input_kind, int batch,
                                       float* d_emb, void* stream) {
  CnxHandle* h = reinterpret_cast<CnxHandle*>(handle);
  if (!h || !d_emb || batch < 0) return fail(EFFOCR_ERR_INVALID, "convnext_forward: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (input_kind == EFFOCR_INPUT_PATCH_BUFFER) {
    if (batch > h->max_batch) return fail(EFFOCR_ERR_INVALID, "convnext_forward: batch exceeds max_batch");
    return batch == 0 ? EFFOCR_OK : cnx_forward_chunk(h, batch, d_emb, s);
  }
  if (input_kind != EFFOCR_INPUT_NCHW_F32 || !d_input) return fail(EFFOCR_ERR_INVALID, "convnext_forward: unsupported input");
  for (int b0 = 0; b0 < batch; b0 += h->max_batch) {
    const int B = batch - b0 < h->max_batch ? batch - b0 : h->max_batch;
    const float* img = reinterpret_cast<const float*>(d_input) + static_cast<size_t>(b0) * 3 * 224 * 224;
    {
      KernelScope ks(PROF_IM2PATCH, s);
      const long long total = static_cast<long long>(B) * 3136 * 12;
      cnx_im2patch4_kernel<<<static_cast<int>((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32), 256, 0, s>>>(img, h->patches, B);
    }
    EFFOCR_CUDA(cudaGetLastError());
    EFFOCR_TRY(cnx_forward_chunk(h, B, d_emb + static_cast<size_t>(b0) * 768, s));
  }
  return EFFOCR_OK;
}
