// YOLOv5s character/word localizer (ultralytics v6/7 yolov5s.yaml) as a C-ABI handle.
// Replaces the onnxruntime session behind EffLocalizer.run
// (/root/reference/onnx_engines/localizer_engine.py:25-29,49-55): letterboxed f32 [B,3,H,W] image ->
// decoded predictions f32 [B, 3*(H/8*W/8 + H/16*W/16 + H/32*W/32), 5+nc].
//
// Layout: activations are NHWC fp16, so every 1x1 convolution IS the tcgen05 GEMM
// [pixels, Cin] x [Cout, Cin]^T with the folded-BatchNorm bias and SiLU in its epilogue, and channel
// concatenation is free: producers write their GEMM output straight into a channel slice of the
// consumer's buffer (output leading dimension = total channels).  k > 1 convolutions gather an
// im2col matrix (fp16) and run the same GEMM.  BatchNorm (eps 1e-3) is folded into the fp16
// weights / fp32 bias on the host at create time.
#include <math.h>

#include <stdlib.h>

#include <vector>

#include "../../include/effocr_b200.h"
#include "gemm.h"
#include "conv3_sm100.cuh"
#include "yolo_split_sm100.cuh"

namespace effocr {

#ifdef EFFOCR_AB
// ------------------------------------------------------------------ gather kernels (HBM-bound)
// layer 0: f32 NCHW image -> im2col rows of Conv(3, 32, k6, s2, p2); col = c*36 + ky*6 + kx (torch
// weight order), zero-padded from 108 to 112 columns.
__global__ void __launch_bounds__(256) yolo_im2col0_kernel(const float* __restrict__ img, __half* __restrict__ col,
                                                           int B, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * 14;  // 14 vectors of 8 columns
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(i % 14);
    const long long row = i / 14;
    const int ox = static_cast<int>(row % Wo);
    const int oy = static_cast<int>((row / Wo) % Ho);
    const int b = static_cast<int>(row / (static_cast<long long>(Wo) * Ho));
    __half v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cidx = vec * 8 + j;
      float x = 0.f;
      if (cidx < 108) {
        const int c = cidx / 36, ky = (cidx % 36) / 6, kx = cidx % 6;
        const int iy = oy * 2 - 2 + ky, ix = ox * 2 - 2 + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) x = img[((static_cast<long long>(b) * 3 + c) * H + iy) * W + ix];
      }
      v[j] = __float2half_rn(x);
    }
    *reinterpret_cast<uint4*>(col + row * 112 + vec * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

#endif  // EFFOCR_AB

// ------------------------------------------------------------------ layer 0 as an implicit GEMM (no im2col buffer)
// Conv(3, 32, k6, s2, p2) + folded BN + SiLU straight from the f32 NCHW image to the fp16 NHWC activation.  With K = 108
// the explicit im2col matrix is 1.47 GB per 64 lines (written, then read back by the GEMM): 2.3 ms of the 8.5 ms forward.
// Here a CTA stages the (2*8+4) x (2*32+4) x 3 input patch of an 8 x 32 output tile in shared memory as fp16 and each
// warp computes one output row (32 pixels x 32 channels) with mma.sync.m16n8k16: the A fragment of k = c*36 + ky*6 + kx
// is ONE 32-bit shared load (two horizontally adjacent pixels: kx is even), the 32 x 112 weight matrix lives in
// registers as B fragments (56 per thread).  The legacy mma.sync path is the right tool for K = 108, N = 32: a tcgen05
// tile would spend its time waiting for the gather, not in the tensor core.
constexpr int kStemTH = 8, kStemTW = 32;                 // output tile
constexpr int kStemPH = 2 * kStemTH + 4, kStemPW = 2 * kStemTW + 4;  // input patch 20 x 68
constexpr int kStemPlane = kStemPH * kStemPW;           // 1360 halves per channel

// same SiLU as the GEMM epilogue (gemm_sm100.cuh)
__device__ __forceinline__ float stem_silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ void mma_m16n8k16_f16f32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) yolo_stem_kernel(const float* __restrict__ img, const __half* __restrict__ w /*[32,112]*/,
                                                        const float* __restrict__ bias, __half* __restrict__ out, int ld_out,
                                                        int B, int H, int W) {
  __shared__ __align__(16) __half patch[3 * kStemPlane];
  __shared__ __align__(16) __half stage[8][32 * 32];  // per warp: 32 pixels x 32 channels
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_x = (Wo + kStemTW - 1) / kStemTW, tiles_y = (Ho + kStemTH - 1) / kStemTH;
  const int num_tiles = B * tiles_y * tiles_x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  // B fragments: for n-tile nt and k-step ks, b0 = W[n = nt*8 + g][k = ks*16 + 2t, +1], b1 = same with k + 8
  uint32_t bf[4][7][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) {
      const __half* wr = w + (nt * 8 + g) * 112 + ks * 16 + 2 * t;
      bf[nt][ks][0] = *reinterpret_cast<const uint32_t*>(wr);
      bf[nt][ks][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
    }
  // patch offset (in halves) of k = ks*16 + 2t (+8): c*plane + ky*PW + kx; padded columns (k >= 108, zero weights) -> 0
  int koff[7][2];
#pragma unroll
  for (int ks = 0; ks < 7; ++ks)
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int k = ks * 16 + 2 * t + 8 * h2;
      koff[ks][h2] = (k < 108) ? (k / 36) * kStemPlane + ((k % 36) / 6) * kStemPW + (k % 6) : 0;
    }
  float bv[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    bv[nt][0] = bias[nt * 8 + 2 * t];
    bv[nt][1] = bias[nt * 8 + 2 * t + 1];
  }

  // software pipeline: the NEXT tile's input patch is fetched into registers (16 values per thread) while the current
  // tile is being multiplied, so the global-load latency is off the load -> sync -> compute -> sync chain
  constexpr int kPerThread = (3 * kStemPlane + 255) / 256;  // 16
  float pre[kPerThread];
  auto fetch = [&](int tile) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int iy0 = 2 * ty * kStemTH - 2, ix0 = 2 * tx * kStemTW - 2;
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      const int i = threadIdx.x + u * 256;
      float v = 0.f;
      if (i < 3 * kStemPlane) {
        const int c = i / kStemPlane, r = (i % kStemPlane) / kStemPW, x = i % kStemPW;
        const int iy = iy0 + r, ix = ix0 + x;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + ((static_cast<long long>(b) * 3 + c) * H + iy) * W + ix);
      }
      pre[u] = v;
    }
  };
  if (blockIdx.x < num_tiles) fetch(blockIdx.x);
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int oy0 = ty * kStemTH, ox0 = tx * kStemTW;
    __syncthreads();  // previous tile's readers are done with the patch
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      const int i = threadIdx.x + u * 256;
      if (i < 3 * kStemPlane) patch[i] = __float2half_rn(pre[u]);
    }
    __syncthreads();
    if (tile + gridDim.x < num_tiles) fetch(tile + gridDim.x);
    const int oy = oy0 + warp;  // this warp's output row
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      // fragment rows g and g + 8 of this m-tile = output pixels ox0 + mt*16 + g (+8)
      const __half* p0 = patch + (2 * warp) * kStemPW + 2 * (mt * 16 + g);
      const __half* p1 = p0 + 16;  // 8 pixels further = 16 input columns
#pragma unroll
      for (int ks = 0; ks < 7; ++ks) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(p0 + koff[ks][0]);
        a[1] = *reinterpret_cast<const uint32_t*>(p1 + koff[ks][0]);
        a[2] = *reinterpret_cast<const uint32_t*>(p0 + koff[ks][1]);
        a[3] = *reinterpret_cast<const uint32_t*>(p1 + koff[ks][1]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_m16n8k16_f16f32(acc[mt][nt], a, bf[nt][ks][0], bf[nt][ks][1]);
      }
    }
    // bias + SiLU -> fp16, staged per warp so every global store is a 16-byte vector of one pixel's channels
    __half* st = stage[warp];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int px = mt * 16 + g, ch = nt * 8 + 2 * t;
        *reinterpret_cast<__half2*>(st + px * 32 + ch) =
            __floats2half2_rn(stem_silu(acc[mt][nt][0] + bv[nt][0]), stem_silu(acc[mt][nt][1] + bv[nt][1]));
        *reinterpret_cast<__half2*>(st + (px + 8) * 32 + ch) =
            __floats2half2_rn(stem_silu(acc[mt][nt][2] + bv[nt][0]), stem_silu(acc[mt][nt][3] + bv[nt][1]));
      }
    __syncwarp();
    if (oy < Ho) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int v = i * 32 + lane;  // vector index: pixel v / 4, 8-channel group v % 4
        const int px = v >> 2, cg = v & 3;
        if (ox0 + px < Wo)
          *reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * Ho + oy) * Wo + ox0 + px) * ld_out + cg * 8) =
              *reinterpret_cast<const uint4*>(st + px * 32 + cg * 8);
      }
    }
    __syncwarp();
  }
}

#ifdef EFFOCR_AB
// 3x3, pad 1, stride s: NHWC fp16 (pixel pitch ld_in) -> [B*Ho*Wo, 9*C], column = (ky*3 + kx)*C + c.
// One 16-byte vector per thread; C / 8 is a power of two and the index fits 32 bits (host checks), so the index
// arithmetic is shifts, one constant division and two 32-bit divisions -- the first version's 64-bit runtime
// divisions (six per vector) held the kernel at ~1.9 TB/s.
template <int LOG_VPT>
__global__ void __launch_bounds__(256) yolo_im2col3_kernel(const __half* __restrict__ in, int ld_in,
                                                           __half* __restrict__ col, int H, int W, int stride,
                                                           unsigned Ho, unsigned Wo, unsigned total) {
  constexpr unsigned VPT = 1u << LOG_VPT;
  const unsigned C = VPT * 8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned cv = i & (VPT - 1);
    const unsigned t = i >> LOG_VPT;
    const unsigned row = t / 9u, tap = t - row * 9u;
    const unsigned r2 = row / Wo, ox = row - r2 * Wo;
    const unsigned b = r2 / Ho, oy = r2 - b * Ho;
    const unsigned ky = tap / 3u, kx = tap - ky * 3u;
    const int iy = static_cast<int>(oy) * stride - 1 + static_cast<int>(ky), ix = static_cast<int>(ox) * stride - 1 + static_cast<int>(kx);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = *reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * H + iy) * W + ix) * ld_in + cv * 8);
    *reinterpret_cast<uint4*>(col + static_cast<long long>(row) * (9u * C) + tap * C + cv * 8) = v;
  }
}

#endif  // EFFOCR_AB

// nearest 2x upsample into a channel slice of the consumer's concat buffer
__global__ void __launch_bounds__(256) yolo_upsample2x_kernel(const __half* __restrict__ in, int ld_in,
                                                              __half* __restrict__ out, int ld_out, int B, int h, int w,
                                                              int C) {
  const int vpt = C / 8;
  const long long total = static_cast<long long>(B) * (2 * h) * (2 * w) * vpt;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % vpt);
    const long long pix = i / vpt;
    const int ox = static_cast<int>(pix % (2 * w));
    const int oy = static_cast<int>((pix / (2 * w)) % (2 * h));
    const int b = static_cast<int>(pix / (4LL * w * h));
    const uint4 v = *reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * h + oy / 2) * w + ox / 2) * ld_in + cv * 8);
    *reinterpret_cast<uint4*>(out + pix * ld_out + cv * 8) = v;
  }
}

// SPPF: three chained 5x5 / stride 1 / pad 2 max-pools == window maxima of 5, 9, 13 (clipped at the border).
// Reads channels [0, C) of the concat buffer, writes [C, 2C), [2C, 3C), [3C, 4C).
__global__ void __launch_bounds__(256) yolo_pool5_kernel(__half* __restrict__ buf, int ld, int B, int h, int w, int C,
                                                         int src_off, int dst_off) {
  // one 5x5 / stride 1 / pad 2 max-pool of channels [src_off, src_off + C) into [dst_off, dst_off + C); SPPF chains three
  // of them (windows 5, 9, 13) exactly as ultralytics defines it -- 75 loads per output instead of the 169 of a direct
  // 13x13 window scan (the first version: 181 us per 64 lines)
  const int vpt = C / 8;
  const unsigned total = static_cast<unsigned>(B) * h * w * vpt;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned cv = i % vpt;
    const unsigned pix = i / vpt;
    const int x = static_cast<int>(pix % w);
    const int y = static_cast<int>((pix / w) % h);
    const long long img = static_cast<long long>(pix / (static_cast<unsigned>(w) * h)) * h * w;
    const __half2 ninf = __float2half2_rn(-INFINITY);
    __half2 m[4] = {ninf, ninf, ninf, ninf};
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(buf + (img + static_cast<long long>(yy) * w + xx) * ld + src_off + cv * 8);
        const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], hv[j]);
      }
    }
    *reinterpret_cast<uint4*>(buf + static_cast<long long>(pix) * ld + dst_off + cv * 8) = *reinterpret_cast<const uint4*>(m);
  }
}

// Detect decode (inference branch of ultralytics Detect.forward): raw[b*ny*nx + y*nx + x, a*no + o] ->
// out[b, off + (a*ny + y)*nx + x, o];  s = sigmoid(raw); xy = (2s + grid - 0.5) * stride; wh = (2s)^2 * anchor_px.
__global__ void __launch_bounds__(256) yolo_decode_kernel(const float* __restrict__ raw, int ldr, float* __restrict__ out,
                                                          int B, int ny, int nx, int no, int total_preds, int off,
                                                          float stride, float aw0, float ah0, float aw1, float ah1,
                                                          float aw2, float ah2) {
  const long long total = static_cast<long long>(B) * ny * nx * 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int a = static_cast<int>(i % 3);
    const long long pix = i / 3;
    const int x = static_cast<int>(pix % nx);
    const int y = static_cast<int>((pix / nx) % ny);
    const int b = static_cast<int>(pix / (static_cast<long long>(nx) * ny));
    const float* r = raw + pix * ldr + a * no;
    float* o = out + (static_cast<long long>(b) * total_preds + off + (static_cast<long long>(a) * ny + y) * nx + x) * no;
    const float aw = a == 0 ? aw0 : (a == 1 ? aw1 : aw2);
    const float ah = a == 0 ? ah0 : (a == 1 ? ah1 : ah2);
    for (int k = 0; k < no; ++k) {
      const float s = 1.0f / (1.0f + expf(-r[k]));
      float v = s;
      if (k == 0) v = (s * 2.0f + (static_cast<float>(x) - 0.5f)) * stride;
      else if (k == 1) v = (s * 2.0f + (static_cast<float>(y) - 0.5f)) * stride;
      else if (k == 2) v = (s * 2.0f) * (s * 2.0f) * aw;
      else if (k == 3) v = (s * 2.0f) * (s * 2.0f) * ah;
      o[k] = v;
    }
  }
}

// ------------------------------------------------------------------ host side
struct ConvW {
  __half* w = nullptr;  // [cout, kdim] fp16, BN folded, k>1: (ky, kx, cin) column order (layer 0: torch order)
  __half* w2 = nullptr; // split mode: [cout, 2 * kdim] = [Whi | Wlo] (layer 0: unused)
  float* wf = nullptr;  // split mode, layer 0 only: [108][32] fp32, k-major
  float* b = nullptr;   // [cout] fp32 folded BN bias
  int cin = 0, cout = 0, k = 1, s = 1, kdim = 0;
};

struct Act {
  __half* p;
  int ld;
};

struct YoloHandle {
  int nc = 2, no = 7, max_batch = 0, max_h = 0, max_w = 0, ldr = 24;
  int mode = EFFOCR_YOLO_SPLIT;  // EFFOCR_YOLO_SPLIT (reference-accurate, default) | EFFOCR_YOLO_FP16 (fast)
  long long lo_off = 0;          // element offset from an activation's hi plane to its lo plane (= arena_elems)
  std::vector<ConvW> convs;
  __half* det_w[3] = {nullptr, nullptr, nullptr};
  __half* det_w2[3] = {nullptr, nullptr, nullptr};  // [3 * no, 2 * C] = [Whi | Wlo]
  float* det_b[3] = {nullptr, nullptr, nullptr};
  float anchors_px[3][3][2];
  std::vector<void*> allocs;
  // workspace
  __half* col = nullptr;
  size_t col_elems = 0;
  __half* arena = nullptr;
  size_t arena_elems = 0;
  float* raw = nullptr;
  size_t raw_elems = 0;
  ~YoloHandle() {
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
};

// conv spec list in graph order (must match effocr_b200/localizer_engine.py::yolo_weight_order)
struct ConvSpec { int cin, cout, k, s; };
static void c3_specs(std::vector<ConvSpec>& v, int c1, int c2, int n) {
  const int c_ = c2 / 2;
  v.push_back({c1, c_, 1, 1});
  v.push_back({c1, c_, 1, 1});
  v.push_back({2 * c_, c2, 1, 1});
  for (int j = 0; j < n; ++j) {
    v.push_back({c_, c_, 1, 1});
    v.push_back({c_, c_, 3, 1});
  }
}
static std::vector<ConvSpec> yolo_specs() {
  std::vector<ConvSpec> v;
  v.push_back({3, 32, 6, 2});      // 0
  v.push_back({32, 64, 3, 2});     // 1
  c3_specs(v, 64, 64, 1);          // 2
  v.push_back({64, 128, 3, 2});    // 3
  c3_specs(v, 128, 128, 2);        // 4
  v.push_back({128, 256, 3, 2});   // 5
  c3_specs(v, 256, 256, 3);        // 6
  v.push_back({256, 512, 3, 2});   // 7
  c3_specs(v, 512, 512, 1);        // 8
  v.push_back({512, 256, 1, 1});   // 9 SPPF cv1
  v.push_back({1024, 512, 1, 1});  // 9 SPPF cv2
  v.push_back({512, 256, 1, 1});   // 10
  c3_specs(v, 512, 256, 1);        // 13
  v.push_back({256, 128, 1, 1});   // 14
  c3_specs(v, 256, 128, 1);        // 17
  v.push_back({128, 128, 3, 2});   // 18
  c3_specs(v, 256, 256, 1);        // 20
  v.push_back({256, 256, 3, 2});   // 21
  c3_specs(v, 512, 512, 1);        // 23
  return v;
}

static inline int grid_for(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = 148LL * 32;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

template <int BN, int CB, bool SPLIT = false>
static int launch_conv3(const __half* in, int ld_in, int B, int H, int W, int C, const __half* w, int kdim, const float* bias,
                        int Cout, int stride, __half* out, int ld_out, const __half* resid, int ld_res, cudaStream_t s,
                        long long lo_off = 0) {
  using Cfg = Conv3Cfg<BN, CB>;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  CUtensorMap tin, tw, tout, tin_lo, tout_lo;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    const uint64_t st[3] = {static_cast<uint64_t>(ld_in) * 2, static_cast<uint64_t>(W) * ld_in * 2, static_cast<uint64_t>(H) * W * ld_in * 2};
    const uint32_t box[4] = {CB, static_cast<uint32_t>(kConvTW * stride), static_cast<uint32_t>(kConvTH * stride), 1};
    const uint32_t es[4] = {1, static_cast<uint32_t>(stride), static_cast<uint32_t>(stride), 1};
    EFFOCR_TRY(make_tmap_4d(&tin, in, 2, dims, st, box, es, CB * 2));
    if (SPLIT) EFFOCR_TRY(make_tmap_4d(&tin_lo, in + lo_off, 2, dims, st, box, es, CB * 2));
    else tin_lo = tin;
  }
  // SPLIT: `w` is [Cout, 2 * kdim] = [Whi | Wlo]
  EFFOCR_TRY(make_tmap_2d(&tw, w, 2, Cout, SPLIT ? 2 * kdim : kdim, SPLIT ? 2 * kdim : kdim, BN, CB, CB * 2));
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho), static_cast<uint64_t>(B)};
    const uint64_t st[3] = {static_cast<uint64_t>(ld_out) * 2, static_cast<uint64_t>(Wo) * ld_out * 2, static_cast<uint64_t>(Ho) * Wo * ld_out * 2};
    const uint32_t box[4] = {32, kConvTW, 2, 1};
    const uint32_t es[4] = {1, 1, 1, 1};
    EFFOCR_TRY(make_tmap_4d(&tout, out, 2, dims, st, box, es, 64));
    if (SPLIT) EFFOCR_TRY(make_tmap_4d(&tout_lo, out + lo_off, 2, dims, st, box, es, 64));
    else tout_lo = tout;
  }
  auto kern = conv3_tc_kernel<BN, CB, SPLIT>;
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  Conv3Params p;
  p.B = B; p.Ho = Ho; p.Wo = Wo; p.C = C; p.Cout = Cout; p.stride = stride; p.bias = bias; p.resid = resid; p.ld_res = ld_res;
  p.lo_off = lo_off; p.kdim = kdim;
  const int tiles = B * ((Ho + kConvTH - 1) / kConvTH) * ((Wo + kConvTW - 1) / kConvTW) * ((Cout + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  {
    KernelScope ks(PROF_GEMM_OTHER, s);
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, s>>>(tin, tw, tout, tin_lo, tout_lo, p);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

// 1x1 convolution / Detect head in split precision: out = act(A . W^T + bias) with A = (hi, lo) planes, W2 = [Whi | Wlo]
template <int BN, bool F32, int ACT>
static int launch_split3_bn(const __half* A, long long lda, long long lo_off, const __half* w2, int M, int N, int K,
                            const float* bias, void* out, long long ldo, cudaStream_t s) {
  using Cfg = GemmTmaCfg<BN, F32>;
  CUtensorMap ta, tal, tw, tc, tcl;
  EFFOCR_TRY(make_tmap_f16_2d(&ta, A, M, K, lda, kBlockM));
  EFFOCR_TRY(make_tmap_f16_2d(&tal, A + lo_off, M, K, lda, kBlockM));
  EFFOCR_TRY(make_tmap_f16_2d(&tw, w2, N, 2 * K, 2 * K, BN));
  if (F32) {
    EFFOCR_TRY(make_tmap_2d(&tc, out, 4, M, N, ldo, 32, 32, 128));
    tcl = tc;
  } else {
    EFFOCR_TRY(make_tmap_2d(&tc, out, 2, M, N, ldo, 32, 32, 64));
    EFFOCR_TRY(make_tmap_2d(&tcl, reinterpret_cast<__half*>(out) + lo_off, 2, M, N, ldo, 32, 32, 64));
  }
  auto kern = gemm_split3_kernel<BN, F32, ACT>;
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int tiles = ((M + kBlockM - 1) / kBlockM) * ((N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  Split3Params p;
  p.M = M; p.N = N; p.K = K; p.bias = bias;
  {
    KernelScope ks(PROF_GEMM_OTHER, s);
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, s>>>(ta, tal, tw, tc, tcl, p);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}
template <bool F32, int ACT>
static int launch_split3(const __half* A, long long lda, long long lo_off, const __half* w2, int M, int N, int K,
                         const float* bias, void* out, long long ldo, cudaStream_t s) {
  switch (choose_block_n(N)) {
    case 64: return launch_split3_bn<64, F32, ACT>(A, lda, lo_off, w2, M, N, K, bias, out, ldo, s);
    case 128: return launch_split3_bn<128, F32, ACT>(A, lda, lo_off, w2, M, N, K, bias, out, ldo, s);
    case 192: return launch_split3_bn<192, F32, ACT>(A, lda, lo_off, w2, M, N, K, bias, out, ldo, s);
    default: return launch_split3_bn<256, F32, ACT>(A, lda, lo_off, w2, M, N, K, bias, out, ldo, s);
  }
}

struct YoloRun {
  YoloHandle* h;
  cudaStream_t s;
  int B;
  size_t conv_i = 0;
  size_t arena_off = 0;
  int status = EFFOCR_OK;

  Act buf(long long pixels, int C) {
    Act a{h->arena + arena_off, C};
    arena_off += static_cast<size_t>(pixels) * C;
    arena_off = (arena_off + 127) / 128 * 128;
    if (arena_off > h->arena_elems && status == EFFOCR_OK) status = fail(EFFOCR_ERR_NOMEM, "yolo: activation arena too small");
    return a;
  }
  static Act slice(Act a, int c0) { return Act{a.p + c0, a.ld}; }

  bool split() const { return h->mode == EFFOCR_YOLO_SPLIT; }
  void gemm(const __half* A, int lda, long long rows, const ConvW& cw, Act out, const Act* resid) {
    if (status) return;
    if (split() && !resid) {
      status = launch_split3<false, ACT_SILU>(A, lda, h->lo_off, cw.w2, static_cast<int>(rows), cw.cout, cw.kdim, cw.b, out.p,
                                              out.ld, s);
      return;
    }
    GemmArgs g;
    g.A = A; g.lda = lda; g.W = cw.w; g.ldw = cw.kdim; g.M = static_cast<int>(rows); g.N = cw.cout; g.K = cw.kdim;
    g.out = out.p; g.ldo = out.ld; g.bias = cw.b; g.act = 2;
    if (resid) { g.resid = resid->p; g.ldr = resid->ld; }
    g.prof_tag = PROF_GEMM_OTHER;
    status = gemm_f16(g, s);
  }
  // 1x1 conv + BN + SiLU (+ residual)
  void conv1(Act in, long long rows, Act out, const Act* resid = nullptr) {
    const ConvW& cw = h->convs[conv_i++];
    gemm(in.p, in.ld, rows, cw, out, resid);
  }
  // 3x3 conv (stride 1 or 2) + BN + SiLU (+ residual); returns nothing, output dims are the caller's
  void conv3(Act in, int H, int W, Act out, const Act* resid = nullptr) {
    const ConvW& cw = h->convs[conv_i++];
    if (status) return;
    const int Ho = (H - 1) / cw.s + 1, Wo = (W - 1) / cw.s + 1;
    const long long rows = static_cast<long long>(B) * Ho * Wo;
#ifdef EFFOCR_AB
    static const bool use_im2col = [] {
      const char* e = getenv("EFFOCR_YOLO_CONV3");  // "im2col" = explicit gather + GEMM (first version; A/B runs)
      return e && e[0] == 'i';
    }();
#else
    constexpr bool use_im2col = false;
#endif
    if (split()) {
      const __half* rp = resid ? resid->p : nullptr;
      const int rl = resid ? resid->ld : 0;
      const long long lo = h->lo_off;
      if (cw.cin == 32) status = launch_conv3<64, 32, true>(in.p, in.ld, B, H, W, cw.cin, cw.w2, cw.kdim, cw.b, cw.cout, cw.s, out.p, out.ld, rp, rl, s, lo);
      else if (cw.cout <= 64) status = launch_conv3<64, 64, true>(in.p, in.ld, B, H, W, cw.cin, cw.w2, cw.kdim, cw.b, cw.cout, cw.s, out.p, out.ld, rp, rl, s, lo);
      else if (cw.cout <= 128) status = launch_conv3<128, 64, true>(in.p, in.ld, B, H, W, cw.cin, cw.w2, cw.kdim, cw.b, cw.cout, cw.s, out.p, out.ld, rp, rl, s, lo);
      else status = launch_conv3<256, 64, true>(in.p, in.ld, B, H, W, cw.cin, cw.w2, cw.kdim, cw.b, cw.cout, cw.s, out.p, out.ld, rp, rl, s, lo);
      return;
    }
    if (!use_im2col) {  // implicit GEMM: the A operand comes straight from the activation tensor through 4-D TMA boxes
      const __half* rp = resid ? resid->p : nullptr;
      const int rl = resid ? resid->ld : 0;
      if (cw.cin == 32) status = launch_conv3<64, 32>(in.p, in.ld, B, H, W, cw.cin, cw.w, cw.kdim, cw.b, cw.cout, cw.s, out.p, out.ld, rp, rl, s);
      else if (cw.cout <= 64) status = launch_conv3<64, 64>(in.p, in.ld, B, H, W, cw.cin, cw.w, cw.kdim, cw.b, cw.cout, cw.s, out.p, out.ld, rp, rl, s);
      else if (cw.cout <= 128) status = launch_conv3<128, 64>(in.p, in.ld, B, H, W, cw.cin, cw.w, cw.kdim, cw.b, cw.cout, cw.s, out.p, out.ld, rp, rl, s);
      else status = launch_conv3<256, 64>(in.p, in.ld, B, H, W, cw.cin, cw.w, cw.kdim, cw.b, cw.cout, cw.s, out.p, out.ld, rp, rl, s);
      return;
    }
#ifdef EFFOCR_AB
    if (static_cast<size_t>(rows) * cw.kdim > h->col_elems) { status = fail(EFFOCR_ERR_NOMEM, "yolo: im2col scratch too small"); return; }
    {
      KernelScope ks(PROF_CONV_IM2COL, s);
      const long long total = rows * 9 * (cw.cin / 8);
      if (total >= (1LL << 32)) { status = fail(EFFOCR_ERR_INVALID, "yolo: batch too large for the 32-bit im2col index"); return; }
      const int grid = grid_for(total);
      const unsigned tot = static_cast<unsigned>(total);
      switch (cw.cin) {
        case 32: yolo_im2col3_kernel<2><<<grid, 256, 0, s>>>(in.p, in.ld, h->col, H, W, cw.s, Ho, Wo, tot); break;
        case 64: yolo_im2col3_kernel<3><<<grid, 256, 0, s>>>(in.p, in.ld, h->col, H, W, cw.s, Ho, Wo, tot); break;
        case 128: yolo_im2col3_kernel<4><<<grid, 256, 0, s>>>(in.p, in.ld, h->col, H, W, cw.s, Ho, Wo, tot); break;
        case 256: yolo_im2col3_kernel<5><<<grid, 256, 0, s>>>(in.p, in.ld, h->col, H, W, cw.s, Ho, Wo, tot); break;
        default: status = fail(EFFOCR_ERR_INVALID, "yolo: 3x3 conv input channels must be 32, 64, 128 or 256"); return;
      }
    }
    gemm(h->col, cw.kdim, rows, cw, out, resid);
#endif
  }
  // C3(c1, c2, n, shortcut) at a level with `pix` pixels of size Hl x Wl
  void c3(Act in, int c2, int n, bool shortcut, int Hl, int Wl, Act out) {
    const long long pix = static_cast<long long>(B) * Hl * Wl;
    const int c_ = c2 / 2;
    const size_t save = arena_off;
    Act cat = buf(pix, 2 * c_);
    Act ya = buf(pix, c_), yb = buf(pix, c_), t = buf(pix, c_);
    // weight order: cv1, cv2, cv3, m.j.cv1, m.j.cv2 -- fetch explicitly
    const size_t i_cv1 = conv_i, i_cv2 = conv_i + 1, i_cv3 = conv_i + 2, i_m = conv_i + 3;
    conv_i = i_cv1;
    Act y = (n == 0) ? slice(cat, 0) : ya;
    conv1(in, pix, y);                       // cv1
    conv_i = i_cv2;
    conv1(in, pix, slice(cat, c_));          // cv2 -> second half of the concat
    conv_i = i_m;
    for (int j = 0; j < n; ++j) {
      conv1(y, pix, t);                      // m.j.cv1
      Act dst = (j == n - 1) ? slice(cat, 0) : (y.p == ya.p ? yb : ya);
      conv3(t, Hl, Wl, dst, shortcut ? &y : nullptr);  // m.j.cv2 (+ y)
      y = dst;
    }
    const size_t i_end = conv_i;
    conv_i = i_cv3;
    conv1(cat, pix, out);                    // cv3
    conv_i = i_end;
    arena_off = save;                        // scratch of this block is dead once cv3 has been enqueued (stream order)
  }
};

static int yolo_forward_impl(YoloHandle* h, const float* img, int B, int H, int W, float* pred, cudaStream_t s) {
  YoloRun r{h, s, B};
  const int H1 = H / 2, W1 = W / 2, H2 = H / 4, W2 = W / 4, H3 = H / 8, W3 = W / 8, H4 = H / 16, W4 = W / 16, H5 = H / 32,
            W5 = W / 32;
  const long long p1 = 1LL * B * H1 * W1, p2 = 1LL * B * H2 * W2, p3 = 1LL * B * H3 * W3, p4 = 1LL * B * H4 * W4,
                  p5 = 1LL * B * H5 * W5;
  // long-lived buffers first (never released)
  Act x0 = r.buf(p1, 32), x1 = r.buf(p2, 64), x2 = r.buf(p2, 64), x3 = r.buf(p3, 128), x5 = r.buf(p4, 256),
      x7 = r.buf(p5, 512), x8 = r.buf(p5, 512), catS = r.buf(p5, 1024), x9 = r.buf(p5, 512);
  Act cat12 = r.buf(p4, 512), cat16 = r.buf(p3, 256), cat19 = r.buf(p4, 256), cat22 = r.buf(p5, 512);
  Act x13 = r.buf(p4, 256), x17 = r.buf(p3, 128), x20 = r.buf(p4, 256), x23 = r.buf(p5, 512);
  if (r.status) return r.status;
  // layer 0
  {
    const ConvW& cw = h->convs[r.conv_i++];
#ifdef EFFOCR_AB
    static const bool stem_gemm = [] {
      const char* e = getenv("EFFOCR_YOLO_STEM");  // "gemm" = explicit im2col + tcgen05 GEMM (first version; A/B runs)
      return e && e[0] == 'g';
    }();
#else
    constexpr bool stem_gemm = false;
#endif
    if (r.split()) {
      const int tiles = B * ((H1 + kStemF32TH - 1) / kStemF32TH) * ((W1 + kStemF32TW - 1) / kStemF32TW);
      const int grid = tiles < 4 * sm_count() ? tiles : 4 * sm_count();
      KernelScope ks(PROF_YOLO_MISC, s);
      yolo_stem_f32_kernel<<<grid, 256, 0, s>>>(img, cw.wf, cw.b, x0.p, x0.p + h->lo_off, x0.ld, B, H, W);
      EFFOCR_CUDA(cudaGetLastError());
    } else if (stem_gemm) {
#ifdef EFFOCR_AB
      if (static_cast<size_t>(p1) * 112 > h->col_elems) return fail(EFFOCR_ERR_NOMEM, "yolo: im2col scratch too small");
      {
        KernelScope ks(PROF_CONV_IM2COL, s);
        yolo_im2col0_kernel<<<grid_for(p1 * 14), 256, 0, s>>>(img, h->col, B, H, W);
      }
      r.gemm(h->col, 112, p1, cw, x0, nullptr);
#endif
    } else {
      const int tiles = B * ((H1 + kStemTH - 1) / kStemTH) * ((W1 + kStemTW - 1) / kStemTW);
      const int grid = tiles < sm_count() ? tiles : sm_count();  // persistent, one CTA per SM (222 registers: B fragments + prefetch)
      KernelScope ks(PROF_YOLO_MISC, s);
      yolo_stem_kernel<<<grid, 256, 0, s>>>(img, cw.w, cw.b, x0.p, x0.ld, B, H, W);
      EFFOCR_CUDA(cudaGetLastError());
    }
  }
  r.conv3(x0, H1, W1, x1);                              // 1
  r.c3(x1, 64, 1, true, H2, W2, x2);                    // 2
  r.conv3(x2, H2, W2, x3);                              // 3
  r.c3(x3, 128, 2, true, H3, W3, YoloRun::slice(cat16, 128));   // 4 -> second half of cat16
  r.conv3(YoloRun::slice(cat16, 128), H3, W3, x5);      // 5
  r.c3(x5, 256, 3, true, H4, W4, YoloRun::slice(cat12, 256));   // 6 -> second half of cat12
  r.conv3(YoloRun::slice(cat12, 256), H4, W4, x7);      // 7
  r.c3(x7, 512, 1, true, H5, W5, x8);                   // 8
  r.conv1(x8, p5, YoloRun::slice(catS, 0));             // 9 SPPF cv1 -> first quarter of catS
  if (!r.status) {
    KernelScope ks(PROF_YOLO_MISC, s);
    for (int k = 0; k < 3; ++k) {  // windows 5, 9, 13 = three chained 5x5 pools
      if (r.split()) yolo_pool5_split_kernel<<<grid_for(p5 * 32), 256, 0, s>>>(catS.p, h->lo_off, 1024, B, H5, W5, 256, k * 256, (k + 1) * 256);
      else yolo_pool5_kernel<<<grid_for(p5 * 32), 256, 0, s>>>(catS.p, 1024, B, H5, W5, 256, k * 256, (k + 1) * 256);
    }
  }
  r.conv1(catS, p5, x9);                                // 9 SPPF cv2
  r.conv1(x9, p5, YoloRun::slice(cat22, 256));          // 10 -> second half of cat22
  if (!r.status) {
    KernelScope ks(PROF_YOLO_MISC, s);
    yolo_upsample2x_kernel<<<grid_for(p4 * 32), 256, 0, s>>>(cat22.p + 256, 512, cat12.p, 512, B, H5, W5, 256);  // 11, 12
    if (r.split()) yolo_upsample2x_kernel<<<grid_for(p4 * 32), 256, 0, s>>>(cat22.p + h->lo_off + 256, 512, cat12.p + h->lo_off, 512, B, H5, W5, 256);
  }
  r.c3(cat12, 256, 1, false, H4, W4, x13);              // 13
  r.conv1(x13, p4, YoloRun::slice(cat19, 128));         // 14 -> second half of cat19
  if (!r.status) {
    KernelScope ks(PROF_YOLO_MISC, s);
    yolo_upsample2x_kernel<<<grid_for(p3 * 16), 256, 0, s>>>(cat19.p + 128, 256, cat16.p, 256, B, H4, W4, 128);  // 15, 16
    if (r.split()) yolo_upsample2x_kernel<<<grid_for(p3 * 16), 256, 0, s>>>(cat19.p + h->lo_off + 128, 256, cat16.p + h->lo_off, 256, B, H4, W4, 128);
  }
  r.c3(cat16, 128, 1, false, H3, W3, x17);              // 17
  r.conv3(x17, H3, W3, YoloRun::slice(cat19, 0));       // 18 -> first half of cat19
  r.c3(cat19, 256, 1, false, H4, W4, x20);              // 20
  r.conv3(x20, H4, W4, YoloRun::slice(cat22, 0));       // 21 -> first half of cat22
  r.c3(cat22, 512, 1, false, H5, W5, x23);              // 23
  if (r.status) return r.status;
  // Detect
  const int total_preds = 3 * (H3 * W3 + H4 * W4 + H5 * W5);
  const Act feats[3] = {x17, x20, x23};
  const int chans[3] = {128, 256, 512};
  const int nys[3] = {H3, H4, H5}, nxs[3] = {W3, W4, W5};
  const float strides[3] = {8.f, 16.f, 32.f};
  int off = 0;
  for (int l = 0; l < 3; ++l) {
    const long long rows = 1LL * B * nys[l] * nxs[l];
    if (static_cast<size_t>(rows) * h->ldr > h->raw_elems) return fail(EFFOCR_ERR_NOMEM, "yolo: detect scratch too small");
    if (r.split()) {
      EFFOCR_TRY((launch_split3<true, ACT_NONE>(feats[l].p, feats[l].ld, h->lo_off, h->det_w2[l], static_cast<int>(rows), 3 * h->no,
                                                chans[l], h->det_b[l], h->raw, h->ldr, s)));
    } else {
      GemmArgs g;
      g.A = feats[l].p; g.lda = feats[l].ld; g.W = h->det_w[l]; g.ldw = chans[l]; g.M = static_cast<int>(rows);
      g.N = 3 * h->no; g.K = chans[l]; g.out = h->raw; g.ldo = h->ldr; g.out_f32 = 1; g.bias = h->det_b[l];
      g.prof_tag = PROF_GEMM_OTHER;
      EFFOCR_TRY(gemm_f16(g, s));
    }
    {
      KernelScope ks(PROF_YOLO_MISC, s);
      yolo_decode_kernel<<<grid_for(rows * 3), 256, 0, s>>>(h->raw, h->ldr, pred, B, nys[l], nxs[l], h->no, total_preds, off,
                                                            strides[l], h->anchors_px[l][0][0], h->anchors_px[l][0][1],
                                                            h->anchors_px[l][1][0], h->anchors_px[l][1][1],
                                                            h->anchors_px[l][2][0], h->anchors_px[l][2][1]);
    }
    off += 3 * nys[l] * nxs[l];
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

}  // namespace effocr

using namespace effocr;

extern "C" int effocr_yolo_create(int nc, int max_batch, int max_h, int max_w, const float* const* h_weights,
                                  int n_weights, effocr_yolo_t* out) {
  if (!out) return fail(EFFOCR_ERR_INVALID, "yolo_create: null out");
  *out = nullptr;
  EFFOCR_TRY(require_sm100());
  const std::vector<ConvSpec> specs = yolo_specs();
  if (n_weights != static_cast<int>(specs.size()) * 5 + 7)
    return fail(EFFOCR_ERR_INVALID, "yolo_create: expected 5 tensors per conv (w, bn.w, bn.b, bn.mean, bn.var) + 3 x (w, b) + anchors");
  if (nc < 1 || nc > 80 || max_batch < 1 || max_h % 32 || max_w % 32 || max_h < 32 || max_w < 32)
    return fail(EFFOCR_ERR_INVALID, "yolo_create: bad nc / batch / input shape (H, W must be multiples of 32)");
  YoloHandle* h = new YoloHandle();
  h->nc = nc; h->no = 5 + nc; h->max_batch = max_batch; h->max_h = max_h; h->max_w = max_w;
  h->ldr = (3 * h->no + 3) / 4 * 4;
  int st = EFFOCR_OK;
  for (size_t i = 0; i < specs.size() && !st; ++i) {
    const ConvSpec& sp = specs[i];
    const float* w = h_weights[5 * i];
    const float *g = h_weights[5 * i + 1], *bb = h_weights[5 * i + 2], *mu = h_weights[5 * i + 3], *var = h_weights[5 * i + 4];
    ConvW cw;
    cw.cin = sp.cin; cw.cout = sp.cout; cw.k = sp.k; cw.s = sp.s;
    const int kk = sp.k * sp.k;
    cw.kdim = (sp.k == 6) ? 112 : sp.cin * kk;
    std::vector<__half> wh(static_cast<size_t>(sp.cout) * cw.kdim, __float2half_rn(0.f));
    std::vector<float> bias(sp.cout);
    for (int o = 0; o < sp.cout; ++o) {
      const float scale = g[o] / sqrtf(var[o] + 1e-3f);
      bias[o] = bb[o] - mu[o] * scale;
      for (int c = 0; c < sp.cin; ++c)
        for (int t = 0; t < kk; ++t) {
          const float v = w[(static_cast<size_t>(o) * sp.cin + c) * kk + t] * scale;
          const size_t col = (sp.k == 6) ? static_cast<size_t>(c) * 36 + t : static_cast<size_t>(t) * sp.cin + c;
          wh[static_cast<size_t>(o) * cw.kdim + col] = __float2half_rn(v);
        }
    }
    if ((st = h->alloc(&cw.w, wh.size()))) break;
    if ((st = h->alloc(&cw.b, bias.size()))) break;
    cudaMemcpy(cw.w, wh.data(), wh.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(cw.b, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
    // split mode: the folded fp32 weight as [Whi | Wlo] (layer 0: the fp32 matrix itself, k-major)
    if (sp.k == 6) {
      std::vector<float> wf(static_cast<size_t>(108) * 32);
      for (int o = 0; o < sp.cout; ++o) {
        const float scale = g[o] / sqrtf(var[o] + 1e-3f);
        for (int c = 0; c < 3; ++c)
          for (int t = 0; t < 36; ++t) wf[static_cast<size_t>(c * 36 + t) * 32 + o] = w[(static_cast<size_t>(o) * 3 + c) * 36 + t] * scale;
      }
      if ((st = h->alloc(&cw.wf, wf.size()))) break;
      cudaMemcpy(cw.wf, wf.data(), wf.size() * 4, cudaMemcpyHostToDevice);
    } else {
      std::vector<__half> w2(static_cast<size_t>(sp.cout) * 2 * cw.kdim);
      for (int o = 0; o < sp.cout; ++o) {
        const float scale = g[o] / sqrtf(var[o] + 1e-3f);
        for (int c = 0; c < sp.cin; ++c)
          for (int t = 0; t < kk; ++t) {
            const float v = w[(static_cast<size_t>(o) * sp.cin + c) * kk + t] * scale;
            const size_t col = static_cast<size_t>(t) * sp.cin + c;
            const __half hi = __float2half_rn(v);
            w2[static_cast<size_t>(o) * 2 * cw.kdim + col] = hi;
            w2[static_cast<size_t>(o) * 2 * cw.kdim + cw.kdim + col] = __float2half_rn(v - __half2float(hi));
          }
      }
      if ((st = h->alloc(&cw.w2, w2.size()))) break;
      cudaMemcpy(cw.w2, w2.data(), w2.size() * 2, cudaMemcpyHostToDevice);
    }
    h->convs.push_back(cw);
  }
  const int chans[3] = {128, 256, 512};
  const float strides[3] = {8.f, 16.f, 32.f};
  const size_t base = specs.size() * 5;
  for (int l = 0; l < 3 && !st; ++l) {
    const float* w = h_weights[base + 2 * l];
    const float* b = h_weights[base + 2 * l + 1];
    const int n = 3 * h->no;
    std::vector<__half> wh(static_cast<size_t>(n) * chans[l]);
    for (size_t i = 0; i < wh.size(); ++i) wh[i] = __float2half_rn(w[i]);
    if ((st = h->alloc(&h->det_w[l], wh.size()))) break;
    if ((st = h->alloc(&h->det_b[l], static_cast<size_t>(h->ldr)))) break;
    cudaMemcpy(h->det_w[l], wh.data(), wh.size() * 2, cudaMemcpyHostToDevice);
    {
      std::vector<__half> w2(static_cast<size_t>(n) * 2 * chans[l]);
      for (int o = 0; o < n; ++o)
        for (int c = 0; c < chans[l]; ++c) {
          const float v = w[static_cast<size_t>(o) * chans[l] + c];
          const __half hi = __float2half_rn(v);
          w2[static_cast<size_t>(o) * 2 * chans[l] + c] = hi;
          w2[static_cast<size_t>(o) * 2 * chans[l] + chans[l] + c] = __float2half_rn(v - __half2float(hi));
        }
      if ((st = h->alloc(&h->det_w2[l], w2.size()))) break;
      cudaMemcpy(h->det_w2[l], w2.data(), w2.size() * 2, cudaMemcpyHostToDevice);
    }
    cudaMemset(h->det_b[l], 0, h->ldr * 4);
    cudaMemcpy(h->det_b[l], b, n * 4, cudaMemcpyHostToDevice);
    const float* an = h_weights[base + 6];  // [3,3,2] in stride units (ultralytics buffer `anchors`)
    for (int a = 0; a < 3; ++a)
      for (int d = 0; d < 2; ++d) h->anchors_px[l][a][d] = an[(l * 3 + a) * 2 + d] * strides[l];
  }
  if (!st) {
    const size_t px = static_cast<size_t>(max_batch) * max_h * max_w;
    h->col_elems = px / 4 * 112 + 4096;  // layer 0 dominates: (H/2 * W/2) rows x 112
    const size_t col2 = px / 16 * 288 + 4096;
    if (col2 > h->col_elems) h->col_elems = col2;
    // activation arena: long-lived buffers (33.5 elements per input pixel) + the largest C3 scratch
    // (10 per pixel, at stride 4) + per-buffer 128-element rounding
    h->arena_elems = px * 48 + (1u << 20);
    h->lo_off = static_cast<long long>(h->arena_elems);  // the lo planes of all activations: a second arena right behind the first
    h->raw_elems = px / 64 * h->ldr + 4096;
    if (!(st = h->alloc(&h->col, h->col_elems)) && !(st = h->alloc(&h->arena, 2 * h->arena_elems))) st = h->alloc(&h->raw, h->raw_elems);
  }
  if (!st && cudaDeviceSynchronize() != cudaSuccess) st = fail(EFFOCR_ERR_CUDA, "yolo_create: upload failed");
  if (st) { delete h; return st; }
  *out = reinterpret_cast<effocr_yolo_t>(h);
  return EFFOCR_OK;
}

extern "C" void effocr_yolo_destroy(effocr_yolo_t h) { delete reinterpret_cast<YoloHandle*>(h); }

extern "C" int effocr_yolo_set_mode(effocr_yolo_t handle, int mode) {
  YoloHandle* h = reinterpret_cast<YoloHandle*>(handle);
  if (!h || (mode != EFFOCR_YOLO_SPLIT && mode != EFFOCR_YOLO_FP16)) return fail(EFFOCR_ERR_INVALID, "yolo_set_mode: bad handle / mode");
  h->mode = mode;
  return EFFOCR_OK;
}

extern "C" int effocr_yolo_num_predictions(int height, int width) {
  if (height % 32 || width % 32) return -1;
  return 3 * ((height / 8) * (width / 8) + (height / 16) * (width / 16) + (height / 32) * (width / 32));
}

extern "C" int effocr_yolo_forward(effocr_yolo_t handle, const float* d_images, int batch, int height, int width,
                                   float* d_pred, void* stream) {
  YoloHandle* h = reinterpret_cast<YoloHandle*>(handle);
  if (!h || !d_images || !d_pred || batch < 0) return fail(EFFOCR_ERR_INVALID, "yolo_forward: bad arguments");
  if (height % 32 || width % 32 || height < 32 || width < 32) return fail(EFFOCR_ERR_INVALID, "yolo_forward: H and W must be multiples of 32");
  if (static_cast<long long>(height) * width > static_cast<long long>(h->max_h) * h->max_w)
    return fail(EFFOCR_ERR_INVALID, "yolo_forward: image larger than the handle's workspace");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int npred = effocr_yolo_num_predictions(height, width);
  for (int b0 = 0; b0 < batch; b0 += h->max_batch) {
    const int B = batch - b0 < h->max_batch ? batch - b0 : h->max_batch;
    EFFOCR_TRY(yolo_forward_impl(h, d_images + static_cast<size_t>(b0) * 3 * height * width, B, height, width,
                                 d_pred + static_cast<size_t>(b0) * npred * h->no, s));
  }
  return EFFOCR_OK;
}
