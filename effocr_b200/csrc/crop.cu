// K1: fused crop -> white square pad -> anti-aliased bilinear resize -> ImageNet normalise -> fp16,
// gathering boxes straight out of the u8 line images into the recognizer's input tensor
// (optionally already in the patch-major layout the ViT patch-embedding GEMM reads).
//
// Replaces the per-crop CPU transform of the reference:
//   /root/reference/infer_effocr.py:284-293 (numpy slice im[y0:y1, x0:x1]) +
//   /root/reference/utils/datasets_utils.py:166-172 create_paired_transform =
//   MedianPad(override white, pad right/bottom) :69-90 -> ToTensor -> Resize((224,224)) (torchvision
//   0.26: bilinear, align_corners=False, antialias=True) -> Normalize(IMAGENET mean/std),
// measured at 281 crops/s/thread on the CPU.
//
// HBM-bound by construction: reads h*w*3 bytes, writes 3*224*224*2 = 301 056 bytes per crop.
#include "../../include/effocr_b200.h"
#include "host_common.h"

namespace effocr {

struct AxisTaps {
  int mn, sz;
  float center, inv_total;
};

// ATen _compute_indices_min_size_weights_aa (triangle filter, interp_size 2), fp32 like ATen.
__device__ __forceinline__ AxisTaps make_taps(int i, float scale, float support, float invscale, int in_size) {
  AxisTaps t;
  t.center = scale * (static_cast<float>(i) + 0.5f);
  int mn = static_cast<int>(t.center - support + 0.5f);
  int mx = static_cast<int>(t.center + support + 0.5f);
  t.mn = mn < 0 ? 0 : mn;
  mx = mx > in_size ? in_size : mx;
  t.sz = mx - t.mn;
  float total = 0.f;
  for (int j = 0; j < t.sz; ++j) {
    const float w = 1.0f - fabsf((static_cast<float>(j + t.mn) - t.center + 0.5f) * invscale);
    total += w > 0.f ? w : 0.f;
  }
  t.inv_total = total != 0.f ? 1.0f / total : 0.f;
  return t;
}
__device__ __forceinline__ float tap_weight(const AxisTaps& t, int j, float invscale) {
  const float w = 1.0f - fabsf((static_cast<float>(j + t.mn) - t.center + 0.5f) * invscale);
  return (w > 0.f ? w : 0.f) * t.inv_total;
}

// numpy basic-slice bounds for a[start:stop] on an axis of length n
__device__ __forceinline__ void numpy_slice(int start, int stop, int n, int* lo, int* len) {
  if (start < 0) { start += n; if (start < 0) start = 0; }
  if (start > n) start = n;
  if (stop < 0) { stop += n; if (stop < 0) stop = 0; }
  if (stop > n) stop = n;
  *lo = start;
  *len = stop > start ? stop - start : 0;
}

constexpr int kCropMaxStage = 3072;  // staged source pixels (float4 each) per block: 48 KB
constexpr int kCropSmemBytes = kCropMaxStage * 16 + 224 * 16 + 224 * 4;  // 53 KB: four blocks per SM

constexpr float kWhiteR = (1.0f - 0.485f) / 0.229f, kWhiteG = (1.0f - 0.456f) / 0.224f, kWhiteB = (1.0f - 0.406f) / 0.225f;

template <int LAYOUT>  // 0 NCHW f16, 1 NCHW f32, 2 patch-major f16 ([n*196, 768], ViT/16), 3 4x4-patch-major f16 ([n*3136, 48])
__global__ void __launch_bounds__(224) crop_resize_kernel(const uint8_t* __restrict__ pixels,
                                                          const effocr_image_desc* __restrict__ images,
                                                          const effocr_crop_box* __restrict__ boxes, int n_boxes,
                                                          void* __restrict__ out) {
  constexpr int OUT = 224;
  const int n = blockIdx.y;
  const int i = blockIdx.x * 8 + threadIdx.x / 28;  // output row
  const int j0 = (threadIdx.x % 28) * 8;            // first of 8 output columns
  const effocr_crop_box bx = boxes[n];
  const effocr_image_desc im = images[bx.image];
  int x_lo, w, y_lo, h;
  numpy_slice(bx.x0, bx.x1, im.width, &x_lo, &w);
  numpy_slice(bx.y0, bx.y1, im.height, &y_lo, &h);
  const bool empty = (w == 0 || h == 0);

  float acc[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[c][k] = 0.f;

  // ---- fast path (block-uniform): the 8 output rows of this block read a small band of source rows; stage it in
  // shared memory as float RGB once (each source pixel feeds ~12 outputs when a 64-px crop is blown up to 224)
  // together with the horizontal tap table, so the inner loop is LDS.128 + FMA only.
  extern __shared__ __align__(16) uint8_t crop_smem[];
  float4* s_src = reinterpret_cast<float4*>(crop_smem);                         // [kCropMaxStage]
  float4* s_xw = s_src + kCropMaxStage;                                         // [224] horizontal weights (<= 4 taps)
  int* s_xmn = reinterpret_cast<int*>(s_xw + OUT);                              // [224] first horizontal tap
  float4* s_h = nullptr;  // [nrows][224] horizontal pass: behind the staged band when both fit the 48 KB staging area
  bool fast = false;
  if (!empty) {
    const int S = h > w ? h : w;
    const float scale = static_cast<float>(S) / static_cast<float>(OUT);
    const float support = scale >= 1.0f ? scale : 1.0f;
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const AxisTaps tf = make_taps(blockIdx.x * 8, scale, support, invscale, S);
    const AxisTaps tl = make_taps(blockIdx.x * 8 + 7, scale, support, invscale, S);
    const int r0 = tf.mn, nrows = tl.mn + tl.sz - tf.mn;
    fast = scale <= 1.5f && nrows > 0 && nrows * S <= kCropMaxStage;
    if (fast) {
      const uint8_t* src = pixels + im.offset + static_cast<long long>(y_lo) * im.pitch + static_cast<long long>(x_lo) * 3;
      for (int idx = threadIdx.x; idx < nrows * S; idx += blockDim.x) {
        const int rr = idx / S, xx = idx - rr * S;
        const int y = r0 + rr;
        float4 px = make_float4(1.f, 1.f, 1.f, 0.f);  // white pad right / bottom
        if (y < h && xx < w) {
          const uint8_t* p = src + static_cast<long long>(y) * im.pitch + xx * 3;
          px.x = static_cast<float>(__ldg(p)) * (1.0f / 255.0f);
          px.y = static_cast<float>(__ldg(p + 1)) * (1.0f / 255.0f);
          px.z = static_cast<float>(__ldg(p + 2)) * (1.0f / 255.0f);
        }
        s_src[idx] = px;
      }
      // tables are stored transposed, index (j % 8) * 28 + j / 8: a warp's lanes own consecutive column octets, so
      // for a fixed k = j % 8 they read consecutive entries (16-byte lane stride) instead of a 128-byte stride
      for (int j = threadIdx.x; j < OUT; j += blockDim.x) {
        const AxisTaps tx = make_taps(j, scale, support, invscale, S);
        const int slot = (j & 7) * 28 + (j >> 3);
        s_xmn[slot] = tx.mn;
        s_xw[slot] = make_float4(tx.sz > 0 ? tap_weight(tx, 0, invscale) : 0.f, tx.sz > 1 ? tap_weight(tx, 1, invscale) : 0.f,
                              tx.sz > 2 ? tap_weight(tx, 2, invscale) : 0.f, tx.sz > 3 ? tap_weight(tx, 3, invscale) : 0.f);
      }
      __syncthreads();
      const AxisTaps ty = make_taps(i, scale, support, invscale, S);
      float wys[4];
#pragma unroll
      for (int yy = 0; yy < 4; ++yy) wys[yy] = yy < ty.sz ? tap_weight(ty, yy, invscale) : 0.f;
      const int oct = j0 >> 3;
      if (nrows * (S + OUT) <= kCropMaxStage) {
        s_h = s_src + nrows * S;
        // separable resampling, the order ATen itself uses (horizontal pass, then vertical): every band row is
        // resampled to the 224 output columns once (4 taps) and each output pixel then reads <= 4 of those values,
        // instead of 16 source pixels -- the same FMA sequence per pixel as the nested loop below (zero-weight taps
        // add exactly 0), so the two paths are bit-identical; shared-memory traffic drops ~2.3x (the kernel ran at
        // 97 % L1/shared pipe utilisation).
        {
          const int j = threadIdx.x;  // one output column per thread (blockDim.x == 224)
          const int slot = (j & 7) * 28 + (j >> 3);
          const int xmn = s_xmn[slot];
          const float4 xw = s_xw[slot];
          const float wxs[4] = {xw.x, xw.y, xw.z, xw.w};
          for (int rr = 0; rr < nrows; ++rr) {
            const float4* rowp = s_src + rr * S;
            float hr = 0.f, hg = 0.f, hb = 0.f;
#pragma unroll
            for (int xx = 0; xx < 4; ++xx) {
              const int x = xmn + xx < S ? xmn + xx : S - 1;
              const float4 px = rowp[x];
              hr = fmaf(wxs[xx], px.x, hr); hg = fmaf(wxs[xx], px.y, hg); hb = fmaf(wxs[xx], px.z, hb);
            }
            s_h[rr * OUT + slot] = make_float4(hr, hg, hb, 0.f);
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
          for (int yy = 0; yy < 4; ++yy) {
            if (yy >= ty.sz) break;
            const float4 hv = s_h[(ty.mn - r0 + yy) * OUT + k * 28 + oct];
            r = fmaf(wys[yy], hv.x, r); g = fmaf(wys[yy], hv.y, g); b = fmaf(wys[yy], hv.z, b);
          }
          acc[0][k] = (r - 0.485f) / 0.229f;
          acc[1][k] = (g - 0.456f) / 0.224f;
          acc[2][k] = (b - 0.406f) / 0.225f;
        }
      } else
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int xmn = s_xmn[k * 28 + oct];
        const float4 xw = s_xw[k * 28 + oct];
        const float wxs[4] = {xw.x, xw.y, xw.z, xw.w};
        float r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
          if (yy >= ty.sz) break;
          const float wy = wys[yy];
          const float4* rowp = s_src + (ty.mn - r0 + yy) * S;
          float rr = 0.f, gg = 0.f, bb = 0.f;
#pragma unroll
          for (int xx = 0; xx < 4; ++xx) {
            const int x = xmn + xx < S ? xmn + xx : S - 1;  // weights beyond the tap count are zero
            const float4 px = rowp[x];
            rr = fmaf(wxs[xx], px.x, rr); gg = fmaf(wxs[xx], px.y, gg); bb = fmaf(wxs[xx], px.z, bb);
          }
          r = fmaf(wy, rr, r); g = fmaf(wy, gg, g); b = fmaf(wy, bb, b);
        }
        acc[0][k] = (r - 0.485f) / 0.229f;
        acc[1][k] = (g - 0.456f) / 0.224f;
        acc[2][k] = (b - 0.406f) / 0.225f;
      }
    }
  }

  if (!empty && !fast) {
    const int S = h > w ? h : w;
    const float scale = static_cast<float>(S) / static_cast<float>(OUT);
    const float support = scale >= 1.0f ? scale : 1.0f;
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const uint8_t* src = pixels + im.offset + static_cast<long long>(y_lo) * im.pitch + static_cast<long long>(x_lo) * 3;
    const AxisTaps ty = make_taps(i, scale, support, invscale, S);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const AxisTaps tx = make_taps(j0 + k, scale, support, invscale, S);
      float r = 0.f, g = 0.f, b = 0.f;
      for (int yy = 0; yy < ty.sz; ++yy) {
        const int y = ty.mn + yy;
        const float wy = tap_weight(ty, yy, invscale);
        float rr = 0.f, gg = 0.f, bb = 0.f;
        for (int xx = 0; xx < tx.sz; ++xx) {
          const int x = tx.mn + xx;
          const float wx = tap_weight(tx, xx, invscale);
          float pr = 1.0f, pg = 1.0f, pb = 1.0f;  // white pad right / bottom
          if (y < h && x < w) {
            const uint8_t* p = src + static_cast<long long>(y) * im.pitch + x * 3;
            pr = static_cast<float>(__ldg(p)) * (1.0f / 255.0f);
            pg = static_cast<float>(__ldg(p + 1)) * (1.0f / 255.0f);
            pb = static_cast<float>(__ldg(p + 2)) * (1.0f / 255.0f);
          }
          rr = fmaf(wx, pr, rr); gg = fmaf(wx, pg, gg); bb = fmaf(wx, pb, bb);
        }
        r = fmaf(wy, rr, r); g = fmaf(wy, gg, g); b = fmaf(wy, bb, b);
      }
      acc[0][k] = (r - 0.485f) / 0.229f;
      acc[1][k] = (g - 0.456f) / 0.224f;
      acc[2][k] = (b - 0.406f) / 0.225f;
    }
  }
  if (LAYOUT >= 2) {
    // patch-major fp16 layouts are WHITE-CENTRED: the value stored is (normalised - white level).  A character crop is
    // mostly white padding / background, and the white level (2.2489, 2.4286, 2.6400) is not an fp16 number: its rounding
    // error (up to 8.6e-4, the same sign on every white pixel) adds up coherently in the patch-embedding sums and was
    // 90 % of the encoder's embedding error.  Centred, white is exactly 0; the encoders add W . white to their
    // patch-embedding bias instead (vit.cu / convnext.cu).  An empty rectangle (the reference's all-zero crop) is -white.
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[0][k] -= kWhiteR;
      acc[1][k] -= kWhiteG;
      acc[2][k] -= kWhiteB;
    }
  }
  // the selection below is resolved at compile time, indices are static after unrolling
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (LAYOUT == 1) {
      float* o = reinterpret_cast<float*>(out) + ((static_cast<long long>(n) * 3 + c) * OUT + i) * OUT + j0;
      *reinterpret_cast<float4*>(o) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
    } else if (LAYOUT == 3) {
      // ConvNeXt stem im2col: row = n*3136 + (i/4)*56 + j/4, col = c*16 + (i%4)*4 + j%4; 8 columns = two patches
      const long long row = static_cast<long long>(n) * 3136 + (i >> 2) * 56 + (j0 >> 2);
      __half* o = reinterpret_cast<__half*>(out) + row * 48 + c * 16 + (i & 3) * 4;
      uint2 p0, p1;
      *reinterpret_cast<__half2*>(&p0.x) = __floats2half2_rn(acc[c][0], acc[c][1]);
      *reinterpret_cast<__half2*>(&p0.y) = __floats2half2_rn(acc[c][2], acc[c][3]);
      *reinterpret_cast<__half2*>(&p1.x) = __floats2half2_rn(acc[c][4], acc[c][5]);
      *reinterpret_cast<__half2*>(&p1.y) = __floats2half2_rn(acc[c][6], acc[c][7]);
      *reinterpret_cast<uint2*>(o) = p0;
      *reinterpret_cast<uint2*>(o + 48) = p1;
    } else {
      uint4 pk;
      *reinterpret_cast<__half2*>(&pk.x) = __floats2half2_rn(acc[c][0], acc[c][1]);
      *reinterpret_cast<__half2*>(&pk.y) = __floats2half2_rn(acc[c][2], acc[c][3]);
      *reinterpret_cast<__half2*>(&pk.z) = __floats2half2_rn(acc[c][4], acc[c][5]);
      *reinterpret_cast<__half2*>(&pk.w) = __floats2half2_rn(acc[c][6], acc[c][7]);
      __half* o;
      if (LAYOUT == 0) {
        o = reinterpret_cast<__half*>(out) + ((static_cast<long long>(n) * 3 + c) * OUT + i) * OUT + j0;
      } else {
        const long long row = static_cast<long long>(n) * 196 + (i >> 4) * 14 + (j0 >> 4);
        o = reinterpret_cast<__half*>(out) + row * 768 + c * 256 + (i & 15) * 16 + (j0 & 15);
      }
      *reinterpret_cast<uint4*>(o) = pk;
    }
  }
}

// a1 on the device for the no-resize case of EffLocalizer.letterbox (localizer_engine.py:107-138 with r == 1, e.g.
// 64 x 1024 lines into a 64 x 1024 or larger model shape): place the u8 RGB line at (top, left) of a grey (114) canvas
// and emit the f32 NCHW tensor of load_localizer_img (:80-85: BGR->RGB, / 255).  Bit-exact: no interpolation happens.
__global__ void __launch_bounds__(256) letterbox_pad_kernel(const uint8_t* __restrict__ pixels,
                                                            const effocr_image_desc* __restrict__ images, int n_images,
                                                            int H, int W, float* __restrict__ out) {
  const long long total = static_cast<long long>(n_images) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const int b = static_cast<int>(i / (static_cast<long long>(W) * H));
    const effocr_image_desc im = images[b];
    // dw, dh split exactly like the reference: top = round(dh/2 - 0.1), left = round(dw/2 - 0.1)
    const int dh = H - im.height, dw = W - im.width;
    const int top = static_cast<int>(rintf(dh * 0.5f - 0.1f)), left = static_cast<int>(rintf(dw * 0.5f - 0.1f));
    const int sy = y - top, sx = x - left;
    float r = 114.0f / 255.0f, g = r, bl = r;
    if (sy >= 0 && sy < im.height && sx >= 0 && sx < im.width) {
      const uint8_t* p = pixels + im.offset + static_cast<long long>(sy) * im.pitch + sx * 3;
      r = static_cast<float>(p[0]) / 255.0f;
      g = static_cast<float>(p[1]) / 255.0f;
      bl = static_cast<float>(p[2]) / 255.0f;
    }
    float* o = out + (static_cast<long long>(b) * 3 * H + y) * W + x;
    o[0] = r;
    o[static_cast<long long>(H) * W] = g;
    o[2LL * H * W] = bl;
  }
}

// a1 on the device, general case: EffLocalizer.letterbox (localizer_engine.py:107-138) = cv2.resize(INTER_LINEAR) of the
// u8 line to (nw, nh) followed by grey (114) padding, then load_localizer_img's CHW / RGB / float32 / 255 (:80-85).
// cv2's 8-bit linear resize is fixed-point and therefore reproducible bit for bit (OpenCV imgproc/resize.cpp,
// HResizeLinear + VResizeLinear<uchar,int,short>): horizontal taps a0 + a1 = 2048 (11 bits) applied to u8 pixels,
// vertical taps b0, b1 applied as ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16), then (+2) >> 2.  The tap
// tables (source index pair + coefficient pair per destination column / row) are built on the host with OpenCV's
// own float arithmetic (localizer_engine.letterbox_plan) -- a few KB per distinct line shape.
__global__ void __launch_bounds__(256) letterbox_resize_kernel(const uint8_t* __restrict__ pixels,
                                                               const effocr_image_desc* __restrict__ images,
                                                               const effocr_letterbox_plan* __restrict__ plans,
                                                               const int4* __restrict__ taps, int n_images, int H, int W,
                                                               float* __restrict__ out) {
  const long long total = static_cast<long long>(n_images) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const int b = static_cast<int>(i / (static_cast<long long>(W) * H));
    const effocr_letterbox_plan pl = plans[b];
    const int dy = y - pl.top, dx = x - pl.left;
    float r = 114.0f / 255.0f, g = r, bl = r;
    if (dy >= 0 && dy < pl.new_height && dx >= 0 && dx < pl.new_width) {
      const effocr_image_desc im = images[b];
      const int4 tx = __ldg(taps + pl.xtap_offset + dx);  // (sx0, sx1, a0, a1)
      const int4 ty = __ldg(taps + pl.ytap_offset + dy);  // (sy0, sy1, b0, b1)
      const uint8_t* r0 = pixels + im.offset + static_cast<long long>(ty.x) * im.pitch;
      const uint8_t* r1 = pixels + im.offset + static_cast<long long>(ty.y) * im.pitch;
      int v[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int s0 = r0[tx.x * 3 + c] * tx.z + r0[tx.y * 3 + c] * tx.w;
        const int s1 = r1[tx.x * 3 + c] * tx.z + r1[tx.y * 3 + c] * tx.w;
        int o = (((ty.z * (s0 >> 4)) >> 16) + ((ty.w * (s1 >> 4)) >> 16) + 2) >> 2;
        v[c] = o < 0 ? 0 : (o > 255 ? 255 : o);
      }
      r = __fdiv_rn(static_cast<float>(v[0]), 255.0f);
      g = __fdiv_rn(static_cast<float>(v[1]), 255.0f);
      bl = __fdiv_rn(static_cast<float>(v[2]), 255.0f);
    }
    float* o = out + (static_cast<long long>(b) * 3 * H + y) * W + x;
    o[0] = r;
    o[static_cast<long long>(H) * W] = g;
    o[2LL * H * W] = bl;
  }
}

}  // namespace effocr

using namespace effocr;

extern "C" int effocr_letterbox_resize(const uint8_t* d_pixels, const effocr_image_desc* d_images,
                                       const effocr_letterbox_plan* d_plans, const void* d_taps, int n_images, int height,
                                       int width, float* d_out, void* stream) {
  EFFOCR_TRY(require_sm100());
  if (n_images < 0 || height <= 0 || width <= 0) return fail(EFFOCR_ERR_INVALID, "letterbox_resize: bad arguments");
  if (n_images == 0) return EFFOCR_OK;
  if (!d_pixels || !d_images || !d_plans || !d_taps || !d_out) return fail(EFFOCR_ERR_INVALID, "letterbox_resize: null buffer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(n_images) * height * width;
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  {
    KernelScope ks(PROF_YOLO_MISC, s);
    letterbox_resize_kernel<<<static_cast<int>(g), 256, 0, s>>>(d_pixels, d_images, d_plans, reinterpret_cast<const int4*>(d_taps),
                                                                 n_images, height, width, d_out);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

extern "C" int effocr_letterbox_pad(const uint8_t* d_pixels, const effocr_image_desc* d_images, int n_images, int height,
                                    int width, float* d_out, void* stream) {
  EFFOCR_TRY(require_sm100());
  if (n_images < 0 || height <= 0 || width <= 0) return fail(EFFOCR_ERR_INVALID, "letterbox_pad: bad arguments");
  if (n_images == 0) return EFFOCR_OK;
  if (!d_pixels || !d_images || !d_out) return fail(EFFOCR_ERR_INVALID, "letterbox_pad: null buffer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(n_images) * height * width;
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  {
    KernelScope ks(PROF_YOLO_MISC, s);
    letterbox_pad_kernel<<<static_cast<int>(g), 256, 0, s>>>(d_pixels, d_images, n_images, height, width, d_out);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

extern "C" int effocr_crop_resize(const uint8_t* d_pixels, const effocr_image_desc* d_images,
                                  const effocr_crop_box* d_boxes, int n_boxes, int layout, void* d_out, void* stream) {
  EFFOCR_TRY(require_sm100());
  if (n_boxes < 0) return fail(EFFOCR_ERR_INVALID, "crop_resize: negative box count");
  if (n_boxes == 0) return EFFOCR_OK;
  if (!d_pixels || !d_images || !d_boxes || !d_out) return fail(EFFOCR_ERR_INVALID, "crop_resize: null buffer");
  if (n_boxes > 65535) return fail(EFFOCR_ERR_INVALID, "crop_resize: at most 65535 boxes per call");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(28, n_boxes);
  static bool attr = false;
  if (!attr) {
    EFFOCR_CUDA(cudaFuncSetAttribute(crop_resize_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCropSmemBytes));
    EFFOCR_CUDA(cudaFuncSetAttribute(crop_resize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCropSmemBytes));
    EFFOCR_CUDA(cudaFuncSetAttribute(crop_resize_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCropSmemBytes));
    EFFOCR_CUDA(cudaFuncSetAttribute(crop_resize_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCropSmemBytes));
    attr = true;
  }
  KernelScope ks(PROF_CROP, s);
  switch (layout) {
    case EFFOCR_CROP_NCHW_F16: crop_resize_kernel<0><<<grid, 224, kCropSmemBytes, s>>>(d_pixels, d_images, d_boxes, n_boxes, d_out); break;
    case EFFOCR_CROP_NCHW_F32: crop_resize_kernel<1><<<grid, 224, kCropSmemBytes, s>>>(d_pixels, d_images, d_boxes, n_boxes, d_out); break;
    case EFFOCR_CROP_PATCH_F16: crop_resize_kernel<2><<<grid, 224, kCropSmemBytes, s>>>(d_pixels, d_images, d_boxes, n_boxes, d_out); break;
    case EFFOCR_CROP_PATCH4_F16: crop_resize_kernel<3><<<grid, 224, kCropSmemBytes, s>>>(d_pixels, d_images, d_boxes, n_boxes, d_out); break;
    default: return fail(EFFOCR_ERR_INVALID, "crop_resize: unknown layout");
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}
