// Reference-accurate ("split") kernels of the YOLOv5s localizer.
//
// The reference runs the detector in fp32 (onnxruntime CPU session, /root/reference/onnx_engines/localizer_engine.py:54).
// With fp16 activations and fp16 folded weights the decoded boxes drift by up to ~1 px and confidences by ~6e-3 against
// fp32 (tests/test_gpu_e2e_identity.py measured it) -- enough to flip `conf > conf_thres`, an NMS decision at IoU 0.01 or
// the rounding of a crop rectangle on a few per cent of the lines, i.e. different transcriptions.  The tensor cores still
// do the work: every activation is kept as a PAIR of fp16 planes (hi = fp16(y), lo = fp16(y - hi): 22 mantissa bits),
// every folded weight matrix as [Whi | Wlo], and a convolution accumulates the three products
//        hi x Whi  +  lo x Whi  +  hi x Wlo        (lo x Wlo is below 2^-22 and dropped)
// into ONE fp32 TMEM accumulator -- the K loop simply runs three segments.  The epilogue (bias, SiLU, shortcut) works on
// the fp32 value and writes both planes.  Same trick as the kNN kernel's split-fp16 scores (knn.cu).
//
//   gemm_split3_kernel   1x1 convolutions and the Detect heads (fp32 output): [pixels, K] x [N, K]^T
//   conv3_tc_kernel<.., SPLIT = true>  (conv3_sm100.cuh)  3x3 convolutions, implicit GEMM through 4-D TMA boxes
//   yolo_stem_f32_kernel layer 0 (6x6 s2, 3 -> 32; 2 % of the FLOPs) in plain fp32 FMAs from the f32 image
//   yolo_pool5_split_kernel / plane-wise upsample for SPPF and the FPN
#pragma once
#include "gemm_sm100_tma_epi.cuh"

namespace effocr {

struct Split3Params {
  int M, N, K;        // K = columns of ONE plane; the weight matrix has 2 * K columns ([Whi | Wlo])
  const float* bias;  // [N] or nullptr
};

// fp32 value -> (hi, lo) fp16 pair
__device__ __forceinline__ void split_f32(float y, __half& hi, __half& lo) {
  hi = __float2half_rn(y);
  lo = __float2half_rn(y - __half2float(hi));
}

template <int BLOCK_N, bool OUT_F32, int ACT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_split3_kernel(const __grid_constant__ CUtensorMap tma_a_hi, const __grid_constant__ CUtensorMap tma_a_lo,
                   const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_c_hi,
                   const __grid_constant__ CUtensorMap tma_c_lo, Split3Params p) {
  using Cfg = GemmTmaCfg<BLOCK_N, OUT_F32>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::kABytes;
  uint8_t* smem_c = smem + STAGES * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_c + Cfg::kNumStaging * Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_a_hi);
    tma_prefetch_desc(&tma_a_lo);
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_c_hi);
    if (!OUT_F32) tma_prefetch_desc(&tma_c_lo);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + kBlockM - 1) / kBlockM;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;

  if (warp_idx == 0) {
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * kBlockM;
        const int n0 = (tile % num_n) * BLOCK_N;
        for (int seg = 0; seg < 3; ++seg)
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            // a k-block that runs past K is zero-filled on the A side (the A maps are K columns wide), so whatever
            // the weight box holds there (the other plane, or nothing) contributes nothing
            tma_load_2d(seg == 1 ? &tma_a_lo : &tma_a_hi, &full_bar[stage], smem_a + stage * Cfg::kABytes, kb * kBlockK, m0);
            tma_load_2d(&tma_w, &full_bar[stage], smem_b + stage * Cfg::kBBytes, (seg == 2 ? p.K : 0) + kb * kBlockK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp_idx == 1) {
    const bool leader_lane = elect_one_sync();
    constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N);
    const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < 3 * num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint64_t da = make_sw128_kmajor_desc(a_base + stage * Cfg::kABytes);
        const uint64_t db = make_sw128_kmajor_desc(b_base + stage * Cfg::kBBytes);
        if (leader_lane) {
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader_lane) umma_commit(&tfull_bar[as]);
      __syncwarp();
    }
  } else if (warp_idx >= 4) {
    const int q = warp_idx & 3;
    const int half = (warp_idx - 4) >> 2;
    uint8_t* stg = smem_c + (warp_idx - 4) * 2 * Cfg::kWarpStagingBytes;
    constexpr int CHUNKS = BLOCK_N / 64;
    int buf = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int m0 = (tile / num_n) * kBlockM + q * 32;
      const int n0 = (tile % num_n) * BLOCK_N + half * (BLOCK_N / 2);
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / 2);
      mbar_wait(&tfull_bar[as], aphase);
      tcgen05_fence_after();
      uint32_t v[2][32];
      tmem_ld_32x32b_x32(tbase, v[0]);
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = n0 + c * 32;
        tmem_ld_wait();
        if (c + 1 < CHUNKS) {
          tmem_ld_32x32b_x32(tbase + (c + 1) * 32, v[(c + 1) & 1]);
        } else {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        if (col0 >= p.N) continue;  // warp-uniform
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = __uint_as_float(v[c & 1][i]);
          if (p.bias && col0 + i < p.N) t += __ldg(p.bias + col0 + i);
          if (ACT == ACT_SILU) t = silu(t);
          y[i] = t;
        }
        if constexpr (OUT_F32) {
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* dst = stg + buf * Cfg::kWarpStagingBytes;
#pragma unroll
          for (int j = 0; j < 8; ++j)  // 128-byte rows, SWIZZLE_128B
            *reinterpret_cast<float4*>(dst + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tma_c_hi, dst, col0, m0);
            tma_store_commit();
          }
          buf ^= 1;
        } else {
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            uint8_t* dst = stg + buf * Cfg::kWarpStagingBytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // 64-byte rows, SWIZZLE_64B
              uint4 pk;
              __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                ph[t] = __floats2half2_rn(y[8 * j + 2 * t], y[8 * j + 2 * t + 1]);
                if (pl == 0) {
                  const float2 f = __half22float2(ph[t]);
                  y[8 * j + 2 * t] -= f.x;
                  y[8 * j + 2 * t + 1] -= f.y;
                }
              }
              *reinterpret_cast<uint4*>(dst + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(pl ? &tma_c_lo : &tma_c_hi, dst, col0, m0);
              tma_store_commit();
            }
            buf ^= 1;
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------ layer 0 in fp32
// Conv(3, 32, k6, s2, p2) + folded BN + SiLU from the f32 NCHW image: K = 108, N = 32 -- 2 % of the network's FLOPs.
// The input pixels (k / 255) are not fp16 numbers, so the split scheme would need three mma.sync passes over a
// register-resident weight matrix that already fills the register file; plain fp32 FMAs from shared memory are simpler
// and exact.  One thread = one output pixel x 32 channels; the 20 x 68 x 3 input patch of an 8 x 32 output tile and the
// [108][32] weight matrix (k-major, broadcast reads) live in shared memory.
constexpr int kStemF32TH = 8, kStemF32TW = 64;                                     // output tile: 8 rows x 64 columns
constexpr int kStemF32PH = 2 * kStemF32TH + 4, kStemF32PW = 2 * kStemF32TW + 4;    // input patch 20 x 132
constexpr int kStemF32PWH = kStemF32PW / 2;                                        // 66 columns per parity plane

// One thread = TWO output pixels (columns tx and tx + 32 of the tile) x 32 channels: per tap the eight broadcast
// LDS.128 of the weight row are shared by 64 FMAs (the one-pixel version issued 9 shared loads per 32 FMAs and ran at
// the LSU rate: 2.17 ms of the 7.5 ms forward).  The patch is stored de-interleaved by column parity, so the stride-2
// reads of a warp (output column ox reads input columns 2 ox + kx) hit 32 consecutive words: no bank conflicts.
__global__ void __launch_bounds__(256) yolo_stem_f32_kernel(const float* __restrict__ img, const float* __restrict__ w /*[108][32]*/,
                                                            const float* __restrict__ bias, __half* __restrict__ out_hi,
                                                            __half* __restrict__ out_lo, int ld_out, int B, int H, int W) {
  __shared__ __align__(16) float patch[3][kStemF32PH][2][kStemF32PWH + 1];
  __shared__ __align__(16) float ws[108][32];
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_x = (Wo + kStemF32TW - 1) / kStemF32TW, tiles_y = (Ho + kStemF32TH - 1) / kStemF32TH;
  const int num_tiles = B * tiles_y * tiles_x;
  for (int i = threadIdx.x; i < 108 * 32; i += 256) ws[i / 32][i % 32] = w[i];
  const int ty_l = threadIdx.x >> 5, tx_l = threadIdx.x & 31;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int iy0 = 2 * ty * kStemF32TH - 2, ix0 = 2 * tx * kStemF32TW - 2;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * kStemF32PH * kStemF32PW; i += 256) {
      const int c = i / (kStemF32PH * kStemF32PW), r = (i / kStemF32PW) % kStemF32PH, x = i % kStemF32PW;
      const int iy = iy0 + r, ix = ix0 + x;
      float v = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + ((static_cast<long long>(b) * 3 + c) * H + iy) * W + ix);
      patch[c][r][x & 1][x >> 1] = v;
    }
    __syncthreads();
    float acc[2][32];
#pragma unroll
    for (int n = 0; n < 32; ++n) { acc[0][n] = 0.f; acc[1][n] = 0.f; }
    for (int c = 0; c < 3; ++c)
      for (int ky = 0; ky < 6; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 6; ++kx) {
          // input column 2 ox + kx: parity kx & 1, index ox + kx / 2
          const float x0 = patch[c][2 * ty_l + ky][kx & 1][tx_l + (kx >> 1)];
          const float x1 = patch[c][2 * ty_l + ky][kx & 1][tx_l + 32 + (kx >> 1)];
          const float4* wr = reinterpret_cast<const float4*>(ws[c * 36 + ky * 6 + kx]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 wv = wr[j];
            acc[0][4 * j] = fmaf(x0, wv.x, acc[0][4 * j]);
            acc[0][4 * j + 1] = fmaf(x0, wv.y, acc[0][4 * j + 1]);
            acc[0][4 * j + 2] = fmaf(x0, wv.z, acc[0][4 * j + 2]);
            acc[0][4 * j + 3] = fmaf(x0, wv.w, acc[0][4 * j + 3]);
            acc[1][4 * j] = fmaf(x1, wv.x, acc[1][4 * j]);
            acc[1][4 * j + 1] = fmaf(x1, wv.y, acc[1][4 * j + 1]);
            acc[1][4 * j + 2] = fmaf(x1, wv.z, acc[1][4 * j + 2]);
            acc[1][4 * j + 3] = fmaf(x1, wv.w, acc[1][4 * j + 3]);
          }
        }
      }
    const int oy = ty * kStemF32TH + ty_l;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int ox = tx * kStemF32TW + tx_l + 32 * px;
      if (oy < Ho && ox < Wo) {
        const long long pix = (static_cast<long long>(b) * Ho + oy) * Wo + ox;
        __half hi[32], lo[32];
#pragma unroll
        for (int n = 0; n < 32; ++n) split_f32(silu(acc[px][n] + __ldg(bias + n)), hi[n], lo[n]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(out_hi + pix * ld_out + 8 * j) = *reinterpret_cast<const uint4*>(hi + 8 * j);
          *reinterpret_cast<uint4*>(out_lo + pix * ld_out + 8 * j) = *reinterpret_cast<const uint4*>(lo + 8 * j);
        }
      }
    }
  }
}

// SPPF max-pool on (hi, lo) pairs: fp16 rounding is monotone, so the larger value has the larger hi plane and, among
// equal hi planes, the larger lo plane -- a lexicographic maximum selects the pair of the true maximum.
__global__ void __launch_bounds__(256) yolo_pool5_split_kernel(__half* __restrict__ buf, long long lo_off, int ld, int B, int h,
                                                               int w, int C, int src_off, int dst_off) {
  const int vpt = C / 8;
  const unsigned total = static_cast<unsigned>(B) * h * w * vpt;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned cv = i % vpt;
    const unsigned pix = i / vpt;
    const int x = static_cast<int>(pix % w);
    const int y = static_cast<int>((pix / w) % h);
    const long long img = static_cast<long long>(pix / (static_cast<unsigned>(w) * h)) * h * w;
    float mh[8], ml[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mh[j] = -INFINITY; ml[j] = -INFINITY; }
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        const __half* src = buf + (img + static_cast<long long>(yy) * w + xx) * ld + src_off + cv * 8;
        const uint4 vh = *reinterpret_cast<const uint4*>(src);
        const uint4 vl = *reinterpret_cast<const uint4*>(src + lo_off);
        const __half* hh = reinterpret_cast<const __half*>(&vh);
        const __half* hl = reinterpret_cast<const __half*>(&vl);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = __half2float(hh[j]), bl = __half2float(hl[j]);
          if (a > mh[j] || (a == mh[j] && bl > ml[j])) { mh[j] = a; ml[j] = bl; }
        }
      }
    }
    __half oh[8], ol[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { oh[j] = __float2half_rn(mh[j]); ol[j] = __float2half_rn(ml[j]); }
    __half* dst = buf + static_cast<long long>(pix) * ld + dst_off + cv * 8;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(oh);
    *reinterpret_cast<uint4*>(dst + lo_off) = *reinterpret_cast<const uint4*>(ol);
  }
}

}  // namespace effocr
