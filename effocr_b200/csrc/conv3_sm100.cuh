// 3x3 convolution (pad 1, stride 1 or 2) + folded-BN bias + SiLU (+ residual) on NHWC fp16 as an IMPLICIT GEMM:
// no im2col matrix.  YOLOv5 `Conv.forward` with k = 3 (ultralytics/yolov5 is un-vendored; the reference reaches it through
// the ONNX session of /root/reference/onnx_engines/localizer_engine.py:54; restated in oracle/yolo.py).
//
// The first version gathered a [pixels, 9*C] matrix in HBM (written once, read once by the GEMM: 3.5 GB per 64 letterboxed
// lines, 1.4 ms of the 5.3 ms forward pass).  Here the GEMM's A operand is loaded straight from the activation tensor:
// a tile of 128 output pixels is an 8 x 16 patch of one image, and for tap (ky, kx) and channel block cb the A k-block is
// the SAME patch shifted by (ky - 1, kx - 1) (times the stride) -- one 4-D TMA box {CB channels, 16, 8, 1} with traversal
// stride `s` in W and H; out-of-image taps (negative or too large coordinates) are zero-filled by the TMA unit, which is
// exactly the convolution's zero padding.  The box lands in shared memory as 128 rows (pixels, w fastest) of CB fp16 with
// the 128-byte (CB = 64) or 64-byte (CB = 32) swizzle, i.e. directly in the K-major operand layout tcgen05.mma reads.
// Weights are [Cout, 9*C] with column = tap * C + c (the layout the im2col GEMM already used).  The epilogue stores each
// warp's 2 x 16 pixel patch x 32 channels through a 4-D TMA box as well (edge tiles are clipped by the TMA unit), into a
// channel slice of the consumer's concat buffer when asked to.
#pragma once
#include "gemm_sm100_tma_epi.cuh"

namespace effocr {

constexpr int kConvTW = 16, kConvTH = 8;  // 128 output pixels per tile

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1, int32_t c2,
                                             int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// K-major operand tile with 64-byte rows (32 fp16) written by TMA with SWIZZLE_64B: 8-row groups are 512 B apart.
__device__ __forceinline__ uint64_t make_sw64_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;  // SWIZZLE_64B
  return d;
}

template <int BLOCK_N, int CB>
struct Conv3Cfg {
  static_assert(CB == 32 || CB == 64, "channel block 32 or 64");
  static constexpr int kABytes = 128 * CB * 2;
  static constexpr int kBBytes = BLOCK_N * CB * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kWarpStagingBytes = 32 * 32 * 2;
  static constexpr int kStagingTotal = 8 * 2 * kWarpStagingBytes;  // 8 warps x 2 buffers
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - kBarrierBytes - kStagingTotal) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingTotal + kBarrierBytes + 1024;
  static constexpr int kTmemCols = GemmCfg<BLOCK_N>::kTmemCols;
  static_assert(kStages >= 3, "pipeline too shallow");
};

struct Conv3Params {
  int B, Ho, Wo, C, Cout, stride;
  const float* bias;      // [Cout] folded BN bias
  const __half* resid;    // optional residual (same pixel grid as the output), pixel pitch ld_res
  int ld_res;
  long long lo_off;       // SPLIT: element offset from a tensor's hi plane to its lo plane (residual)
  int kdim;               // SPLIT: columns of one weight plane (9 * C); the lo plane follows at column kdim
};

// SPLIT (the localizer's reference-accurate mode, see yolo.cu): activations are (hi, lo) fp16 plane pairs, weights are
// [Whi | Wlo]; the K loop runs three segments -- hi x Whi, lo x Whi, hi x Wlo -- into the same fp32 accumulator, i.e. the
// fp32 product up to 2^-22, and the epilogue splits its fp32 result into the two output planes.  Only the producer's
// choice of tensor map / weight column and the epilogue differ; the MMA loop just runs 3x as many k-blocks.
template <int BLOCK_N, int CB, bool SPLIT = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv3_tc_kernel(const __grid_constant__ CUtensorMap tma_in, const __grid_constant__ CUtensorMap tma_w,
                const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_in_lo,
                const __grid_constant__ CUtensorMap tma_out_lo, Conv3Params p) {
  using Cfg = Conv3Cfg<BLOCK_N, CB>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::kABytes;
  uint8_t* smem_c = smem + STAGES * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_c + Cfg::kStagingTotal);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_in);
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_out);
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

  const int tiles_w = (p.Wo + kConvTW - 1) / kConvTW, tiles_h = (p.Ho + kConvTH - 1) / kConvTH;
  const int num_n = (p.Cout + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = p.B * tiles_h * tiles_w * num_n;
  const int cblocks = p.C / CB;
  const int num_kb = 9 * cblocks;
  constexpr int SEGS = SPLIT ? 3 : 1;

  auto decode = [&](int tile, int& b, int& oh0, int& ow0, int& n0) {
    n0 = (tile % num_n) * BLOCK_N;
    int t = tile / num_n;
    ow0 = (t % tiles_w) * kConvTW;
    t /= tiles_w;
    oh0 = (t % tiles_h) * kConvTH;
    b = t / tiles_h;
  };

  if (warp_idx == 0) {
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int b, oh0, ow0, n0;
        decode(tile, b, oh0, ow0, n0);
        for (int seg = 0; seg < SEGS; ++seg)
          for (int kb = 0; kb < num_kb; ++kb) {
            const int tap = kb / cblocks, cb = kb - tap * cblocks;
            const int ky = tap / 3, kx = tap - ky * 3;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_4d(seg == 1 ? &tma_in_lo : &tma_in, &full_bar[stage], smem_a + stage * Cfg::kABytes, cb * CB,
                        ow0 * p.stride - 1 + kx, oh0 * p.stride - 1 + ky, b);
            tma_load_2d(&tma_w, &full_bar[stage], smem_b + stage * Cfg::kBBytes, (seg == 2 ? p.kdim : 0) + tap * p.C + cb * CB, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp_idx == 1) {
    // warp-uniform loop, tcgen05 issue predicated on one elected lane
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
      for (int kb = 0; kb < SEGS * num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint64_t da = CB == 64 ? make_sw128_kmajor_desc(a_base + stage * Cfg::kABytes)
                                     : make_sw64_kmajor_desc(a_base + stage * Cfg::kABytes);
        const uint64_t db = CB == 64 ? make_sw128_kmajor_desc(b_base + stage * Cfg::kBBytes)
                                     : make_sw64_kmajor_desc(b_base + stage * Cfg::kBBytes);
        if (leader_lane) {
#pragma unroll
          for (int k = 0; k < CB / kUmmaK; ++k) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader_lane) umma_commit(&tfull_bar[as]);
      __syncwarp();
    }
  } else if (warp_idx >= 4) {
    // epilogue warp (q, half): pixels q*32 .. +32 of the tile = output rows oh0 + 2q, +1 (16 columns each), column half `half`
    const int q = warp_idx & 3;
    const int half = (warp_idx - 4) >> 2;
    uint8_t* stg = smem_c + (warp_idx - 4) * 2 * Cfg::kWarpStagingBytes;
    constexpr int CHUNKS = BLOCK_N / 64;  // 32-column chunks per half
    int buf = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      int b, oh0, ow0, n0;
      decode(tile, b, oh0, ow0, n0);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int oh = oh0 + 2 * q + (lane >> 4), ow = ow0 + (lane & 15);  // this lane's pixel
      const bool pix_ok = oh < p.Ho && ow < p.Wo;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / 2);
      mbar_wait(&tfull_bar[as], aphase);
      tcgen05_fence_after();
      uint32_t v[2][32];
      tmem_ld_32x32b_x32(tbase, v[0]);
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = n0 + half * (BLOCK_N / 2) + c * 32;
        tmem_ld_wait();
        if (c + 1 < CHUNKS) {
          tmem_ld_32x32b_x32(tbase + (c + 1) * 32, v[(c + 1) & 1]);
        } else {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        if (col0 >= p.Cout) continue;  // warp-uniform: this chunk lies beyond the layer's output channels
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
          y[i] = silu(__uint_as_float(v[c & 1][i]) + bb.x);
          y[i + 1] = silu(__uint_as_float(v[c & 1][i + 1]) + bb.y);
          y[i + 2] = silu(__uint_as_float(v[c & 1][i + 2]) + bb.z);
          y[i + 3] = silu(__uint_as_float(v[c & 1][i + 3]) + bb.w);
        }
        if (p.resid && pix_ok) {
          const __half* r = p.resid + ((static_cast<long long>(b) * p.Ho + oh) * p.Wo + ow) * p.ld_res + col0;
#pragma unroll
          for (int pl = 0; pl < (SPLIT ? 2 : 1); ++pl) {  // SPLIT: the residual is hi + lo
            const __half* rp = r + (pl ? p.lo_off : 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = *reinterpret_cast<const uint4*>(rp + 8 * j);
              const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __half22float2(hp[t]);
                y[8 * j + 2 * t] += f.x;
                y[8 * j + 2 * t + 1] += f.y;
              }
            }
          }
        }
#pragma unroll
        for (int pl = 0; pl < (SPLIT ? 2 : 1); ++pl) {
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* dst = stg + buf * Cfg::kWarpStagingBytes;
          // 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              ph[t] = __floats2half2_rn(y[8 * j + 2 * t], y[8 * j + 2 * t + 1]);
              if (SPLIT && pl == 0) {  // what the hi plane could not hold goes to the lo plane
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
            tma_store_4d(pl ? &tma_out_lo : &tma_out, dst, col0, ow0, oh0 + 2 * q, b);
            tma_store_commit();
          }
          buf ^= 1;
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace effocr
