// LayerNorm (norm1) + QKV projection in ONE kernel (ViT-S width, K = D = 384):
//
//     qkv[M, N] (fp16)  <-  LayerNorm(x[M, 384]) * gamma + beta  .  Wqkv[N, 384]^T + bias          (timm Block: attn.qkv(norm1(x)))
//
// timm is un-vendored; call site /root/reference/models/encoders.py:58,62-64, restated in oracle/vit.py.  Unfused, norm1 is
// a stand-alone pass that reads the fp32 residual stream (310 MB at batch 1024) and writes the fp16 operand h (155 MB),
// which the QKV GEMM reads back: 81 us per layer of pure HBM time in front of a tensor-bound GEMM.  Here the GEMM's A
// operand is produced on the SM: the kernel is the A-stationary schedule (gemm_tn_astat_kernel: a CTA keeps the
// 128 x 384 A block in shared memory and walks all N tiles against it) with the TMA load of A replaced by eight LayerNorm
// warps.
//
//   while the tensor core works on the current row block, each LN warp reads its 16 rows of the NEXT block (pulled into L2
//   one block earlier by prefetch.global.L2; one warp per row, 3 x float4 per lane, the arithmetic of layernorm_rows_kernel:
//   two-pass mean / variance in registers), normalises them and HOLDS the fp16 result in registers (96 per thread);
//   as the A k-blocks of the current block are released, two at a time, by the MMAs of the last N tile, the held rows are
//   stored straight into the swizzled K-major operand layout (SWIZZLE_128B, 16 lanes = one 128-byte row of one k-block),
//   fence.proxy.async, arrive -- a few hundred cycles per k-block pair, so the tensor core never waits for LayerNorm.
//
// h never exists in HBM; the result is bit-identical to layernorm_rows_kernel + GEMM (same operations per element).
// Warp roles (640 threads): warp 0 TMA producer (weight tiles), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11
// epilogue (bias, fp16, per-warp TMA stores), warps 12-19 LayerNorm.  Registers are re-distributed with setmaxnreg: the
// LayerNorm warps take 144 (96 held + two rows in flight), the epilogue warps 72, warps 0-3 40 (640 x 96 in total).
#pragma once
#include "gemm_sm100_tma_epi.cuh"
#include "ln_math.cuh"

namespace effocr {

#ifndef LNQ_ROWS_IN_FLIGHT
#define LNQ_ROWS_IN_FLIGHT 4
#endif
constexpr int kLnqD = 384;
constexpr int kLnqKB = kLnqD / 64;       // 6 k-blocks
constexpr int kLnqThreads = 640;
constexpr int kLnqLnWarps = 8;
constexpr int kLnqRowsPerWarp = 128 / kLnqLnWarps;  // 16

template <int BLOCK_N>
struct LnQkvCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB per k-block
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kWarpStagingBytes = 32 * 32 * 2;
  static constexpr int kStagingTotal = 8 * 2 * kWarpStagingBytes;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - kBarrierBytes - kStagingTotal - kLnqKB * kABytes) / kBBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kLnqKB * kABytes + kStages * kBBytes + kStagingTotal + kBarrierBytes + 1024;
  static constexpr int kTmemCols = GemmCfg<BLOCK_N>::kTmemCols;
  static_assert(kStages >= 3, "pipeline too shallow");
};

__device__ __forceinline__ float lnq_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kLnqThreads, 1)
ln_gemm_astat_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_c, int M, int N, EpiTmaParams ep) {
  using Cfg = LnQkvCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int KB = kLnqKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + KB * Cfg::kABytes;
  uint8_t* smem_c = smem_b + STAGES * Cfg::kBBytes;
  uint64_t* bfull_bar = reinterpret_cast<uint64_t*>(smem_c + Cfg::kStagingTotal);
  uint64_t* bempty_bar = bfull_bar + STAGES;
  uint64_t* afull_bar = bempty_bar + STAGES;
  uint64_t* aempty_bar = afull_bar + KB;
  uint64_t* tfull_bar = aempty_bar + KB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&bfull_bar[i], 1); mbar_init(&bempty_bar[i], 1); }
    for (int i = 0; i < KB; ++i) { mbar_init(&afull_bar[i], kLnqLnWarps); mbar_init(&aempty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
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

  const int num_m = (M + kBlockM - 1) / kBlockM;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;

  if (warp_idx < 4) setmaxnreg_dec<40>();
  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer: weight tiles only
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int mb = blockIdx.x; mb < num_m; mb += gridDim.x) {
        for (int nb = 0; nb < num_n; ++nb) {
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&bempty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&bfull_bar[stage], Cfg::kBBytes);
            tma_load_2d(&tma_b, &bfull_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBlockK, nb * BLOCK_N);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, predicated issue)
    const bool leader_lane = elect_one_sync();
    constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N);
    const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t a_phase = 0;
    int local = 0;
    for (int mb = blockIdx.x; mb < num_m; mb += gridDim.x, a_phase ^= 1) {
      for (int nb = 0; nb < num_n; ++nb, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < KB; ++kb) {
          if (nb == 0) mbar_wait(&afull_bar[kb], a_phase);
          mbar_wait(&bfull_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(a_base + kb * Cfg::kABytes);
          const uint64_t db = make_sw128_kmajor_desc(b_base + stage * Cfg::kBBytes);
          if (leader_lane) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
            umma_commit(&bempty_bar[stage]);
            if (nb == num_n - 1) umma_commit(&aempty_bar[kb]);  // A k-block free for the next row block
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (leader_lane) umma_commit(&tfull_bar[as]);
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4 && warp_idx < 12) {
    // ------------------------------------------------------------------ epilogue warps (lean: one 32-column chunk in registers)
    setmaxnreg_dec<48>();
    const int q = warp_idx & 3;
    const int half = (warp_idx - 4) >> 2;
    uint8_t* stg = smem_c + (warp_idx - 4) * 2 * Cfg::kWarpStagingBytes;
    constexpr int CHUNKS = BLOCK_N / 64;
    int buf = 0;
    int local = 0;
    for (int mb = blockIdx.x; mb < num_m; mb += gridDim.x) {
      for (int nb = 0; nb < num_n; ++nb, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        const int m0 = mb * kBlockM + q * 32;
        const int n0 = nb * BLOCK_N + half * (BLOCK_N / 2);
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / 2);
        mbar_wait(&tfull_bar[as], aphase);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tbase + c * 32, v);
          tmem_ld_wait();
          if (c == CHUNKS - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
          }
          const int col0 = n0 + c * 32;
          if (col0 >= N) continue;  // warp-uniform
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* dst = stg + buf * Cfg::kWarpStagingBytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (ep.bias && col0 + 8 * j + 8 <= N) {  // N is a multiple of 8
              b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 8 * j));
              b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 8 * j + 4));
            }
            uint4 pk;
            __half2* ph = reinterpret_cast<__half2*>(&pk);
            ph[0] = __floats2half2_rn(__uint_as_float(v[8 * j]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y);
            ph[1] = __floats2half2_rn(__uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w);
            ph[2] = __floats2half2_rn(__uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y);
            ph[3] = __floats2half2_rn(__uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w);
            // 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3)
            *reinterpret_cast<uint4*>(dst + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tma_c, dst, col0, m0);
            tma_store_commit();
          }
          buf ^= 1;
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else if (warp_idx >= 12) {
    // ------------------------------------------------------------------ LayerNorm warps: rows lw*16 .. lw*16+15 of every block
    setmaxnreg_inc<168>();
    const int lw = warp_idx - 12;
    constexpr int R = kLnqRowsPerWarp;
    uint32_t hold[R][6];  // normalised fp16 pairs of this lane's 12 columns of each row
    auto prefetch_l2 = [&](int mb) {  // this warp's rows of block mb: 12 x 128-byte lines per row
      const int row0 = mb * kBlockM + lw * R;
      if (lane < 12) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int row = row0 + r;
          if (row < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + static_cast<long long>(row) * ldx + lane * 32));
        }
      }
    };
    auto load_norm = [&](int mb) {
      const int row0 = mb * kBlockM + lw * R;
#pragma unroll
      for (int r0 = 0; r0 < R; r0 += LNQ_ROWS_IN_FLIGHT) {
        float4 t[LNQ_ROWS_IN_FLIGHT][3];
#pragma unroll
        for (int i = 0; i < LNQ_ROWS_IN_FLIGHT; ++i) {
          const int row = row0 + r0 + i;
          const float* xr = x + static_cast<long long>(row < M ? row : M - 1) * ldx;
#pragma unroll
          for (int j = 0; j < 3; ++j) t[i][j] = *reinterpret_cast<const float4*>(xr + (j * 32 + lane) * 4);
        }
#pragma unroll
        for (int i = 0; i < LNQ_ROWS_IN_FLIGHT; ++i) {
          // layernorm_rows_kernel's arithmetic: sum in column order, butterfly, mean; sum of squared deviations likewise
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) { s += t[i][j].x; s += t[i][j].y; s += t[i][j].z; s += t[i][j].w; }
          const float mu = __fmul_rn(lnq_warp_sum(s), 1.0f / kLnqD);
          float qv = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float d = __fsub_rn(t[i][j].x, mu); qv = __fmaf_rn(d, d, qv);
            d = __fsub_rn(t[i][j].y, mu); qv = __fmaf_rn(d, d, qv);
            d = __fsub_rn(t[i][j].z, mu); qv = __fmaf_rn(d, d, qv);
            d = __fsub_rn(t[i][j].w, mu); qv = __fmaf_rn(d, d, qv);
          }
          const float rs = ln_rstd(lnq_warp_sum(qv), 1.0f / kLnqD, eps);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int c = (j * 32 + lane) * 4;
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
            const __half2 lo = __floats2half2_rn(ln_affine(t[i][j].x, mu, rs, g4.x, b4.x), ln_affine(t[i][j].y, mu, rs, g4.y, b4.y));
            const __half2 hi = __floats2half2_rn(ln_affine(t[i][j].z, mu, rs, g4.z, b4.z), ln_affine(t[i][j].w, mu, rs, g4.w, b4.w));
            hold[r0 + i][2 * j] = *reinterpret_cast<const uint32_t*>(&lo);
            hold[r0 + i][2 * j + 1] = *reinterpret_cast<const uint32_t*>(&hi);
          }
        }
      }
    };
    uint32_t a_phase = 0;
    if (static_cast<int>(blockIdx.x) < num_m) {
      if (static_cast<int>(blockIdx.x + gridDim.x) < num_m) prefetch_l2(blockIdx.x + gridDim.x);
      load_norm(blockIdx.x);
    }
    const int piece = (lane & 15) >> 1, sub = (lane & 1) * 8;
    for (int mb = blockIdx.x; mb < num_m; mb += gridDim.x, a_phase ^= 1) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {  // columns j*128 .. j*128+127 = k-blocks 2j (lanes 0-15) and 2j+1 (lanes 16-31)
        mbar_wait(&aempty_bar[2 * j], a_phase ^ 1);
        mbar_wait(&aempty_bar[2 * j + 1], a_phase ^ 1);
        uint8_t* kb_base = smem_a + (2 * j + (lane >> 4)) * Cfg::kABytes;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const int r = lw * R + i;  // row inside the 128-row block
          // K-major SWIZZLE_128B tile: 16-byte piece p of row r lives at piece p ^ (r & 7)
          *reinterpret_cast<uint2*>(kb_base + r * 128 + ((piece ^ (r & 7)) << 4) + sub) = make_uint2(hold[i][2 * j], hold[i][2 * j + 1]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&afull_bar[2 * j]);
          mbar_arrive(&afull_bar[2 * j + 1]);
        }
      }
      // the next block of this CTA: normalise it under the MMAs of the block just handed over; pull the one after into L2
      const int nxt = mb + static_cast<int>(gridDim.x);
      if (nxt < num_m) {
        if (nxt + static_cast<int>(gridDim.x) < num_m) prefetch_l2(nxt + gridDim.x);
        load_norm(nxt);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace effocr
