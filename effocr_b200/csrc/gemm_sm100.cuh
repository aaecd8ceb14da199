// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//
//     C[M,N] = epilogue( A[M,K] . W[N,K]^T )          fp16 operands, fp32 accumulation in TMEM
//
// A is a row-major activation matrix (tokens / pixels / queries x features) and W a row-major
// weight matrix in torch.nn.Linear layout, so both operands are K-major and are staged by TMA
// (128-byte swizzle) straight into the layout tcgen05.mma reads.  One CTA per SM loops over
// 128 x BLOCK_N output tiles:
//
//     warp 0      TMA producer      (one elected lane)    global -> smem ring, STAGES deep
//     warp 1      MMA issuer        (one elected lane)    smem -> TMEM accumulator (2 buffers)
//     warp 2      TMEM allocator
//     warps 4-11  epilogue          TMEM -> registers -> fused op -> global
//
// The two TMEM accumulator buffers let the epilogue of tile i overlap the MMAs of tile i+1.
// Epilogues are functors (bias / GELU / SiLU / residual / layer-scale / patch-embed row remap),
// which is how the reference's per-layer torch ops (timm Block.forward, yolov5 Conv.forward --
// both outside /root/reference) are fused into the producing GEMM.
#pragma once
#include "sm100_ptx.cuh"

namespace effocr {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 fp16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;
constexpr int kSmemLimit = 227 * 1024;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - kBarrierBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32)    ? 32
                                   : (2 * BLOCK_N <= 64)  ? 64
                                   : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256
                                                          : 512;
  static_assert(BLOCK_N % 64 == 0 && BLOCK_N >= 64 && BLOCK_N <= 256, "BLOCK_N in {64,128,192,256}");
  static_assert(kStages >= 3, "pipeline too shallow");
};

// ------------------------------------------------------------------ epilogue helpers
__device__ __forceinline__ float gelu_erf(float x) {
  // exact-erf GELU (timm nn.GELU default), x * Phi(x), written as
  //     gelu(x) = relu(x) - |x| * Phi(-|x|),     Phi(-u) = 2^P(u),  u = min(|x|, 6)
  // with P a degree-5 fit of log2(0.5 erfc(u / sqrt 2)) weighted by u Phi(-u) (tools/fit_gelu.py).
  // |gelu - exact| <= 8.6e-7 over all x (fp16 output ulp is >= 6e-8 .. 1e-3); ONE MUFU op and
  // 9 FMA/ALU ops per element -- the epilogue shares issue slots with nothing else but must keep
  // up with the tensor pipe (128 x 256 outputs per ~3000 cycles).
  const float a = fabsf(x);
  const float u = fminf(a, 6.0f);
  float p = fmaf(-0.00047329580411314964f, u, 0.007084473501890898f);
  p = fmaf(p, u, -0.05182719975709915f);
  p = fmaf(p, u, -0.4599926173686981f);
  p = fmaf(p, u, -1.1507878303527832f);
  p = fmaf(p, u, -1.000037670135498f);
  return fmaf(-a, ex2_approx(p), fmaxf(x, 0.0f));
}
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2 };

// out[row, col] = act(acc + bias[col]) (* gamma[col]) (+ resid[row, col])
template <int ACT, typename OutT, bool RESID>
struct EpiStore {
  struct Params {
    OutT* out;
    long long ldo;
    const float* bias;   // [N] or nullptr
    const float* gamma;  // [N] layer-scale or nullptr
    const OutT* resid;   // may alias out
    long long ldr;
  };
  __device__ static __forceinline__ void apply(const Params& p, int row, int col0, const uint32_t (&v)[32],
                                               int M, int N) {
    if (row >= M) return;
    float y[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(v[i]);
    const bool full = (col0 + 32 <= N);
    if (full) {
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
          y[i] += b.x; y[i + 1] += b.y; y[i + 2] += b.z; y[i + 3] += b.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (ACT == ACT_GELU) y[i] = gelu_erf(y[i]);
        if (ACT == ACT_SILU) y[i] = silu(y[i]);
      }
      if (p.gamma) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + i));
          y[i] *= g.x; y[i + 1] *= g.y; y[i + 2] *= g.z; y[i + 3] *= g.w;
        }
      }
      if constexpr (sizeof(OutT) == 4) {
        float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col0;
        if (RESID) {
          const float* r = reinterpret_cast<const float*>(p.resid) + static_cast<long long>(row) * p.ldr + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 rv = *reinterpret_cast<const float4*>(r + i);
            y[i] += rv.x; y[i + 1] += rv.y; y[i + 2] += rv.z; y[i + 3] += rv.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(o + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
      } else {
        __half* o = reinterpret_cast<__half*>(p.out) + static_cast<long long>(row) * p.ldo + col0;
        if (RESID) {
          const __half* r = reinterpret_cast<const __half*>(p.resid) + static_cast<long long>(row) * p.ldr + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const uint4 rv = *reinterpret_cast<const uint4*>(r + i);
            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(rh[j]);
              y[i + 2 * j] += f.x; y[i + 2 * j + 1] += f.y;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
          for (int j = 0; j < 4; ++j) ph[j] = __floats2half2_rn(y[i + 2 * j], y[i + 2 * j + 1]);
          *reinterpret_cast<uint4*>(o + i) = pk;
        }
      }
    } else {
      // ragged right edge: scalar, guarded
      for (int i = 0; i < 32; ++i) {
        const int col = col0 + i;
        if (col >= N) break;
        float t = y[i];
        if (p.bias) t += __ldg(p.bias + col);
        if (ACT == ACT_GELU) t = gelu_erf(t);
        if (ACT == ACT_SILU) t = silu(t);
        if (p.gamma) t *= __ldg(p.gamma + col);
        if constexpr (sizeof(OutT) == 4) {
          if (RESID) t += reinterpret_cast<const float*>(p.resid)[static_cast<long long>(row) * p.ldr + col];
          reinterpret_cast<float*>(p.out)[static_cast<long long>(row) * p.ldo + col] = t;
        } else {
          if (RESID) t += __half2float(reinterpret_cast<const __half*>(p.resid)[static_cast<long long>(row) * p.ldr + col]);
          reinterpret_cast<__half*>(p.out)[static_cast<long long>(row) * p.ldo + col] = __float2half_rn(t);
        }
      }
    }
  }
};

// ViT patch embedding: GEMM row r = b * P + p (P patches per image) lands in token row
// b * (P + 1) + 1 + p of the fp32 residual stream, with conv bias and position embedding
// added (timm VisionTransformer._pos_embed).  N must be a multiple of 32.
struct EpiPatchEmbed {
  struct Params {
    float* x;          // [B * (P + 1), N]
    const float* bias; // [N]
    const float* pos;  // [(P + 1), N]
    int patches;       // P
  };
  __device__ static __forceinline__ void apply(const Params& p, int row, int col0, const uint32_t (&v)[32],
                                               int M, int N) {
    if (row >= M || col0 + 32 > N) return;
    const int b = row / p.patches;
    const int pi = row - b * p.patches;
    float* o = p.x + (static_cast<long long>(b) * (p.patches + 1) + 1 + pi) * N + col0;
    const float* pe = p.pos + static_cast<long long>(1 + pi) * N + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
      const float4 pp = __ldg(reinterpret_cast<const float4*>(pe + i));
      float4 r;
      r.x = __uint_as_float(v[i]) + bb.x + pp.x;
      r.y = __uint_as_float(v[i + 1]) + bb.y + pp.y;
      r.z = __uint_as_float(v[i + 2]) + bb.z + pp.z;
      r.w = __uint_as_float(v[i + 3]) + bb.w + pp.w;
      *reinterpret_cast<float4*>(o + i) = r;
    }
  }
};

// ------------------------------------------------------------------ the kernel
template <int BLOCK_N, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M, int N,
               int K, typename Epi::Params ep) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);  // one arrival per epilogue warp
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

  const int num_m = (M + kBlockM - 1) / kBlockM;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + kBlockK - 1) / kBlockK;

  if (warp_idx == 0) {
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * kBlockM;
        const int n0 = (tile % num_n) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(&tma_a, &full_bar[stage], smem_a + stage * Cfg::kABytes, kb * kBlockK, m0);
          tma_load_2d(&tma_b, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBlockK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // warp-uniform loop, tcgen05 issue predicated on one elected lane (descriptors stay in uniform registers)
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
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint64_t da = make_sw128_kmajor_desc(a_base + stage * Cfg::kABytes);
        const uint64_t db = make_sw128_kmajor_desc(b_base + stage * Cfg::kBBytes);
        if (leader_lane) {
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 fp16 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4) units
            umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader_lane) umma_commit(&tfull_bar[as]);
      __syncwarp();
    }
  } else if (warp_idx >= 4) {
    const int q = warp_idx & 3;           // TMEM lane quarter this warp may read
    const int half = (warp_idx - 4) >> 2; // which half of the tile's columns
    constexpr int CHUNKS = BLOCK_N / 64;  // 32-column chunks per half
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int m0 = (tile / num_n) * kBlockM;
      const int n0 = (tile % num_n) * BLOCK_N;
      mbar_wait(&tfull_bar[as], aphase);
      tcgen05_fence_after();
      const int row = m0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int cc = half * CHUNKS + c;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + cc * 32, v);
        tmem_ld_wait();
        Epi::apply(ep, row, n0 + cc * 32, v, M, N);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace effocr
