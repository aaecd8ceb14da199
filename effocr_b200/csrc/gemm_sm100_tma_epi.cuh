// Second-generation epilogue for the tcgen05 GEMM (gemm_sm100.cuh holds the mainloop description):
// accumulators go TMEM -> registers -> fused op -> swizzled shared-memory staging tile -> TMA store,
// so every global write is a full 128-byte line issued by the TMA engine instead of 32 scattered
// 16-byte stores per warp instruction (the first-round epilogue spent > 3000 L1 wavefront-cycles per
// tile on those, more than the MMAs of the tile).  The fp32 residual update of the ViT stream,
//        x <- x + (A.W^T + bias) * gamma,
// is done by the TMA reduce-add (cp.reduce.async.bulk.tensor ... .add): the read-modify-write of x
// happens in L2 and the SM never loads the residual.
//
// Each of the 8 epilogue warps owns a 32-row band (its TMEM lane quarter) of one column half and
// walks its 32-column chunks through a private pair of staging buffers with private TMA stores, so
// there is no cross-warp barrier in the epilogue; the TMEM load of chunk c+1 is in flight while
// chunk c is being processed.
#pragma once
#include "gemm_sm100.cuh"

namespace effocr {

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- cluster / cta_group::2 helpers (a CTA pair = two SMs of one TPC sharing one 256-row MMA)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> rank 0's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes are credited to the LEADER's mbarrier.
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive on the same barrier offset in BOTH CTAs of the pair once the issued MMAs have retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int BLOCK_N, bool OUT_F32>
struct GemmTmaCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kWarpStagingBytes = 32 * 32 * (OUT_F32 ? 4 : 2);  // one warp's 32 x 32 chunk
  static constexpr int kStagingBytes = 4 * kWarpStagingBytes;            // per 4-warp column half
  static constexpr int kNumStaging = 4;                                  // 8 warps x 2 buffers
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - kBarrierBytes - kNumStaging * kStagingBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kNumStaging * kStagingBytes + kBarrierBytes + 1024;
  static constexpr int kTmemCols = GemmCfg<BLOCK_N>::kTmemCols;
  static_assert(kStages >= 3, "pipeline too shallow");
};

struct EpiTmaParams {
  const float* bias;   // [N] or nullptr
  const float* gamma;  // [N] or nullptr
  // ViT patch embedding (fp32 output only): GEMM row m = image b, patch pi goes to row b * (patches + 1) + 1 + pi of
  // x_out with pos[1 + pi] added.  The row remap is not a TMA box, so the staged chunk leaves through coalesced row
  // copies (four full 128-byte rows per warp instruction) instead of the TMA store.  nullptr = ordinary output.
  const float* pos = nullptr;
  float* x_out = nullptr;
  long long ldx = 0;
  int patches = 0;
};


// Epilogue of ONE output tile for ONE epilogue warp (TMEM lane quarter q, column half `half`).
template <int BLOCK_N, int ACT, bool OUT_F32, bool REDUCE, int NBUF = 2>
__device__ __forceinline__ void tma_epilogue_tile(const CUtensorMap& tma_c, const EpiTmaParams& ep, uint32_t tmem_base,
                                                  uint64_t* tfull_bar, uint64_t* tempty_bar, int as, uint32_t aphase,
                                                  int tile_m0, int tile_n0, int N, int q, int half, int lane,
                                                  uint8_t* stg, int warp_stg_bytes, int& buf,
                                                  bool tempty_on_leader = false, int M_rows = 0x7fffffff) {
  constexpr int CHUNKS = BLOCK_N / 64;  // 32-column chunks per half
  const int m0 = tile_m0 + q * 32;
  const int n0 = tile_n0 + half * (BLOCK_N / 2);
  const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / 2);
  mbar_wait(&tfull_bar[as], aphase);
  tcgen05_fence_after();
  uint32_t v[2][32];
  tmem_ld_32x32b_x32(tbase, v[0]);
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int col0 = n0 + c * 32;
    // bias for this chunk: issued before the TMEM wait so both latencies overlap
    float bv[32];
    const bool full = (col0 + 32 <= N);
    if (ep.bias) {
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i));
          bv[i] = b.x; bv[i + 1] = b.y; bv[i + 2] = b.z; bv[i + 3] = b.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) bv[i] = (col0 + i < N) ? __ldg(ep.bias + col0 + i) : 0.f;
      }
    }
    tmem_ld_wait();
    if (c + 1 < CHUNKS) {
      tmem_ld_32x32b_x32(tbase + (c + 1) * 32, v[(c + 1) & 1]);  // prefetch the next chunk
    } else {
      tcgen05_fence_before();  // accumulator fully read: hand the TMEM buffer back to the MMA warp
      __syncwarp();
      if (lane == 0) {
        if (tempty_on_leader) mbar_arrive_leader(&tempty_bar[as]);  // 2-CTA pair: the issuing CTA owns the barrier
        else mbar_arrive(&tempty_bar[as]);
      }
    }
    float y[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float t = __uint_as_float(v[c & 1][i]);
      if (ep.bias) t += bv[i];
      if (ACT == ACT_GELU) t = gelu_erf(t);
      if (ACT == ACT_SILU) t = silu(t);
      y[i] = t;
    }
    if (ep.gamma) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < N) y[i] *= __ldg(ep.gamma + col0 + i);
    }
    // this warp's staging buffer `buf` was last read by the TMA store it issued NBUF chunks ago
    if (lane == 0) tma_store_wait_read<NBUF - 1>();
    __syncwarp();
    uint8_t* dst = stg + buf * warp_stg_bytes;
    if constexpr (OUT_F32) {
      // 128-byte rows, SWIZZLE_128B: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(dst + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
    } else {
      // 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
        for (int t = 0; t < 4; ++t) ph[t] = __floats2half2_rn(y[8 * j + 2 * t], y[8 * j + 2 * t + 1]);
        *reinterpret_cast<uint4*>(dst + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
      }
    }
    if constexpr (OUT_F32 && !REDUCE) {
      if (ep.pos) {  // warp-uniform: remapped rows + position embedding, coalesced copies out of the staged chunk
        __syncwarp();
        const int piece = lane & 7;
        int grow = m0 + (lane >> 3);              // GEMM row of this lane's first staged row
        int b = grow / ep.patches;
        int pi = grow - b * ep.patches;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const float4 val = *reinterpret_cast<const float4*>(dst + r * 128 + ((piece ^ (r & 7)) << 4));
          if (grow < M_rows && col0 + piece * 4 + 4 <= N) {
            const float4 pp = __ldg(reinterpret_cast<const float4*>(ep.pos + static_cast<long long>(1 + pi) * N + col0 + piece * 4));
            float* o = ep.x_out + (static_cast<long long>(b) * (ep.patches + 1) + 1 + pi) * ep.ldx + col0 + piece * 4;
            *reinterpret_cast<float4*>(o) = make_float4(val.x + pp.x, val.y + pp.y, val.z + pp.z, val.w + pp.w);
          }
          grow += 4;
          pi += 4;
          if (pi >= ep.patches) { pi -= ep.patches; ++b; }
        }
        __syncwarp();  // the staging tile may be rewritten
        buf = (buf + 1 == NBUF) ? 0 : buf + 1;
        continue;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (REDUCE) tma_reduce_add_2d(&tma_c, dst, col0, m0);
      else tma_store_2d(&tma_c, dst, col0, m0);
      tma_store_commit();
    }
    buf = (buf + 1 == NBUF) ? 0 : buf + 1;
  }
}

// Epilogue of one tile for ONE of SIXTEEN epilogue warps (lane quarter q, column quarter cq), 16-column chunks:
// twice the warps of tma_epilogue_tile with half the registers each, for epilogues whose math (GELU) would otherwise
// pace the tile -- four warps per scheduler hide the TMEM-load and MUFU latencies that two cannot.  fp16 output only.
template <int BLOCK_N, int ACT>
__device__ __forceinline__ void tma_epilogue_tile16(const CUtensorMap& tma_c, const EpiTmaParams& ep, uint32_t tmem_base,
                                                    uint64_t* tfull_bar, uint64_t* tempty_bar, int as, uint32_t aphase,
                                                    int tile_m0, int tile_n0, int N, int q, int cq, int lane, uint8_t* stg,
                                                    int& buf) {
  constexpr int CHUNKS = BLOCK_N / 64;  // 16-column chunks per warp
  const int m0 = tile_m0 + q * 32;
  const int n0 = tile_n0 + cq * (BLOCK_N / 4);
  const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + cq * (BLOCK_N / 4);
  mbar_wait(&tfull_bar[as], aphase);
  tcgen05_fence_after();
  uint32_t v[2][16];
  tmem_ld_32x32b_x16(tbase, v[0]);
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int col0 = n0 + c * 16;
    float bv[16];
    if (ep.bias) {
      if (col0 + 16 <= N) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i));
          bv[i] = b.x; bv[i + 1] = b.y; bv[i + 2] = b.z; bv[i + 3] = b.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) bv[i] = (col0 + i < N) ? __ldg(ep.bias + col0 + i) : 0.f;
      }
    }
    tmem_ld_wait();
    if (c + 1 < CHUNKS) {
      tmem_ld_32x32b_x16(tbase + (c + 1) * 16, v[(c + 1) & 1]);
    } else {
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    float y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float t = __uint_as_float(v[c & 1][i]);
      if (ep.bias) t += bv[i];
      if (ACT == ACT_GELU) t = gelu_erf(t);
      if (ACT == ACT_SILU) t = silu(t);
      y[i] = t;
    }
    if (ep.gamma) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (col0 + i < N) y[i] *= __ldg(ep.gamma + col0 + i);
    }
    if (lane == 0) tma_store_wait_read<1>();
    __syncwarp();
    uint8_t* dst = stg + buf * 1024;
    // 32-byte rows, SWIZZLE_32B: 16-byte piece j of row r lives at piece j ^ ((r >> 2) & 1)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint4 pk;
      __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
      for (int t = 0; t < 4; ++t) ph[t] = __floats2half2_rn(y[8 * j + 2 * t], y[8 * j + 2 * t + 1]);
      *reinterpret_cast<uint4*>(dst + lane * 32 + ((j ^ ((lane >> 2) & 1)) << 4)) = pk;
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

// ACT: activation; OUT_F32: output element type; REDUCE: TMA reduce-add into the output (in-place residual)
template <int BLOCK_N, int ACT, bool OUT_F32, bool REDUCE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_tma_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   const __grid_constant__ CUtensorMap tma_c, int M, int N, int K, EpiTmaParams ep) {
  using Cfg = GemmTmaCfg<BLOCK_N, OUT_F32>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::kABytes;
  uint8_t* smem_c = smem + STAGES * Cfg::kStageBytes;  // 1024-aligned: stage sizes are multiples of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_c + Cfg::kNumStaging * Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
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
        {  // pull the A block of this CTA's NEXT tile from HBM into L2 while this tile computes
          const int nt = tile + gridDim.x;
          if (nt < num_tiles && nt / num_n != tile / num_n) {
            const int pm0 = (nt / num_n) * kBlockM;
            for (int kb = 0; kb < num_kb && kb < 24; ++kb) tma_prefetch_l2_2d(&tma_a, kb * kBlockK, pm0);
          }
        }
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
    // The whole warp walks the loop so waits and descriptor arithmetic stay warp-uniform (uniform registers); only the
    // tcgen05 instructions are predicated on one elected lane.  Issuing from inside an `if (elect_one)` region costs
    // a R2UR waterfall of ~16 instructions per MMA (see mlp_sm100.cuh).
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
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader_lane) umma_commit(&tfull_bar[as]);
      __syncwarp();
    }
  } else if (warp_idx >= 4) {
    // Each epilogue warp owns a 32-row band (its TMEM lane quarter q) of one column half and works
    // alone: private double-buffered staging tile, private TMA stores -- no cross-warp barriers.
    const int q = warp_idx & 3;
    const int half = (warp_idx - 4) >> 2;
    uint8_t* stg = smem_c + (warp_idx - 4) * 2 * Cfg::kWarpStagingBytes;
    int buf = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      tma_epilogue_tile<BLOCK_N, ACT, OUT_F32, REDUCE>(tma_c, ep, tmem_base, tfull_bar, tempty_bar, as, aphase,
                                                        (tile / num_n) * kBlockM, (tile % num_n) * BLOCK_N, N, q, half, lane,
                                                        stg, Cfg::kWarpStagingBytes, buf);
    }
    if (lane == 0) tma_store_wait_all<0>();  // global writes complete before the CTA exits
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}


// ------------------------------------------------------------------ A-stationary variant (K <= 384)
// With K = 384 a 128 x BLOCK_N tile re-reads its 96 KB A block for every N tile, and the kernel runs
// at the L2 -> SM bandwidth cap (measured: ~13 TB/s of TMA traffic at 55 % tensor-pipe activity).
// Here a CTA walks ALL N tiles of one 128-row block before moving on, keeping the A block (up to six
// 16 KB k-blocks) resident in shared memory and streaming only the weight tiles, which cuts operand
// traffic per tile from (A + B) to B.  A k-block slot j is released for the next row block as soon as
// the last N tile's MMAs on it retire, so the refill overlaps the tail of the current block.
template <int BLOCK_N, bool OUT_F32>
struct GemmAStatCfg {
  static constexpr int kMaxKB = 6;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kWarpStagingBytes = 32 * 32 * (OUT_F32 ? 4 : 2);
  static constexpr int kNumBuf = OUT_F32 ? 1 : 2;  // fp32 tiles are single-buffered to leave room for the A block
  static constexpr int kStagingTotal = 8 * kNumBuf * kWarpStagingBytes;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - kBarrierBytes - kStagingTotal - kMaxKB * kABytes) / kBBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kMaxKB * kABytes + kStages * kBBytes + kStagingTotal + kBarrierBytes + 1024;
  static constexpr int kTmemCols = GemmCfg<BLOCK_N>::kTmemCols;
  static_assert(kStages >= 3, "pipeline too shallow");
};

template <int BLOCK_N, int ACT, bool OUT_F32, bool REDUCE, int EPI_WARPS = 8>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, 1)
gemm_tn_astat_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_c, int M, int N, int K, EpiTmaParams ep) {
  using Cfg = GemmAStatCfg<BLOCK_N, OUT_F32>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int MAXKB = Cfg::kMaxKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + MAXKB * Cfg::kABytes;
  uint8_t* smem_c = smem_b + STAGES * Cfg::kBBytes;
  uint64_t* bfull_bar = reinterpret_cast<uint64_t*>(smem_c + Cfg::kStagingTotal);
  uint64_t* bempty_bar = bfull_bar + STAGES;
  uint64_t* afull_bar = bempty_bar + STAGES;
  uint64_t* aempty_bar = afull_bar + MAXKB;
  uint64_t* tfull_bar = aempty_bar + MAXKB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&bfull_bar[i], 1); mbar_init(&bempty_bar[i], 1); }
    for (int i = 0; i < MAXKB; ++i) { mbar_init(&afull_bar[i], 1); mbar_init(&aempty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], EPI_WARPS); }
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
  const int num_kb = (K + kBlockK - 1) / kBlockK;  // <= MAXKB (host checks)

  if (warp_idx == 0) {
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t a_phase = 0;  // one A-ring revolution per row block
      for (int mb = blockIdx.x; mb < num_m; mb += gridDim.x, a_phase ^= 1) {
        const int m0 = mb * kBlockM;
        if (mb + static_cast<int>(gridDim.x) < num_m)  // next row block of this CTA: HBM -> L2 now
          for (int kb = 0; kb < num_kb; ++kb) tma_prefetch_l2_2d(&tma_a, kb * kBlockK, m0 + gridDim.x * kBlockM);
        for (int nb = 0; nb < num_n; ++nb) {
          for (int kb = 0; kb < num_kb; ++kb) {
            if (nb == 0) {
              mbar_wait(&aempty_bar[kb], a_phase ^ 1);
              mbar_arrive_expect_tx(&afull_bar[kb], Cfg::kABytes);
              tma_load_2d(&tma_a, &afull_bar[kb], smem_a + kb * Cfg::kABytes, kb * kBlockK, m0);
            }
            mbar_wait(&bempty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&bfull_bar[stage], Cfg::kBBytes);
            tma_load_2d(&tma_b, &bfull_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBlockK, nb * BLOCK_N);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    const bool leader_lane = elect_one_sync();  // warp-uniform loop, predicated tcgen05 issue (see gemm_tn_tma_kernel)
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
        for (int kb = 0; kb < num_kb; ++kb) {
          if (nb == 0) mbar_wait(&afull_bar[kb], a_phase);
          mbar_wait(&bfull_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(a_base + kb * Cfg::kABytes);
          const uint64_t db = make_sw128_kmajor_desc(b_base + stage * Cfg::kBBytes);
          if (leader_lane) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
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
  } else if (warp_idx >= 4 && EPI_WARPS == 16) {
    if constexpr (!OUT_F32 && !REDUCE) {
      const int q = warp_idx & 3;
      const int cq = (warp_idx - 4) >> 2;
      uint8_t* stg = smem_c + (warp_idx - 4) * 2048;
      int buf = 0;
      int local = 0;
      for (int mb = blockIdx.x; mb < num_m; mb += gridDim.x) {
        for (int nb = 0; nb < num_n; ++nb, ++local) {
          tma_epilogue_tile16<BLOCK_N, ACT>(tma_c, ep, tmem_base, tfull_bar, tempty_bar, local & 1, (local >> 1) & 1,
                                            mb * kBlockM, nb * BLOCK_N, N, q, cq, lane, stg, buf);
        }
      }
      if (lane == 0) tma_store_wait_all<0>();
    }
  } else if (warp_idx >= 4) {
    const int q = warp_idx & 3;
    const int half = (warp_idx - 4) >> 2;
    uint8_t* stg = smem_c + (warp_idx - 4) * Cfg::kNumBuf * Cfg::kWarpStagingBytes;
    int buf = 0;
    int local = 0;
    for (int mb = blockIdx.x; mb < num_m; mb += gridDim.x) {
      for (int nb = 0; nb < num_n; ++nb, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        tma_epilogue_tile<BLOCK_N, ACT, OUT_F32, REDUCE, Cfg::kNumBuf>(tma_c, ep, tmem_base, tfull_bar, tempty_bar, as, aphase,
                                                          mb * kBlockM, nb * BLOCK_N, N, q, half, lane, stg,
                                                          Cfg::kWarpStagingBytes, buf);
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}


// ------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) compute one 256 x BLOCK_N tile: each loads its own 128 rows of A
// and HALF of the weight tile, the leader issues tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' shared
// memory, and each CTA drains its own 128 accumulator rows from its own TMEM.  Per CTA and k-block only
// 16 KB + BLOCK_N/2 x 128 B enter shared memory (vs 16 KB + BLOCK_N x 128 B), which halves the B-operand share of
// the smem write + read traffic that bounds the single-CTA kernel at ~55 % tensor-pipe activity, and deepens the
// TMA look-ahead (6-7 stages instead of 4).
template <int BLOCK_N, bool OUT_F32>
struct GemmPairCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;  // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kWarpStagingBytes = 32 * 32 * (OUT_F32 ? 4 : 2);
  static constexpr int kStagingTotal = 16 * kWarpStagingBytes;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - kBarrierBytes - kStagingTotal) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingTotal + kBarrierBytes + 1024;
  static constexpr int kTmemCols = GemmCfg<BLOCK_N>::kTmemCols;
  static_assert(BLOCK_N % 32 == 0 && kStages >= 3, "bad pair configuration");
};

template <int BLOCK_N, int ACT, bool OUT_F32, bool REDUCE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, int M, int N, int K, EpiTmaParams ep) {
  using Cfg = GemmPairCfg<BLOCK_N, OUT_F32>;
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
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's arrive.expect_tx; bytes arrive from both CTAs' TMA loads
      mbar_init(&empty_bar[i], 1);  // one multicast commit per round
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);  // 8 epilogue warps of each CTA arrive on the LEADER's barrier
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tcgen05_fence_before();
  cluster_sync_all();  // barrier inits of both CTAs visible before any remote arrive / multicast commit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (M + 2 * kBlockM - 1) / (2 * kBlockM);
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + kBlockK - 1) / kBlockK;

  if (warp_idx == 0) {
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / num_n) * 2 * kBlockM + rank * kBlockM;
        const int n0 = (tile % num_n) * BLOCK_N + rank * (BLOCK_N / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          tma_load_2d_2sm(&tma_a, &full_bar[stage], smem_a + stage * Cfg::kABytes, kb * kBlockK, m0);
          tma_load_2d_2sm(&tma_b, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBlockK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    if (rank == 0) {
      const bool leader_lane = elect_one_sync();  // warp-uniform loop, predicated tcgen05 issue (see gemm_tn_tma_kernel)
      constexpr uint32_t idesc = make_idesc_f16(2 * kBlockM, BLOCK_N);
      const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
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
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_f16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (leader_lane) umma_commit_2sm(&tfull_bar[as]);
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4) {
    const int q = warp_idx & 3;
    const int half = (warp_idx - 4) >> 2;
    uint8_t* stg = smem_c + (warp_idx - 4) * 2 * Cfg::kWarpStagingBytes;
    int buf = 0;
    int local = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      tma_epilogue_tile<BLOCK_N, ACT, OUT_F32, REDUCE>(tma_c, ep, tmem_base, tfull_bar, tempty_bar, as, aphase,
                                                        (tile / num_n) * 2 * kBlockM + rank * kBlockM,
                                                        (tile % num_n) * BLOCK_N, N, q, half, lane, stg,
                                                        Cfg::kWarpStagingBytes, buf, /*tempty_on_leader=*/true, M);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  cluster_sync_all();  // the peer's barriers / smem stay alive until the leader's last multicast has landed
  if (warp_idx == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

}  // namespace effocr
