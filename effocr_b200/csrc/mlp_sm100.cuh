// Fused transformer MLP block for sm_100a (ViT-Ti / ViT-S widths):
//
//     x[M, D] (fp32 residual stream)  +=  GELU( h[M, D] . W1[HID, D]^T + b1 ) . W2[D, HID]^T + b2
//
// i.e. timm Block.forward's `x = x + mlp(norm2(x))` minus the LayerNorm (timm is un-vendored; call site
// /root/reference/models/encoders.py:58,62-64; restated in oracle/vit.py).  The unfused path writes the
// [M, HID] fp16 hidden activations to HBM and reads them back (2 x 620 MB per layer at ViT-S, batch 1024 --
// half of the block's DRAM traffic); here they never leave the SM:
//
//   * a CTA PAIR (cluster 2x1, tcgen05 cta_group::2) owns 256 token rows; each CTA keeps its 128 x D slice of h
//     resident in shared memory for the whole tile (A-stationary) and streams HALF of every weight chunk;
//   * the hidden dimension is walked in 64-wide chunks j:   S_j = h . W1_j^T        (M=256, N=64,  K=D)   -> TMEM
//                                                            P_j = GELU(S_j + b1_j)  16 epilogue warps    -> smem, fp16, swizzled
//                                                            O  += P_j . W2_j^T      (M=256, N=D, K=64)   -> TMEM
//     with S double-buffered in TMEM (D + 2*64 <= 512 columns) and P double-buffered in smem, so GELU of chunk j
//     overlaps the MMAs of S_{j+1} and O += P_{j-1};
//   * O (+ b2) leaves through swizzled staging tiles and TMA reduce-add into x (read-modify-write in L2).
//
// Warp roles (640 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator,
// warps 4-19 epilogue (warp w: TMEM lane quarter w & 3, column quarter (w - 4) >> 2).
// Barriers that gate the leader's MMAs collect arrivals from BOTH CTAs on the leader's copy; barriers signalled by
// tcgen05.commit are multicast to both CTAs.
#pragma once
#include "gemm_sm100_tma_epi.cuh"

namespace effocr {

constexpr int kMlpThreads = 640;
constexpr int kMlpEpiWarps = 16;

template <int D>
struct MlpCfg {
  static_assert(D == 192 || D == 384, "fused MLP: model width 192 or 384 (O accumulator + 2 S buffers must fit 512 TMEM columns)");
  static constexpr int kKB = D / 64;                    // k-blocks of the fc1 contraction
  static constexpr int kABytes = 128 * 64 * 2;          // one k-block of this CTA's 128 rows of h
  static constexpr int kATotal = kKB * kABytes;
  static constexpr int kPBytes = 128 * 64 * 2;          // one 64-wide hidden chunk of this CTA's rows
  static constexpr int kW1KbBytes = 32 * 64 * 2;        // this CTA's 32 of the chunk's 64 W1 rows, one k-block
  static constexpr int kW2SubBytes = 96 * 64 * 2;       // this CTA's 96 of 192 W2 rows (one N = 192 MMA)
  static constexpr int kNSub = D / 192;                 // N = 192 MMAs per K step of O += P . W2^T
  static constexpr int kStageBytes = kKB * kW1KbBytes;  // == kNSub * kW2SubBytes
  static_assert(kStageBytes == kNSub * kW2SubBytes, "W1 and W2 chunk halves share one ring slot size");
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - kBarrierBytes - kATotal - 2 * kPBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kATotal + 2 * kPBytes + kStages * kStageBytes + kBarrierBytes + 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kSCol = D;  // S buffers live right after the O accumulator
  static_assert(kStages >= 3, "weight ring too shallow");
};

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2): one issue slot per two elements in the GELU epilogue
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_erf (gemm_sm100.cuh) on two elements, same polynomial evaluated in n = -min(|x|, 6) (odd coefficients negated):
//   gelu(x) = relu(x) + n * 2^P(-n);   7 issue slots per element instead of 10.5
__device__ __forceinline__ __half2 gelu_erf2(uint64_t acc, uint64_t bias) {
  float x0, x1;
  unpack2(add2(acc, bias), x0, x1);
  const uint64_t n = pack2(fmaxf(-fabsf(x0), -6.0f), fmaxf(-fabsf(x1), -6.0f));
  uint64_t p = fma2(pack2(0.00047329580411314964f, 0.00047329580411314964f), n, pack2(0.007084473501890898f, 0.007084473501890898f));
  p = fma2(p, n, pack2(0.05182719975709915f, 0.05182719975709915f));
  p = fma2(p, n, pack2(-0.4599926173686981f, -0.4599926173686981f));
  p = fma2(p, n, pack2(1.1507878303527832f, 1.1507878303527832f));
  p = fma2(p, n, pack2(-1.000037670135498f, -1.000037670135498f));
  float p0, p1, y0, y1;
  unpack2(p, p0, p1);
  unpack2(fma2(n, pack2(ex2_approx(p0), ex2_approx(p1)), pack2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f))), y0, y1);
  return __floats2half2_rn(y0, y1);
}

// AHEAD: how many chunks the fc1 MMAs run ahead of the fc2 MMAs (the S buffer of chunk j is free again as soon as
// the epilogue has pulled it into registers, so AHEAD = 2 needs no third buffer); it is the latency budget of the
// GELU epilogue in units of one chunk's MMA time.
template <int D, int AHEAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_fused_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w1,
                      const __grid_constant__ CUtensorMap tma_w2, const __grid_constant__ CUtensorMap tma_x, int M,
                      int HID, const float* __restrict__ b1, const float* __restrict__ b2, long long* __restrict__ dbg) {
  using Cfg = MlpCfg<D>;
  // dbg != nullptr: CTA 0 logs clock64() at the pipeline hand-offs of its third tile (tools/mlp_timeline.py)
#define MLP_DBG(slot) do { if (dbg && blockIdx.x == 0 && local == 2) dbg[(slot)] = clock64(); } while (0)
  constexpr int STAGES = Cfg::kStages;
  constexpr int KB = Cfg::kKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_p = smem_a + Cfg::kATotal;
  uint8_t* smem_w = smem_p + 2 * Cfg::kPBytes;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem_w + STAGES * Cfg::kStageBytes);
  uint64_t* wempty = wfull + STAGES;
  uint64_t* afull = wempty + STAGES;
  uint64_t* aempty = afull + KB;
  uint64_t* sfull = aempty + 1;
  uint64_t* sempty = sfull + 2;
  uint64_t* pfull = sempty + 2;
  uint64_t* pempty = pfull + 2;
  uint64_t* ofull = pempty + 2;
  uint64_t* oempty = ofull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(oempty + 1);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = (M + 255) / 256;
  const int NCH = HID / 64;  // even (host checks HID % 128 == 0)

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_w1);
    tma_prefetch_desc(&tma_w2);
    tma_prefetch_desc(&tma_x);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < KB; ++i) mbar_init(&afull[i], 1);
    mbar_init(aempty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sfull[i], 1);
      mbar_init(&sempty[i], 2 * kMlpEpiWarps);
      mbar_init(&pfull[i], 2 * kMlpEpiWarps);
      mbar_init(&pempty[i], 1);
    }
    mbar_init(ofull, 1);
    mbar_init(oempty, kMlpEpiWarps);  // the eight draining warps of each CTA
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        const int m0 = tile * 256 + static_cast<int>(rank) * 128;
        mbar_wait(aempty, (local & 1) ^ 1);  // fc1 MMAs of the previous tile have retired
        for (int kb = 0; kb < KB; ++kb) {
          if (rank == 0) mbar_arrive_expect_tx(&afull[kb], 2 * Cfg::kABytes);
          tma_load_2d_2sm(&tma_a, &afull[kb], smem_a + kb * Cfg::kABytes, kb * 64, m0);
        }
        if (tile + num_pairs < num_tiles)  // this CTA's next row block: HBM -> L2 now
          for (int kb = 0; kb < KB; ++kb) tma_prefetch_l2_2d(&tma_a, kb * 64, m0 + num_pairs * 256);
        for (int j = 0; j < NCH + AHEAD; ++j) {
          if (j < NCH) {  // W1 chunk j: rows j*64 + rank*32 .. +32, all of K
            mbar_wait(&wempty[stage], phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&wfull[stage], 2 * Cfg::kStageBytes);
            uint8_t* dst = smem_w + stage * Cfg::kStageBytes;
            for (int kb = 0; kb < KB; ++kb)
              tma_load_2d_2sm(&tma_w1, &wfull[stage], dst + kb * Cfg::kW1KbBytes, kb * 64, j * 64 + static_cast<int>(rank) * 32);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (j >= AHEAD) {  // W2 chunk j-AHEAD: columns (j-AHEAD)*64 .. +64, rows s*192 + rank*96 .. +96
            mbar_wait(&wempty[stage], phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&wfull[stage], 2 * Cfg::kStageBytes);
            uint8_t* dst = smem_w + stage * Cfg::kStageBytes;
            for (int s = 0; s < Cfg::kNSub; ++s)
              tma_load_2d_2sm(&tma_w2, &wfull[stage], dst + s * Cfg::kW2SubBytes, (j - AHEAD) * 64, s * 192 + static_cast<int>(rank) * 96);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // The whole warp walks the loop (waits, descriptor arithmetic stay warp-uniform, so the compiler keeps them in
    // uniform registers); only the tcgen05 instructions themselves are predicated on one elected lane.  Issuing from
    // inside an `if (elect_one)` region instead costs a R2UR waterfall of ~16 instructions per MMA, more than the
    // 32 cycles an N = 64 MMA runs for.
    if (rank == 0) {
      const bool leader_lane = elect_one_sync();
      constexpr uint32_t idesc1 = make_idesc_f16(256, 64);
      constexpr uint32_t idesc2 = make_idesc_f16(256, 192);
      const uint32_t a_base = smem_u32(smem_a), p_base = smem_u32(smem_p), w_base = smem_u32(smem_w);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        for (int j = 0; j < NCH + AHEAD; ++j) {
          if (j < NCH) {  // S[j & 1] = h . W1_j^T
            const int b = j & 1;
            const uint32_t u = (local * static_cast<uint32_t>(NCH) + j) >> 1;
            mbar_wait(&sempty[b], (u & 1) ^ 1);
            mbar_wait(&wfull[stage], phase);
            if (j == 0) {
#pragma unroll
              for (int kb = 0; kb < KB; ++kb) mbar_wait(&afull[kb], local & 1);
            }
            tcgen05_fence_after();
            if (leader_lane) MLP_DBG(0 * 64 + j);
            const uint32_t tmem_s = tmem_base + Cfg::kSCol + b * 64;
            const uint64_t da0 = make_sw128_kmajor_desc(a_base);
            const uint64_t db0 = make_sw128_kmajor_desc(w_base + stage * Cfg::kStageBytes);
            if (leader_lane) {
#pragma unroll
              for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_2sm(tmem_s, da0 + (kb * Cfg::kABytes >> 4) + 2 * k, db0 + (kb * Cfg::kW1KbBytes >> 4) + 2 * k, idesc1,
                               (kb | k) ? 1u : 0u);
              }
              umma_commit_2sm(&wempty[stage]);
              umma_commit_2sm(&sfull[b]);
              if (j == NCH - 1) umma_commit_2sm(aempty);  // h tile free for the next row block
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (j >= AHEAD) {  // O += P[c & 1] . W2_c^T
            const int c = j - AHEAD;
            const int b = c & 1;
            const uint32_t u = (local * static_cast<uint32_t>(NCH) + c) >> 1;
            if (c == 0) mbar_wait(oempty, (local & 1) ^ 1);  // previous tile's O drained by both CTAs
            mbar_wait(&pfull[b], u & 1);
            mbar_wait(&wfull[stage], phase);
            tcgen05_fence_after();
            if (leader_lane) MLP_DBG(1 * 64 + c);
            const uint64_t dp0 = make_sw128_kmajor_desc(p_base + b * Cfg::kPBytes);
            const uint64_t dw0 = make_sw128_kmajor_desc(w_base + stage * Cfg::kStageBytes);
            if (leader_lane) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int s = 0; s < Cfg::kNSub; ++s)
                  umma_f16_2sm(tmem_base + s * 192, dp0 + 2 * k, dw0 + (s * Cfg::kW2SubBytes >> 4) + 2 * k, idesc2, (c | k) ? 1u : 0u);
              }
              umma_commit_2sm(&wempty[stage]);
              umma_commit_2sm(&pempty[b]);
              if (c == NCH - 1) umma_commit_2sm(ofull);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp_idx & 3;
    const int cq = (warp_idx - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t local = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
      for (int j = 0; j < NCH; ++j) {
        const int b = j & 1;
        const uint32_t u = (local * static_cast<uint32_t>(NCH) + j) >> 1;
        uint64_t bv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(b1 + j * 64 + cq * 16 + 4 * i));
          bv[2 * i] = pack2(t.x, t.y);
          bv[2 * i + 1] = pack2(t.z, t.w);
        }
        mbar_wait(&sfull[b], u & 1);
        tcgen05_fence_after();
        if (warp_idx == 4 && lane == 0) MLP_DBG(2 * 64 + j);
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_lane + Cfg::kSCol + b * 64 + cq * 16, v);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&sempty[b]);  // S buffer back to the MMA issuer
        uint4 pk[2];
        __half2* ph = reinterpret_cast<__half2*>(pk);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          ph[i] = gelu_erf2(pack2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bv[i]);
        if (warp_idx == 4 && lane == 0) MLP_DBG(3 * 64 + j);
        mbar_wait(&pempty[b], (u & 1) ^ 1);  // O += P . W2^T of two chunks ago has read this buffer
        if (warp_idx == 4 && lane == 0) MLP_DBG(4 * 64 + j);
        uint8_t* prow = smem_p + b * Cfg::kPBytes + row * 128;
        // K-major SWIZZLE_128B tile: 16-byte piece c of row r lives at piece c ^ (r & 7)
        *reinterpret_cast<uint4*>(prow + (((2 * cq) ^ (row & 7)) << 4)) = pk[0];
        *reinterpret_cast<uint4*>(prow + (((2 * cq + 1) ^ (row & 7)) << 4)) = pk[1];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&pfull[b]);
        if (warp_idx == 4 && lane == 0) MLP_DBG(5 * 64 + j);
      }
      // ---- output: O + b2 -> TMA reduce-add into x.  Eight warps (column halves 0 / 1 of each lane quarter) drain
      // 32 x 32 fp32 boxes (128-byte rows, 4 KB) through the idle P buffers; the other eight only join the barrier.
      // The kernel's own clock64 timeline (tools/mlp_timeline.py) puts this drain at ~11 000 cycles per tile -- 18 B/clk
      // per SM, the L2 reduction rate with all CTAs draining together -- against ~1 800 cycles per hidden chunk in
      // steady state (MMA-bound).  The rate does not depend on how many CTAs drain (4, 16 or 74 pairs: 9 700 / 9 200 /
      // 10 300 cycles), i.e. it is the per-SM TMA reduction path.  Tried and rejected: 16 warps x 2 KB boxes (12 300
      // cycles), prefetching the residual rows into L2 (no change), residual add in registers with per-thread 64-byte
      // row slices (43 000 cycles: 32 cache lines per load / store instruction), residual add on the SM with the chunk
      // transposed through shared memory so that every global access is a full 128-byte row segment (~17 000 cycles:
      // eight warps of dependent TMEM-load / transpose / load / store steps), half of the columns through
      // red.global.add.v4.f32 issued by the otherwise idle warps in parallel with the TMA reductions (no gain: the two
      // paths share the per-SM reduction rate).
      constexpr int OCH = D / 2 / 32;  // 32-column chunks per draining warp
      const int m_row0 = tile * 256 + static_cast<int>(rank) * 128 + q * 32;
      if (warp_idx == 4 && lane == 0) MLP_DBG(6 * 64 + 0);
      if (cq < 2) {
        uint8_t* stg = smem_p + ((warp_idx - 4) & 7) * 4096;
        mbar_wait(ofull, local & 1);
        tcgen05_fence_after();
        if (warp_idx == 4 && lane == 0) MLP_DBG(6 * 64 + 1);
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < OCH; ++c) {
          const int col0 = cq * (D / 2) + c * 32;
          tmem_ld_32x32b_x32(tmem_lane + col0, v);
          float bv[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(b2 + col0 + i));
            bv[i] = t.x; bv[i + 1] = t.y; bv[i + 2] = t.z; bv[i + 3] = t.w;
          }
          tmem_ld_wait();
          if (c + 1 == OCH) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(oempty);  // O accumulator back to the MMA issuer
          }
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          // 128-byte rows, SWIZZLE_128B: 16-byte piece j of row r lives at piece j ^ (r & 7)
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((jj ^ (lane & 7)) << 4)) =
                make_float4(__uint_as_float(v[4 * jj]) + bv[4 * jj], __uint_as_float(v[4 * jj + 1]) + bv[4 * jj + 1],
                            __uint_as_float(v[4 * jj + 2]) + bv[4 * jj + 2], __uint_as_float(v[4 * jj + 3]) + bv[4 * jj + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tma_x, stg, col0, m_row0);
            tma_store_commit();
          }
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
      if (warp_idx == 4 && lane == 0) MLP_DBG(6 * 64 + 2);
      named_bar_sync(1, kMlpEpiWarps * 32);  // every warp's staging reads are done before P is written again
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp_idx == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
#undef MLP_DBG
}

// ------------------------------------------------------------------ 128-wide fc1 chunks
// Same block, but S = h . W1^T is issued for 128 hidden columns at a time (M = 256, N = 128): at N = 64 every MMA
// re-reads the whole 128 x 16 A slice from shared memory for 32 cycles of math and the tensor core waits for operands
// (960 instead of 768 cycles per 64 columns in the clock64 timeline).  S is single-buffered (D + 128 <= 512 TMEM columns):
// the epilogue pulls both 64-column halves of it into registers at once and hands the buffer back before doing the GELU
// math, while the tensor pipe is busy with the two O += P_half . W2_half^T steps of the previous chunk.  P stays
// double-buffered as two 64-wide halves; the weight ring keeps its 24 KB granularity (4 slots): a W1 chunk (64 rows x 384
// per CTA) is two slots of three k-blocks, a W2 chunk two slots of one 64-column slice each.  D = 384 only (ViT-S/16, the
// benchmarked model); ViT-Tiny uses the 64-wide schedule above.
template <int D>
struct Mlp128Cfg {
  static_assert(D == 384, "128-wide fc1 chunks: model width 384");
  static constexpr int kKB = D / 64;
  static constexpr int kABytes = 128 * 64 * 2;
  static constexpr int kATotal = kKB * kABytes;
  static constexpr int kPBytes = 128 * 64 * 2;
  static constexpr int kW1KbBytes = 64 * 64 * 2;   // this CTA's 64 of the chunk's 128 W1 rows, one k-block: 8 KB
  static constexpr int kW2SubBytes = 96 * 64 * 2;  // this CTA's 96 of 192 W2 rows, 64 hidden columns: 12 KB
  static constexpr int kNSub = D / 192;
  static constexpr int kStageBytes = (kKB / 2) * kW1KbBytes;  // three W1 k-blocks == one W2 column slice: 24 KB
  static_assert(kStageBytes == kNSub * kW2SubBytes, "W1 half chunks and W2 slices share one ring slot size");
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - kBarrierBytes - kATotal - 2 * kPBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 4 ? 4 : kStagesRaw;
  static constexpr int kSmemBytes = kATotal + 2 * kPBytes + kStages * kStageBytes + kBarrierBytes + 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kSCol = D;
  static_assert(kStages >= 2, "weight ring too shallow");
};

template <int D>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_fused_pair128_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w1,
                         const __grid_constant__ CUtensorMap tma_w2, const __grid_constant__ CUtensorMap tma_x, int M,
                         int HID, const float* __restrict__ b1, const float* __restrict__ b2) {
  using Cfg = Mlp128Cfg<D>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int KB = Cfg::kKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_p = smem_a + Cfg::kATotal;
  uint8_t* smem_w = smem_p + 2 * Cfg::kPBytes;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem_w + STAGES * Cfg::kStageBytes);
  uint64_t* wempty = wfull + STAGES;
  uint64_t* afull = wempty + STAGES;
  uint64_t* aempty = afull + KB;
  uint64_t* sfull = aempty + 1;
  uint64_t* sempty = sfull + 1;
  uint64_t* pfull = sempty + 1;   // [2] halves
  uint64_t* pempty = pfull + 2;   // [2]
  uint64_t* ofull = pempty + 2;
  uint64_t* oempty = ofull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(oempty + 1);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = (M + 255) / 256;
  const int NCH = HID / 128;  // 128-wide chunks

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_w1);
    tma_prefetch_desc(&tma_w2);
    tma_prefetch_desc(&tma_x);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < KB; ++i) mbar_init(&afull[i], 1);
    mbar_init(aempty, 1);
    mbar_init(sfull, 1);
    mbar_init(sempty, 2 * kMlpEpiWarps);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pfull[i], 2 * kMlpEpiWarps);
      mbar_init(&pempty[i], 1);
    }
    mbar_init(ofull, 1);
    mbar_init(oempty, kMlpEpiWarps);  // the eight draining warps of each CTA
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        const int m0 = tile * 256 + static_cast<int>(rank) * 128;
        mbar_wait(aempty, (local & 1) ^ 1);  // fc1 MMAs of the previous tile have retired
        for (int kb = 0; kb < KB; ++kb) {
          if (rank == 0) mbar_arrive_expect_tx(&afull[kb], 2 * Cfg::kABytes);
          tma_load_2d_2sm(&tma_a, &afull[kb], smem_a + kb * Cfg::kABytes, kb * 64, m0);
        }
        if (tile + num_pairs < num_tiles)  // this CTA's next row block: HBM -> L2 now
          for (int kb = 0; kb < KB; ++kb) tma_prefetch_l2_2d(&tma_a, kb * 64, m0 + num_pairs * 256);
        for (int j = 0; j <= NCH; ++j) {
          if (j < NCH) {  // W1 chunk j: rows j*128 + rank*64 .. +64; k-blocks 0..2 and 3..5 in two slots
            for (int kh = 0; kh < 2; ++kh) {
              mbar_wait(&wempty[stage], phase ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&wfull[stage], 2 * Cfg::kStageBytes);
              uint8_t* dst = smem_w + stage * Cfg::kStageBytes;
              for (int i = 0; i < KB / 2; ++i)
                tma_load_2d_2sm(&tma_w1, &wfull[stage], dst + i * Cfg::kW1KbBytes, (kh * (KB / 2) + i) * 64,
                                j * 128 + static_cast<int>(rank) * 64);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          if (j >= 1) {  // W2 chunk j-1: one slot per 64-column slice (2(j-1)+h)*64, rows s*192 + rank*96 .. +96
            for (int hh = 0; hh < 2; ++hh) {
              mbar_wait(&wempty[stage], phase ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&wfull[stage], 2 * Cfg::kStageBytes);
              uint8_t* dst = smem_w + stage * Cfg::kStageBytes;
              for (int s = 0; s < Cfg::kNSub; ++s)
                tma_load_2d_2sm(&tma_w2, &wfull[stage], dst + s * Cfg::kW2SubBytes, (2 * (j - 1) + hh) * 64,
                                s * 192 + static_cast<int>(rank) * 96);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform loop)
    if (rank == 0) {
      const bool leader_lane = elect_one_sync();
      constexpr uint32_t idesc1 = make_idesc_f16(256, 128);
      constexpr uint32_t idesc2 = make_idesc_f16(256, 192);
      const uint32_t a_base = smem_u32(smem_a), p_base = smem_u32(smem_p), w_base = smem_u32(smem_w);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        for (int j = 0; j <= NCH; ++j) {
          if (j < NCH) {  // S = h . W1_j^T  (128 hidden columns)
            const uint32_t u = local * static_cast<uint32_t>(NCH) + j;
            mbar_wait(sempty, (u & 1) ^ 1);
            if (j == 0) {
#pragma unroll
              for (int kb = 0; kb < KB; ++kb) mbar_wait(&afull[kb], local & 1);
            }
            const uint32_t tmem_s = tmem_base + Cfg::kSCol;
            const uint64_t da0 = make_sw128_kmajor_desc(a_base);
#pragma unroll
            for (int kh = 0; kh < 2; ++kh) {
              mbar_wait(&wfull[stage], phase);
              tcgen05_fence_after();
              const uint64_t db0 = make_sw128_kmajor_desc(w_base + stage * Cfg::kStageBytes);
              if (leader_lane) {
#pragma unroll
                for (int i = 0; i < KB / 2; ++i) {
                  const int kb = kh * (KB / 2) + i;
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16_2sm(tmem_s, da0 + (kb * Cfg::kABytes >> 4) + 2 * k, db0 + (i * Cfg::kW1KbBytes >> 4) + 2 * k, idesc1,
                                 (kb | k) ? 1u : 0u);
                }
                umma_commit_2sm(&wempty[stage]);
                if (kh == 1) {
                  umma_commit_2sm(sfull);
                  if (j == NCH - 1) umma_commit_2sm(aempty);  // h tile free for the next row block
                }
              }
              __syncwarp();
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          if (j >= 1) {  // O += P_half . W2_half^T for both halves of chunk j-1
            const int c = j - 1;
            const uint32_t u = local * static_cast<uint32_t>(NCH) + c;
            if (c == 0) mbar_wait(oempty, (local & 1) ^ 1);  // previous tile's O drained by both CTAs
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              mbar_wait(&pfull[hh], u & 1);
              mbar_wait(&wfull[stage], phase);
              tcgen05_fence_after();
              const uint64_t dp0 = make_sw128_kmajor_desc(p_base + hh * Cfg::kPBytes);
              const uint32_t wst = w_base + stage * Cfg::kStageBytes;
              if (leader_lane) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                  for (int s = 0; s < Cfg::kNSub; ++s) {
                    const uint64_t dw = make_sw128_kmajor_desc(wst + s * Cfg::kW2SubBytes);
                    umma_f16_2sm(tmem_base + s * 192, dp0 + 2 * k, dw + 2 * k, idesc2, (c | hh | k) ? 1u : 0u);
                  }
                }
                umma_commit_2sm(&pempty[hh]);
                umma_commit_2sm(&wempty[stage]);
                if (c == NCH - 1 && hh == 1) umma_commit_2sm(ofull);
              }
              __syncwarp();
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp_idx & 3;
    const int cq = (warp_idx - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t local = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
      for (int j = 0; j < NCH; ++j) {
        const uint32_t u = local * static_cast<uint32_t>(NCH) + j;
        mbar_wait(sfull, u & 1);
        tcgen05_fence_after();
        uint32_t v[2][16];
        tmem_ld_32x32b_x16(tmem_lane + Cfg::kSCol + cq * 16, v[0]);
        tmem_ld_32x32b_x16(tmem_lane + Cfg::kSCol + 64 + cq * 16, v[1]);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(sempty);  // S buffer back to the MMA issuer: both halves are in registers
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint64_t bv[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(b1 + j * 128 + hh * 64 + cq * 16 + 4 * i));
            bv[2 * i] = pack2(t.x, t.y);
            bv[2 * i + 1] = pack2(t.z, t.w);
          }
          uint4 pk[2];
          __half2* ph = reinterpret_cast<__half2*>(pk);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            ph[i] = gelu_erf2(pack2(__uint_as_float(v[hh][2 * i]), __uint_as_float(v[hh][2 * i + 1])), bv[i]);
          mbar_wait(&pempty[hh], (u & 1) ^ 1);  // O += P . W2^T of the previous chunk has read this half buffer
          uint8_t* prow = smem_p + hh * Cfg::kPBytes + row * 128;
          // K-major SWIZZLE_128B tile: 16-byte piece c of row r lives at piece c ^ (r & 7)
          *reinterpret_cast<uint4*>(prow + (((2 * cq) ^ (row & 7)) << 4)) = pk[0];
          *reinterpret_cast<uint4*>(prow + (((2 * cq + 1) ^ (row & 7)) << 4)) = pk[1];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&pfull[hh]);
        }
      }
      // ---- output: O + b2 -> TMA reduce-add into x (eight warps, 32 x 32 fp32 boxes through the idle P buffers)
      constexpr int OCH = D / 2 / 32;
      const int m_row0 = tile * 256 + static_cast<int>(rank) * 128 + q * 32;
      if (cq < 2) {
        uint8_t* stg = smem_p + ((warp_idx - 4) & 7) * 4096;
        mbar_wait(ofull, local & 1);
        tcgen05_fence_after();
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < OCH; ++c) {
          const int col0 = cq * (D / 2) + c * 32;
          tmem_ld_32x32b_x32(tmem_lane + col0, v);
          float bv[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(b2 + col0 + i));
            bv[i] = t.x; bv[i + 1] = t.y; bv[i + 2] = t.z; bv[i + 3] = t.w;
          }
          tmem_ld_wait();
          if (c + 1 == OCH) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(oempty);  // O accumulator back to the MMA issuer
          }
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((jj ^ (lane & 7)) << 4)) =
                make_float4(__uint_as_float(v[4 * jj]) + bv[4 * jj], __uint_as_float(v[4 * jj + 1]) + bv[4 * jj + 1],
                            __uint_as_float(v[4 * jj + 2]) + bv[4 * jj + 2], __uint_as_float(v[4 * jj + 3]) + bv[4 * jj + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tma_x, stg, col0, m_row0);
            tma_store_commit();
          }
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
      named_bar_sync(1, kMlpEpiWarps * 32);  // every warp's staging reads are done before P is written again
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp_idx == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

}  // namespace effocr
