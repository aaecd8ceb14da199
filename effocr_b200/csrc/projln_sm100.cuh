// Attention output projection + residual + LayerNorm in ONE kernel (ViT-S width):
//
//     x[M, D] (fp32 residual stream)  <-  x + att[M, D] . Wp[D, D]^T + bp            (timm Block: x = x + attn(norm1(x)))
//     h[M, D] (fp16)                  <-  LayerNorm(x) * gamma + beta                 (the norm2 that feeds the MLP block)
//
// timm is un-vendored; call site /root/reference/models/encoders.py:58,62-64, restated in oracle/vit.py.  Unfused, the
// projection GEMM updates x with TMA reduce-adds (718 MB of DRAM traffic, bound by the ~20 B/clk per-SM L2 reduction
// path) and a separate LayerNorm kernel reads x again and writes h (465 MB): 218 us per layer at batch 1024.  Here a
// CTA pair (tcgen05 cta_group::2) computes FULL rows (256 x 384 accumulator: every CTA holds 128 rows x 384 fp32 columns
// in TMEM), so the epilogue can normalise them:
//
//   pass 1  v = acc + bp + x_old, x_old arriving through a six-slot shared-memory ring of 128 x 32 fp32 boxes loaded by
//           TMA (coalesced, asynchronous, six chunks in flight, issued by a dedicated warp); v goes back into the ring slot in place and
//           leaves with a plain TMA store (no L2 reduction), and back into TMEM (tcgen05.st) for pass 2; per-thread
//           shifted sums give mean and M2 of the thread's 96 columns;
//   stats   the four warps that share a row (same TMEM lane quarter, different column groups) merge their partial
//           statistics through shared memory (Chan's parallel variance formula: no cancellation);
//   pass 2  v from TMEM -> (v - mean) * rstd * gamma + beta -> fp16 -> swizzled staging -> TMA store into h.
//
// DRAM traffic: att 155 MB + x 310 MB in, x 310 MB + h 155 MB out = 930 MB instead of 1 183 MB, with no reductions in L2.
// Warp roles (640 threads): warp 0 TMA producer (att / weight k-blocks), warp 1 MMA issuer (leader CTA), warp 2 TMEM
// allocator, warp 3 residual-ring producer, warps 4-19 epilogue (lane quarter w & 3, column group (w - 4) >> 2 which owns
// the 32-column chunks cg, cg + 4, cg + 8; chunk c travels through ring slot c % 6).
#pragma once
#include "gemm_sm100_tma_epi.cuh"

namespace effocr {

constexpr int kPlnThreads = 640;
constexpr int kPlnD = 384;
constexpr int kPlnKB = kPlnD / 64;                      // 6 k-blocks
constexpr int kPlnABytes = 128 * 64 * 2;                // 16 KB: this CTA's rows, one k-block
constexpr int kPlnWSubBytes = 96 * 64 * 2;              // 12 KB: this CTA's 96 of 192 weight rows (one N = 192 MMA)
constexpr int kPlnStageBytes = kPlnABytes + 2 * kPlnWSubBytes;  // 40 KB
constexpr int kPlnStages = 2;  // K = 384 is six k-blocks: two stages in flight cover the short MMA phase
constexpr int kPlnXSlotBytes = 128 * 32 * 4;            // 16 KB: 128 rows x 32 fp32 columns
constexpr int kPlnXSlots = 6;  // half of a tile's twelve 32-column chunks in flight
constexpr int kPlnHStageBytes = 32 * 32 * 2;            // 2 KB per epilogue warp
constexpr int kPlnStatBytes = 128 * 4 * 2 * 4;          // [row][column group] (mean, M2)
constexpr int kPlnSmemBytes = kPlnStages * kPlnStageBytes + kPlnXSlots * kPlnXSlotBytes + 16 * kPlnHStageBytes +
                              kPlnStatBytes + 512 + 1024;

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPlnThreads, 1)
proj_ln_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
                    const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_h, int M,
                    const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float eps, long long* __restrict__ dbg, int prefetch_flags) {
#define PLN_DBG(slot) do { if (dbg && blockIdx.x == 0 && local == 2 && warp_idx == 4 && lane == 0) dbg[(slot)] = clock64(); } while (0)
  constexpr int D = kPlnD, KB = kPlnKB, STAGES = kPlnStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_ring = smem;                                          // [STAGES] x (A | W sub 0 | W sub 1)
  uint8_t* smem_x = smem_ring + STAGES * kPlnStageBytes;              // [4] residual slots
  uint8_t* smem_h = smem_x + kPlnXSlots * kPlnXSlotBytes;             // [16] fp16 staging tiles
  float* stats = reinterpret_cast<float*>(smem_h + 16 * kPlnHStageBytes);  // [128][4][2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + kPlnStatBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull = empty_bar + STAGES;
  uint64_t* tempty = tfull + 1;
  uint64_t* xfull = tempty + 1;                 // [kPlnXSlots]
  uint64_t* xempty = xfull + kPlnXSlots;        // [kPlnXSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty + kPlnXSlots);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = (M + 255) / 256;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_h);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 32);  // 16 epilogue warps of each CTA arrive on the LEADER's barrier
    for (int i = 0; i < kPlnXSlots; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer: att and weight k-blocks (both CTAs)
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = tile * 256 + static_cast<int>(rank) * 128;
        if ((prefetch_flags & 2) && tile + num_pairs < num_tiles) {  // the next tile's att rows into L2 as well
          const int pm0 = (tile + num_pairs) * 256 + static_cast<int>(rank) * 128;
          for (int kb = 0; kb < KB; ++kb) tma_prefetch_l2_2d(&tma_a, kb * 64, pm0);
        }
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kPlnStageBytes);
          uint8_t* dst = smem_ring + stage * kPlnStageBytes;
          tma_load_2d_2sm(&tma_a, &full_bar[stage], dst, kb * 64, m0);
          for (int s = 0; s < 2; ++s)
            tma_load_2d_2sm(&tma_w, &full_bar[stage], dst + kPlnABytes + s * kPlnWSubBytes, kb * 64,
                            s * 192 + static_cast<int>(rank) * 96);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 3) {
    // ------------------------------------------------------------------ residual ring producer (per CTA, local barriers)
    if (elect_one_sync()) {
      uint32_t local = 0;
      const bool l2_prefetch = (prefetch_flags & 1) != 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        const int m0 = tile * 256 + static_cast<int>(rank) * 128;
        if (l2_prefetch && tile + num_pairs < num_tiles) {
          // pull the NEXT tile's residual rows from HBM into L2 now: its ring loads (issued as slots free up, on the
          // epilogue's critical path) then pay an L2 hit instead of an HBM round trip
          const int pm0 = (tile + num_pairs) * 256 + static_cast<int>(rank) * 128;
          for (int c = 0; c < 12; ++c) tma_prefetch_l2_2d(&tma_x, c * 32, pm0);
        }
        for (int c = 0; c < 12; ++c) {  // chunk c -> slot c % 6, the slot's use number is 2 * local + c / 6
          const int sl = c % kPlnXSlots;
          const uint32_t u = 2 * local + c / kPlnXSlots;
          mbar_wait(&xempty[sl], (u & 1) ^ 1);
          mbar_arrive_expect_tx(&xfull[sl], kPlnXSlotBytes);
          tma_load_2d(&tma_x, &xfull[sl], smem_x + sl * kPlnXSlotBytes, c * 32, m0);
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform loop)
    if (rank == 0) {
      const bool leader_lane = elect_one_sync();
      constexpr uint32_t idesc = make_idesc_f16(256, 192);
      const uint32_t ring_base = smem_u32(smem_ring);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        mbar_wait(tempty, (local & 1) ^ 1);  // previous tile's rows have left TMEM (both CTAs)
        tcgen05_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sb = ring_base + stage * kPlnStageBytes;
          const uint64_t da = make_sw128_kmajor_desc(sb);
          const uint64_t dw0 = make_sw128_kmajor_desc(sb + kPlnABytes);
          const uint64_t dw1 = make_sw128_kmajor_desc(sb + kPlnABytes + kPlnWSubBytes);
          if (leader_lane) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_2sm(tmem_base, da + 2 * k, dw0 + 2 * k, idesc, (kb | k) ? 1u : 0u);
              umma_f16_2sm(tmem_base + 192, da + 2 * k, dw1 + 2 * k, idesc, (kb | k) ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == KB - 1) umma_commit_2sm(tfull);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp_idx & 3;
    const int cg = (warp_idx - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* hst = smem_h + (warp_idx - 4) * kPlnHStageBytes;
    uint32_t local = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
      const int m0 = tile * 256 + static_cast<int>(rank) * 128;
      PLN_DBG(0);
      mbar_wait(tfull, local & 1);
      tcgen05_fence_after();
      PLN_DBG(1);
      // ---- pass 1: residual add, x_new out, shifted sums
      float sh = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int k = 0; k < 3; ++k) {
        const int c = k * 4 + cg, col0 = c * 32;
        const int sl = c % kPlnXSlots;
        const uint32_t u = 2 * local + c / kPlnXSlots;
        uint8_t* xslot = smem_x + sl * kPlnXSlotBytes;
        uint8_t* xrow = xslot + row * 128;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_lane + col0, v);
        PLN_DBG(2 + 4 * k);
        mbar_wait(&xfull[sl], u & 1);
        tmem_ld_wait();
        PLN_DBG(3 + 4 * k);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4* px = reinterpret_cast<float4*>(xrow + ((j ^ (row & 7)) << 4));  // SWIZZLE_128B
          const float4 xo = *px;
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col0 + 4 * j));
          float4 o;
          o.x = xo.x + (__uint_as_float(v[4 * j]) + bb.x);
          o.y = xo.y + (__uint_as_float(v[4 * j + 1]) + bb.y);
          o.z = xo.z + (__uint_as_float(v[4 * j + 2]) + bb.z);
          o.w = xo.w + (__uint_as_float(v[4 * j + 3]) + bb.w);
          if (k == 0 && j == 0) sh = o.x;
          const float d0 = o.x - sh, d1 = o.y - sh, d2 = o.z - sh, d3 = o.w - sh;
          s1 += (d0 + d1) + (d2 + d3);
          s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
          *px = o;
          v[4 * j] = __float_as_uint(o.x); v[4 * j + 1] = __float_as_uint(o.y);
          v[4 * j + 2] = __float_as_uint(o.z); v[4 * j + 3] = __float_as_uint(o.w);
        }
        tmem_st_32x32b_x32(tmem_lane + col0, v);  // keep the updated row for pass 2
        fence_proxy_async_smem();
        PLN_DBG(4 + 4 * k);
        named_bar_sync(1 + cg, 128);  // the four lane quarters of this column group have rewritten the slot
        PLN_DBG(5 + 4 * k);
        if (q == 0 && lane == 0) {
          tma_store_2d(&tma_x, xslot, col0, m0);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(&xempty[sl]);  // slot free for the chunk six further
          // (releasing the slot one chunk late -- to take this ~2 000-cycle wait off the column group's path -- was
          //  measured slower: the release chains through the ring and every hop then costs an HBM round trip)
        }
      }
      // ---- row statistics: this thread's 96 columns -> (mean, M2), merged over the four column groups (Chan)
      {
        const float mean_w = sh + s1 * (1.0f / 96.0f);
        const float m2_w = s2 - s1 * s1 * (1.0f / 96.0f);
        stats[(row * 4 + cg) * 2] = mean_w;
        stats[(row * 4 + cg) * 2 + 1] = m2_w;
      }
      PLN_DBG(14);
      tmem_st_wait();
      named_bar_sync(5 + q, 128);  // the four column groups of this lane quarter
      PLN_DBG(15);
      float mean, rstd;
      {
        float mw[4], m2 = 0.f;
        mean = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          mw[i] = stats[(row * 4 + i) * 2];
          m2 += stats[(row * 4 + i) * 2 + 1];
          mean += mw[i];
        }
        mean *= 0.25f;
#pragma unroll
        for (int i = 0; i < 4; ++i) m2 = fmaf(96.0f * (mw[i] - mean), mw[i] - mean, m2);
        rstd = rsqrtf(m2 * (1.0f / D) + eps);
      }
      // ---- pass 2: normalise, fp16, out
#pragma unroll 1
      for (int k = 0; k < 3; ++k) {
        const int col0 = (k * 4 + cg) * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_lane + col0, v);
        tmem_ld_wait();
        if (k == 2) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(tempty);  // accumulator columns free for the next tile's MMAs
        }
        PLN_DBG(16 + 2 * k);
        if (lane == 0) tma_store_wait_read<0>();  // staging tile free (previous chunk's store has read it)
        __syncwarp();
        PLN_DBG(17 + 2 * k);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk;
          __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int c = 8 * j + 4 * t;
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + col0 + c));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + col0 + c));
            ph[2 * t] = __floats2half2_rn(fmaf((__uint_as_float(v[c]) - mean) * rstd, g4.x, b4.x),
                                          fmaf((__uint_as_float(v[c + 1]) - mean) * rstd, g4.y, b4.y));
            ph[2 * t + 1] = __floats2half2_rn(fmaf((__uint_as_float(v[c + 2]) - mean) * rstd, g4.z, b4.z),
                                              fmaf((__uint_as_float(v[c + 3]) - mean) * rstd, g4.w, b4.w));
          }
          // 64-byte rows, SWIZZLE_64B: 16-byte piece j of row r lives at piece j ^ ((r >> 1) & 3)
          *reinterpret_cast<uint4*>(hst + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tma_h, hst, col0, m0 + q * 32);
          tma_store_commit();
        }
      }
      PLN_DBG(22);
      named_bar_sync(5 + q, 128);  // stats of this tile are consumed before the next tile overwrites them
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp_idx == 2) tmem_dealloc_2sm(tmem_base, 512);
#undef PLN_DBG
}

}  // namespace effocr
