// The second half of a ViT-S encoder block in ONE kernel:
//
//     x'  = x + att[M, D] . Wp[D, D]^T + bp                                   (timm Block: x = x + attn(norm1(x)), projection part)
//     x   = x' + GELU( LayerNorm(x') * gamma + beta . W1[HID, D]^T + b1 ) . W2[D, HID]^T + b2          (x = x + mlp(norm2(x)))
//
// timm is un-vendored; call site /root/reference/models/encoders.py:58,62-64, restated in oracle/vit.py.  As two kernels
// (projln_sm100.cuh + mlp_sm100.cuh) the residual stream crosses HBM four times per layer (read + write in each kernel),
// the normalised fp16 operand h is written and read back, and fc2 leaves through TMA reduce-adds: 1 705 MB of DRAM
// traffic per layer at batch 1024.  Here a CTA pair (tcgen05 cta_group::2) owns 256 full rows from the projection to the
// second residual and x is read once and written once (775 MB with att):
//
//   proj     acc[256 x 384] = att . Wp^T      A = the att tile in the 96 KB A region (TMA), Wp k-blocks through the weight ring
//   pass 1   x' = acc + bp + x, x arriving through a six-slot ring of 128 x 32 fp32 boxes that lives IN THE A REGION (the att
//            tile is dead once the projection MMAs have retired); x' goes back into TMEM -- it is the initial value of
//            the fc2 accumulator, so the residual add of the MLP costs nothing and x' never exists in HBM; row statistics
//            as in projln_sm100.cuh (shifted sums per thread, Chan merge over the four column groups)
//   pass 2   LayerNorm(x') -> fp16 -> straight into the A region in the swizzled K-major operand layout: h never exists in HBM
//   mlp      S = h . W1_j^T (128 hidden columns at a time) -> GELU -> P (smem) -> O += P . W2_j^T, O = the x' columns
//   drain    O + b2 -> plain TMA stores into x
//
// The schedule, barriers and warp roles of the mlp phase are those of mlp_fused_pair128_kernel; the projection and the two
// passes are those of proj_ln_pair_kernel.  Warp roles (640 threads): warp 0 TMA producer (att, Wp, W1, W2), warp 1 MMA
// issuer (leader CTA), warp 2 TMEM allocator, warp 3 residual-ring producer, warps 4-19 epilogue (TMEM lane quarter
// w & 3, column group (w - 4) >> 2).
// Numerics: the two kernels' arithmetic except that fc2 accumulates ON TOP of x' in the fp32 TMEM accumulator instead of
// being added to it afterwards, and that the row sums of pass 1 are taken two lanes at a time (packed fp32x2), which flips
// the fp16 rounding of an occasional h element: ~1e-5 relative on x per layer (tests/test_gpu_blocks.py), ~1e-4 on the
// final embedding (tests/test_gpu_recognizer.py::test_three_kernel_layer_equals_separate_kernels), oracle bound 1e-3.
#pragma once
#include "mlp_sm100.cuh"
#include "projln_sm100.cuh"

namespace effocr {

__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

struct BlockTailCfg {
  static constexpr int D = 384;
  static constexpr int kKB = D / 64;
  static constexpr int kABytes = 128 * 64 * 2;     // 16 KB: one k-block of this CTA's 128 rows == one residual-ring slot
  static constexpr int kATotal = kKB * kABytes;    // 96 KB
  static constexpr int kPBytes = 128 * 64 * 2;
  static constexpr int kW1KbBytes = 64 * 64 * 2;   // this CTA's 64 of a chunk's 128 W1 rows, one k-block
  static constexpr int kWSubBytes = 96 * 64 * 2;   // this CTA's 96 of 192 Wp / W2 rows, 64 columns
  static constexpr int kStageBytes = 2 * kWSubBytes;  // 24 KB: one Wp k-block == three W1 k-blocks == one W2 column slice
  static constexpr int kStages = 4;
  static constexpr int kXSlots = 6;
  static constexpr int kStatBytes = 128 * 4 * 2 * 4;  // [row][column group] (mean, M2), in the idle P region
  static constexpr int kBarrierBytes = 512;
  static constexpr int kSmemBytes = kATotal + 2 * kPBytes + kStages * kStageBytes + kBarrierBytes + 1024;
  static constexpr int kSCol = D;
  static_assert(kStageBytes == 3 * kW1KbBytes && kStatBytes <= 2 * kPBytes && kSmemBytes <= kSmemLimit - 1024, "layout");
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
block_tail_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_wp,
                       const __grid_constant__ CUtensorMap tma_w1, const __grid_constant__ CUtensorMap tma_w2,
                       const __grid_constant__ CUtensorMap tma_xl, const __grid_constant__ CUtensorMap tma_xs, int M, int HID,
                       const float* __restrict__ bp, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       const float* __restrict__ b1, const float* __restrict__ b2, int l2_prefetch,
                       int stagger, int reverse, long long* __restrict__ dbg) {
  using Cfg = BlockTailCfg;
  // timeline probe (tools/tail_timeline.py): CTA 0 stamps clock64 into dbg[role * 512 + 8 * tile + slot]
  // (compiled in only with -DEFFOCR_TAIL_TIMELINE: the stamps cost 8 % of the kernel's time through register pressure)
#ifdef EFFOCR_TAIL_TIMELINE
#define BT_STAMP(role, slot) do { if (dbg && blockIdx.x == 0 && local < 60) dbg[(role) * 512 + 8 * local + (slot)] = clock64(); } while (0)
#else
#define BT_STAMP(role, slot) do { } while (0)
#endif
#ifdef EFFOCR_TAIL_TIMELINE
#define BT_WAIT(acc, ...) do { const long long _t = clock64(); __VA_ARGS__; (acc) += clock64() - _t; } while (0)
#else
#define BT_WAIT(acc, ...) do { __VA_ARGS__; } while (0)
#endif
  constexpr int D = Cfg::D, KB = Cfg::kKB, STAGES = Cfg::kStages, XS = Cfg::kXSlots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                      // att tile -> residual ring -> h tile
  uint8_t* smem_p = smem_a + Cfg::kATotal;     // row statistics -> P halves -> output staging
  uint8_t* smem_w = smem_p + 2 * Cfg::kPBytes;
  float* stats = reinterpret_cast<float*>(smem_p);
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem_w + STAGES * Cfg::kStageBytes);
  uint64_t* wempty = wfull + STAGES;
  uint64_t* afull = wempty + STAGES;   // [KB] att k-blocks landed
  uint64_t* aempty = afull + KB;       // fc1 MMAs of the tile retired: the A region takes the next att tile
  uint64_t* tfull1 = aempty + 1;       // projection accumulators complete (and the att tile dead)
  uint64_t* hfull = tfull1 + 1;        // [3] k-blocks 2k, 2k + 1 of h written by the epilogue warps of both CTAs
  uint64_t* sfull = hfull + 3;
  uint64_t* sempty = sfull + 1;
  uint64_t* pfull = sempty + 1;        // [2]
  uint64_t* pempty = pfull + 2;        // [2]
  uint64_t* ofull = pempty + 2;
  uint64_t* oempty = ofull + 1;
  uint64_t* xfull = oempty + 1;        // [12] one per 32-column chunk (a per-slot barrier lets a fast column group ask for a
                                       // slot's second use before its first has landed and read the parity of the use before)
  uint64_t* xempty = xfull + 12;       // [XS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty + XS);
  static_assert((2 * STAGES + KB + 13 + 12 + XS) * 8 + 4 <= Cfg::kBarrierBytes, "barrier area too small");

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = (M + 255) / 256;
  const int NCH = HID / 128;  // 128-wide hidden chunks

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_wp);
    tma_prefetch_desc(&tma_w1);
    tma_prefetch_desc(&tma_w2);
    tma_prefetch_desc(&tma_xl);
    tma_prefetch_desc(&tma_xs);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < KB; ++i) mbar_init(&afull[i], 1);
    mbar_init(aempty, 1);
    mbar_init(tfull1, 1);
    for (int i = 0; i < 3; ++i) mbar_init(&hfull[i], 2 * kMlpEpiWarps);
    mbar_init(sfull, 1);
    mbar_init(sempty, 2 * kMlpEpiWarps);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pfull[i], 2 * kMlpEpiWarps);
      mbar_init(&pempty[i], 1);
    }
    mbar_init(ofull, 1);
    mbar_init(oempty, kMlpEpiWarps);  // the eight draining warps of each CTA
    for (int i = 0; i < 12; ++i) mbar_init(&xfull[i], 1);
    for (int i = 0; i < XS; ++i) mbar_init(&xempty[i], 1);
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
  if (stagger > 0) {
    // Every pair walks the same phases in the same time, so without this all 148 SMs read x (pass 1) and write it (drain)
    // at the same moments and leave HBM idle during the long mlp phase.  `stagger` = estimated cycles per tile.  Pairs
    // that own one tile fewer than the others (pair >= rem) have a tile's worth of slack at the end: their starts are
    // spread over most of a tile; the pairs on the critical path are spread over 15 % of one.
    const int rem = num_tiles % num_pairs;
    long long delay = 0;
    if (num_tiles > num_pairs) {
      if (rem == 0 || pair < rem) delay = static_cast<long long>(stagger) * 15 / 100 * pair / (rem == 0 ? num_pairs : rem);
      else delay = static_cast<long long>(stagger) * (20 + 70 * (pair - rem) / (num_pairs - rem)) / 100;
    }
    const long long t_go = clock64() + delay;
    while (clock64() < t_go) __nanosleep(200);
  }
  // (setmaxnreg re-distribution -- 64 / 104 or 40 / 104 registers for the control / epilogue warps -- was measured 18 % and
  //  95 % SLOWER: ptxas then spills in the MMA issuer and the epilogue alike; every warp keeps the launch value of 96)
  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        const int m0 = (reverse ? num_tiles - 1 - tile : tile) * 256 + static_cast<int>(rank) * 128;
        mbar_wait(aempty, (local & 1) ^ 1);  // fc1 MMAs of the previous tile have retired: the A region is free
        for (int kb = 0; kb < KB; ++kb) {
          if (rank == 0) mbar_arrive_expect_tx(&afull[kb], 2 * Cfg::kABytes);
          tma_load_2d_2sm(&tma_a, &afull[kb], smem_a + kb * Cfg::kABytes, kb * 64, m0);
        }
        if (l2_prefetch & 2)  // THIS tile's residual rows towards L2 while the projection runs (pass 1 needs them next)
          for (int c = 0; c < 12; ++c) tma_prefetch_l2_2d(&tma_xl, c * 32, m0);
        for (int kb = 0; kb < KB; ++kb) {  // Wp k-block kb: rows s*192 + rank*96 .. +96 of both N = 192 halves
          mbar_wait(&wempty[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&wfull[stage], 2 * Cfg::kStageBytes);
          uint8_t* dst = smem_w + stage * Cfg::kStageBytes;
          for (int s = 0; s < 2; ++s)
            tma_load_2d_2sm(&tma_wp, &wfull[stage], dst + s * Cfg::kWSubBytes, kb * 64, s * 192 + static_cast<int>(rank) * 96);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
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
              for (int s = 0; s < 2; ++s)
                tma_load_2d_2sm(&tma_w2, &wfull[stage], dst + s * Cfg::kWSubBytes, (2 * (j - 1) + hh) * 64,
                                s * 192 + static_cast<int>(rank) * 96);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp_idx == 3) {
    // ------------------------------------------------------------------ residual ring producer (per CTA, local barriers)
    if (elect_one_sync()) {
      uint32_t local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        const int m0 = (reverse ? num_tiles - 1 - tile : tile) * 256 + static_cast<int>(rank) * 128;
        mbar_wait(tfull1, local & 1);  // the projection MMAs have read the att tile: its region becomes the ring
        if ((l2_prefetch & 1) && tile + num_pairs < num_tiles) {
          // the NEXT tile's residual rows: HBM -> L2 now, while this tile's long tensor-bound mlp phase leaves HBM idle;
          // its ring loads (on the epilogue's critical path, the tensor pipe waits for them) then pay an L2 hit
          const int pm0 = (tile + num_pairs) * 256 + static_cast<int>(rank) * 128;
          for (int c = 0; c < 12; ++c) tma_prefetch_l2_2d(&tma_xl, c * 32, pm0);
        }
        for (int c = 0; c < 12; ++c) {  // chunk c -> slot c % 6, the slot's use number is 2 * local + c / 6
          const int sl = c % XS;
          const uint32_t u = 2 * local + c / XS;
          mbar_wait(&xempty[sl], (u & 1) ^ 1);
          mbar_arrive_expect_tx(&xfull[c], Cfg::kABytes);
          tma_load_2d(&tma_xl, &xfull[c], smem_a + sl * Cfg::kABytes, c * 32, m0);
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
      long long w_o = 0, w_a = 0, w_w = 0, w_h = 0, w_s = 0, w_p = 0;  // cycles waited per barrier class (timeline build)
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        // ---- projection: acc = att . Wp^T into the O columns (drained by both CTAs for the previous tile)
        if (leader_lane) BT_STAMP(0, 0);
        BT_WAIT(w_o, mbar_wait(oempty, (local & 1) ^ 1));
        tcgen05_fence_after();
        if (leader_lane) BT_STAMP(0, 1);   // O columns drained
        for (int kb = 0; kb < KB; ++kb) {
          BT_WAIT(w_a, mbar_wait(&afull[kb], local & 1));
          BT_WAIT(w_w, mbar_wait(&wfull[stage], phase));
          tcgen05_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(a_base + kb * Cfg::kABytes);
          const uint64_t dw0 = make_sw128_kmajor_desc(w_base + stage * Cfg::kStageBytes);
          const uint64_t dw1 = make_sw128_kmajor_desc(w_base + stage * Cfg::kStageBytes + Cfg::kWSubBytes);
          if (leader_lane) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_2sm(tmem_base, da + 2 * k, dw0 + 2 * k, idesc2, (kb | k) ? 1u : 0u);
              umma_f16_2sm(tmem_base + 192, da + 2 * k, dw1 + 2 * k, idesc2, (kb | k) ? 1u : 0u);
            }
            umma_commit_2sm(&wempty[stage]);
            if (kb == KB - 1) umma_commit_2sm(tfull1);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // ---- mlp: h is in the A region once the epilogue warps of both CTAs have written it; x' sits in the O columns
        if (leader_lane) BT_STAMP(0, 2);   // projection issued
        for (int j = 0; j <= NCH; ++j) {
          if (j < NCH) {  // S = h . W1_j^T  (128 hidden columns)
            const uint32_t u = local * static_cast<uint32_t>(NCH) + j;
            BT_WAIT(w_s, mbar_wait(sempty, (u & 1) ^ 1));
            const uint32_t tmem_s = tmem_base + Cfg::kSCol;
            const uint64_t da0 = make_sw128_kmajor_desc(a_base);
#pragma unroll
            for (int kh = 0; kh < 2; ++kh) {
              BT_WAIT(w_w, mbar_wait(&wfull[stage], phase));
              tcgen05_fence_after();
              const uint64_t db0 = make_sw128_kmajor_desc(w_base + stage * Cfg::kStageBytes);
#pragma unroll
              for (int i = 0; i < KB / 2; ++i) {
                const int kb = kh * (KB / 2) + i;
                if (j == 0 && (kb & 1) == 0) {  // first chunk of the tile: h arrives two k-blocks at a time
                  BT_WAIT(w_h, mbar_wait(&hfull[kb >> 1], local & 1));
                  tcgen05_fence_after();
                  if (leader_lane && kb == 4) BT_STAMP(0, 3);   // h complete
                }
                if (leader_lane) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16_2sm(tmem_s, da0 + (kb * Cfg::kABytes >> 4) + 2 * k, db0 + (i * Cfg::kW1KbBytes >> 4) + 2 * k, idesc1,
                                 (kb | k) ? 1u : 0u);
                }
              }
              if (leader_lane) {
                umma_commit_2sm(&wempty[stage]);
                if (kh == 1) {
                  umma_commit_2sm(sfull);
                  if (j == NCH - 1) umma_commit_2sm(aempty);  // h tile dead: the A region takes the next att tile
                }
              }
              __syncwarp();
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          if (j >= 1) {  // O += P_half . W2_half^T for both halves of chunk j-1, always on top of what O holds (x')
            const int c = j - 1;
            const uint32_t u = local * static_cast<uint32_t>(NCH) + c;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              BT_WAIT(w_p, mbar_wait(&pfull[hh], u & 1));
              BT_WAIT(w_w, mbar_wait(&wfull[stage], phase));
              tcgen05_fence_after();
              const uint64_t dp0 = make_sw128_kmajor_desc(p_base + hh * Cfg::kPBytes);
              const uint32_t wst = w_base + stage * Cfg::kStageBytes;
              if (leader_lane) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                  for (int s = 0; s < 2; ++s) {
                    const uint64_t dw = make_sw128_kmajor_desc(wst + s * Cfg::kWSubBytes);
                    umma_f16_2sm(tmem_base + s * 192, dp0 + 2 * k, dw + 2 * k, idesc2, 1u);
                  }
                }
                umma_commit_2sm(&pempty[hh]);
                umma_commit_2sm(&wempty[stage]);
                if (c == NCH - 1 && hh == 1) { umma_commit_2sm(ofull); BT_STAMP(0, 4); }  // last fc2 MMAs issued
#ifdef EFFOCR_TAIL_TIMELINE
                if (dbg && blockIdx.x == 0 && c == NCH - 1 && hh == 1) {
                  dbg[1000] = w_o; dbg[1001] = w_a; dbg[1002] = w_w; dbg[1003] = w_h; dbg[1004] = w_s; dbg[1005] = w_p; dbg[1006] = local + 1;
                }
#endif
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
    const int cg = (warp_idx - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t local = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
      if (warp_idx == 4 && lane == 0) BT_STAMP(1, 0);
      mbar_wait(tfull1, local & 1);
      tcgen05_fence_after();
      if (warp_idx == 4 && lane == 0) BT_STAMP(1, 1);   // projection complete
      // ---- pass 1: x' = acc + bp + x, back into TMEM; shifted sums of this thread's 96 columns
      // (packed fp32x2 arithmetic: the two passes are issue-bound, sixteen warps share four schedulers)
      float sh = 0.f;
      uint64_t s1p = pack2(0.f, 0.f), s2p = pack2(0.f, 0.f), nsh2 = pack2(0.f, 0.f);
#pragma unroll 1
      for (int k = 0; k < 3; ++k) {
        const int c = k * 4 + cg, col0 = c * 32;
        const int sl = c % XS;
        const uint8_t* xrow = smem_a + sl * Cfg::kABytes + row * 128;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_lane + col0, v);
        mbar_wait(&xfull[c], local & 1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 xo = *reinterpret_cast<const float4*>(xrow + ((j ^ (row & 7)) << 4));  // SWIZZLE_128B
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bp + col0 + 4 * j));
          const uint64_t o01 = add2(pack2(xo.x, xo.y), add2(pack2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), pack2(bb.x, bb.y)));
          const uint64_t o23 = add2(pack2(xo.z, xo.w), add2(pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack2(bb.z, bb.w)));
          float o0, o1, o2, o3;
          unpack2(o01, o0, o1);
          unpack2(o23, o2, o3);
          if (k == 0 && j == 0) { sh = o0; nsh2 = pack2(-sh, -sh); }
          const uint64_t d01 = add2(o01, nsh2), d23 = add2(o23, nsh2);
          s1p = add2(s1p, add2(d01, d23));
          s2p = fma2(d01, d01, s2p);
          s2p = fma2(d23, d23, s2p);
          v[4 * j] = __float_as_uint(o0); v[4 * j + 1] = __float_as_uint(o1);
          v[4 * j + 2] = __float_as_uint(o2); v[4 * j + 3] = __float_as_uint(o3);
        }
        tmem_st_32x32b_x32(tmem_lane + col0, v);  // x': pass 2 reads it, the fc2 MMAs accumulate onto it
        named_bar_sync(1 + cg, 128);  // the four lane quarters of this column group have read the slot
        if (q == 0 && lane == 0) mbar_arrive(&xempty[sl]);
      }
      // ---- row statistics: this thread's 96 columns -> (mean, M2), merged over the four column groups (Chan)
      {
        float s1a, s1b, s2a, s2b;
        unpack2(s1p, s1a, s1b);
        unpack2(s2p, s2a, s2b);
        const float s1 = s1a + s1b, s2 = s2a + s2b;
        const float mean_w = sh + s1 * (1.0f / 96.0f);
        const float m2_w = s2 - s1 * s1 * (1.0f / 96.0f);
        stats[(row * 4 + cg) * 2] = mean_w;
        stats[(row * 4 + cg) * 2 + 1] = m2_w;
      }
      tmem_st_wait();
      // every warp is through pass 1: the statistics are complete and no ring slot is in use, so the A region can take h
      named_bar_sync(5, kMlpEpiWarps * 32);
      if (warp_idx == 4 && lane == 0) BT_STAMP(1, 2);   // pass 1 done in every warp
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
      // ---- pass 2: normalise, fp16, into the A region (K-major SWIZZLE_128B: piece p of row r at piece p ^ (r & 7));
      //      chunk k of every warp completes k-blocks 2k and 2k + 1, handed to the MMA issuer pair by pair
      const uint64_t nmean2 = pack2(-mean, -mean), rstd2 = pack2(rstd, rstd);
#pragma unroll 1
      for (int k = 0; k < 3; ++k) {
        const int col0 = (k * 4 + cg) * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_lane + col0, v);
        tmem_ld_wait();
        uint8_t* hrow = smem_a + (col0 >> 6) * Cfg::kABytes + row * 128;
        const int p0 = ((col0 >> 5) & 1) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk;
          __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int c = 8 * j + 4 * t;
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + col0 + c));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + col0 + c));
            // ((v - mean) * rstd) * gamma + beta, the scalar kernels' operation order, two elements per instruction
            const uint64_t n01 = mul2(add2(pack2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), nmean2), rstd2);
            const uint64_t n23 = mul2(add2(pack2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])), nmean2), rstd2);
            float h0, h1, h2, h3;
            unpack2(fma2(n01, pack2(g4.x, g4.y), pack2(b4.x, b4.y)), h0, h1);
            unpack2(fma2(n23, pack2(g4.z, g4.w), pack2(b4.z, b4.w)), h2, h3);
            ph[2 * t] = __floats2half2_rn(h0, h1);
            ph[2 * t + 1] = __floats2half2_rn(h2, h3);
          }
          *reinterpret_cast<uint4*>(hrow + (((p0 + j) ^ (row & 7)) << 4)) = pk;
        }
        fence_proxy_async_smem();
        if (k == 2) tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&hfull[k]);
      }
      if (warp_idx == 4 && lane == 0) BT_STAMP(1, 3);   // pass 2 done (this warp)
      // ---- mlp: GELU of every S chunk into the P halves
      for (int j = 0; j < NCH; ++j) {
        const uint32_t u = local * static_cast<uint32_t>(NCH) + j;
        mbar_wait(sfull, u & 1);
        tcgen05_fence_after();
        if (j == 0 && warp_idx == 4 && lane == 0) BT_STAMP(1, 4);   // first S chunk complete
        uint32_t v[2][16];
        tmem_ld_32x32b_x16(tmem_lane + Cfg::kSCol + cg * 16, v[0]);
        tmem_ld_32x32b_x16(tmem_lane + Cfg::kSCol + 64 + cg * 16, v[1]);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(sempty);  // S buffer back to the MMA issuer: both halves are in registers
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint64_t bv[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(b1 + j * 128 + hh * 64 + cg * 16 + 4 * i));
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
          *reinterpret_cast<uint4*>(prow + (((2 * cg) ^ (row & 7)) << 4)) = pk[0];
          *reinterpret_cast<uint4*>(prow + (((2 * cg + 1) ^ (row & 7)) << 4)) = pk[1];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&pfull[hh]);
        }
      }
      // ---- drain: O (= x' + fc2) + b2 -> plain TMA stores into x (eight warps, 32 x 32 fp32 boxes through the idle P buffers)
      if (warp_idx == 4 && lane == 0) BT_STAMP(1, 5);   // last GELU chunk written
      constexpr int OCH = D / 2 / 32;
      const int m_row0 = (reverse ? num_tiles - 1 - tile : tile) * 256 + static_cast<int>(rank) * 128 + q * 32;
      if (cg < 2) {
        uint8_t* stg = smem_p + ((warp_idx - 4) & 7) * 4096;
        mbar_wait(ofull, local & 1);
        tcgen05_fence_after();
        if (warp_idx == 4 && lane == 0) BT_STAMP(1, 6);   // fc2 complete
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < OCH; ++c) {
          const int col0 = cg * (D / 2) + c * 32;
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
            if (lane == 0) mbar_arrive_leader(oempty);  // O columns back to the MMA issuer (next tile's projection)
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
            tma_store_2d(&tma_xs, stg, col0, m_row0);
            tma_store_commit();
          }
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
      named_bar_sync(5, kMlpEpiWarps * 32);  // staging reads done before the P region holds the next tile's statistics
      if (warp_idx == 4 && lane == 0) BT_STAMP(1, 7);   // tile drained
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp_idx == 2) tmem_dealloc_2sm(tmem_base, 512);
#undef BT_STAMP
}

}  // namespace effocr
