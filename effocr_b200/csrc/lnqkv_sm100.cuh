// LayerNorm (norm1) + QKV projection in ONE kernel (ViT-S width, K = D = 384):
//
//     qkv[M, N] (fp16)  <-  LayerNorm(x[M, 384]) * gamma + beta  .  Wqkv[N, 384]^T + bias          (timm Block: attn.qkv(norm1(x)))
//
// timm is un-vendored; call site /root/reference/models/encoders.py:58,62-64, restated in oracle/vit.py.  Unfused, norm1 is
// a stand-alone pass that reads the fp32 residual stream (310 MB at batch 1024) and writes the fp16 operand h (155 MB),
// which the QKV GEMM reads back: 80 us per layer of pure HBM time in front of a tensor-bound GEMM.  Here the GEMM's A
// operand is produced on the SM and h never exists in HBM.
//
// Schedule: A-stationary CTA PAIR (cta_group::2).  The two CTAs of a cluster own 256 consecutive rows, 128 each; a CTA
// keeps its normalised 128 x 384 fp16 block in shared memory (six SWIZZLE_128B K-major k-blocks, 96 KB) and the pair
// walks all N tiles against it with M = 256 MMAs issued by the leader CTA.  Each CTA streams only HALF of every weight
// tile (BLOCK_N / 2 rows x 64 columns, 12 KB), which is what makes the schedule feasible: a single CTA walking the whole
// 885 KB weight matrix per 128 rows asks its SM for ~61 bytes per clock from L2 with three 24 KB stages in flight, and the
// MMA issuer then sat on the weight barrier for half of every tile (clock64 timeline, tools/lnq_timeline.py).
//
//   LayerNorm warps (8 per CTA, 16 rows each): the fp32 rows of the NEXT block stream through per-warp shared-memory slots
//   (1-D bulk copies, kLnqXSlots rows ahead, running on across blocks); two rows at a time go through the arithmetic of
//   layernorm_rows_kernel (two-pass mean / variance, one warp per row, 3 x float4 per lane) and the fp16 result is HELD in
//   registers (96 per thread).  As the A k-blocks of the current block are released, two at a time, by the MMAs of its
//   last N tile, the held rows are stored straight into the swizzled operand layout, fence.proxy.async, arrive on the
//   LEADER's barrier (a plain remote arrive: mbarrier.arrive.release.cluster compiles to MEMBAR.ALL.GPU and cost ~3 000
//   cycles per hand-over).
//   What did not work for the row stream (CUDA-event times at M = 201 728): plain loads -- with 222 KB
//   of the SM's 256 KB carved out as shared memory the L1 holds ~28 KB of pending lines, ~6 bytes per clock per SM;
//   prefetch.global.L2 / cp.async.bulk.prefetch.L2 one block ahead -- x was then read from DRAM 1.6 times.
//
// The result is bit-identical to layernorm_rows_kernel + GEMM (same operations per element, same MMA order).
// Warp roles (640 threads): warp 0 TMA producer (weight half-tiles), warp 1 MMA issuer (leader CTA only), warp 2 TMEM
// allocator, warps 4-11 epilogue (bias, fp16, per-warp TMA stores), warps 12-19 LayerNorm.  Registers are re-distributed
// with setmaxnreg (see kLnqRegs*).
#pragma once
#include "gemm_sm100_tma_epi.cuh"
#include "ln_math.cuh"

namespace effocr {

// timeline probe (tools/lnq_timeline.py): CTA 0 stamps clock64 into dbg[role * 1024 + index]
#define LNQ_STAMP(role, idx) do { if (dbg && blockIdx.x == 0 && (idx) < 1024) dbg[(role) * 1024 + (idx)] = clock64(); } while (0)

#ifndef LNQ_X_SLOTS
#define LNQ_X_SLOTS 3
#endif
#ifndef LNQ_ILP
#define LNQ_ILP 2
#endif
#ifndef LNQ_BN
#define LNQ_BN 192
#endif
constexpr int kLnqD = 384;
constexpr int kLnqKB = kLnqD / 64;       // 6 k-blocks
constexpr int kLnqThreads = 640;
constexpr int kLnqLnWarps = 8;
// setmaxnreg budget: the CTA's pool is what it was launched with (640 threads x 96 registers); an increase that the
// decreases do not cover would wait forever
constexpr int kLnqRegsLaunch = 96, kLnqRegsCtl = 40, kLnqRegsEpi = 80, kLnqRegsLn = 136;
static_assert(4 * kLnqRegsCtl + 8 * kLnqRegsEpi + kLnqLnWarps * kLnqRegsLn <= 20 * kLnqRegsLaunch, "register pool exceeded");
constexpr int kLnqRowsPerWarp = 128 / kLnqLnWarps;  // 16
constexpr int kLnqXSlots = LNQ_X_SLOTS;            // rows each LayerNorm warp keeps in flight (shared-memory slots)

template <int BLOCK_N>
struct LnQkvCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;          // 16 KB per k-block
  static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;    // this CTA's half of a weight tile
  static constexpr int kStoreCols = (BLOCK_N / 2) % 64 == 0 ? 64 : 32;   // columns per TMA store (128- or 64-byte rows)
  static constexpr int kWarpStagingBytes = 32 * kStoreCols * 2;
  static constexpr int kStagingTotal = 8 * kWarpStagingBytes;    // one store buffer per epilogue warp
  static constexpr int kXRowBytes = kLnqD * 4;
  static constexpr int kXBytes = kLnqLnWarps * kLnqXSlots * kXRowBytes;  // fp32 rows in flight for the LayerNorm warps
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesRaw = (kSmemLimit - 2048 - kBarrierBytes - kStagingTotal - kXBytes - kLnqKB * kABytes) / kBBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kLnqKB * kABytes + kStages * kBBytes + kStagingTotal + kXBytes + kBarrierBytes + 1024;
  static constexpr int kTmemCols = GemmCfg<BLOCK_N>::kTmemCols;
  static_assert(BLOCK_N % 32 == 0 && kStages >= 3, "pipeline too shallow");
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kLnqThreads, 1)
ln_gemm_astat_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_c, int M, int N, EpiTmaParams ep, int flags,
                     long long* __restrict__ dbg) {
  using Cfg = LnQkvCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int KB = kLnqKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + KB * Cfg::kABytes;
  uint8_t* smem_c = smem_b + STAGES * Cfg::kBBytes;
  uint8_t* smem_x = smem_c + Cfg::kStagingTotal;
  uint64_t* bfull_bar = reinterpret_cast<uint64_t*>(smem_x + Cfg::kXBytes);
  uint64_t* bempty_bar = bfull_bar + STAGES;
  uint64_t* afull_bar = bempty_bar + STAGES;
  uint64_t* aempty_bar = afull_bar + KB;
  uint64_t* tfull_bar = aempty_bar + KB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* xfull_bar = tempty_bar + 2;  // [LN warp][slot]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull_bar + kLnqLnWarps * kLnqXSlots);
  static_assert((2 * STAGES + 2 * KB + 4 + kLnqLnWarps * kLnqXSlots) * 8 + 4 <= Cfg::kBarrierBytes, "barrier area too small");

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&bfull_bar[i], 1);   // leader's arrive.expect_tx; bytes arrive from both CTAs' TMA loads
      mbar_init(&bempty_bar[i], 1);  // one multicast commit per round
    }
    for (int i = 0; i < KB; ++i) {
      mbar_init(&afull_bar[i], 2 * kLnqLnWarps);  // the LayerNorm warps of BOTH CTAs arrive on the leader's barrier
      mbar_init(&aempty_bar[i], 1);               // multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);  // 8 epilogue warps of each CTA arrive on the leader's barrier
    }
    for (int i = 0; i < kLnqLnWarps * kLnqXSlots; ++i) mbar_init(&xfull_bar[i], 1);
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

  const int num_sb = (M + 2 * kBlockM - 1) / (2 * kBlockM);  // 256-row super-blocks, one per pair and round
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;

  if (warp_idx < 4) setmaxnreg_dec<kLnqRegsCtl>();
  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer: this CTA's half of every weight tile
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int sb = pair; sb < num_sb; sb += num_pairs) {
        for (int nb = 0; nb < num_n; ++nb) {
          const int n0 = nb * BLOCK_N + rank * (BLOCK_N / 2);
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&bempty_bar[stage], phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&bfull_bar[stage], 2 * Cfg::kBBytes);
            tma_load_2d_2sm(&tma_b, &bfull_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBlockK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform loop, predicated issue)
    if (rank == 0) {
      const bool leader_lane = elect_one_sync();
      constexpr uint32_t idesc = make_idesc_f16(2 * kBlockM, BLOCK_N);
      const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t a_phase = 0;
      int local = 0;
      for (int sb = pair; sb < num_sb; sb += num_pairs, a_phase ^= 1) {
        for (int nb = 0; nb < num_n; ++nb, ++local) {
          const int as = local & 1;
          const uint32_t aphase = (local >> 1) & 1;
          if (leader_lane) LNQ_STAMP(0, 4 * local);          // tile: about to wait for the TMEM buffer
          mbar_wait(&tempty_bar[as], aphase ^ 1);
          tcgen05_fence_after();
          if (leader_lane) LNQ_STAMP(0, 4 * local + 1);      // TMEM buffer free
          const uint32_t tmem_d = tmem_base + as * BLOCK_N;
          for (int kb = 0; kb < KB; ++kb) {
            if (nb == 0 && (kb & 1) == 0) mbar_wait(&afull_bar[kb >> 1], a_phase);  // one barrier per k-block pair
            mbar_wait(&bfull_bar[stage], phase);
            tcgen05_fence_after();
            if (leader_lane && kb == 0) LNQ_STAMP(0, 4 * local + 2);  // first k-block's operands ready
            const uint64_t da = make_sw128_kmajor_desc(a_base + kb * Cfg::kABytes);
            const uint64_t db = make_sw128_kmajor_desc(b_base + stage * Cfg::kBBytes);
            if (leader_lane) {
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_f16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
              umma_commit_2sm(&bempty_bar[stage]);
              if (nb == num_n - 1) umma_commit_2sm(&aempty_bar[kb]);  // A k-block free for the next block, in both CTAs
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (leader_lane) { umma_commit_2sm(&tfull_bar[as]); LNQ_STAMP(0, 4 * local + 3); }  // tile issued
          __syncwarp();
        }
      }
    }
  } else if (warp_idx >= 4 && warp_idx < 12) {
    // ------------------------------------------------------------------ epilogue warps (one 32-column chunk in registers)
    setmaxnreg_dec<kLnqRegsEpi>();
    const int q = warp_idx & 3;
    const int half = (warp_idx - 4) >> 2;
    uint8_t* dst = smem_c + (warp_idx - 4) * Cfg::kWarpStagingBytes;
    constexpr int CHUNKS = BLOCK_N / 64;
    int local = 0;
    for (int sb = pair; sb < num_sb; sb += num_pairs) {
      for (int nb = 0; nb < num_n; ++nb, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        const int m0 = ((flags & 256) ? num_sb - 1 - sb : sb) * 2 * kBlockM + rank * kBlockM + q * 32;  // bit 8: walk the row blocks backwards
        const int n0 = nb * BLOCK_N + half * (BLOCK_N / 2);
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / 2);
        if (warp_idx == 4 && lane == 0) LNQ_STAMP(1, 3 * local);
        mbar_wait(&tfull_bar[as], aphase);
        tcgen05_fence_after();
        if (warp_idx == 4 && lane == 0) LNQ_STAMP(1, 3 * local + 1);   // accumulator complete
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int col0 = n0 + c * 32;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tbase + c * 32, v);
          // the chunk's 32 bias values, all eight loads in flight under the TMEM read (N is a multiple of 8)
          float4 bv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.bias && col0 + 4 * j + 4 <= N) bv[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 4 * j));
          }
          tmem_ld_wait();
          if (c == CHUNKS - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
          }
          if (col0 >= N || (flags & 4)) continue;  // warp-uniform (flag 4: timing experiment without the stores)
          constexpr int PER = Cfg::kStoreCols / 32;  // chunks per store
          if (c % PER == 0) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0 = bv[2 * j], b1 = bv[2 * j + 1];
            uint4 pk;
            __half2* ph = reinterpret_cast<__half2*>(&pk);
            ph[0] = __floats2half2_rn(__uint_as_float(v[8 * j]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y);
            ph[1] = __floats2half2_rn(__uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w);
            ph[2] = __floats2half2_rn(__uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y);
            ph[3] = __floats2half2_rn(__uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w);
            if constexpr (PER == 2) {
              // 128-byte rows, SWIZZLE_128B: 16-byte piece p of row r lives at piece p ^ (r & 7)
              *reinterpret_cast<uint4*>(dst + lane * 128 + ((((c & 1) * 4 + j) ^ (lane & 7)) << 4)) = pk;
            } else {
              // 64-byte rows, SWIZZLE_64B: 16-byte piece j of row r lives at piece j ^ ((r >> 1) & 3)
              *reinterpret_cast<uint4*>(dst + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
            }
          }
          if (c % PER == PER - 1 || col0 + 32 >= N) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tma_c, dst, col0 - (c % PER) * 32, m0);
              tma_store_commit();
            }
          }
        }
        if (warp_idx == 4 && lane == 0) LNQ_STAMP(1, 3 * local + 2);   // tile drained
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else if (warp_idx >= 12) {
    // ------------------------------------------------------------------ LayerNorm warps: rows lw*16 .. lw*16+15 of every block
    setmaxnreg_inc<kLnqRegsLn>();
    const int lw = warp_idx - 12;
    constexpr int R = kLnqRowsPerWarp;
    uint32_t hold[R][6];  // normalised fp16 pairs of this lane's 12 columns of each row
    const uint32_t xs = smem_u32(smem_x + lw * kLnqXSlots * Cfg::kXRowBytes);
    uint64_t* xfull = xfull_bar + lw * kLnqXSlots;
    const int nblk = pair < num_sb ? (num_sb - 1 - pair) / num_pairs + 1 : 0;
    const int total = nblk * R;  // rows in this warp's stream
    auto issue = [&](int n, int slot) {  // lane 0: row n of the stream into `slot`
      const int sb = pair + (n / R) * num_pairs;
      int row = ((flags & 256) ? num_sb - 1 - sb : sb) * 2 * kBlockM + rank * kBlockM + lw * R + (n % R);
      row = row < M ? row : M - 1;
      mbar_arrive_expect_tx(&xfull[slot], Cfg::kXRowBytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       xs + slot * Cfg::kXRowBytes),
                   "l"(x + static_cast<long long>(row) * ldx), "r"(Cfg::kXRowBytes), "r"(smem_u32(&xfull[slot]))
                   : "memory");
    };
    const bool skip_ln = flags & 2;  // timing experiment: hand over whatever the registers hold
    if (lane == 0 && !skip_ln)
      for (int n = 0; n < kLnqXSlots && n < total; ++n) issue(n, n);
    int xn = 0, xslot = 0;
    uint32_t xphase = 0;
    auto load_norm = [&]() {  // the next 16 rows of the stream -> hold[]
      // U rows at a time, their instruction streams interleaved: one row alone is a ~1300-cycle dependent chain (two
      // butterflies, the divide and square root)
      constexpr int U = LNQ_ILP;
      static_assert(R % U == 0 && U < kLnqXSlots, "rows per step");
#pragma unroll
      for (int r = 0; r < R; r += U) {
        int sl[U];
        float4 c[U][3];
        float s[U], mu[U], qv[U], rs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          sl[u] = xslot;
          mbar_wait(&xfull[xslot], xphase);
          if (++xslot == kLnqXSlots) { xslot = 0; xphase ^= 1; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t src = xs + sl[u] * Cfg::kXRowBytes + lane * 16;
#pragma unroll
          for (int j = 0; j < 3; ++j)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(c[u][j].x), "=f"(c[u][j].y), "=f"(c[u][j].z), "=f"(c[u][j].w)
                         : "r"(src + j * 512)
                         : "memory");
        }
        // layernorm_rows_kernel's arithmetic: sum in column order, butterfly, mean; sum of squared deviations likewise
#pragma unroll
        for (int u = 0; u < U; ++u) {
          s[u] = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) { s[u] += c[u][j].x; s[u] += c[u][j].y; s[u] += c[u][j].z; s[u] += c[u][j].w; }
        }
        // every lane's slot reads fed its sum; once the warp has converged the slots are free: refill kLnqXSlots rows ahead
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (xn + u + kLnqXSlots < total) issue(xn + u + kLnqXSlots, sl[u]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int u = 0; u < U; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          mu[u] = __fmul_rn(s[u], 1.0f / kLnqD);
          qv[u] = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float d = __fsub_rn(c[u][j].x, mu[u]); qv[u] = __fmaf_rn(d, d, qv[u]);
            d = __fsub_rn(c[u][j].y, mu[u]); qv[u] = __fmaf_rn(d, d, qv[u]);
            d = __fsub_rn(c[u][j].z, mu[u]); qv[u] = __fmaf_rn(d, d, qv[u]);
            d = __fsub_rn(c[u][j].w, mu[u]); qv[u] = __fmaf_rn(d, d, qv[u]);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int u = 0; u < U; ++u) qv[u] += __shfl_xor_sync(0xffffffffu, qv[u], o);
#pragma unroll
        for (int u = 0; u < U; ++u) rs[u] = ln_rstd(qv[u], 1.0f / kLnqD, eps);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int col = (j * 32 + lane) * 4;
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + col));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + col));
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const __half2 lo = __floats2half2_rn(ln_affine(c[u][j].x, mu[u], rs[u], g4.x, b4.x), ln_affine(c[u][j].y, mu[u], rs[u], g4.y, b4.y));
            const __half2 hi = __floats2half2_rn(ln_affine(c[u][j].z, mu[u], rs[u], g4.z, b4.z), ln_affine(c[u][j].w, mu[u], rs[u], g4.w, b4.w));
            hold[r + u][2 * j] = *reinterpret_cast<const uint32_t*>(&lo);
            hold[r + u][2 * j + 1] = *reinterpret_cast<const uint32_t*>(&hi);
          }
        }
        xn += U;
      }
    };
    uint32_t a_phase = 0;
    if (pair < num_sb && !skip_ln) load_norm();
    const int piece = (lane & 15) >> 1, sub = (lane & 1) * 8;
    int blk = 0;
    for (int sb = pair; sb < num_sb; sb += num_pairs, a_phase ^= 1, ++blk) {
      if (lw == 0 && lane == 0 && (flags & 8)) LNQ_STAMP(2, 8 * blk);       // rows normalised, waiting for the A buffer
#pragma unroll
      for (int j = 0; j < 3; ++j) {  // columns j*128 .. j*128+127 = k-blocks 2j (lanes 0-15) and 2j+1 (lanes 16-31)
        mbar_wait(&aempty_bar[2 * j], a_phase ^ 1);
        mbar_wait(&aempty_bar[2 * j + 1], a_phase ^ 1);
        if (lw == 0 && lane == 0 && (flags & 8)) LNQ_STAMP(2, 8 * blk + 1 + 2 * j);   // k-block pair j released by the MMAs
        uint8_t* kb_base = smem_a + (2 * j + (lane >> 4)) * Cfg::kABytes;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const int r = lw * R + i;  // row inside the 128-row block
          // K-major SWIZZLE_128B tile: 16-byte piece p of row r lives at piece p ^ (r & 7)
          *reinterpret_cast<uint2*>(kb_base + r * 128 + ((piece ^ (r & 7)) << 4) + sub) = make_uint2(hold[i][2 * j], hold[i][2 * j + 1]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&afull_bar[j]);
        if (lw == 0 && lane == 0 && (flags & 8)) LNQ_STAMP(2, 8 * blk + 2 + 2 * j);   // k-block pair j handed over
      }
      // the next block of this CTA: normalise it under the MMAs of the block just handed over
      if (sb + num_pairs < num_sb && !skip_ln) load_norm();
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();  // the peer's barriers / shared memory stay alive until the leader's last multicast has landed
  if (warp_idx == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

}  // namespace effocr
