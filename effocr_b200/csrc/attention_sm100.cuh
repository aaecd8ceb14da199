// Fused multi-head attention for the fixed ViT/16 @ 224 sequence (197 tokens, head dim 64) on tcgen05.
//
// One persistent CTA per SM walks (image, head) pairs; the two 128-row query tiles of a pair are processed
// concurrently by two groups of four softmax warps:
//
//   S_g = Q_g K^T   tcgen05.mma  M=128, N=208 (197 keys padded to a multiple of 16), K=64; fp32 S in TMEM
//   softmax         one thread per query row, ONE pass over the S row in TMEM (tcgen05.ld, next chunk prefetched):
//                   running max + fp16 deltas staged in the P tile, then exp2 / row sum from shared memory in place
//   O_g = P_g V     tcgen05.mma  M=128, N=64, K=208; A = P (K-major smem: three 64-key SWIZZLE_128B chunks +
//                   one 16-key SWIZZLE_32B tail), B = V used in place as an MN-major operand (V is
//                   [key][dim] in memory: no transpose pass); fp32 O in TMEM, aliased onto S_g's columns
//   epilogue        O / rowsum -> fp16 -> global, rows < 197 only
//
// The MMA warp issues S0(i), PV1(i-1), S1(i), PV0(i): the two groups run half a period out of phase, so the MUFU-bound
// softmax of one tile overlaps the MMAs / TMEM waits / epilogue of the other.
//
// Q, K, V tiles come straight out of the fused QKV GEMM's output ([B*197, 3*H*64] fp16) by TMA; K of the next
// pair is prefetched (double buffer), V is single-buffered (its reload hides behind the next pair's S + softmax).
// Keys 197..207 of a K/V box belong to the next image (or are zero-filled at the end of the tensor); they are
// masked before the softmax, so P is exactly 0 there.
//
// Replaces the attention inside timm's Block.forward reached from /root/reference/models/encoders.py:62-64.
#pragma once
#include "sm100_ptx.cuh"

namespace effocr {

constexpr int kAtT = 197;                 // tokens
constexpr int kAtN = 208;                 // keys padded to a multiple of 16 (UMMA N granularity at M=128)
constexpr int kAtQBytes = 128 * 128;      // 128 rows x 64 fp16
constexpr int kAtKVBytes = kAtN * 128;    // 208 rows x 64 fp16
constexpr int kAtPMain = 3 * 128 * 128;   // keys 0..191: three SWIZZLE_128B chunks
constexpr int kAtPTail = 128 * 32;        // keys 192..207: one SWIZZLE_32B chunk
constexpr int kAtPBytes = kAtPMain + kAtPTail;
constexpr int kAtThreads = 384;
constexpr int kAtSmemBytes = 2 * kAtQBytes + 3 * kAtKVBytes + 2 * kAtPBytes + 256 + 1024;
constexpr uint32_t kAtTmemRegion = 256;   // columns per query-tile group (S: 208, O aliases the first 64)

// MN-major operand tile with 128-byte rows written by TMA (SWIZZLE_128B): row = K index (key), the 128 bytes of
// a row are 64 consecutive MN elements (head dims).  SBO = 1024 B between groups of 8 K rows; one MN atom only.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;              // LBO: stride between 64-element MN atoms (single atom here)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;      // SBO
  d |= static_cast<uint64_t>(1) << 46;              // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;              // SWIZZLE_128B
  return d;
}
// K-major operand tile with 32-byte rows (16 fp16 = one UMMA K step), SWIZZLE_32B: 8-row atoms of 256 B.
__device__ __forceinline__ uint64_t make_sw32_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;       // SBO: 8 rows x 32 B
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;              // SWIZZLE_32B
  return d;
}
// kind::f16, fp16 A (K-major) x fp16 B (MN-major), fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_f16_bmn(int M, int N) {
  return make_idesc_f16(M, N) | (1u << 16);
}

#ifdef EFFOCR_AB  // A/B variants (two-pass / fp16-delta single pass / 16 softmax warps), not part of the product build
template <bool kSinglePass>
__global__ void __launch_bounds__(kAtThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                    __half* __restrict__ out, int batch, int H, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                              // [2 groups] x 16 KB
  uint8_t* smem_k = smem_q + 2 * kAtQBytes;            // [2 pair parities] x 26 KB
  uint8_t* smem_v = smem_k + 2 * kAtKVBytes;           // 26 KB
  uint8_t* smem_p = smem_v + kAtKVBytes;               // [2 groups] x 52 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + 2 * kAtPBytes);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* k_full = bars + 4;     // [2]
  uint64_t* k_empty = bars + 6;    // [2]
  uint64_t* v_full = bars + 8;
  uint64_t* v_empty = bars + 9;
  uint64_t* s_full = bars + 10;    // [2]
  uint64_t* p_full = bars + 12;    // [2]
  uint64_t* o_full = bars + 14;    // [2]
  uint64_t* t_free = bars + 16;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int D = H * 64;
  const int num_bh = batch * H;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&t_free[i], 4);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ---------------- TMA producer
    if (elect_one_sync()) {
      int it = 0;  // pair counter of this CTA
      for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
        const int b = bh / H, h = bh % H;
        const int row0 = b * kAtT;
        const int s = it & 1;
        const uint32_t par = it & 1, par2 = (it >> 1) & 1;
        mbar_wait(&k_empty[s], par2 ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kAtKVBytes);
        tma_load_2d(&tma_kv, &k_full[s], smem_k + s * kAtKVBytes, D + h * 64, row0);
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&q_empty[g], par ^ 1);
          mbar_arrive_expect_tx(&q_full[g], kAtQBytes);
          tma_load_2d(&tma_q, &q_full[g], smem_q + g * kAtQBytes, h * 64, row0 + g * 128);
        }
        mbar_wait(v_empty, par ^ 1);
        mbar_arrive_expect_tx(v_full, kAtKVBytes);
        tma_load_2d(&tma_kv, v_full, smem_v, 2 * D + h * 64, row0);
      }
    }
  } else if (warp_idx == 1) {
    // ---------------- MMA issuer.  Issue order per pair i:  S0(i), PV1(i-1), S1(i), PV0(i)  -- the two query-tile
    // groups run half a period out of phase, so one group's MUFU-bound softmax overlaps the other group's MMAs,
    // TMEM waits and epilogue instead of both groups competing for the MUFU pipe and then idling together.
    // The whole warp walks the loop (waits and descriptor arithmetic stay warp-uniform, i.e. in uniform registers);
    // only the tcgen05 instructions are predicated on one elected lane -- issuing from inside an `if (elect_one)`
    // region costs a R2UR waterfall of ~16 instructions per MMA, three times the 32 cycles an N = 64 MMA runs for.
    {
      const bool leader_lane = elect_one_sync();
      constexpr uint32_t idesc_s = make_idesc_f16(128, kAtN);
      constexpr uint32_t idesc_o = make_idesc_f16_bmn(128, 64);
      const uint32_t q_base = smem_u32(smem_q), k_base = smem_u32(smem_k), v_base = smem_u32(smem_v), p_base = smem_u32(smem_p);
      auto issue_s = [&](int g, int it) {
        const uint32_t par = it & 1;
        const int s = it & 1;
        mbar_wait(&q_full[g], par);
        mbar_wait(&t_free[g], par ^ 1);  // region g (S/O columns) drained by the previous pair's epilogue
        tcgen05_fence_after();
        const uint64_t dq = make_sw128_kmajor_desc(q_base + g * kAtQBytes);
        const uint64_t dk = make_sw128_kmajor_desc(k_base + s * kAtKVBytes);
        if (leader_lane) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + g * kAtTmemRegion, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
          umma_commit(&q_empty[g]);
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int it) {
        const uint32_t par = it & 1;
        mbar_wait(&p_full[g], par);  // all four warps of the group have read S_g and written P_g
        tcgen05_fence_after();
        const uint32_t pbase = p_base + g * kAtPBytes;
        if (leader_lane) {
#pragma unroll
          for (int ks = 0; ks < 12; ++ks) {
            const uint64_t dp = make_sw128_kmajor_desc(pbase + (ks >> 2) * kAtQBytes) + 2 * (ks & 3);
            const uint64_t dv = make_sw128_mnmajor_desc(v_base + ks * 2048);
            umma_f16(tmem_base + g * kAtTmemRegion, dp, dv, idesc_o, ks ? 1u : 0u);
          }
          umma_f16(tmem_base + g * kAtTmemRegion, make_sw32_kmajor_desc(pbase + kAtPMain),
                   make_sw128_mnmajor_desc(v_base + 12 * 2048), idesc_o, 1u);
          umma_commit(&o_full[g]);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (leader_lane) umma_commit(bar);
        __syncwarp();
      };
      int it = 0;
      for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&k_full[s], (it >> 1) & 1);
        issue_s(0, it);
        if (it > 0) {
          issue_pv(1, it - 1);   // V(it-1) is still resident: its buffer is released right here
          commit(v_empty);
        }
        issue_s(1, it);
        commit(&k_empty[s]);
        mbar_wait(v_full, it & 1);  // V(it), reloaded after PV1(it-1) retired
        issue_pv(0, it);
      }
      if (it > 0) {
        issue_pv(1, it - 1);
        commit(v_empty);
      }
    }
  } else if (warp_idx >= 4) {
    // ---------------- softmax + epilogue: group g = query tile, thread = query row
    const int g = (warp_idx - 4) >> 2;
    const int q = warp_idx & 3;
    const int r = q * 32 + lane;  // row inside the 128-row tile
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * kAtTmemRegion;
    uint8_t* pmain = smem_p + g * kAtPBytes + r * 128;
    uint8_t* ptail = smem_p + g * kAtPBytes + kAtPMain + r * 32;
    int it = 0;
    for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
      const int b = bh / H, h = bh % H;
      const uint32_t par = it & 1;
      mbar_wait(&s_full[g], par);
      tcgen05_fence_after();
      float sum;
      if constexpr (kSinglePass) {
        // Pass 1 -- the ONLY pass over TMEM (tcgen05.ld moves 64 B/clk/SM; reading S twice was the kernel's bound):
        // t = s * scale (log2 domain), running row maximum m, and the DELTAS d = t - m (<= 0) go to the P tile as
        // fp16, each 32-key chunk relative to the running maximum at that point (ref[c]).  fp16 keeps |d| * 2^-11
        // absolute precision, i.e. the entries that matter (d near 0) are accurate to ~1e-4 in the exponent.
        float ref[7];
        float mrun = -INFINITY;
        {
          uint32_t v[2][32];
          uint32_t vt[16];
          tmem_ld_32x32b_x32(trow, v[0]);
  #pragma unroll
          for (int c = 0; c < 6; ++c) {
            tmem_ld_wait();
            if (c + 1 < 6) tmem_ld_32x32b_x32(trow + (c + 1) * 32, v[(c + 1) & 1]);
            else tmem_ld_32x32b_x16(trow + 192, vt);
            float t[32];
            float m4[4] = {mrun, mrun, mrun, mrun};
  #pragma unroll
            for (int i = 0; i < 32; ++i) {
              t[i] = __uint_as_float(v[c & 1][i]) * scale_log2e;
              m4[i & 3] = fmaxf(m4[i & 3], t[i]);
            }
            mrun = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            ref[c] = mrun;
            uint8_t* chunk = pmain + (c >> 1) * kAtQBytes;
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              __half2* ph = reinterpret_cast<__half2*>(&pk);
  #pragma unroll
              for (int u = 0; u < 4; ++u) ph[u] = __floats2half2_rn(t[8 * j + 2 * u] - mrun, t[8 * j + 2 * u + 1] - mrun);
              const int piece = (c & 1) * 4 + j;  // 16-byte piece inside the 64-key chunk
              *reinterpret_cast<uint4*>(chunk + ((piece ^ (r & 7)) << 4)) = pk;
            }
          }
          tmem_ld_wait();
          float t[16];
  #pragma unroll
          for (int i = 0; i < 16; ++i) {
            t[i] = (i < 5) ? __uint_as_float(vt[i]) * scale_log2e : -INFINITY;  // keys >= 197 masked
            mrun = fmaxf(mrun, t[i]);
          }
          ref[6] = mrun;
  #pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4 pk;
            __half2* ph = reinterpret_cast<__half2*>(&pk);
  #pragma unroll
            for (int u = 0; u < 4; ++u) ph[u] = __floats2half2_rn(t[8 * j + 2 * u] - mrun, t[8 * j + 2 * u + 1] - mrun);
            *reinterpret_cast<uint4*>(ptail + ((j ^ ((r >> 2) & 1)) << 4)) = pk;  // SWIZZLE_32B: bit 4 ^= bit 7
          }
        }
        // S fully read: the accumulator columns may be overwritten by P V (which aliases them) once P is ready
        tcgen05_fence_before();
        // Pass 2 -- shared memory only (each thread re-reads what it wrote): p = 2^(d + ref[c] - m), row sum, in place
        {
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
  #pragma unroll
          for (int c = 0; c < 6; ++c) {
            const float off = ref[c] - mrun;
            uint8_t* chunk = pmain + (c >> 1) * kAtQBytes;
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int piece = (c & 1) * 4 + j;
              uint4* slot = reinterpret_cast<uint4*>(chunk + ((piece ^ (r & 7)) << 4));
              uint4 pk = *slot;
              __half2* ph = reinterpret_cast<__half2*>(&pk);
  #pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 d = __half22float2(ph[u]);
                const float p0 = ex2_approx(d.x + off), p1 = ex2_approx(d.y + off);
                s4[(2 * u) & 3] += p0;
                s4[(2 * u + 1) & 3] += p1;
                ph[u] = __floats2half2_rn(p0, p1);
              }
              *slot = pk;
            }
          }
          {
            const float off = ref[6] - mrun;  // == 0
  #pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint4* slot = reinterpret_cast<uint4*>(ptail + ((j ^ ((r >> 2) & 1)) << 4));
              uint4 pk = *slot;
              __half2* ph = reinterpret_cast<__half2*>(&pk);
  #pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 d = __half22float2(ph[u]);
                const float p0 = ex2_approx(d.x + off), p1 = ex2_approx(d.y + off);  // 2^-inf = 0 for masked keys
                s4[(2 * u) & 3] += p0;
                s4[(2 * u + 1) & 3] += p1;
                ph[u] = __floats2half2_rn(p0, p1);
              }
              *slot = pk;
            }
          }
          sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        }
      } else {
        // pass 1: row maximum over the 197 valid keys (chunk c+1 in flight while chunk c is reduced)
        float mx;
        {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains: FMNMX latency, not count
          uint32_t v[2][32];
          uint32_t vt[16];
          tmem_ld_32x32b_x32(trow, v[0]);
  #pragma unroll
          for (int c = 0; c < 6; ++c) {
            tmem_ld_wait();
            if (c + 1 < 6) tmem_ld_32x32b_x32(trow + (c + 1) * 32, v[(c + 1) & 1]);
            else tmem_ld_32x32b_x16(trow + 192, vt);
  #pragma unroll
            for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[c & 1][i]));
          }
          tmem_ld_wait();
  #pragma unroll
          for (int i = 0; i < 5; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(vt[i]));  // keys 192..196
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        }
        const float moff = mx * scale_log2e;
        // pass 2: p = 2^(s * scale - max * scale), row sum, P -> smem (fp16, K-major, swizzled)
        {
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t v[2][32];
          uint32_t vt[16];
          tmem_ld_32x32b_x32(trow, v[0]);
  #pragma unroll
          for (int c = 0; c < 6; ++c) {
            tmem_ld_wait();
            if (c + 1 < 6) tmem_ld_32x32b_x32(trow + (c + 1) * 32, v[(c + 1) & 1]);
            else tmem_ld_32x32b_x16(trow + 192, vt);
            float p[32];
  #pragma unroll
            for (int i = 0; i < 32; ++i) {
              p[i] = ex2_approx(fmaf(__uint_as_float(v[c & 1][i]), scale_log2e, -moff));
              s4[i & 3] += p[i];
            }
            uint8_t* chunk = pmain + (c >> 1) * kAtQBytes;
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              __half2* ph = reinterpret_cast<__half2*>(&pk);
  #pragma unroll
              for (int t = 0; t < 4; ++t) ph[t] = __floats2half2_rn(p[8 * j + 2 * t], p[8 * j + 2 * t + 1]);
              const int piece = (c & 1) * 4 + j;  // 16-byte piece inside the 64-key chunk
              *reinterpret_cast<uint4*>(chunk + ((piece ^ (r & 7)) << 4)) = pk;
            }
          }
          tmem_ld_wait();
          float p[16];
  #pragma unroll
          for (int i = 0; i < 16; ++i) {
            p[i] = (i < 5) ? ex2_approx(fmaf(__uint_as_float(vt[i]), scale_log2e, -moff)) : 0.f;  // keys >= 197 masked
            s4[i & 3] += p[i];
          }
          sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  #pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4 pk;
            __half2* ph = reinterpret_cast<__half2*>(&pk);
  #pragma unroll
            for (int t = 0; t < 4; ++t) ph[t] = __floats2half2_rn(p[8 * j + 2 * t], p[8 * j + 2 * t + 1]);
            *reinterpret_cast<uint4*>(ptail + ((j ^ ((r >> 2) & 1)) << 4)) = pk;  // SWIZZLE_32B: bit 4 ^= bit 7
          }
        }
        tcgen05_fence_before();
      }
      // P written: make the generic-proxy writes visible to the MMA (async proxy), then signal
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      // epilogue: O / sum -> fp16 -> global
      mbar_wait(&o_full[g], par);
      tcgen05_fence_after();
      const float inv = 1.0f / sum;
      const int tok = g * 128 + r;
      __half* dst = out + (static_cast<long long>(b) * kAtT + tok) * D + h * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(trow + hh * 32, o);
        tmem_ld_wait();
        if (hh == 1) {  // O fully read: region g is free for the next pair's S
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_free[g]);
        }
        if (tok < kAtT) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              ph[t] = __floats2half2_rn(__uint_as_float(o[8 * j + 2 * t]) * inv, __uint_as_float(o[8 * j + 2 * t + 1]) * inv);
            *reinterpret_cast<uint4*>(dst + hh * 32 + 8 * j) = pk;
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ 16-softmax-warp variant (default)
constexpr int kAt16Threads = 128 + 16 * 32;
constexpr int kAt16SmemBytes = kAtSmemBytes + 2 * 2 * 2 * 128 * 4;
__device__ __forceinline__ void named_bar_sync_at(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kAt16Threads, 1)
attention_tc16_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                    __half* __restrict__ out, int batch, int H, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                              // [2 groups] x 16 KB
  uint8_t* smem_k = smem_q + 2 * kAtQBytes;            // [2 pair parities] x 26 KB
  uint8_t* smem_v = smem_k + 2 * kAtKVBytes;           // 26 KB
  uint8_t* smem_p = smem_v + kAtKVBytes;               // [2 groups] x 52 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + 2 * kAtPBytes);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* k_full = bars + 4;     // [2]
  uint64_t* k_empty = bars + 6;    // [2]
  uint64_t* v_full = bars + 8;
  uint64_t* v_empty = bars + 9;
  uint64_t* s_full = bars + 10;    // [2]
  uint64_t* p_full = bars + 12;    // [2]
  uint64_t* o_full = bars + 14;    // [2]
  uint64_t* t_free = bars + 16;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* xch = reinterpret_cast<float*>(bars + 32);  // [max | sum][group][column half][128 rows]

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int D = H * 64;
  const int num_bh = batch * H;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);   // both column halves of the group
      mbar_init(&o_full[i], 1);
      mbar_init(&t_free[i], 8);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ---------------- TMA producer
    if (elect_one_sync()) {
      int it = 0;  // pair counter of this CTA
      for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
        const int b = bh / H, h = bh % H;
        const int row0 = b * kAtT;
        const int s = it & 1;
        const uint32_t par = it & 1, par2 = (it >> 1) & 1;
        mbar_wait(&k_empty[s], par2 ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kAtKVBytes);
        tma_load_2d(&tma_kv, &k_full[s], smem_k + s * kAtKVBytes, D + h * 64, row0);
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&q_empty[g], par ^ 1);
          mbar_arrive_expect_tx(&q_full[g], kAtQBytes);
          tma_load_2d(&tma_q, &q_full[g], smem_q + g * kAtQBytes, h * 64, row0 + g * 128);
        }
        mbar_wait(v_empty, par ^ 1);
        mbar_arrive_expect_tx(v_full, kAtKVBytes);
        tma_load_2d(&tma_kv, v_full, smem_v, 2 * D + h * 64, row0);
      }
    }
  } else if (warp_idx == 1) {
    // ---------------- MMA issuer.  Issue order per pair i:  S0(i), PV1(i-1), S1(i), PV0(i)  -- the two query-tile
    // groups run half a period out of phase, so one group's MUFU-bound softmax overlaps the other group's MMAs,
    // TMEM waits and epilogue instead of both groups competing for the MUFU pipe and then idling together.
    // The whole warp walks the loop (waits and descriptor arithmetic stay warp-uniform, i.e. in uniform registers);
    // only the tcgen05 instructions are predicated on one elected lane -- issuing from inside an `if (elect_one)`
    // region costs a R2UR waterfall of ~16 instructions per MMA, three times the 32 cycles an N = 64 MMA runs for.
    {
      const bool leader_lane = elect_one_sync();
      constexpr uint32_t idesc_s = make_idesc_f16(128, kAtN);
      constexpr uint32_t idesc_o = make_idesc_f16_bmn(128, 64);
      const uint32_t q_base = smem_u32(smem_q), k_base = smem_u32(smem_k), v_base = smem_u32(smem_v), p_base = smem_u32(smem_p);
      auto issue_s = [&](int g, int it) {
        const uint32_t par = it & 1;
        const int s = it & 1;
        mbar_wait(&q_full[g], par);
        mbar_wait(&t_free[g], par ^ 1);  // region g (S/O columns) drained by the previous pair's epilogue
        tcgen05_fence_after();
        const uint64_t dq = make_sw128_kmajor_desc(q_base + g * kAtQBytes);
        const uint64_t dk = make_sw128_kmajor_desc(k_base + s * kAtKVBytes);
        if (leader_lane) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + g * kAtTmemRegion, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
          umma_commit(&q_empty[g]);
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int it) {
        const uint32_t par = it & 1;
        mbar_wait(&p_full[g], par);  // all four warps of the group have read S_g and written P_g
        tcgen05_fence_after();
        const uint32_t pbase = p_base + g * kAtPBytes;
        if (leader_lane) {
#pragma unroll
          for (int ks = 0; ks < 12; ++ks) {
            const uint64_t dp = make_sw128_kmajor_desc(pbase + (ks >> 2) * kAtQBytes) + 2 * (ks & 3);
            const uint64_t dv = make_sw128_mnmajor_desc(v_base + ks * 2048);
            umma_f16(tmem_base + g * kAtTmemRegion, dp, dv, idesc_o, ks ? 1u : 0u);
          }
          umma_f16(tmem_base + g * kAtTmemRegion, make_sw32_kmajor_desc(pbase + kAtPMain),
                   make_sw128_mnmajor_desc(v_base + 12 * 2048), idesc_o, 1u);
          umma_commit(&o_full[g]);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (leader_lane) umma_commit(bar);
        __syncwarp();
      };
      int it = 0;
      for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&k_full[s], (it >> 1) & 1);
        issue_s(0, it);
        if (it > 0) {
          issue_pv(1, it - 1);   // V(it-1) is still resident: its buffer is released right here
          commit(v_empty);
        }
        issue_s(1, it);
        commit(&k_empty[s]);
        mbar_wait(v_full, it & 1);  // V(it), reloaded after PV1(it-1) retired
        issue_pv(0, it);
      }
      if (it > 0) {
        issue_pv(1, it - 1);
        commit(v_empty);
      }
    }
  } else if (warp_idx >= 4) {
    // ---------------- softmax + epilogue, SIXTEEN warps: a query row is shared by two threads (same TMEM lane, warps
    // with equal warp_idx % 4), each owning one half of the key columns (keys 0..111 / 112..207, 16-column chunks) and
    // one half of the output columns.  Row maximum and row sum are exchanged through shared memory under a 64-thread
    // named barrier.  Four softmax warps per scheduler instead of two: the MUFU pipe and the TMEM loads of one warp
    // hide behind the arithmetic of the others (the 8-warp version ran at XU 41 %, issue 29 %: latency-bound).
    const int w = warp_idx - 4;
    const int q = warp_idx & 3;
    const int g = (w >> 2) & 1;
    const int hc = w >> 3;
    const int r = q * 32 + lane;  // row inside the 128-row tile
    const int pair_bar = 1 + g * 4 + q;
    const int c0 = hc ? 7 : 0, nch = hc ? 6 : 7;  // my 16-column chunks: [c0, c0 + nch)
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * kAtTmemRegion;
    uint8_t* pmain = smem_p + g * kAtPBytes + r * 128;
    uint8_t* ptail = smem_p + g * kAtPBytes + kAtPMain + r * 32;
    float* my_max = xch + ((0 * 2 + g) * 2 + hc) * 128 + r;
    float* peer_max = xch + ((0 * 2 + g) * 2 + (hc ^ 1)) * 128 + r;
    float* my_sum = xch + ((1 * 2 + g) * 2 + hc) * 128 + r;
    float* peer_sum = xch + ((1 * 2 + g) * 2 + (hc ^ 1)) * 128 + r;
    int it = 0;
    for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
      const int b = bh / H, h = bh % H;
      const uint32_t par = it & 1;
      mbar_wait(&s_full[g], par);
      tcgen05_fence_after();
      // pass 1: maximum over my valid keys (chunk i+1 in flight while chunk i is reduced)
      float mx;
      {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        uint32_t v[2][16];
        tmem_ld_32x32b_x16(trow + c0 * 16, v[0]);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          if (i < nch) {
            tmem_ld_wait();
            if (i + 1 < nch) tmem_ld_32x32b_x16(trow + (c0 + i + 1) * 16, v[(i + 1) & 1]);
            const int nvalid = (c0 + i == 12) ? 5 : 16;  // keys 192..196 are the last valid ones
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (e < nvalid) m4[e & 3] = fmaxf(m4[e & 3], __uint_as_float(v[i & 1][e]));
          }
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      *my_max = mx;
      named_bar_sync_at(pair_bar, 64);
      mx = fmaxf(mx, *peer_max);
      const float moff = mx * scale_log2e;
      // pass 2: p = 2^(s * scale - max * scale), partial row sum, P -> smem (fp16, K-major, swizzled)
      float sum;
      {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t v[2][16];
        tmem_ld_32x32b_x16(trow + c0 * 16, v[0]);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          if (i < nch) {
            const int c = c0 + i;
            tmem_ld_wait();
            if (i + 1 < nch) tmem_ld_32x32b_x16(trow + (c + 1) * 16, v[(i + 1) & 1]);
            float p[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              p[e] = (c == 12 && e >= 5) ? 0.f : ex2_approx(fmaf(__uint_as_float(v[i & 1][e]), scale_log2e, -moff));
              s4[e & 3] += p[e];
            }
            uint4 pk[2];
            __half2* ph = reinterpret_cast<__half2*>(pk);
#pragma unroll
            for (int t = 0; t < 8; ++t) ph[t] = __floats2half2_rn(p[2 * t], p[2 * t + 1]);
            if (c < 12) {
              uint8_t* chunk = pmain + (c >> 2) * kAtQBytes;
              const int piece = (c & 3) * 2;  // 16-byte pieces inside the 64-key SWIZZLE_128B chunk
              *reinterpret_cast<uint4*>(chunk + ((piece ^ (r & 7)) << 4)) = pk[0];
              *reinterpret_cast<uint4*>(chunk + (((piece + 1) ^ (r & 7)) << 4)) = pk[1];
            } else {
              *reinterpret_cast<uint4*>(ptail + ((0 ^ ((r >> 2) & 1)) << 4)) = pk[0];  // SWIZZLE_32B: bit 4 ^= bit 7
              *reinterpret_cast<uint4*>(ptail + ((1 ^ ((r >> 2) & 1)) << 4)) = pk[1];
            }
          }
        }
        sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      *my_sum = sum;
      tcgen05_fence_before();
      // P written: make the generic-proxy writes visible to the MMA (async proxy), then signal
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      // epilogue: O / sum -> fp16 -> global; this thread owns output columns hc*32 .. +32
      mbar_wait(&o_full[g], par);
      tcgen05_fence_after();
      named_bar_sync_at(pair_bar, 64);  // partner's partial sum is visible; its reads of my max are done
      const float inv = 1.0f / (sum + *peer_sum);
      const int tok = g * 128 + r;
      __half* dst = out + (static_cast<long long>(b) * kAtT + tok) * D + h * 64 + hc * 32;
      uint32_t o[2][16];
      tmem_ld_32x32b_x16(trow + hc * 32, o[0]);
      tmem_ld_32x32b_x16(trow + hc * 32 + 16, o[1]);
      tmem_ld_wait();
      tcgen05_fence_before();  // my part of O is read: region g is free for the next pair's S once all 8 warps arrive
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_free[g]);
      if (tok < kAtT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk;
          __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            ph[t] = __floats2half2_rn(__uint_as_float(o[j >> 1][(8 * j + 2 * t) & 15]) * inv,
                                      __uint_as_float(o[j >> 1][(8 * j + 2 * t + 1) & 15]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * j) = pk;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, 512);
}

#endif  // EFFOCR_AB
// ------------------------------------------------------------------ one TMEM pass, two key blocks (default)
// Output tile through TMA (kTmaOut): each softmax warp stages its 32 rows x 64 fp16 in the (dead) first P chunk of its group,
// SWIZZLE_128B, and one lane issues a 4-D tensor store {64 columns, 32 tokens, 1 image} -- tokens >= 197 are clipped by the
// tensor map.  The per-thread row stores it replaces (eight 16-byte pieces 768 B apart per warp instruction) cost 20 % of the
// kernel (tools/att_ablate.py: 160 -> 129 us without them).
__device__ __forceinline__ void at_tma_store_4d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1, int32_t c2,
                                                int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// 16-byte store to shared memory by 32-bit shared address (the generic-pointer form compiles to ST.E.128 with 64-bit
// address arithmetic per store)
__device__ __forceinline__ void at_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void at_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void at_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void at_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// kAblate (timing experiments only, results are wrong): 1 = no global stores, 2 = no P stores to shared memory,
// 4 = exponentials replaced by an FMA (no MUFU), 8 = no maximum search
template <bool kDbg, int kAblate = 0, bool kTmaOut = true>
__global__ void __launch_bounds__(kAtThreads, 1)
attention_tc2b_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                      const __grid_constant__ CUtensorMap tma_o,
                    __half* __restrict__ out, int batch, int H, float scale_log2e, long long* dbg, int reverse) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                              // [2 groups] x 16 KB
  // K single-buffered, V double-buffered: K(i+1) is fetched between S1(i) and S0(i+1), half a unit apart, while V(i) stays
  // live until P.V of group 1 retires in the NEXT iteration -- with one V buffer the issuing warp waited 1 700 cycles per
  // unit for V (tools/att_timeline.py)
  uint8_t* smem_k = smem_q + 2 * kAtQBytes;            // 26 KB
  uint8_t* smem_v = smem_k + kAtKVBytes;               // [2 unit parities] x 26 KB
  uint8_t* smem_p = smem_v + 2 * kAtKVBytes;           // [2 groups] x 52 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + 2 * kAtPBytes);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* k_full = bars + 4;
  uint64_t* k_empty = bars + 5;
  uint64_t* v_full = bars + 6;     // [2]
  uint64_t* v_empty = bars + 8;    // [2]
  uint64_t* s_full = bars + 10;    // [2]
  uint64_t* p_full = bars + 12;    // [2]
  uint64_t* o_full = bars + 14;    // [2]
  uint64_t* t_free = bars + 16;    // [2]
  uint64_t* pb_full = bars + 18;   // [2] second key block of P written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int D = H * 64;
  const int num_bh = batch * H;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    if (kTmaOut) tma_prefetch_desc(&tma_o);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pb_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&t_free[i], 4);
    }
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ---------------- TMA producer
    if (elect_one_sync()) {
      int it = 0;  // pair counter of this CTA
      for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
        const int bhr = reverse ? num_bh - 1 - bh : bh;  // walk the units backwards: see vit.cu (L2-friendly kernel order)
        const int b = bhr / H, h = bhr % H;
        const int row0 = b * kAtT;
        const int s = it & 1;
        const uint32_t par = it & 1, par2 = (it >> 1) & 1;
        mbar_wait(k_empty, par ^ 1);
        mbar_arrive_expect_tx(k_full, kAtKVBytes);
        tma_load_2d(&tma_kv, k_full, smem_k, D + h * 64, row0);
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&q_empty[g], par ^ 1);
          mbar_arrive_expect_tx(&q_full[g], kAtQBytes);
          tma_load_2d(&tma_q, &q_full[g], smem_q + g * kAtQBytes, h * 64, row0 + g * 128);
        }
        mbar_wait(&v_empty[s], par2 ^ 1);
        mbar_arrive_expect_tx(&v_full[s], kAtKVBytes);
        tma_load_2d(&tma_kv, &v_full[s], smem_v + s * kAtKVBytes, 2 * D + h * 64, row0);
      }
    }
  } else if (warp_idx == 1) {
    // ---------------- MMA issuer (warp-uniform loop, tcgen05 issue predicated on one elected lane).
    // Issue order per pair i:  S0(i), PV1a(i-1), PV1b(i-1), S1(i), PV0a(i), PV0b(i).
    {
      const bool leader_lane = elect_one_sync();
      constexpr uint32_t idesc_s = make_idesc_f16(128, kAtN);
      constexpr uint32_t idesc_o = make_idesc_f16_bmn(128, 64);
      const uint32_t q_base = smem_u32(smem_q), k_base = smem_u32(smem_k), v_base = smem_u32(smem_v), p_base = smem_u32(smem_p);
      // optional wait-time accounting (tools/att_timeline.py): cycles the issuing warp spends in each mbarrier wait
      long long acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      auto twait = [&](uint64_t* bar, uint32_t par, int slot) {
        if constexpr (kDbg) {
          const long long t0 = clock64();
          mbar_wait(bar, par);
          acc[slot] += clock64() - t0;
        } else {
          mbar_wait(bar, par);
        }
      };
      auto issue_s = [&](int g, int it) {
        const uint32_t par = it & 1;
        twait(&q_full[g], par, 1 + 4 * g);
        twait(&t_free[g], par ^ 1, 2 + 4 * g);  // region g (S / O1 / O2 columns) drained by the previous pair's epilogue
        tcgen05_fence_after();
        const uint64_t dq = make_sw128_kmajor_desc(q_base + g * kAtQBytes);
        const uint64_t dk = make_sw128_kmajor_desc(k_base);
        if (leader_lane) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + g * kAtTmemRegion, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
          umma_commit(&q_empty[g]);
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      };
      // O1 (TMEM columns 0..63 of the region) = P[:, 0:128] V[0:128];  O2 (columns 64..127) = P[:, 128:208] V[128:208]
      auto issue_pv = [&](int g, int it) {
        const uint32_t par = it & 1;
        const uint32_t pbase = p_base + g * kAtPBytes;
        const uint32_t vb = v_base + (it & 1) * kAtKVBytes;  // V of this unit
        twait(&p_full[g], par, 3 + 4 * g);  // key block A of P written, S columns 0..127 consumed by all four warps
        tcgen05_fence_after();
        if (leader_lane) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t dp = make_sw128_kmajor_desc(pbase + (ks >> 2) * kAtQBytes) + 2 * (ks & 3);
            const uint64_t dv = make_sw128_mnmajor_desc(vb + ks * 2048);
            umma_f16(tmem_base + g * kAtTmemRegion, dp, dv, idesc_o, ks ? 1u : 0u);
          }
        }
        __syncwarp();
        twait(&pb_full[g], par, 4 + 4 * g);  // key block B written, all of S consumed
        tcgen05_fence_after();
        if (leader_lane) {
#pragma unroll
          for (int ks = 8; ks < 12; ++ks) {
            const uint64_t dp = make_sw128_kmajor_desc(pbase + (ks >> 2) * kAtQBytes) + 2 * (ks & 3);
            const uint64_t dv = make_sw128_mnmajor_desc(vb + ks * 2048);
            umma_f16(tmem_base + g * kAtTmemRegion + 64, dp, dv, idesc_o, ks > 8 ? 1u : 0u);
          }
          umma_f16(tmem_base + g * kAtTmemRegion + 64, make_sw32_kmajor_desc(pbase + kAtPMain),
                   make_sw128_mnmajor_desc(vb + 12 * 2048), idesc_o, 1u);
          umma_commit(&o_full[g]);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (leader_lane) umma_commit(bar);
        __syncwarp();
      };
      int it = 0;
      const long long t_begin = kDbg ? clock64() : 0;
      for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
        twait(k_full, it & 1, 0);
        issue_s(0, it);
        if (it > 0) {
          issue_pv(1, it - 1);   // V(it-1) lives in the other V buffer, released right here
          commit(&v_empty[(it - 1) & 1]);
        }
        issue_s(1, it);
        commit(k_empty);         // K(it+1) may land now; it is needed at S0(it+1)
        twait(&v_full[it & 1], (it >> 1) & 1, 9);
        issue_pv(0, it);
      }
      if (it > 0) {
        issue_pv(1, it - 1);
        commit(&v_empty[(it - 1) & 1]);
      }
      if (kDbg && blockIdx.x == 0 && leader_lane) {
        for (int i = 0; i < 10; ++i) dbg[i] = acc[i];
        dbg[10] = clock64() - t_begin;
        dbg[11] = it;
      }
    }
  } else if (warp_idx >= 4) {
    // ---------------- softmax + epilogue: group g = query tile, thread = query row.  ONE pass over S in TMEM
    // (tcgen05.ld moves ~64 B/clk/SM and reading S twice bounded the two-pass kernel): the keys are split into block A
    // (0..127) and block B (128..196); a block is pulled into registers once, reduced to its own maximum, exponentiated
    // against THAT maximum and written to the P tile; P_A V_A and P_B V_B accumulate into separate TMEM tiles
    // (flash-attention style) and the epilogue combines them with the factors 2^(m_A - m), 2^(m_B - m).
    const int g = (warp_idx - 4) >> 2;
    const int q = warp_idx & 3;
    const int r = q * 32 + lane;  // row inside the 128-row tile
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * kAtTmemRegion;
    uint8_t* pmain = smem_p + g * kAtPBytes + r * 128;
    uint8_t* ptail = smem_p + g * kAtPBytes + kAtPMain + r * 32;
    const uint32_t pmain_s = smem_u32(pmain), ptail_s = smem_u32(ptail);
    auto ex2x = [&](float x) { return (kAblate & 4) ? fmaf(x, 0.001f, 0.5f) : ex2_approx(x); };
    int it = 0;
    long long a_s = 0, a_c = 0, a_o = 0, a_e = 0;  // dbg: cycles in the s_full wait / softmax / o_full wait / epilogue
    for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
      const int bhr = reverse ? num_bh - 1 - bh : bh;
      const int b = bhr / H, h = bhr % H;
      const uint32_t par = it & 1;
      const long long ts0 = kDbg ? clock64() : 0;
      mbar_wait(&s_full[g], par);
      const long long ts1 = kDbg ? clock64() : 0;
      tcgen05_fence_after();
      if (kTmaOut) {  // the previous unit's output tile has left the first P chunk
        if (lane == 0) at_store_wait_read0();
        __syncwarp();
      }
      float mA = 0.f, mB = 0.f, lA = 1.f, lB = 1.f;
      // rows 224..255 (group 1, lane quarter 3) do not exist (197 tokens): that warp only keeps the hand-shakes going;
      // whatever its P rows hold feeds only output rows that are never stored
      const bool live = !(g == 1 && q == 3);
      if (live) {  // ---- block A: keys 0..127
        uint32_t v[8][16];
#pragma unroll
        for (int c = 0; c < 8; ++c) tmem_ld_32x32b_x16(trow + c * 16, v[c]);
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int e = 0; e < 16; ++e) m4[e & 3] = fmaxf(m4[e & 3], __uint_as_float(v[c][e]));
        mA = (kAblate & 8) ? __uint_as_float(v[0][0]) : fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float moff = mA * scale_log2e;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 pk[2];
          __half2* ph = reinterpret_cast<__half2*>(pk);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float p0 = ex2x(fmaf(__uint_as_float(v[c][2 * t]), scale_log2e, -moff));
            const float p1 = ex2x(fmaf(__uint_as_float(v[c][2 * t + 1]), scale_log2e, -moff));
            s4[(2 * t) & 3] += p0;
            s4[(2 * t + 1) & 3] += p1;
            ph[t] = __floats2half2_rn(p0, p1);
          }
          const uint32_t chunk = pmain_s + (c >> 2) * kAtQBytes;
          const int piece = (c & 3) * 2;
          if (!(kAblate & 2)) {
            at_sts128(chunk + ((piece ^ (r & 7)) << 4), pk[0]);
            at_sts128(chunk + (((piece + 1) ^ (r & 7)) << 4), pk[1]);
          } else if (pk[0].x == 0x12345u) {
            at_sts128(chunk, pk[1]);
          }
        }
        lA = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (live) {  // ---- block B: keys 128..207 (valid up to 196)
        uint32_t v[5][16];
#pragma unroll
        for (int c = 0; c < 5; ++c) tmem_ld_32x32b_x16(trow + 128 + c * 16, v[c]);
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 5; ++c)
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c < 4 || e < 5) m4[e & 3] = fmaxf(m4[e & 3], __uint_as_float(v[c][e]));
        mB = (kAblate & 8) ? __uint_as_float(v[0][0]) : fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float moff = mB * scale_log2e;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          uint4 pk[2];
          __half2* ph = reinterpret_cast<__half2*>(pk);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const bool ok0 = c < 4 || 2 * t < 5, ok1 = c < 4 || 2 * t + 1 < 5;  // keys >= 197 masked
            const float p0 = ok0 ? ex2x(fmaf(__uint_as_float(v[c][2 * t]), scale_log2e, -moff)) : 0.f;
            const float p1 = ok1 ? ex2x(fmaf(__uint_as_float(v[c][2 * t + 1]), scale_log2e, -moff)) : 0.f;
            s4[(2 * t) & 3] += p0;
            s4[(2 * t + 1) & 3] += p1;
            ph[t] = __floats2half2_rn(p0, p1);
          }
          if (kAblate & 2) {
            if (pk[0].x == 0x12345u) *reinterpret_cast<uint4*>(pmain) = pk[1];
          } else if (c < 4) {
            const uint32_t chunk = pmain_s + 2 * kAtQBytes;
            const int piece = c * 2;
            at_sts128(chunk + ((piece ^ (r & 7)) << 4), pk[0]);
            at_sts128(chunk + (((piece + 1) ^ (r & 7)) << 4), pk[1]);
          } else {
            at_sts128(ptail_s + ((0 ^ ((r >> 2) & 1)) << 4), pk[0]);  // SWIZZLE_32B: bit 4 ^= bit 7
            at_sts128(ptail_s + ((1 ^ ((r >> 2) & 1)) << 4), pk[1]);
          }
        }
        lB = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pb_full[g]);
      // ---- epilogue: (O1 * aA + O2 * aB) / (lA * aA + lB * aB) -> fp16 -> global
      const float m = fmaxf(mA, mB);
      const float aA = ex2_approx((mA - m) * scale_log2e), aB = ex2_approx((mB - m) * scale_log2e);
      const float inv = 1.0f / fmaf(lA, aA, lB * aB);
      const float wA = aA * inv, wB = aB * inv;
      const long long ts2 = kDbg ? clock64() : 0;
      mbar_wait(&o_full[g], par);
      const long long ts3 = kDbg ? clock64() : 0;
      tcgen05_fence_after();
      const int tok = g * 128 + r;
      __half* dst = out + (static_cast<long long>(b) * kAtT + tok) * D + h * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o1[32], o2[32];
        tmem_ld_32x32b_x32(trow + hh * 32, o1);
        tmem_ld_32x32b_x32(trow + 64 + hh * 32, o2);
        tmem_ld_wait();
        if (hh == 1) {  // O fully read: region g is free for the next pair's S
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_free[g]);
        }
        if (kTmaOut || tok < kAtT) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int e = 8 * j + 2 * t;
              ph[t] = __floats2half2_rn(fmaf(__uint_as_float(o1[e]), wA, __uint_as_float(o2[e]) * wB),
                                        fmaf(__uint_as_float(o1[e + 1]), wA, __uint_as_float(o2[e + 1]) * wB));
            }
            if (kTmaOut) at_sts128(pmain_s + (((hh * 4 + j) ^ (r & 7)) << 4), pk);
            else if (!(kAblate & 1) || pk.x == 0x12345u) *reinterpret_cast<uint4*>(dst + hh * 32 + 8 * j) = pk;
          }
        }
      }
      if (kTmaOut) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && live) {
          at_tma_store_4d(&tma_o, smem_p + g * kAtPBytes + q * 32 * 128, h * 64, g * 128 + q * 32, b, 0);
          at_store_commit();
        }
      }
      if constexpr (kDbg) {
        a_s += ts1 - ts0; a_c += ts2 - ts1; a_o += ts3 - ts2; a_e += clock64() - ts3;
      }
    }
    if (kTmaOut && lane == 0) at_store_wait_all0();
    if (kDbg && blockIdx.x == 0 && q == 0 && lane == 0) {
      dbg[12 + 4 * g] = a_s; dbg[13 + 4 * g] = a_c; dbg[14 + 4 * g] = a_o; dbg[15 + 4 * g] = a_e;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, 512);
}

#ifdef EFFOCR_AB  // A/B variant, not part of the product build (python -m effocr_b200.build with EFFOCR_AB=1)
// ------------------------------------------------------------------ one TMEM pass, three key blocks, 16 softmax warps
// The two-block kernel above is bound by its softmax warps: tools/att_timeline.py shows each of them busy ~5 700 of the
// 7 300 cycles a (image, head) unit takes (3 900 softmax + 1 800 epilogue), one exponential every 8 cycles per warp being
// the MUFU rate a single warp can draw.  Here every query row is served by TWO threads in different warps: warp X takes
// keys 0..63 and then 64..127, warp Y keys 128..196, each key block with its own maximum and its own P.V accumulator
// (O1 | O2 | O3 alias the S columns of their own key block, so the TMEM budget is unchanged); the epilogue is split by
// output columns (X: 0..31, Y: 32..63) after the two threads exchanged (max, sum) per block through shared memory.
// 64 live scores per thread instead of 128 keeps the 640-thread CTA inside 102 registers.
constexpr int kAt3Threads = 128 + 16 * 32;
constexpr int kAt3StatBytes = 8192;  // [2 groups][6 values: m1 l1 m2 l2 m3 l3][128 rows] fp32 (6 KB used)
constexpr int kAt3SmemBytes = kAtSmemBytes + kAt3StatBytes;

template <bool kDbg>
__global__ void __launch_bounds__(kAt3Threads, 1)
attention_tc3b_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                      __half* __restrict__ out, int batch, int H, float scale_log2e, long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                              // [2 groups] x 16 KB
  uint8_t* smem_k = smem_q + 2 * kAtQBytes;            // [2 unit parities] x 26 KB
  uint8_t* smem_v = smem_k + 2 * kAtKVBytes;           // 26 KB
  uint8_t* smem_p = smem_v + kAtKVBytes;               // [2 groups] x 52 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + 2 * kAtPBytes);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* k_full = bars + 4;     // [2]
  uint64_t* k_empty = bars + 6;    // [2]
  uint64_t* v_full = bars + 8;
  uint64_t* v_empty = bars + 9;
  uint64_t* s_full = bars + 10;    // [2]
  uint64_t* p1_full = bars + 12;   // [2] keys 0..63 of P written
  uint64_t* p2_full = bars + 14;   // [2] keys 64..127
  uint64_t* p3_full = bars + 16;   // [2] keys 128..207
  uint64_t* o_full = bars + 18;    // [2]
  uint64_t* t_free = bars + 20;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  float* stats = reinterpret_cast<float*>(smem_p + 2 * kAtPBytes + 256);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int D = H * 64;
  const int num_bh = batch * H;

  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p1_full[i], 4);
      mbar_init(&p2_full[i], 4);
      mbar_init(&p3_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&t_free[i], 8);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ---------------- TMA producer (same schedule as the two-block kernel)
    if (elect_one_sync()) {
      int it = 0;
      for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
        const int b = bh / H, h = bh % H;
        const int row0 = b * kAtT;
        const int s = it & 1;
        const uint32_t par = it & 1, par2 = (it >> 1) & 1;
        mbar_wait(&k_empty[s], par2 ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kAtKVBytes);
        tma_load_2d(&tma_kv, &k_full[s], smem_k + s * kAtKVBytes, D + h * 64, row0);
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&q_empty[g], par ^ 1);
          mbar_arrive_expect_tx(&q_full[g], kAtQBytes);
          tma_load_2d(&tma_q, &q_full[g], smem_q + g * kAtQBytes, h * 64, row0 + g * 128);
        }
        mbar_wait(v_empty, par ^ 1);
        mbar_arrive_expect_tx(v_full, kAtKVBytes);
        tma_load_2d(&tma_kv, v_full, smem_v, 2 * D + h * 64, row0);
      }
    }
  } else if (warp_idx == 1) {
    // ---------------- MMA issuer (warp-uniform loop).  Issue order per unit i:  S0(i), PV1(i-1), S1(i), PV0(i).
    const bool leader_lane = elect_one_sync();
    constexpr uint32_t idesc_s = make_idesc_f16(128, kAtN);
    constexpr uint32_t idesc_o = make_idesc_f16_bmn(128, 64);
    const uint32_t q_base = smem_u32(smem_q), k_base = smem_u32(smem_k), v_base = smem_u32(smem_v), p_base = smem_u32(smem_p);
    long long acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    auto twait = [&](uint64_t* bar, uint32_t par, int slot) {
      if constexpr (kDbg) {
        const long long t0 = clock64();
        mbar_wait(bar, par);
        acc[slot] += clock64() - t0;
      } else {
        mbar_wait(bar, par);
      }
    };
    auto issue_s = [&](int g, int it) {
      const uint32_t par = it & 1;
      const int s = it & 1;
      twait(&q_full[g], par, 1 + 4 * g);
      twait(&t_free[g], par ^ 1, 2 + 4 * g);  // region g (S / O1 / O2 / O3 columns) drained by the previous unit's epilogue
      tcgen05_fence_after();
      const uint64_t dq = make_sw128_kmajor_desc(q_base + g * kAtQBytes);
      const uint64_t dk = make_sw128_kmajor_desc(k_base + s * kAtKVBytes);
      if (leader_lane) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + g * kAtTmemRegion, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit(&q_empty[g]);
        umma_commit(&s_full[g]);
      }
      __syncwarp();
    };
    // O1 (columns 0..63 of the region) = P[:, 0:64] V[0:64], O2 (64..127) = P[:, 64:128] V[64:128],
    // O3 (128..191) = P[:, 128:208] V[128:208]; each waits only for its own key block of P
    auto issue_pv = [&](int g, int it) {
      const uint32_t par = it & 1;
      const uint32_t pbase = p_base + g * kAtPBytes;
      const uint32_t tO = tmem_base + g * kAtTmemRegion;
      twait(&p1_full[g], par, 3 + 4 * g);
      tcgen05_fence_after();
      if (leader_lane) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tO, make_sw128_kmajor_desc(pbase) + 2 * ks, make_sw128_mnmajor_desc(v_base + ks * 2048), idesc_o, ks ? 1u : 0u);
      }
      __syncwarp();
      twait(&p3_full[g], par, 4 + 4 * g);
      tcgen05_fence_after();
      if (leader_lane) {
#pragma unroll
        for (int ks = 8; ks < 12; ++ks)
          umma_f16(tO + 128, make_sw128_kmajor_desc(pbase + 2 * kAtQBytes) + 2 * (ks & 3), make_sw128_mnmajor_desc(v_base + ks * 2048),
                   idesc_o, ks > 8 ? 1u : 0u);
        umma_f16(tO + 128, make_sw32_kmajor_desc(pbase + kAtPMain), make_sw128_mnmajor_desc(v_base + 12 * 2048), idesc_o, 1u);
      }
      __syncwarp();
      twait(&p2_full[g], par, 4 + 4 * g);
      tcgen05_fence_after();
      if (leader_lane) {
#pragma unroll
        for (int ks = 4; ks < 8; ++ks)
          umma_f16(tO + 64, make_sw128_kmajor_desc(pbase + kAtQBytes) + 2 * (ks & 3), make_sw128_mnmajor_desc(v_base + ks * 2048),
                   idesc_o, ks > 4 ? 1u : 0u);
        umma_commit(&o_full[g]);
      }
      __syncwarp();
    };
    auto commit = [&](uint64_t* bar) {
      if (leader_lane) umma_commit(bar);
      __syncwarp();
    };
    int it = 0;
    const long long t_begin = kDbg ? clock64() : 0;
    for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
      const int s = it & 1;
      twait(&k_full[s], (it >> 1) & 1, 0);
      issue_s(0, it);
      if (it > 0) {
        issue_pv(1, it - 1);  // V(it-1) is still resident: its buffer is released right here
        commit(v_empty);
      }
      issue_s(1, it);
      commit(&k_empty[s]);
      twait(v_full, it & 1, 9);  // V(it), reloaded after PV1(it-1) retired
      issue_pv(0, it);
    }
    if (it > 0) {
      issue_pv(1, it - 1);
      commit(v_empty);
    }
    if (kDbg && blockIdx.x == 0 && leader_lane) {
      for (int i = 0; i < 10; ++i) dbg[i] = acc[i];
      dbg[10] = clock64() - t_begin;
      dbg[11] = it;
    }
  } else if (warp_idx >= 4) {
    // ---------------- softmax + epilogue: group g = query tile, (warp X, warp Y) = the two threads of a query row
    const int idx = warp_idx - 4;
    const int g = idx >> 3;
    const int half = (idx >> 2) & 1;  // 0: warp X (keys 0..127, output columns 0..31), 1: warp Y (keys 128..196, columns 32..63)
    const int q = idx & 3;            // == warp_idx & 3: the TMEM lane quarter this warp may read
    const int r = q * 32 + lane;      // row inside the 128-row tile
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * kAtTmemRegion;
    uint8_t* pmain = smem_p + g * kAtPBytes + r * 128;
    uint8_t* ptail = smem_p + g * kAtPBytes + kAtPMain + r * 32;
    float* st = stats + g * 6 * 128 + r;  // value s of this row at st[s * 128]
    const int pair_bar = 1 + g * 4 + q;
    // rows 224..255 (group 1, lane quarter 3) do not exist (197 tokens): those warps only keep the hand-shakes going
    const bool live = !(g == 1 && q == 3);
    long long a_s = 0, a_c = 0, a_o = 0, a_e = 0;  // dbg: cycles in the s_full wait / softmax / o_full wait / epilogue
    // one 64-key block: scores from TMEM columns tcol.., P to the SWIZZLE_128B chunk, returns (max, sum)
    auto block64 = [&](uint32_t tcol, uint8_t* chunk, float& m_out, float& l_out) {
      uint32_t v[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x16(trow + tcol + c * 16, v[c]);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 16; ++e) m4[e & 3] = fmaxf(m4[e & 3], __uint_as_float(v[c][e]));
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const float moff = m * scale_log2e;
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 pk[2];
        __half2* ph = reinterpret_cast<__half2*>(pk);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[c][2 * t]), scale_log2e, -moff));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[c][2 * t + 1]), scale_log2e, -moff));
          s4[(2 * t) & 3] += p0;
          s4[(2 * t + 1) & 3] += p1;
          ph[t] = __floats2half2_rn(p0, p1);
        }
        *reinterpret_cast<uint4*>(chunk + (((2 * c) ^ (r & 7)) << 4)) = pk[0];
        *reinterpret_cast<uint4*>(chunk + (((2 * c + 1) ^ (r & 7)) << 4)) = pk[1];
      }
      m_out = m;
      l_out = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    };
    auto publish = [&](uint64_t* bar) {  // P block written: visible to the MMA (async proxy), then signal
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    int it = 0;
    for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x, ++it) {
      const int b = bh / H, h = bh % H;
      const uint32_t par = it & 1;
      const long long ts0 = kDbg ? clock64() : 0;
      mbar_wait(&s_full[g], par);
      const long long ts1 = kDbg ? clock64() : 0;
      tcgen05_fence_after();
      if (half == 0) {
        float m1 = 0.f, l1 = 1.f, m2 = 0.f, l2 = 1.f;
        if (live) block64(0, pmain, m1, l1);
        publish(&p1_full[g]);
        if (live) block64(64, pmain + kAtQBytes, m2, l2);
        publish(&p2_full[g]);
        st[0] = m1; st[128] = l1; st[256] = m2; st[384] = l2;
      } else {
        float m3 = 0.f, l3 = 1.f;
        if (live) {  // keys 128..191 as a full block, then the five valid keys 192..196 of the SWIZZLE_32B tail
          uint32_t v[4][16], tl[8];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x16(trow + 128 + c * 16, v[c]);
          tmem_ld_32x32b_x8(trow + 192, tl);
          tmem_ld_wait();
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int e = 0; e < 16; ++e) m4[e & 3] = fmaxf(m4[e & 3], __uint_as_float(v[c][e]));
#pragma unroll
          for (int e = 0; e < 5; ++e) m4[e & 3] = fmaxf(m4[e & 3], __uint_as_float(tl[e]));
          m3 = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          const float moff = m3 * scale_log2e;
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
          uint8_t* chunk = pmain + 2 * kAtQBytes;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 pk[2];
            __half2* ph = reinterpret_cast<__half2*>(pk);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[c][2 * t]), scale_log2e, -moff));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[c][2 * t + 1]), scale_log2e, -moff));
              s4[(2 * t) & 3] += p0;
              s4[(2 * t + 1) & 3] += p1;
              ph[t] = __floats2half2_rn(p0, p1);
            }
            *reinterpret_cast<uint4*>(chunk + (((2 * c) ^ (r & 7)) << 4)) = pk[0];
            *reinterpret_cast<uint4*>(chunk + (((2 * c + 1) ^ (r & 7)) << 4)) = pk[1];
          }
          {
            float pt[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              pt[e] = e < 5 ? ex2_approx(fmaf(__uint_as_float(tl[e]), scale_log2e, -moff)) : 0.f;  // keys >= 197 masked
              s4[e & 3] += pt[e];
            }
            uint4 pk0;
            __half2* ph = reinterpret_cast<__half2*>(&pk0);
#pragma unroll
            for (int t = 0; t < 4; ++t) ph[t] = __floats2half2_rn(pt[2 * t], pt[2 * t + 1]);
            *reinterpret_cast<uint4*>(ptail + ((0 ^ ((r >> 2) & 1)) << 4)) = pk0;  // SWIZZLE_32B: bit 4 ^= bit 7
            *reinterpret_cast<uint4*>(ptail + ((1 ^ ((r >> 2) & 1)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
          }
          l3 = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        }
        publish(&p3_full[g]);
        st[512] = m3; st[640] = l3;
      }
      named_bar_sync_at(pair_bar, 64);  // both threads of the row have published their (max, sum) pairs
      // ---- epilogue: (O1 a1 + O2 a2 + O3 a3) / (l1 a1 + l2 a2 + l3 a3) -> fp16 -> global, 32 output columns per thread
      const float m1 = st[0], l1 = st[128], m2 = st[256], l2 = st[384], m3 = st[512], l3 = st[640];
      const float m = fmaxf(fmaxf(m1, m2), m3);
      const float a1 = ex2_approx((m1 - m) * scale_log2e), a2 = ex2_approx((m2 - m) * scale_log2e),
                  a3 = ex2_approx((m3 - m) * scale_log2e);
      const float inv = 1.0f / fmaf(l1, a1, fmaf(l2, a2, l3 * a3));
      const float w1 = a1 * inv, w2 = a2 * inv, w3 = a3 * inv;
      const long long ts2 = kDbg ? clock64() : 0;
      mbar_wait(&o_full[g], par);
      const long long ts3 = kDbg ? clock64() : 0;
      tcgen05_fence_after();
      const int tok = g * 128 + r;
      __half* dst = out + (static_cast<long long>(b) * kAtT + tok) * D + h * 64 + half * 32;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o1[16], o2[16], o3[16];
        tmem_ld_32x32b_x16(trow + half * 32 + hh * 16, o1);
        tmem_ld_32x32b_x16(trow + 64 + half * 32 + hh * 16, o2);
        tmem_ld_32x32b_x16(trow + 128 + half * 32 + hh * 16, o3);
        tmem_ld_wait();
        if (hh == 1) {  // my part of O is read: region g is free for the next unit's S once all eight warps arrive
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_free[g]);
        }
        if (tok < kAtT) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4 pk;
            __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int e = 8 * j + 2 * t;
              ph[t] = __floats2half2_rn(
                  fmaf(__uint_as_float(o1[e]), w1, fmaf(__uint_as_float(o2[e]), w2, __uint_as_float(o3[e]) * w3)),
                  fmaf(__uint_as_float(o1[e + 1]), w1, fmaf(__uint_as_float(o2[e + 1]), w2, __uint_as_float(o3[e + 1]) * w3)));
            }
            *reinterpret_cast<uint4*>(dst + hh * 16 + 8 * j) = pk;
          }
        }
      }
      if constexpr (kDbg) {
        a_s += ts1 - ts0; a_c += ts2 - ts1; a_o += ts3 - ts2; a_e += clock64() - ts3;
      }
    }
    if (kDbg && blockIdx.x == 0 && q == 0 && lane == 0) {
      const int o = 12 + 4 * g + 8 * half;
      dbg[o] = a_s; dbg[o + 1] = a_c; dbg[o + 2] = a_o; dbg[o + 3] = a_e;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, 512);
}

#endif  // EFFOCR_AB
}  // namespace effocr
