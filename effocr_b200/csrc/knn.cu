// Exact inner-product top-k against a flat glyph-prototype index (K10 in SURVEY.md section 2a).
// Replaces faiss.IndexFlatIP.search reached through pytorch_metric_learning's FaissKNN at
// /root/reference/infer_effocr.py:184-187,317 (k = 10) and infer_effocr_onnx_multi.py:372 (k = 1).
//
// Scores must be fp32-accurate: with random-init encoders the top-1 / top-2 margin is ~1e-5
// (SURVEY.md App. D), far below fp16 or tf32 resolution.  Each fp32 vector is therefore split into
// two fp16 terms  x = hi + lo * 2^-11  (22 mantissa bits) and the score is evaluated on the tensor
// cores as
//        q . x  ~=  qhi.xhi  +  2^-11 * ( qhi.xlo + qlo.xhi )          (dropped term ~ 2^-22)
// by ONE tcgen05 GEMM over the concatenated operands  A' = [qhi | qhi | qlo],  B' = [xhi | xlo | xhi]
// whose first K/3 feeds TMEM accumulator 1 and the rest accumulator 2.  The epilogue keeps a running
// top-32 short-list per query in shared memory (the N x B score matrix never exists in HBM), a
// second kernel merges the per-CTA short-lists, re-evaluates the 32 survivors with plain fp32 FMAs
// against the fp32 index and emits the top-k by (score desc, id asc).
#include <math.h>

#include <vector>

#include "../../include/effocr_b200.h"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace effocr {

constexpr int kKnnBM = 128, kKnnBN = 128, kKnnBK = 64;
constexpr int kKnnCap = 32;  // short-list length per (query, CTA half)
constexpr int kKnnStages = 4;
constexpr int kKnnThreads = 384;
constexpr int kKnnStageBytes = (kKnnBM + kKnnBN) * kKnnBK * 2;            // 32 KB
constexpr int kKnnListBytes = 256 * kKnnCap * 8;                           // 64 KB
constexpr int kKnnScoreBytes = 256 * 32 * 4;                               // 32 KB: one 32-column chunk of scores per epilogue thread
constexpr int kKnnSmemBytes = kKnnStages * kKnnStageBytes + kKnnListBytes + kKnnScoreBytes + 512 + 1024;

// x fp32 [rows, D] -> fp16 [rows_pad, 3 * Dp]:  mode 0 (queries) [hi | hi | lo], mode 1 (index) [hi | lo | hi];
// lo = fp16((x - hi) * 2048); padding rows / columns are zero.
__global__ void knn_split_kernel(const float* __restrict__ x, __half* __restrict__ out, int rows, int rows_pad, int D,
                                 int Dp, int mode) {
  const long long total = static_cast<long long>(rows_pad) * Dp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Dp);
    const long long r = i / Dp;
    float v = 0.f;
    if (r < rows && c < D) v = x[r * D + c];
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn((v - __half2float(hi)) * 2048.0f);
    __half* o = out + r * 3 * Dp + c;
    o[0] = hi;
    o[Dp] = mode == 0 ? hi : lo;
    o[2 * Dp] = mode == 0 ? lo : hi;
  }
}

// Per-thread short-list (one shared-memory column per epilogue thread, entries 256 floats apart).
// Candidates arrive in ascending id order; a full list evicts its minimum (largest id among equal
// minima) only for a strictly larger score, so equal scores keep the lower id.
struct KnnListState {
  int count;
  float minv;
  int minpos;
};
template <int CAP>
__device__ __noinline__ KnnListState knn_insert(KnnListState st, float sc, int n, float* lv, int* li) {
  if (st.count < CAP) {
    lv[st.count * 256] = sc;
    li[st.count * 256] = n;
    ++st.count;
    if (st.count < CAP) return st;
  } else {
    lv[st.minpos * 256] = sc;
    li[st.minpos * 256] = n;
  }
  st.minv = INFINITY;
  st.minpos = 0;
  int mid = -1;
#pragma unroll 4
  for (int s2 = 0; s2 < CAP; ++s2) {
    const float vv = lv[s2 * 256];
    const int ii = li[s2 * 256];
    if (vv < st.minv || (vv == st.minv && ii > mid)) { st.minv = vv; st.minpos = s2; mid = ii; }
  }
  return st;
}

// CAP: short-list length per (query, CTA half).  The union of the per-partition lists must contain the exact top-k:
// CAP = 16 for k <= 10 (six spare places absorb re-orderings by the ~2^-22 relative error of the split-fp16 scores),
// CAP = 32 otherwise.  Insertions cost O(CAP) and their number grows with CAP, so the short list is the epilogue's cost.
template <int CAP>
__global__ void __launch_bounds__(kKnnThreads, 1)
knn_gemm_topk_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_x, int B,
                     int N, int Dp, int n_tiles, int splits, float* __restrict__ part_val,
                     int* __restrict__ part_idx) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kKnnStages * kKnnBM * kKnnBK * 2;
  float* list_val = reinterpret_cast<float*>(smem + kKnnStages * kKnnStageBytes);
  int* list_idx = reinterpret_cast<int*>(list_val + 256 * kKnnCap);
  float* score_buf = reinterpret_cast<float*>(smem + kKnnStages * kKnnStageBytes + kKnnListBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kKnnStages * kKnnStageBytes + kKnnListBytes + kKnnScoreBytes);
  uint64_t* empty_bar = full_bar + kKnnStages;
  uint64_t* tfull_bar = empty_bar + kKnnStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (warp_idx == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_x);
  }
  if (warp_idx == 1 && elect_one_sync()) {
    for (int i = 0; i < kKnnStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    fence_barrier_init();
  }
  if (warp_idx == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int mt = blockIdx.x / splits, sp = blockIdx.x % splits;
  const int per = (n_tiles + splits - 1) / splits;
  const int nt0 = sp * per;
  const int nt1 = (nt0 + per < n_tiles) ? nt0 + per : n_tiles;
  const int m0 = mt * kKnnBM;
  const int kb1 = Dp / kKnnBK;     // k-blocks feeding accumulator 1 (hi.hi)
  const int num_kb = 3 * kb1;

  if (warp_idx == 0) {
    if (elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      for (int nt = nt0; nt < nt1; ++nt) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kKnnStageBytes);
          tma_load_2d(&tma_q, &full_bar[stage], smem_a + stage * kKnnBM * kKnnBK * 2, kb * kKnnBK, m0);
          tma_load_2d(&tma_x, &full_bar[stage], smem_b + stage * kKnnBN * kKnnBK * 2, kb * kKnnBK, nt * kKnnBN);
          if (++stage == kKnnStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // warp-uniform loop, tcgen05 issue predicated on one elected lane (descriptors stay in uniform registers)
    const bool leader_lane = elect_one_sync();
    constexpr uint32_t idesc = make_idesc_f16(kKnnBM, kKnnBN);
    const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
    int stage = 0; uint32_t phase = 0; int local = 0;
    for (int nt = nt0; nt < nt1; ++nt, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tcgen05_fence_after();
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint64_t da = make_sw128_kmajor_desc(a_base + stage * kKnnBM * kKnnBK * 2);
        const uint64_t db = make_sw128_kmajor_desc(b_base + stage * kKnnBN * kKnnBK * 2);
        const bool second = kb >= kb1;
        const uint32_t tmem_d = tmem_base + as * 256 + (second ? 128 : 0);
        const int kk = second ? kb - kb1 : kb;
        if (leader_lane) {
#pragma unroll
          for (int k = 0; k < kKnnBK / 16; ++k) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kk | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kKnnStages) { stage = 0; phase ^= 1; }
      }
      if (leader_lane) umma_commit(&tfull_bar[as]);
      __syncwarp();
    }
  } else if (warp_idx >= 4) {
    const int q = warp_idx & 3;
    const int half = (warp_idx - 4) >> 2;
    const int e = threadIdx.x - 128;  // epilogue thread id 0..255 -> list column
    float* lv = list_val + e;
    int* li = list_idx + e;
    KnnListState st;
    st.count = 0; st.minv = INFINITY; st.minpos = 0;
    int local = 0;
    for (int nt = nt0; nt < nt1; ++nt, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = half * 64 + c * 32;
        uint32_t v1[32], v2[32];
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256 + col0;
        tmem_ld_32x32b_x32(ta, v1);
        tmem_ld_32x32b_x32(ta + 128, v2);
        tmem_ld_wait();
        const int n0 = nt * kKnnBN + col0;
        // Candidates that can still enter this thread's list, judged against the list as it stands now (its minimum only
        // rises): a bit mask, the scores parked in shared memory.  The warp then loops while ANY lane has candidates left,
        // every lane taking its next one in ascending id order -- the same sequence of insertions per thread as testing
        // the 32 candidates one by one, but max-over-lanes(popcount) warp-wide insert calls per chunk instead of one for
        // every candidate that any of the 32 lanes wants (4x fewer at 10 000 prototypes).
        uint32_t cand = 0;
        float* sb = score_buf + e;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sc = fmaf(__uint_as_float(v2[i]), 4.8828125e-4f, __uint_as_float(v1[i]));
          sb[i * 256] = sc;
          if (n0 + i < N && (st.count < CAP || sc > st.minv)) cand |= 1u << i;
        }
        while (__any_sync(0xffffffffu, cand != 0)) {
          if (cand) {
            const int i = __ffs(cand) - 1;
            cand &= cand - 1;
            const float sc = sb[i * 256];
            if (st.count < CAP || sc > st.minv) st = knn_insert<CAP>(st, sc, n0 + i, lv, li);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    // flush this thread's short-list: partial slot = sp * 2 + half
    const int row = m0 + q * 32 + lane;
    if (row < B) {
      const long long o = (static_cast<long long>(row) * (splits * 2) + sp * 2 + half) * CAP;
      for (int s2 = 0; s2 < CAP; ++s2) {
        part_val[o + s2] = s2 < st.count ? lv[s2 * 256] : -INFINITY;
        part_idx[o + s2] = s2 < st.count ? li[s2 * 256] : -1;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc(tmem_base, 512);
}

// One warp per query: merge `nparts` short-lists down to the best `cap` (16 for k <= 10, else 32) by (approx score desc,
// id asc) -- the same margin argument as for the per-partition lists: the spare places absorb re-orderings by the ~2^-22
// relative error of the split-fp16 scores -- re-score those with fp32 FMAs against the fp32 index, order by (exact score
// desc, id asc), emit k.  The candidates are pulled into registers once (PER per lane) when they fit: the selection rounds
// are a dependent chain, and re-reading the lists from memory in every round made the kernel pure load latency.
constexpr int kKnnMergePer = 24;  // candidates per lane held in registers: nparts * cap <= 768
__global__ void __launch_bounds__(128) knn_merge_rerank_kernel(const float* __restrict__ part_val,
                                                               const int* __restrict__ part_idx, int nparts, int cap,
                                                               const float* __restrict__ q, const float* __restrict__ xb,
                                                               int B, int D, int k, float* __restrict__ out_val,
                                                               long long* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= B) return;
  const int total = nparts * cap;
  const float* pv = part_val + static_cast<long long>(row) * total;
  const int* pi = part_idx + static_cast<long long>(row) * total;
  const bool in_regs = total <= 32 * kKnnMergePer;
  float cv[kKnnMergePer];
  int ci[kKnnMergePer];
  if (in_regs) {
#pragma unroll
    for (int t = 0; t < kKnnMergePer; ++t) {
      const int j = lane + 32 * t;
      cv[t] = j < total ? pv[j] : -INFINITY;
      ci[t] = j < total ? pi[j] : -1;
    }
  }
  // iterative selection: `cap` rounds of warp arg-max with "already taken" tracked by last (value, id)
  float last_v = INFINITY;
  int last_i = -1;
  float my_v = -INFINITY;  // lane r ends up holding the r-th best candidate
  int my_i = -1;
  for (int r = 0; r < cap; ++r) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    if (in_regs) {
#pragma unroll
      for (int t = 0; t < kKnnMergePer; ++t) {
        const float v = cv[t];
        const int id = ci[t];
        // a live candidate strictly after (last_v, last_i) in (value desc, id asc) order that beats the running best
        const bool after = id >= 0 && ((v < last_v) || (v == last_v && id > last_i));
        if (after && (v > bv || (v == bv && id < bi))) { bv = v; bi = id; }
      }
    } else {
      for (int j = lane; j < total; j += 32) {
        const float v = pv[j];
        const int id = pi[j];
        if (id < 0) continue;
        const bool after = (v < last_v) || (v == last_v && id > last_i);
        if (!after) continue;
        if (v > bv || (v == bv && id < bi)) { bv = v; bi = id; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) break;  // candidates exhausted
    if (lane == r) { my_v = bv; my_i = bi; }
    last_v = bv;
    last_i = bi;
  }
  // exact fp32 re-score, one candidate at a time, warp-cooperative dot product
  const float* qr = q + static_cast<long long>(row) * D;
  float exact = -INFINITY;
  for (int c = 0; c < cap; ++c) {
    const int id = __shfl_sync(0xffffffffu, my_i, c);
    if (id < 0) continue;
    const float* xr = xb + static_cast<long long>(id) * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(qr[d], xr[d], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == c) exact = acc;
  }
  // rank among the 32 by (exact desc, id asc)
  int rank = 0;
  for (int c = 0; c < cap; ++c) {
    const float ov = __shfl_sync(0xffffffffu, exact, c);
    const int oi = __shfl_sync(0xffffffffu, my_i, c);
    if (oi < 0 || c == lane) continue;
    if (ov > exact || (ov == exact && oi < my_i)) ++rank;
  }
  if (my_i >= 0 && rank < k) {
    out_val[static_cast<long long>(row) * k + rank] = exact;
    out_idx[static_cast<long long>(row) * k + rank] = my_i;
  }
}

__global__ void knn_fill_kernel(float* v, long long* i, long long n) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t < n) { v[t] = -3.4028234663852886e38f; i[t] = -1; }  // faiss' "no result" convention for IP
}

// x / max(||x||_2, eps) per row (torch.nn.functional.normalize; infer_effocr.py:316)
__global__ void l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(xr[d], xr[d], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.0f / fmaxf(sqrtf(s), eps);
  for (int d = lane; d < D; d += 32) out[static_cast<long long>(row) * D + d] = xr[d] * inv;
}

struct KnnHandle {
  int N = 0, D = 0, Dp = 0, Npad = 0;
  float* xb = nullptr;      // [N, D] fp32 (re-rank)
  __half* xsplit = nullptr; // [Npad, 3 Dp]
  // search workspace (grown on demand)
  __half* qsplit = nullptr; size_t qsplit_cap = 0;
  float* part_val = nullptr; int* part_idx = nullptr; size_t part_cap = 0;
  ~KnnHandle() {
    cudaFree(xb); cudaFree(xsplit); cudaFree(qsplit); cudaFree(part_val); cudaFree(part_idx);
  }
};

}  // namespace effocr

using namespace effocr;

extern "C" int effocr_knn_create(const float* d_vectors, int n, int dim, effocr_knn_t* out) {
  if (!out) return fail(EFFOCR_ERR_INVALID, "knn_create: null out");
  *out = nullptr;
  EFFOCR_TRY(require_sm100());
  if (n < 0 || dim <= 0 || (n > 0 && !d_vectors)) return fail(EFFOCR_ERR_INVALID, "knn_create: bad arguments");
  KnnHandle* h = new KnnHandle();
  h->N = n; h->D = dim; h->Dp = (dim + 63) / 64 * 64; h->Npad = (n + kKnnBN - 1) / kKnnBN * kKnnBN;
  if (n > 0) {
    cudaError_t e;
    if ((e = cudaMalloc(&h->xb, size_t(n) * dim * 4)) != cudaSuccess ||
        (e = cudaMalloc(&h->xsplit, size_t(h->Npad) * 3 * h->Dp * 2)) != cudaSuccess ||
        (e = cudaMemcpy(h->xb, d_vectors, size_t(n) * dim * 4, cudaMemcpyDeviceToDevice)) != cudaSuccess) {
      delete h;
      return cuda_fail(e, "knn_create alloc/copy");
    }
    knn_split_kernel<<<148 * 8, 256>>>(h->xb, h->xsplit, n, h->Npad, dim, h->Dp, 1);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) { delete h; return cuda_fail(e, "knn_split_kernel"); }
  }
  *out = reinterpret_cast<effocr_knn_t>(h);
  return EFFOCR_OK;
}

extern "C" void effocr_knn_destroy(effocr_knn_t h) { delete reinterpret_cast<KnnHandle*>(h); }
extern "C" int effocr_knn_ntotal(effocr_knn_t h) { return h ? reinterpret_cast<KnnHandle*>(h)->N : 0; }
extern "C" int effocr_knn_dim(effocr_knn_t h) { return h ? reinterpret_cast<KnnHandle*>(h)->D : 0; }
extern "C" const float* effocr_knn_vectors(effocr_knn_t h) { return h ? reinterpret_cast<KnnHandle*>(h)->xb : nullptr; }

extern "C" int effocr_knn_search(effocr_knn_t handle, const float* d_queries, int nq, int k, float* d_dist,
                                 long long* d_idx, void* stream) {
  KnnHandle* h = reinterpret_cast<KnnHandle*>(handle);
  if (!h || nq < 0 || k <= 0) return fail(EFFOCR_ERR_INVALID, "knn_search: bad arguments");
  if (k > kKnnCap) return fail(EFFOCR_ERR_INVALID, "knn_search: k must be <= 32");
  if (nq == 0) return EFFOCR_OK;
  if (!d_queries || !d_dist || !d_idx) return fail(EFFOCR_ERR_INVALID, "knn_search: null buffer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long nout = static_cast<long long>(nq) * k;
  {
    KernelScope ks(PROF_MISC, s);
    knn_fill_kernel<<<static_cast<int>((nout + 255) / 256), 256, 0, s>>>(d_dist, d_idx, nout);
  }
  EFFOCR_CUDA(cudaGetLastError());
  if (h->N == 0) return EFFOCR_OK;

  const int m_tiles = (nq + kKnnBM - 1) / kKnnBM;
  const int n_tiles = h->Npad / kKnnBN;
  int splits = sm_count() / m_tiles;
  if (splits < 1) splits = 1;
  if (splits > n_tiles) splits = n_tiles;
  // drop empty splits: ceil-div distribution must leave every split at least one tile
  { const int per = (n_tiles + splits - 1) / splits; splits = (n_tiles + per - 1) / per; }
  const int nq_pad = m_tiles * kKnnBM;
  const size_t need_q = size_t(nq_pad) * 3 * h->Dp;
  if (need_q > h->qsplit_cap) {
    cudaFree(h->qsplit); h->qsplit = nullptr; h->qsplit_cap = 0;
    EFFOCR_CUDA(cudaMalloc(&h->qsplit, need_q * 2));
    h->qsplit_cap = need_q;
  }
  const size_t need_p = size_t(nq) * splits * 2 * kKnnCap;
  if (need_p > h->part_cap) {
    cudaFree(h->part_val); cudaFree(h->part_idx); h->part_val = nullptr; h->part_idx = nullptr; h->part_cap = 0;
    EFFOCR_CUDA(cudaMalloc(&h->part_val, need_p * 4));
    EFFOCR_CUDA(cudaMalloc(&h->part_idx, need_p * 4));
    h->part_cap = need_p;
  }
  {
    KernelScope ks(PROF_KNN_SPLIT, s);
    knn_split_kernel<<<148 * 4, 256, 0, s>>>(d_queries, h->qsplit, nq, nq_pad, h->D, h->Dp, 0);
  }
  EFFOCR_CUDA(cudaGetLastError());
  CUtensorMap tq, tx;
  EFFOCR_TRY(make_tmap_f16_2d(&tq, h->qsplit, nq_pad, 3 * h->Dp, 3 * h->Dp, kKnnBM));
  EFFOCR_TRY(make_tmap_f16_2d(&tx, h->xsplit, h->Npad, 3 * h->Dp, 3 * h->Dp, kKnnBN));
  static bool attr = false;
  if (!attr) {
    EFFOCR_CUDA(cudaFuncSetAttribute(knn_gemm_topk_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kKnnSmemBytes));
    EFFOCR_CUDA(cudaFuncSetAttribute(knn_gemm_topk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kKnnSmemBytes));
    attr = true;
  }
  const int cap = k <= 10 ? 16 : kKnnCap;
  {
    KernelScope ks(PROF_KNN_GEMM, s);
    if (cap == 16)
      knn_gemm_topk_kernel<16><<<m_tiles * splits, kKnnThreads, kKnnSmemBytes, s>>>(tq, tx, nq, h->N, h->Dp, n_tiles, splits,
                                                                                  h->part_val, h->part_idx);
    else
      knn_gemm_topk_kernel<32><<<m_tiles * splits, kKnnThreads, kKnnSmemBytes, s>>>(tq, tx, nq, h->N, h->Dp, n_tiles, splits,
                                                                                  h->part_val, h->part_idx);
  }
  EFFOCR_CUDA(cudaGetLastError());
  {
    KernelScope ks(PROF_KNN_MERGE, s);
    knn_merge_rerank_kernel<<<(nq + 3) / 4, 128, 0, s>>>(h->part_val, h->part_idx, splits * 2, cap, d_queries, h->xb, nq,
                                                         h->D, k, d_dist, d_idx);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

extern "C" int effocr_l2_normalize(const float* d_x, float* d_out, int rows, int dim, float eps, void* stream) {
  EFFOCR_TRY(require_sm100());
  if (rows <= 0) return EFFOCR_OK;
  {
    KernelScope ks(PROF_L2NORM, reinterpret_cast<cudaStream_t>(stream));
    l2_normalize_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_x, d_out, rows, dim, eps);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}
