#include "gemm.h"

#include "../../include/effocr_b200.h"
#include "gemm_sm100.cuh"
#include "gemm_sm100_tma_epi.cuh"
#include "lnqkv_sm100.cuh"

namespace effocr {

int choose_block_n(int N) {
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  // ragged N: least padding wins, ties to the wider tile
  int best = 256, best_pad = (256 - N % 256) % 256;
  const int cands[3] = {192, 128, 64};
  for (int c : cands) {
    int pad = (c - N % c) % c;
    if (pad < best_pad) { best = c; best_pad = pad; }
  }
  return best;
}

template <int BN, class Epi>
static int launch(const GemmArgs& a, const typename Epi::Params& ep, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  EFFOCR_TRY(make_tmap_f16_2d(&ta, a.A, a.M, a.K, a.lda, kBlockM));
  EFFOCR_TRY(make_tmap_f16_2d(&tb, a.W, a.N, a.K, a.ldw, BN));
  auto kern = gemm_tn_kernel<BN, Epi>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int tiles = ((a.M + kBlockM - 1) / kBlockM) * ((a.N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  {
    KernelScope ks(a.prof_tag, stream);
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, a.M, a.N, a.K, ep);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

template <class Epi>
static int launch_bn(int bn, const GemmArgs& a, const typename Epi::Params& ep, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch<64, Epi>(a, ep, stream);
    case 128: return launch<128, Epi>(a, ep, stream);
    case 192: return launch<192, Epi>(a, ep, stream);
    case 256: return launch<256, Epi>(a, ep, stream);
  }
  return fail(EFFOCR_ERR_INVALID, "block_n must be 64, 128, 192 or 256");
}

#ifdef EFFOCR_AB
template <int ACT, typename OutT, bool RES>
static int launch_store(int bn, const GemmArgs& a, cudaStream_t stream) {
  using Epi = EpiStore<ACT, OutT, RES>;
  typename Epi::Params ep;
  ep.out = reinterpret_cast<OutT*>(a.out);
  ep.ldo = a.ldo;
  ep.bias = a.bias;
  ep.gamma = a.gamma;
  ep.resid = reinterpret_cast<const OutT*>(a.resid);
  ep.ldr = a.ldr;
  return launch_bn<Epi>(bn, a, ep, stream);
}
#endif

// ---- TMA-store / TMA-reduce epilogue path (aligned outputs, no separate residual buffer)
template <int BN, int ACT, bool F32, bool RED>
static int launch_tma(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmTmaCfg<BN, F32>;
  CUtensorMap ta, tb, tc;
  EFFOCR_TRY(make_tmap_f16_2d(&ta, a.A, a.M, a.K, a.lda, kBlockM));
  EFFOCR_TRY(make_tmap_f16_2d(&tb, a.W, a.N, a.K, a.ldw, BN));
  EFFOCR_TRY(make_tmap_2d(&tc, a.out, F32 ? 4 : 2, a.M, a.N, a.ldo, 32, 32, F32 ? 128 : 64));
  auto kern = gemm_tn_tma_kernel<BN, ACT, F32, RED>;
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int tiles = ((a.M + kBlockM - 1) / kBlockM) * ((a.N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  EpiTmaParams ep;
  ep.bias = a.bias;
  ep.gamma = a.gamma;
  {
    KernelScope ks(a.prof_tag, stream);
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tc, a.M, a.N, a.K, ep);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

// A-stationary schedule: K <= 384 and at least two N tiles per row block
template <int BN, int ACT, bool F32, bool RED>
static int launch_astat(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmAStatCfg<BN, F32>;
  // sixteen epilogue warps (16-column chunks) when the epilogue has real math to hide: GELU / SiLU on fp16 outputs
  constexpr int EPI_WARPS = (!F32 && !RED && ACT != ACT_NONE) ? 16 : 8;
  CUtensorMap ta, tb, tc;
  EFFOCR_TRY(make_tmap_f16_2d(&ta, a.A, a.M, a.K, a.lda, kBlockM));
  EFFOCR_TRY(make_tmap_f16_2d(&tb, a.W, a.N, a.K, a.ldw, BN));
  if (EPI_WARPS == 16) EFFOCR_TRY(make_tmap_2d(&tc, a.out, 2, a.M, a.N, a.ldo, 32, 16, 32));
  else EFFOCR_TRY(make_tmap_2d(&tc, a.out, F32 ? 4 : 2, a.M, a.N, a.ldo, 32, 32, F32 ? 128 : 64));
  auto kern = gemm_tn_astat_kernel<BN, ACT, F32, RED, EPI_WARPS>;
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int num_m = (a.M + kBlockM - 1) / kBlockM;
  const int grid = num_m < sm_count() ? num_m : sm_count();
  EpiTmaParams ep;
  ep.bias = a.bias;
  ep.gamma = a.gamma;
  {
    KernelScope ks(a.prof_tag, stream);
    kern<<<grid, 128 + 32 * EPI_WARPS, Cfg::kSmemBytes, stream>>>(ta, tb, tc, a.M, a.N, a.K, ep);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

// CTA-pair (cta_group::2) schedule: 256 x BN tiles, BN in {192, 256}
template <int BN, int ACT, bool F32, bool RED>
static int launch_pair(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmPairCfg<BN, F32>;
  CUtensorMap ta, tb, tc;
  EFFOCR_TRY(make_tmap_f16_2d(&ta, a.A, a.M, a.K, a.lda, kBlockM));
  EFFOCR_TRY(make_tmap_f16_2d(&tb, a.W, a.N, a.K, a.ldw, BN / 2));
  EFFOCR_TRY(make_tmap_2d(&tc, a.out, F32 ? 4 : 2, a.M, a.N, a.ldo, 32, 32, F32 ? 128 : 64));
  auto kern = gemm_tn_pair_kernel<BN, ACT, F32, RED>;
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int tiles = ((a.M + 2 * kBlockM - 1) / (2 * kBlockM)) * ((a.N + BN - 1) / BN);
  int pairs = sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  EpiTmaParams ep;
  ep.bias = a.bias;
  ep.gamma = a.gamma;
  if (a.pos) {  // ViT patch embedding: row remap + position embedding in the epilogue
    ep.pos = a.pos;
    ep.x_out = reinterpret_cast<float*>(a.out);
    ep.ldx = a.ldo;
    ep.patches = a.patches;
  }
  {
    KernelScope ks(a.prof_tag, stream);
    kern<<<2 * pairs, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tc, a.M, a.N, a.K, ep);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

static bool use_pair(int bn, const GemmArgs& a) {
  if (a.epilogue == 2) return false;  // 2 = plain single-CTA TMA-epilogue schedule (A/B testing)
  const int num_m = (a.M + 2 * kBlockM - 1) / (2 * kBlockM);
  // measured on B200 (tools/gpu_pair_selftest.py, M = 201 728): QKV +7 %, fc1 (no act) +15 %, fc2 (K = 1536) +20 %,
  // GELU fc1 and the HBM-bound proj unchanged
  if (a.act == ACT_GELU && a.epilogue != 3) return false;  // the GELU epilogue paces fc1; the pair schedule is 10 % slower there
  return (bn == 192 || bn == 256) && num_m * ((a.N + bn - 1) / bn) >= sm_count() / 2;
}

static bool use_astat(int bn, const GemmArgs& a) {
  if (a.epilogue == 2) return false;
  const int num_n = (a.N + bn - 1) / bn;
  const int num_m = (a.M + kBlockM - 1) / kBlockM;
  // measured on B200 (tools/gpu_gemm_selftest.py): +4..7 % for fp16 outputs, -8 % for the single-buffered fp32 path
  return !a.out_f32 && a.K <= 384 && num_n >= 2 && (bn == 192 || bn == 256) && num_m >= 2 * sm_count();
}

template <int ACT, bool F32, bool RED>
static int launch_tma_bn(int bn, const GemmArgs& a, cudaStream_t stream) {
  if (use_pair(bn, a)) {
    if (bn == 192) return launch_pair<192, ACT, F32, RED>(a, stream);
    return launch_pair<256, ACT, F32, RED>(a, stream);
  }
  if (use_astat(bn, a)) {
    if (bn == 192) return launch_astat<192, ACT, F32, RED>(a, stream);
    return launch_astat<256, ACT, F32, RED>(a, stream);
  }
  switch (bn) {
    case 64: return launch_tma<64, ACT, F32, RED>(a, stream);
    case 128: return launch_tma<128, ACT, F32, RED>(a, stream);
    case 192: return launch_tma<192, ACT, F32, RED>(a, stream);
    case 256: return launch_tma<256, ACT, F32, RED>(a, stream);
  }
  return fail(EFFOCR_ERR_INVALID, "block_n must be 64, 128, 192 or 256");
}

// Returns -1 when the TMA epilogue does not apply (caller falls back to the direct-store epilogue).
static int try_gemm_tma(int bn, const GemmArgs& a, cudaStream_t stream) {
  if (a.pos || a.epilogue == 1) return -1;  // epilogue: 0 auto, 1 direct store, 2 TMA/no A-stat, 3 CTA pair
  const bool inplace = a.resid != nullptr && a.resid == a.out && a.ldr == a.ldo;
  if (a.resid && !inplace) return -1;
  if (a.out_f32) {
    if (a.act != ACT_NONE) return -1;
    return inplace ? launch_tma_bn<ACT_NONE, true, true>(bn, a, stream) : launch_tma_bn<ACT_NONE, true, false>(bn, a, stream);
  }
  if (inplace) return -1;  // fp16 in-place residual: not needed by any model here
  switch (a.act) {
    case ACT_NONE: return launch_tma_bn<ACT_NONE, false, false>(bn, a, stream);
    case ACT_GELU: return launch_tma_bn<ACT_GELU, false, false>(bn, a, stream);
    case ACT_SILU: return launch_tma_bn<ACT_SILU, false, false>(bn, a, stream);
  }
  return -1;
}

bool ln_gemm_supported(int D, int N) { return D == kLnqD && N % 8 == 0 && N >= 192; }

// out[M, N] (fp16) = (LayerNorm(x[M, 384]) * gamma + beta) . W[N, 384]^T + bias  -- one kernel (lnqkv_sm100.cuh)
int ln_gemm_f16(const LnGemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return EFFOCR_OK;
  if (!ln_gemm_supported(a.D, a.N)) return fail(EFFOCR_ERR_INVALID, "ln_gemm: width must be 384 and N a multiple of 8, >= 192");
  if (!a.x || !a.gamma || !a.beta || !a.W || !a.out) return fail(EFFOCR_ERR_INVALID, "ln_gemm: null operand");
  if (a.ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.gamma) & 15) ||
      (reinterpret_cast<uintptr_t>(a.beta) & 15) || (a.bias && (reinterpret_cast<uintptr_t>(a.bias) & 15)) || a.ldo % 8 != 0 ||
      (reinterpret_cast<uintptr_t>(a.out) & 15) || a.ldw % 8 != 0)
    return fail(EFFOCR_ERR_INVALID, "ln_gemm: operands must be 16-byte aligned with 16-byte multiple pitches");
  constexpr int BN = LNQ_BN;
  using Cfg = LnQkvCfg<BN>;
  CUtensorMap tb, tc;
  EFFOCR_TRY(make_tmap_f16_2d(&tb, a.W, a.N, kLnqD, a.ldw, BN / 2));  // each CTA of a pair loads half of a weight tile
  EFFOCR_TRY(make_tmap_2d(&tc, a.out, 2, a.M, a.N, a.ldo, 32, Cfg::kStoreCols, Cfg::kStoreCols * 2));
  auto kern = ln_gemm_astat_kernel<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int num_sb = (a.M + 2 * kBlockM - 1) / (2 * kBlockM);
  int pairs = sm_count() / 2;
  if (num_sb < pairs) pairs = num_sb;
  EpiTmaParams ep;
  ep.bias = a.bias;
  ep.gamma = nullptr;
  {
    KernelScope ks(a.prof_tag, stream);
    // EFFOCR_LNQ_FLAGS: timing experiments (2: skip the LayerNorm stream, 4: skip the output stores; results are wrong)
    static const int flags = [] { const char* e = getenv("EFFOCR_LNQ_FLAGS"); return e ? atoi(e) : 0; }();
    long long* dbg = nullptr;  // EFFOCR_LNQ_DBG_PTR: device buffer of 3 x 1024 int64 for tools/lnq_timeline.py
    if (const char* e = getenv("EFFOCR_LNQ_DBG_PTR")) dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
    kern<<<2 * pairs, kLnqThreads, Cfg::kSmemBytes, stream>>>(a.x, a.ldx, a.gamma, a.beta, a.eps, tb, tc, a.M, a.N, ep,
                                                              flags | (a.reverse ? 256 : 0), dbg);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

int gemm_f16(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return fail(EFFOCR_ERR_INVALID, "gemm: empty problem");
  if (!a.A || !a.W || !a.out) return fail(EFFOCR_ERR_INVALID, "gemm: null operand");
  if (a.K % 8 != 0 || a.lda % 8 != 0 || a.ldw % 8 != 0)
    return fail(EFFOCR_ERR_INVALID, "gemm: K and leading dimensions must be multiples of 8 (16-byte TMA pitch)");
  const int esz = a.out_f32 ? 4 : 2;
  if ((a.ldo * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(a.out) & 15) != 0)
    return fail(EFFOCR_ERR_INVALID, "gemm: output must be 16-byte aligned with a 16-byte multiple pitch");
  if (a.resid && ((a.ldr * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(a.resid) & 15) != 0))
    return fail(EFFOCR_ERR_INVALID, "gemm: residual must be 16-byte aligned with a 16-byte multiple pitch");
  if ((a.bias && (reinterpret_cast<uintptr_t>(a.bias) & 15)) || (a.gamma && (reinterpret_cast<uintptr_t>(a.gamma) & 15)))
    return fail(EFFOCR_ERR_INVALID, "gemm: bias / gamma must be 16-byte aligned");
  const int bn = a.block_n ? a.block_n : choose_block_n(a.N);
  {
    const int st = try_gemm_tma(bn, a, stream);
    if (st >= 0) return st;
  }

  if (a.pos) {
    if (a.N % 32 != 0 || a.patches <= 0 || !a.bias || !a.out_f32)
      return fail(EFFOCR_ERR_INVALID, "gemm: patch-embed epilogue needs N % 32 == 0, bias and fp32 output");
    // large batches: CTA-pair kernel, the staged chunk leaves through coalesced remapped row copies (EpiTmaParams::pos)
    if (use_pair(bn, a) && a.act == ACT_NONE && !a.gamma && !a.resid && a.epilogue != 1 && a.N % 4 == 0)
      return bn == 192 ? launch_pair<192, ACT_NONE, true, false>(a, stream) : launch_pair<256, ACT_NONE, true, false>(a, stream);
    EpiPatchEmbed::Params ep;
    ep.x = reinterpret_cast<float*>(a.out);
    ep.bias = a.bias;
    ep.pos = a.pos;
    ep.patches = a.patches;
    return launch_bn<EpiPatchEmbed>(bn, a, ep, stream);
  }
#ifndef EFFOCR_AB
  // every product GEMM takes the TMA-store / TMA-reduce epilogue above or the patch-embed epilogue; the generic
  // direct-store epilogues (separate residual buffer, forced `epilogue = 1`) exist in A/B builds only
  return fail(EFFOCR_ERR_INVALID, "gemm: this shape needs the direct-store epilogue, which is compiled out of this build (EFFOCR_AB=1)");
#else
  const bool res = a.resid != nullptr;
  if (a.out_f32) {
    if (a.act != ACT_NONE) return fail(EFFOCR_ERR_INVALID, "gemm: fp32 output supports act=none only");
    return res ? launch_store<ACT_NONE, float, true>(bn, a, stream) : launch_store<ACT_NONE, float, false>(bn, a, stream);
  }
  switch (a.act) {
    case ACT_NONE:
      return res ? launch_store<ACT_NONE, __half, true>(bn, a, stream)
                 : launch_store<ACT_NONE, __half, false>(bn, a, stream);
    case ACT_GELU:
      if (res) break;
      return launch_store<ACT_GELU, __half, false>(bn, a, stream);
    case ACT_SILU:
      return res ? launch_store<ACT_SILU, __half, true>(bn, a, stream)
                 : launch_store<ACT_SILU, __half, false>(bn, a, stream);
  }
  return fail(EFFOCR_ERR_INVALID, "gemm: unsupported activation / residual combination");
#endif
}

}  // namespace effocr

extern "C" int effocr_ln_gemm_f16(const float* d_x, long long ldx, const float* d_gamma, const float* d_beta, float eps,
                                  const void* d_w, long long ldw, const float* d_bias, void* d_out, long long ldo, int M, int N,
                                  int D, void* stream) {
  EFFOCR_TRY(effocr::require_sm100());
  effocr::LnGemmArgs a;
  a.x = d_x; a.ldx = ldx; a.gamma = d_gamma; a.beta = d_beta; a.eps = eps;
  a.W = reinterpret_cast<const __half*>(d_w); a.ldw = ldw; a.bias = d_bias;
  a.out = reinterpret_cast<__half*>(d_out); a.ldo = ldo; a.M = M; a.N = N; a.D = D;
  return effocr::ln_gemm_f16(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int effocr_gemm_f16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                               const float* bias, const float* gamma, const void* resid, long long ldr, void* out,
                               long long ldo, int act, int out_f32, int block_n, void* stream) {
  EFFOCR_TRY(effocr::require_sm100());
  effocr::GemmArgs a;
  a.A = reinterpret_cast<const __half*>(A);
  a.lda = lda;
  a.W = reinterpret_cast<const __half*>(W);
  a.ldw = ldw;
  a.M = M; a.N = N; a.K = K;
  a.bias = bias; a.gamma = gamma; a.resid = resid; a.ldr = ldr;
  a.out = out; a.ldo = ldo; a.act = act; a.out_f32 = out_f32;
  a.block_n = block_n & 0xffff;
  a.epilogue = (block_n >> 16) & 3;  // 1: force the direct-store epilogue, 2: TMA epilogue without the A-stationary schedule
  return effocr::gemm_f16(a, reinterpret_cast<cudaStream_t>(stream));
}
