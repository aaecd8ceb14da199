// Host side of the fused projection + residual + LayerNorm kernel (projln_sm100.cuh).
#include "mlp.h"

#include <stdlib.h>

#include "../../include/effocr_b200.h"
#include "projln_sm100.cuh"

namespace effocr {

bool proj_ln_supported(int D) { return D == kPlnD; }

int proj_ln_f16(const ProjLnArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return EFFOCR_OK;
  if (!proj_ln_supported(a.D)) return fail(EFFOCR_ERR_INVALID, "proj_ln: width must be 384");
  if (!a.att || !a.w || !a.bias || !a.gamma || !a.beta || !a.x || !a.h) return fail(EFFOCR_ERR_INVALID, "proj_ln: null operand");
  if (a.lda % 8 != 0 || a.ldh % 8 != 0 || a.ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(a.x) & 15) ||
      (reinterpret_cast<uintptr_t>(a.h) & 15) || (reinterpret_cast<uintptr_t>(a.att) & 15) ||
      (reinterpret_cast<uintptr_t>(a.bias) & 15) || (reinterpret_cast<uintptr_t>(a.gamma) & 15) ||
      (reinterpret_cast<uintptr_t>(a.beta) & 15))
    return fail(EFFOCR_ERR_INVALID, "proj_ln: operands must be 16-byte aligned with 16-byte multiple pitches");
  CUtensorMap ta, tw, tx, th;
  EFFOCR_TRY(make_tmap_f16_2d(&ta, a.att, a.M, kPlnD, a.lda, 128));
  EFFOCR_TRY(make_tmap_f16_2d(&tw, a.w, kPlnD, kPlnD, kPlnD, 96));
  EFFOCR_TRY(make_tmap_2d(&tx, a.x, 4, a.M, kPlnD, a.ldx, 128, 32, 128));
  EFFOCR_TRY(make_tmap_2d(&th, a.h, 2, a.M, kPlnD, a.ldh, 32, 32, 64));
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(proj_ln_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPlnSmemBytes));
    attr_done = true;
  }
  long long* dbg = nullptr;
  {
    const char* e = getenv("EFFOCR_PLN_DBG_PTR");  // profiling aid: device buffer of 64 int64 for clock64 stamps
    if (e) dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  }
  const int tiles = (a.M + 255) / 256;
  int pairs = sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  static const int prefetch_flags = [] {
    // A/B: bit 0 = L2 prefetch of the next tile's residual rows, bit 1 = of its att rows.  Measured stand-alone at batch 1024:
    // 186.5 us without, 214 / 194 / 219 us with flags 1 / 2 / 3 -- the prefetches compete with the ring's own loads for
    // the same HBM bandwidth at the wrong time; off by default
    const char* e = getenv("EFFOCR_PLN_PREFETCH");
    return e ? atoi(e) : 0;
  }();
  {
    KernelScope ks(PROF_PROJ_LN, stream);
    proj_ln_pair_kernel<<<2 * pairs, kPlnThreads, kPlnSmemBytes, stream>>>(ta, tw, tx, th, a.M, a.bias, a.gamma, a.beta, a.eps, dbg,
                                                                           prefetch_flags);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

}  // namespace effocr

extern "C" int effocr_proj_ln_f16(const void* d_att, long long lda, const void* d_w, const float* d_bias, float* d_x,
                                  long long ldx, const float* d_gamma, const float* d_beta, float eps, void* d_h, long long ldh,
                                  int M, int D, void* stream) {
  EFFOCR_TRY(effocr::require_sm100());
  effocr::ProjLnArgs a;
  a.att = reinterpret_cast<const __half*>(d_att); a.lda = lda;
  a.w = reinterpret_cast<const __half*>(d_w); a.bias = d_bias;
  a.x = d_x; a.ldx = ldx; a.gamma = d_gamma; a.beta = d_beta; a.eps = eps;
  a.h = reinterpret_cast<__half*>(d_h); a.ldh = ldh; a.M = M; a.D = D;
  return effocr::proj_ln_f16(a, reinterpret_cast<cudaStream_t>(stream));
}
