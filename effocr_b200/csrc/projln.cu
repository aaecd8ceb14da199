// Host side of the fused projection + residual + LayerNorm kernel (projln_sm100.cuh).
#include "mlp.h"

#include <stdlib.h>

#include "../../include/effocr_b200.h"
#include "blocktail_sm100.cuh"
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

bool block_tail_supported(int D, int HID) { return D == BlockTailCfg::D && HID % 128 == 0 && HID >= 128; }

int block_tail_f16(const BlockTailArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return EFFOCR_OK;
  if (!block_tail_supported(a.D, a.HID)) return fail(EFFOCR_ERR_INVALID, "block_tail: width must be 384 and the hidden size a multiple of 128");
  if (!a.att || !a.wp || !a.bp || !a.gamma || !a.beta || !a.w1 || !a.b1 || !a.w2 || !a.b2 || !a.x)
    return fail(EFFOCR_ERR_INVALID, "block_tail: null operand");
  const uintptr_t bits = reinterpret_cast<uintptr_t>(a.att) | reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.bp) |
                         reinterpret_cast<uintptr_t>(a.gamma) | reinterpret_cast<uintptr_t>(a.beta) |
                         reinterpret_cast<uintptr_t>(a.b1) | reinterpret_cast<uintptr_t>(a.b2);
  if (a.lda % 8 != 0 || a.ldx % 4 != 0 || (bits & 15))
    return fail(EFFOCR_ERR_INVALID, "block_tail: operands must be 16-byte aligned with 16-byte multiple pitches");
  using Cfg = BlockTailCfg;
  CUtensorMap ta, twp, tw1, tw2, txl, txs;
  EFFOCR_TRY(make_tmap_f16_2d(&ta, a.att, a.M, Cfg::D, a.lda, 128));
  EFFOCR_TRY(make_tmap_f16_2d(&twp, a.wp, Cfg::D, Cfg::D, Cfg::D, 96));
  EFFOCR_TRY(make_tmap_f16_2d(&tw1, a.w1, a.HID, Cfg::D, Cfg::D, 64));
  EFFOCR_TRY(make_tmap_f16_2d(&tw2, a.w2, Cfg::D, a.HID, a.HID, 96));
  EFFOCR_TRY(make_tmap_2d(&txl, a.x, 4, a.M, Cfg::D, a.ldx, 128, 32, 128));  // residual ring: 128 x 32 fp32 boxes in
  EFFOCR_TRY(make_tmap_2d(&txs, a.x, 4, a.M, Cfg::D, a.ldx, 32, 32, 128));   // drain: 32 x 32 fp32 boxes out
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(block_tail_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int tiles = (a.M + 255) / 256;
  int pairs = sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  {
    KernelScope ks(PROF_BLOCK_TAIL, stream);
    static const int stagger = [] { const char* e = getenv("EFFOCR_TAIL_STAGGER"); return e ? atoi(e) : 72000; }();  // estimated cycles per tile (0 = no stagger)
    long long* dbg = nullptr;  // EFFOCR_TAIL_DBG_PTR: device buffer of 2 x 512 int64 for tools/tail_timeline.py
    if (const char* e = getenv("EFFOCR_TAIL_DBG_PTR")) dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
    static const int l2_prefetch = [] { const char* e = getenv("EFFOCR_TAIL_PREFETCH"); return e ? atoi(e) : 0; }();  // A/B: 466 us without, 488 us with (stand-alone, batch 1024)
    block_tail_pair_kernel<<<2 * pairs, kMlpThreads, Cfg::kSmemBytes, stream>>>(ta, twp, tw1, tw2, txl, txs, a.M, a.HID, a.bp, a.gamma,
                                                                                a.beta, a.eps, a.b1, a.b2, l2_prefetch, stagger, a.reverse, dbg);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

}  // namespace effocr

extern "C" int effocr_block_tail_f16(const void* d_att, long long lda, const void* d_wp, const float* d_bp, const float* d_gamma,
                                     const float* d_beta, float eps, const void* d_w1, const float* d_b1, const void* d_w2,
                                     const float* d_b2, float* d_x, long long ldx, int M, int D, int HID, void* stream) {
  EFFOCR_TRY(effocr::require_sm100());
  effocr::BlockTailArgs a;
  a.att = reinterpret_cast<const __half*>(d_att); a.lda = lda;
  a.wp = reinterpret_cast<const __half*>(d_wp); a.bp = d_bp; a.gamma = d_gamma; a.beta = d_beta; a.eps = eps;
  a.w1 = reinterpret_cast<const __half*>(d_w1); a.b1 = d_b1;
  a.w2 = reinterpret_cast<const __half*>(d_w2); a.b2 = d_b2;
  a.x = d_x; a.ldx = ldx; a.M = M; a.D = D; a.HID = HID;
  return effocr::block_tail_f16(a, reinterpret_cast<cudaStream_t>(stream));
}

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
