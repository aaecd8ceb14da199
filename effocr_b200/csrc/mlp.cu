// Host side of the fused MLP block (mlp_sm100.cuh): tensor maps, launch, C-ABI test entry.
#include "mlp.h"

#include <stdlib.h>

#include "../../include/effocr_b200.h"
#include "mlp_sm100.cuh"

namespace effocr {

bool mlp_fused_supported(int D, int HID) { return (D == 192 || D == 384) && HID % 128 == 0 && HID >= 128; }

template <int D, int AHEAD>
static int launch_mlp(const MlpArgs& a, cudaStream_t stream) {
  using Cfg = MlpCfg<D>;
  CUtensorMap ta, tw1, tw2, tx;
  EFFOCR_TRY(make_tmap_f16_2d(&ta, a.h, a.M, D, a.ldh, 128));
  EFFOCR_TRY(make_tmap_f16_2d(&tw1, a.w1, a.HID, D, D, 32));
  EFFOCR_TRY(make_tmap_f16_2d(&tw2, a.w2, D, a.HID, a.HID, 96));
  EFFOCR_TRY(make_tmap_2d(&tx, a.x, 4, a.M, D, a.ldx, 32, 32, 128));
  auto kern = mlp_fused_pair_kernel<D, AHEAD>;
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int tiles = (a.M + 255) / 256;
  int pairs = sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  {
    static const int cap = [] { const char* e = getenv("EFFOCR_MLP_MAX_PAIRS"); return e ? atoi(e) : 0; }();  // experiments only
    if (cap > 0 && cap < pairs) pairs = cap;
  }
  {
    KernelScope ks(PROF_MLP_FUSED, stream);
    kern<<<2 * pairs, kMlpThreads, Cfg::kSmemBytes, stream>>>(ta, tw1, tw2, tx, a.M, a.HID, a.b1, a.b2, a.dbg);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

template <int D>
static int launch_mlp128(const MlpArgs& a, cudaStream_t stream) {
  using Cfg = Mlp128Cfg<D>;
  CUtensorMap ta, tw1, tw2, tx;
  EFFOCR_TRY(make_tmap_f16_2d(&ta, a.h, a.M, D, a.ldh, 128));
  EFFOCR_TRY(make_tmap_f16_2d(&tw1, a.w1, a.HID, D, D, 64));
  EFFOCR_TRY(make_tmap_f16_2d(&tw2, a.w2, D, a.HID, a.HID, 96));
  EFFOCR_TRY(make_tmap_2d(&tx, a.x, 4, a.M, D, a.ldx, 32, 32, 128));
  auto kern = mlp_fused_pair128_kernel<D>;
  static bool attr_done = false;
  if (!attr_done) {
    EFFOCR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int tiles = (a.M + 255) / 256;
  int pairs = sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  {
    KernelScope ks(PROF_MLP_FUSED, stream);
    kern<<<2 * pairs, kMlpThreads, Cfg::kSmemBytes, stream>>>(ta, tw1, tw2, tx, a.M, a.HID, a.b1, a.b2);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

int mlp_fused_f16(const MlpArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return EFFOCR_OK;
  if (!mlp_fused_supported(a.D, a.HID))
    return fail(EFFOCR_ERR_INVALID, "mlp_fused: width must be 192 or 384 and the hidden size a multiple of 128");
  if (!a.h || !a.w1 || !a.w2 || !a.b1 || !a.b2 || !a.x) return fail(EFFOCR_ERR_INVALID, "mlp_fused: null operand");
  if (a.ldh % 8 != 0 || a.ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.h) & 15) ||
      (reinterpret_cast<uintptr_t>(a.b1) & 15) || (reinterpret_cast<uintptr_t>(a.b2) & 15))
    return fail(EFFOCR_ERR_INVALID, "mlp_fused: operands must be 16-byte aligned with 16-byte multiple pitches");
  static const int chunk = [] {
    const char* e = getenv("EFFOCR_MLP_CHUNK");  // A/B: 64 = the first schedule (N = 64 fc1 MMAs, two S buffers), 128 (default)
    return e ? atoi(e) : 128;
  }();
#ifdef EFFOCR_AB  // the 64-wide schedule at D = 384, the one-chunk look-ahead and the clock64 timeline are A/B-only
  if (chunk == 128 && a.D == 384 && !a.dbg) return launch_mlp128<384>(a, stream);
  static const int ahead = [] {
    const char* e = getenv("EFFOCR_MLP_AHEAD");  // A/B: 1 = fc1 one chunk ahead of fc2, 2 (default) = two chunks
    return e ? atoi(e) : 2;
  }();
  if (ahead == 1) return a.D == 192 ? launch_mlp<192, 1>(a, stream) : launch_mlp<384, 1>(a, stream);
  return a.D == 192 ? launch_mlp<192, 2>(a, stream) : launch_mlp<384, 2>(a, stream);
#else
  (void)chunk;
  if (a.dbg) return fail(EFFOCR_ERR_INVALID, "mlp_fused: the timeline variant is compiled out of this build (EFFOCR_AB=1)");
  return a.D == 384 ? launch_mlp128<384>(a, stream) : launch_mlp<192, 2>(a, stream);
#endif
}

}  // namespace effocr

extern "C" int effocr_mlp_fused_f16(const void* d_h, long long ldh, const void* d_w1, const float* d_b1, const void* d_w2,
                                    const float* d_b2, float* d_x, long long ldx, int M, int D, int HID, void* stream) {
  EFFOCR_TRY(effocr::require_sm100());
  effocr::MlpArgs a;
  a.h = reinterpret_cast<const __half*>(d_h); a.ldh = ldh;
  a.w1 = reinterpret_cast<const __half*>(d_w1); a.b1 = d_b1;
  a.w2 = reinterpret_cast<const __half*>(d_w2); a.b2 = d_b2;
  a.x = d_x; a.ldx = ldx; a.M = M; a.D = D; a.HID = HID;
  {
    const char* e = getenv("EFFOCR_MLP_DBG_PTR");  // tools/mlp_timeline.py: device buffer of 512 int64 for the hand-off timeline
    if (e) a.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  }
  return effocr::mlp_fused_f16(a, reinterpret_cast<cudaStream_t>(stream));
}
