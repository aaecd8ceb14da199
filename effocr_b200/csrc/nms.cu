// K9: YOLOv5 post-processing on the device -- confidence filter, best-class selection, sort by confidence and
// class-aware greedy IoU suppression.  Replaces EffLocalizer.non_max_suppression
// (/root/reference/onnx_engines/localizer_engine.py:171-277; single-label, non-agnostic path as called at :62
// with max_det = 1000) including its torchvision.ops.nms call (:262).
//
// Arithmetic follows the reference operation by operation in fp32 (no FMA contraction), so the kept set is
// identical, not just close:  conf = obj * cls;  box = (x -+ w/2, y -+ h/2);  candidates sorted by conf
// descending (ties: lower prediction index first);  boxes offset by cls * 7680 before IoU;  a box is dropped
// when IoU with an already kept box is strictly greater than iou_thres;  at most max_det survivors.
#include "../../include/effocr_b200.h"
#include "host_common.h"

namespace effocr {

constexpr int kNmsSharedFlags = 8192;  // suppression flags live in shared memory up to this many candidates, else in HBM
constexpr int kNmsMaxNms = 30000;      // the reference keeps the max_nms = 30000 best candidates (:198, :257)

struct NmsCand {
  float x1, y1, x2, y2, conf, cls;
  int idx, pad;
};

// Every prediction that passes the confidence filter becomes a candidate: the per-image capacity `cap` is >= npred, so
// nothing is ever dropped here (the reference never truncates before its sort either).  slot_of maps a prediction
// index to its candidate slot so that the sort below can carry (confidence, prediction index) keys only.
__global__ void __launch_bounds__(256) nms_filter_kernel(const float* __restrict__ pred, int B, int npred, int no,
                                                         float conf_thres, int cap, NmsCand* __restrict__ cand,
                                                         int* __restrict__ slot_of, int* __restrict__ count) {
  const long long total = static_cast<long long>(B) * npred;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / npred);
    const int k = static_cast<int>(i % npred);
    const float* p = pred + i * no;
    const float obj = p[4];
    if (!(obj > conf_thres)) continue;
    float best = -INFINITY;
    int bc = 0;
    for (int c = 5; c < no; ++c) {
      const float v = __fmul_rn(p[c], obj);
      if (v > best) { best = v; bc = c - 5; }  // first maximum wins, like torch.max
    }
    if (!(best > conf_thres)) continue;
    const int slot = atomicAdd(&count[b], 1);  // slot < npred <= cap: every survivor has a slot
    NmsCand c;
    const float hw = __fmul_rn(p[2], 0.5f), hh = __fmul_rn(p[3], 0.5f);  // x / 2 is exact in fp32
    c.x1 = __fsub_rn(p[0], hw);
    c.y1 = __fsub_rn(p[1], hh);
    c.x2 = __fadd_rn(p[0], hw);
    c.y2 = __fadd_rn(p[1], hh);
    c.conf = best;
    c.cls = static_cast<float>(bc);
    c.idx = k;
    c.pad = 0;
    cand[static_cast<long long>(b) * cap + slot] = c;
    slot_of[i] = slot;
  }
}

// One CTA per image: bitonic sort of the candidate keys (global scratch), then greedy suppression.
__global__ void __launch_bounds__(1024) nms_select_kernel(const NmsCand* __restrict__ cand_all,
                                                          const int* __restrict__ count_all,
                                                          const int* __restrict__ slot_of_all, int npred, int cap,
                                                          unsigned long long* __restrict__ keys_all,
                                                          unsigned char* __restrict__ removed_all, float iou_thres,
                                                          float max_wh, int max_det, float* __restrict__ out,
                                                          int* __restrict__ out_count) {
  __shared__ unsigned char s_removed[kNmsSharedFlags];
  const int b = blockIdx.x;
  const NmsCand* cand = cand_all + static_cast<long long>(b) * cap;
  const int* slot_of = slot_of_all + static_cast<long long>(b) * npred;
  unsigned long long* keys = keys_all + static_cast<long long>(b) * cap;
  int n = count_all[b];
  int np2 = 1;
  while (np2 < n) np2 <<= 1;  // <= cap (a power of two >= npred >= n)
  // key = confidence bits (positive floats order like unsigned ints) : inverted prediction index (ties: lower index first)
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    unsigned long long k = 0ull;
    if (i < n) {
      const unsigned int cb = __float_as_uint(cand[i].conf);
      k = (static_cast<unsigned long long>(cb) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned int>(cand[i].idx));
    }
    keys[i] = k;
  }
  __syncthreads();
  // descending bitonic sort
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], c = keys[ixj];
          const bool desc = ((i & k) == 0);
          if (desc ? (a < c) : (a > c)) { keys[i] = c; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (n > kNmsMaxNms) n = kNmsMaxNms;  // x[argsort(conf, descending)[:max_nms]]
  unsigned char* removed = n <= kNmsSharedFlags ? s_removed : removed_all + static_cast<long long>(b) * cap;
  // sorted position -> candidate slot (through the prediction index the key carries)
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const unsigned int idx = 0xFFFFFFFFu - static_cast<unsigned int>(keys[r] & 0xFFFFFFFFull);
    keys[r] = static_cast<unsigned long long>(slot_of[idx]);
    removed[r] = 0;
  }
  __syncthreads();
  // greedy suppression in sorted order; `kept` advances identically in every thread (removed[i] was written before
  // the barrier that ended the previous kept iteration), so no shared counter is read while another thread writes it
  int kept = 0;
  for (int i = 0; i < n && kept < max_det; ++i) {
    if (removed[i]) continue;
    const NmsCand ci = cand[keys[i]];
    const float off_i = __fmul_rn(ci.cls, max_wh);
    const float ix1 = __fadd_rn(ci.x1, off_i), iy1 = __fadd_rn(ci.y1, off_i), ix2 = __fadd_rn(ci.x2, off_i),
                iy2 = __fadd_rn(ci.y2, off_i);
    const float iarea = __fmul_rn(__fsub_rn(ix2, ix1), __fsub_rn(iy2, iy1));
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      if (removed[j]) continue;
      const NmsCand cj = cand[keys[j]];
      const float off_j = __fmul_rn(cj.cls, max_wh);
      const float jx1 = __fadd_rn(cj.x1, off_j), jy1 = __fadd_rn(cj.y1, off_j), jx2 = __fadd_rn(cj.x2, off_j),
                  jy2 = __fadd_rn(cj.y2, off_j);
      const float jarea = __fmul_rn(__fsub_rn(jx2, jx1), __fsub_rn(jy2, jy1));
      const float xx1 = fmaxf(ix1, jx1), yy1 = fmaxf(iy1, jy1), xx2 = fminf(ix2, jx2), yy2 = fminf(iy2, jy2);
      const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
      const float inter = __fmul_rn(w, h);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
      if (ovr > iou_thres) removed[j] = 1;
    }
    if (threadIdx.x == 0) {
      float* o = out + (static_cast<long long>(b) * max_det + kept) * 6;
      o[0] = ci.x1; o[1] = ci.y1; o[2] = ci.x2; o[3] = ci.y2; o[4] = ci.conf; o[5] = ci.cls;
    }
    ++kept;
    __syncthreads();
  }
  if (threadIdx.x == 0) out_count[b] = kept;
}

struct NmsWorkspace {
  NmsCand* cand = nullptr;
  unsigned long long* keys = nullptr;
  unsigned char* removed = nullptr;
  int* slot_of = nullptr;
  int* count = nullptr;
  size_t cap_cand = 0;  // batch * cap elements
  size_t cap_pred = 0;  // batch * npred elements
  int cap_images = 0;
};
static thread_local NmsWorkspace g_ws;  // one per calling host thread (the shim's worker threads)

}  // namespace effocr

using namespace effocr;

extern "C" int effocr_nms(const float* d_pred, int batch, int npred, int no, float conf_thres, float iou_thres,
                          int max_det, float* d_out, int* d_count, void* stream) {
  EFFOCR_TRY(require_sm100());
  if (batch < 0 || npred < 0 || no < 6 || max_det < 1) return fail(EFFOCR_ERR_INVALID, "nms: bad arguments");
  if (!(conf_thres >= 0.f && conf_thres <= 1.f) || !(iou_thres >= 0.f && iou_thres <= 1.f))
    return fail(EFFOCR_ERR_INVALID, "nms: thresholds must be in [0, 1]");  // the reference asserts the same (:194-195)
  if (batch == 0) return EFFOCR_OK;
  if (!d_pred || !d_out || !d_count) return fail(EFFOCR_ERR_INVALID, "nms: null buffer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  NmsWorkspace& ws = g_ws;
  int cap = 1;
  while (cap < npred) cap <<= 1;  // per-image candidate capacity: every prediction can be a candidate (no silent truncation)
  const size_t need_cand = static_cast<size_t>(batch) * cap, need_pred = static_cast<size_t>(batch) * (npred > 0 ? npred : 1);
  if (need_cand > ws.cap_cand || need_pred > ws.cap_pred || batch > ws.cap_images) {
    cudaFree(ws.cand); cudaFree(ws.keys); cudaFree(ws.removed); cudaFree(ws.slot_of); cudaFree(ws.count);
    ws = NmsWorkspace();
    EFFOCR_CUDA(cudaMalloc(&ws.cand, sizeof(NmsCand) * need_cand));
    EFFOCR_CUDA(cudaMalloc(&ws.keys, sizeof(unsigned long long) * need_cand));
    EFFOCR_CUDA(cudaMalloc(&ws.removed, need_cand));
    EFFOCR_CUDA(cudaMalloc(&ws.slot_of, sizeof(int) * need_pred));
    EFFOCR_CUDA(cudaMalloc(&ws.count, sizeof(int) * static_cast<size_t>(batch)));
    ws.cap_cand = need_cand; ws.cap_pred = need_pred; ws.cap_images = batch;
  }
  EFFOCR_CUDA(cudaMemsetAsync(ws.count, 0, sizeof(int) * batch, s));
  if (npred > 0) {
    const long long total = static_cast<long long>(batch) * npred;
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    {
      KernelScope ks(PROF_NMS, s);
      nms_filter_kernel<<<static_cast<int>(g), 256, 0, s>>>(d_pred, batch, npred, no, conf_thres, cap, ws.cand, ws.slot_of, ws.count);
    }
  }
  {
    KernelScope ks(PROF_NMS, s);
    nms_select_kernel<<<batch, 1024, 0, s>>>(ws.cand, ws.count, ws.slot_of, npred, cap, ws.keys, ws.removed, iou_thres, 7680.0f,
                                              max_det, d_out, d_count);
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}
