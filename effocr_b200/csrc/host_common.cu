#include "host_common.h"

#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/effocr_b200.h"

namespace effocr {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return EFFOCR_ERR_CUDA;
}

static thread_local int g_sm_limit = 0;
void set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; }

int sm_count() {
  // EFFOCR_SM_LIMIT (experiments) / set_sm_limit (two half-batches on two streams, each on half of the SMs): persistent
  // kernels size their grids from this value
  static const int env_limit = [] {
    const char* e = getenv("EFFOCR_SM_LIMIT");
    return e ? atoi(e) : 0;
  }();
  const int lim = g_sm_limit > 0 ? g_sm_limit : env_limit;
  const int n = sm_count_physical();
  return (lim > 0 && lim < n) ? lim : n;
}

int sm_count_physical() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int require_sm100() {
  int dev = 0;
  EFFOCR_CUDA(cudaGetDevice(&dev));
  int major = 0;
  EFFOCR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(EFFOCR_ERR_NO_DEVICE, "effocr_b200 kernels are built for sm_100a only");
  return EFFOCR_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(EFFOCR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(EFFOCR_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  if (box_rows == 0 || box_rows > 256) return fail(EFFOCR_ERR_INVALID, "TMA box rows out of range");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EFFOCR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
  return EFFOCR_OK;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int esz, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(EFFOCR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * esz) % 16 != 0)
    return fail(EFFOCR_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  if (box_rows == 0 || box_rows > 256 || box_cols == 0 || box_cols > 256) return fail(EFFOCR_ERR_INVALID, "TMA box out of range");
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if (swizzle_bytes && static_cast<int>(box_cols) * esz != swizzle_bytes)
    return fail(EFFOCR_ERR_INVALID, "TMA box width must equal the swizzle span");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * static_cast<uint64_t>(esz)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EFFOCR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
  return EFFOCR_OK;
}

// 4-D tensor map over an NHWC-like tensor: dims / box / traversal strides innermost first; byte strides of dims 1..3.
int make_tmap_4d(CUtensorMap* out, const void* base, int esz, const uint64_t dims[4], const uint64_t strides_bytes[3],
                 const uint32_t box[4], const uint32_t elem_strides[4], int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(EFFOCR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(EFFOCR_ERR_INVALID, "TMA operand must be 16-byte aligned");
  for (int i = 0; i < 3; ++i)
    if (strides_bytes[i] % 16 != 0) return fail(EFFOCR_ERR_INVALID, "TMA strides must be multiples of 16 bytes");
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {elem_strides[0], elem_strides[1], elem_strides[2], elem_strides[3]};
  CUresult r = enc(out, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EFFOCR_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed with CUresult " + std::to_string(int(r)));
  return EFFOCR_OK;
}

// ------------------------------------------------------------------ launch accounting / profiling
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRecord { int tag; cudaEvent_t start, stop; };
static std::mutex g_prof_mu;
static std::vector<ProfRecord> g_prof_records;   // in flight (not yet read)
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
static double g_prof_ms[PROF_NUM_TAGS] = {0};
static long long g_prof_n[PROF_NUM_TAGS] = {0};

KernelScope::KernelScope(int tag, cudaStream_t stream) : slot_(-1), stream_(stream) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRecord r;
  r.tag = tag;
  if (!g_prof_pool.empty()) {
    r.start = g_prof_pool.back().first; r.stop = g_prof_pool.back().second; g_prof_pool.pop_back();
  } else {
    if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
  }
  cudaEventRecord(r.start, stream_);
  slot_ = static_cast<int>(g_prof_records.size());
  g_prof_records.push_back(r);
}
KernelScope::~KernelScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot_ < static_cast<int>(g_prof_records.size())) cudaEventRecord(g_prof_records[slot_].stop, stream_);
}

static void prof_drain() {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_records) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess && r.tag >= 0 && r.tag < PROF_NUM_TAGS) {
      g_prof_ms[r.tag] += ms;
      g_prof_n[r.tag] += 1;
    }
    g_prof_pool.emplace_back(r.start, r.stop);
  }
  g_prof_records.clear();
}

static const char* kProfNames[PROF_NUM_TAGS] = {
    "crop_resize", "im2patch", "gemm_patch_embed", "layernorm", "gemm_qkv", "attention", "gemm_proj", "gemm_fc1_gelu",
    "gemm_fc2", "final_layernorm", "l2_normalize", "knn_split", "knn_gemm_topk", "knn_merge_rerank", "misc",
    "gemm_other", "conv_im2col", "yolo_misc", "nms", "dwconv_ln", "mlp_fused", "proj_ln", "block_tail"};

}  // namespace effocr

extern "C" int effocr_build_flags() {
#ifdef EFFOCR_AB
  return 1;
#else
  return 0;
#endif
}

extern "C" long long effocr_launch_count(void) { return effocr::g_launches.load(); }
extern "C" void effocr_profile_enable(int on) {
  effocr::prof_drain();
  effocr::g_prof_on.store(on ? 1 : 0);
}
extern "C" void effocr_profile_reset(void) {
  effocr::prof_drain();
  for (int i = 0; i < effocr::PROF_NUM_TAGS; ++i) { effocr::g_prof_ms[i] = 0; effocr::g_prof_n[i] = 0; }
}
extern "C" int effocr_profile_num_tags(void) { return effocr::PROF_NUM_TAGS; }
extern "C" const char* effocr_profile_tag_name(int tag) {
  return (tag >= 0 && tag < effocr::PROF_NUM_TAGS) ? effocr::kProfNames[tag] : "";
}
extern "C" int effocr_profile_read(int tag, long long* launches, double* total_ms) {
  if (tag < 0 || tag >= effocr::PROF_NUM_TAGS) return EFFOCR_ERR_INVALID;
  effocr::prof_drain();
  if (launches) *launches = effocr::g_prof_n[tag];
  if (total_ms) *total_ms = effocr::g_prof_ms[tag];
  return EFFOCR_OK;
}

extern "C" const char* effocr_last_error(void) { return effocr::g_last_error.c_str(); }
extern "C" int effocr_abi_version(void) { return EFFOCR_B200_ABI_VERSION; }
extern "C" int effocr_device_ok(void) { return effocr::require_sm100(); }
