// ViT recognizer encoder (timm vit_{tiny,small,base}_patch16_224, num_classes=0) as a C-ABI handle.
// Replaces `timm.create_model(...)(x)` reached from /root/reference/models/encoders.py:62-64
// (infer_effocr.py:314) and the ORT session of onnx_engines/recognizer_engine.py:23-27.
//
// Numerics: fp16 GEMM operands, fp32 accumulation (TMEM), fp32 residual stream, fp32 LayerNorm /
// softmax statistics -- SURVEY.md section 0.5 (bf16 or an fp16 residual flips top-1 ids).
#include <stdlib.h>

#include <vector>

#include "../../include/effocr_b200.h"
#include "gemm.h"
#include "mlp.h"
#include "attention_sm100.cuh"
#include "vit_kernels.cuh"

namespace effocr {

struct VitLayer {
  float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  __half *w_qkv, *w_proj, *w_fc1, *w_fc2;
  float *b_qkv, *b_proj, *b_fc1, *b_fc2;
};

struct VitHandle {
  int D = 0, H = 0, depth = 0, mlp = 0, max_batch = 0;
  float eps = 1e-6f;
  static constexpr int T = 197, P = 196, PK = 768;
  __half* w_patch = nullptr;
  float *b_patch = nullptr, *cls = nullptr, *pos = nullptr, *lnf_w = nullptr, *lnf_b = nullptr;
  std::vector<VitLayer> layers;
  // workspace, sized for max_batch
  __half *patches = nullptr, *h16 = nullptr, *qkv = nullptr, *att = nullptr, *mid = nullptr;
  float* x = nullptr;
  std::vector<void*> allocs;

  ~VitHandle() {
    for (void* p : allocs) cudaFree(p);
  }
  template <typename T_>
  int alloc(T_** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T_));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    allocs.push_back(q);
    *p = reinterpret_cast<T_*>(q);
    return EFFOCR_OK;
  }
  int upload_f32(float** dst, const float* src, size_t n) {
    EFFOCR_TRY(alloc(dst, n));
    EFFOCR_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return EFFOCR_OK;
  }
  int upload_f16(__half** dst, const float* src, size_t n) {
    std::vector<__half> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = __float2half_rn(src[i]);
    EFFOCR_TRY(alloc(dst, n));
    EFFOCR_CUDA(cudaMemcpy(*dst, tmp.data(), n * sizeof(__half), cudaMemcpyHostToDevice));
    return EFFOCR_OK;
  }
};

template <typename OutT>
static int layernorm_launch(const float* x, long long ldx, const float* g, const float* b, OutT* out, long long ldo,
                            int rows, int D, float eps, cudaStream_t s, int tag) {
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  KernelScope ks(tag, s);
  switch (D) {
    case 96: layernorm_rows_kernel<96, OutT><<<grid, wpb * 32, 0, s>>>(x, ldx, g, b, out, ldo, rows, eps); break;
    case 192: layernorm_rows_kernel<192, OutT><<<grid, wpb * 32, 0, s>>>(x, ldx, g, b, out, ldo, rows, eps); break;
    case 384: layernorm_rows_kernel<384, OutT><<<grid, wpb * 32, 0, s>>>(x, ldx, g, b, out, ldo, rows, eps); break;
    case 768: layernorm_rows_kernel<768, OutT><<<grid, wpb * 32, 0, s>>>(x, ldx, g, b, out, ldo, rows, eps); break;
    default: return fail(EFFOCR_ERR_INVALID, "layernorm: unsupported width (96/192/384/768)");
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

int layernorm_f16(const float* x, long long ldx, const float* g, const float* b, __half* out, long long ldo, int rows,
                  int D, float eps, cudaStream_t s, int tag) {
  return layernorm_launch<__half>(x, ldx, g, b, out, ldo, rows, D, eps, s, tag);
}
int layernorm_f32(const float* x, long long ldx, const float* g, const float* b, float* out, long long ldo, int rows,
                  int D, float eps, cudaStream_t s, int tag) {
  return layernorm_launch<float>(x, ldx, g, b, out, ldo, rows, D, eps, s, tag);
}

// impl 0: tcgen05 kernel, one TMEM pass over S with two key blocks combined flash-attention style (default); A/B
// variants via EFFOCR_ATTENTION_SOFTMAX or impl: 5 = three key blocks / 16 softmax warps, 2 = two TMEM passes, 3 (env 1) = single pass via fp16 deltas in the P
// tile, 4 = two passes with 16 softmax warps; impl 1: first-generation mma.sync kernel
int attention_f16(const __half* qkv, __half* out, int batch, int T, int H, cudaStream_t s, int impl = 0, int reverse = 0) {
  if (T != 197) return fail(EFFOCR_ERR_INVALID, "attention: sequence length must be 197 (ViT/16 @ 224)");
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
#ifdef EFFOCR_AB
  if (impl == 1) {
    static bool attr = false;
    if (!attr) {
      EFFOCR_CUDA(cudaFuncSetAttribute(attention_197x64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
      attr = true;
    }
    KernelScope ks(PROF_ATTENTION, s);
    attention_197x64_kernel<<<batch * H, kAttnThreads, kAttnSmemBytes, s>>>(qkv, out, T, H, scale_log2e);
  } else
#else
  if (impl != 0) return fail(EFFOCR_ERR_INVALID, "attention: A/B variants are compiled out of this build (EFFOCR_AB=1 python -m effocr_b200.build --force)");
#endif
  {
    static bool attr = false;
    if (!attr) {
#ifdef EFFOCR_AB
      EFFOCR_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes));
      EFFOCR_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes));
      EFFOCR_CUDA(cudaFuncSetAttribute(attention_tc16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAt16SmemBytes));
      EFFOCR_CUDA(cudaFuncSetAttribute(attention_tc3b_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAt3SmemBytes));
      EFFOCR_CUDA(cudaFuncSetAttribute(attention_tc3b_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAt3SmemBytes));
#endif
      EFFOCR_CUDA(cudaFuncSetAttribute(attention_tc2b_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes));
      EFFOCR_CUDA(cudaFuncSetAttribute(attention_tc2b_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes));
      EFFOCR_CUDA(cudaFuncSetAttribute(attention_tc2b_kernel<false, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes));
      attr = true;
    }
    static const int env_variant = [] {
      const char* e = getenv("EFFOCR_ATTENTION_SOFTMAX");
      return e ? atoi(e) : 0;
    }();
    int variant = 0;  // 0: one TMEM pass / two key blocks, 1: single pass via fp16 deltas, 2: two passes, 4: 16 warps two passes
#ifdef EFFOCR_AB
    if (impl == 2) variant = 2;
    else if (impl == 3) variant = 1;
    else if (impl == 4) variant = 4;
    else if (impl == 5) variant = 5;
    else if (env_variant == 1 || env_variant == 2 || env_variant == 4 || env_variant == 5) variant = env_variant;
#else
    (void)env_variant;
#endif
    const long long rows = static_cast<long long>(batch) * T;
    const int D = H * 64;
    CUtensorMap tq, tkv;
    EFFOCR_TRY(make_tmap_f16_2d(&tq, qkv, rows, 3 * D, 3 * D, 128));
    EFFOCR_TRY(make_tmap_f16_2d(&tkv, qkv, rows, 3 * D, 3 * D, kAtN));
    const int pairs = batch * H;
    const int grid = pairs < sm_count() ? pairs : sm_count();
    KernelScope ks(PROF_ATTENTION, s);
#ifdef EFFOCR_AB
    if (variant == 1) attention_tc_kernel<true><<<grid, kAtThreads, kAtSmemBytes, s>>>(tq, tkv, out, batch, H, scale_log2e);
    else if (variant == 2) attention_tc_kernel<false><<<grid, kAtThreads, kAtSmemBytes, s>>>(tq, tkv, out, batch, H, scale_log2e);
    else if (variant == 4) attention_tc16_kernel<<<grid, kAt16Threads, kAt16SmemBytes, s>>>(tq, tkv, out, batch, H, scale_log2e);
    else if (variant == 5) {
      const char* e = getenv("EFFOCR_ATT_DBG_PTR");
      long long* dbg = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
      if (dbg) attention_tc3b_kernel<true><<<grid, kAt3Threads, kAt3SmemBytes, s>>>(tq, tkv, out, batch, H, scale_log2e, dbg);
      else attention_tc3b_kernel<false><<<grid, kAt3Threads, kAt3SmemBytes, s>>>(tq, tkv, out, batch, H, scale_log2e, nullptr);
    } else
#endif
    {
      const char* e = getenv("EFFOCR_ATT_DBG_PTR");  // tools/att_timeline.py: device buffer of 32 int64 receiving wait-time totals
      long long* dbg = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
      // output tile as a 4-D tensor {D, 197 tokens, batch, 1}: a 32-token box that runs past token 196 is clipped
      CUtensorMap to;
      {
        const uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(T), static_cast<uint64_t>(batch), 1};
        const uint64_t strides[3] = {static_cast<uint64_t>(D) * 2, static_cast<uint64_t>(T) * D * 2, static_cast<uint64_t>(batch) * T * D * 2};
        const uint32_t box[4] = {64, 32, 1, 1}, es[4] = {1, 1, 1, 1};
        EFFOCR_TRY(make_tmap_4d(&to, out, 2, dims, strides, box, es, 128));
      }
      static const bool tma_out = [] {
        const char* e = getenv("EFFOCR_ATT_TMA_OUT");  // "0" = per-thread row stores (A/B)
        return !(e && e[0] == '0');
      }();
#ifdef EFFOCR_ATT_ABLATION
      // timing experiments (tools/att_ablate.py): kernels whose results are wrong on purpose -- only in a library built with
      // EFFOCR_NVCC_EXTRA=-DEFFOCR_ATT_ABLATION, never in the product build
      const char* ab = getenv("EFFOCR_ATT_ABLATE");
      const int abl = ab ? atoi(ab) : 0;
#define EFFOCR_ATT_ABL(m)                                                                                              \
  if (abl == m) {                                                                                                      \
    EFFOCR_CUDA(cudaFuncSetAttribute(attention_tc2b_kernel<false, m, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes)); \
    attention_tc2b_kernel<false, m, false><<<grid, kAtThreads, kAtSmemBytes, s>>>(tq, tkv, to, out, batch, H, scale_log2e, nullptr, reverse);        \
  } else
      EFFOCR_ATT_ABL(1) EFFOCR_ATT_ABL(2) EFFOCR_ATT_ABL(4) EFFOCR_ATT_ABL(8) EFFOCR_ATT_ABL(3) EFFOCR_ATT_ABL(7) EFFOCR_ATT_ABL(15)
#undef EFFOCR_ATT_ABL
#endif
      if (dbg) attention_tc2b_kernel<true><<<grid, kAtThreads, kAtSmemBytes, s>>>(tq, tkv, to, out, batch, H, scale_log2e, dbg, reverse);
      else if (!tma_out) attention_tc2b_kernel<false, 0, false><<<grid, kAtThreads, kAtSmemBytes, s>>>(tq, tkv, to, out, batch, H, scale_log2e, nullptr, reverse);
      else attention_tc2b_kernel<false><<<grid, kAtThreads, kAtSmemBytes, s>>>(tq, tkv, to, out, batch, H, scale_log2e, nullptr, reverse);
    }
  }
  EFFOCR_CUDA(cudaGetLastError());
  return EFFOCR_OK;
}

static int vit_forward_chunk(VitHandle* v, int B, float* emb, cudaStream_t s) {
  const int D = v->D, T = VitHandle::T;
  const int M = B * T;
  GemmArgs g;
  // patch embedding + bias + position embedding -> token rows 1..196 of x
  g = GemmArgs();
  g.A = v->patches; g.lda = VitHandle::PK; g.W = v->w_patch; g.ldw = VitHandle::PK;
  g.M = B * VitHandle::P; g.N = D; g.K = VitHandle::PK;
  g.out = v->x; g.ldo = D; g.out_f32 = 1; g.bias = v->b_patch; g.pos = v->pos; g.patches = VitHandle::P;
  g.prof_tag = PROF_GEMM_PATCH;
  EFFOCR_TRY(gemm_f16(g, s));
  {
    KernelScope ks(PROF_MISC, s);
    cls_pos_kernel<<<(B * D + 255) / 256, 256, 0, s>>>(v->x, v->cls, v->pos, B, T, D);
  }
  EFFOCR_CUDA(cudaGetLastError());
  static const bool fused_env = [] {
    const char* e = getenv("EFFOCR_MLP_FUSED");  // "0" = unfused fc1 / fc2 GEMMs (A/B runs)
    return !(e && e[0] == '0');
  }();
  const bool fused_mlp = fused_env && mlp_fused_supported(D, v->mlp);
  static const bool pln_env = [] {
    const char* e = getenv("EFFOCR_PROJ_LN");  // "0" = separate projection GEMM + LayerNorm kernels (A/B runs)
    return !(e && e[0] == '0');
  }();
  const bool fused_pln = pln_env && proj_ln_supported(D);
  static const bool tail_env = [] {
    const char* e = getenv("EFFOCR_BLOCK_TAIL");  // "0" = proj_ln + mlp_fused as two kernels (A/B runs)
    return !(e && e[0] == '0');
  }();
  const bool fused_tail = tail_env && fused_mlp && fused_pln && block_tail_supported(D, v->mlp);
  static const bool cls_env = [] {
    const char* e = getenv("EFFOCR_VIT_LAST_BLOCK_FULL");  // "1" = run the last block on all 197 tokens (A/B runs)
    return !(e && e[0] == '1');
  }();
  // norm1 + QKV in one kernel (lnqkv_sm100.cuh, CTA-pair A-stationary schedule): bit-identical to the two-kernel path and
  // 198 us instead of 81 + 161 us per layer inside the batch-1024 step.  EFFOCR_LN_QKV=0 selects the two-kernel path.
  static const bool lnq_env = [] {
    const char* e = getenv("EFFOCR_LN_QKV");
    return !(e && e[0] == '0');
  }();
  // EFFOCR_SNAKE=1 (A/B, off): every persistent kernel finishes with the highest (or lowest) rows of its output still in the
  // 126 MB L2; the next kernel then walks its tiles from THAT end, the direction flipping with every kernel of the
  // three-kernel layer.  Measured: no difference (11.07 ms per step either way) -- two or three waves of reads are not
  // where the kernels wait.
  static const bool snake_env = [] {
    const char* e = getenv("EFFOCR_SNAKE");
    return e && e[0] == '1';
  }();
  int dir = 1;  // 1: the previous kernel walked forwards (the patch-embedding GEMM does), so the next one walks backwards
  auto next_reverse = [&]() { const int r = snake_env ? dir : 0; dir ^= 1; return r; };
  for (int l = 0; l < v->depth; ++l) {
    const VitLayer& L = v->layers[l];
    const bool last_cls = cls_env && l == v->depth - 1;
    if (lnq_env && !last_cls && ln_gemm_supported(D, 3 * D)) {
      // norm1 + QKV projection in one kernel: the normalised fp16 operand is produced on the SM (lnqkv_sm100.cuh)
      LnGemmArgs q;
      q.x = v->x; q.ldx = D; q.gamma = L.ln1_w; q.beta = L.ln1_b; q.eps = v->eps; q.W = L.w_qkv; q.ldw = D; q.bias = L.b_qkv;
      q.out = v->qkv; q.ldo = 3 * D; q.M = M; q.N = 3 * D; q.D = D; q.prof_tag = PROF_GEMM_QKV;
      q.reverse = next_reverse();
      EFFOCR_TRY(ln_gemm_f16(q, s));
    } else {
    EFFOCR_TRY(layernorm_f16(v->x, D, L.ln1_w, L.ln1_b, v->h16, D, M, D, v->eps, s, PROF_LAYERNORM));
    if (cls_env && l == v->depth - 1) {
      // last block: K and V for every token, Q for the class tokens only (see below)
      g = GemmArgs();
      g.A = v->h16; g.lda = D; g.W = L.w_qkv + static_cast<size_t>(D) * D; g.ldw = D; g.M = M; g.N = 2 * D; g.K = D;
      g.out = v->qkv + D; g.ldo = 3 * D; g.bias = L.b_qkv + D; g.prof_tag = PROF_GEMM_QKV;
      EFFOCR_TRY(gemm_f16(g, s));
      g = GemmArgs();
      g.A = v->h16; g.lda = static_cast<long long>(T) * D; g.W = L.w_qkv; g.ldw = D; g.M = B; g.N = D; g.K = D;
      g.out = v->qkv; g.ldo = static_cast<long long>(T) * 3 * D; g.bias = L.b_qkv; g.prof_tag = PROF_GEMM_OTHER;
      EFFOCR_TRY(gemm_f16(g, s));
    } else {
      g = GemmArgs();
      g.A = v->h16; g.lda = D; g.W = L.w_qkv; g.ldw = D; g.M = M; g.N = 3 * D; g.K = D;
      g.out = v->qkv; g.ldo = 3 * D; g.bias = L.b_qkv; g.prof_tag = PROF_GEMM_QKV;
      EFFOCR_TRY(gemm_f16(g, s));
    }
    }
    // Last block: only the class token reaches the embedding (timm pools x[:, 0]), so the rows of the 196 patch tokens
    // are dead after K and V have been formed: CLS-query attention, then projection / norm2 / MLP on the B class rows
    // of x only (row pitch T * D).  Same kernels, same per-row arithmetic; 1/12 of the projection and MLP work goes away.
    const bool cls_only = cls_env && l == v->depth - 1;
    const int Mr = cls_only ? B : M;                                   // rows that continue through this block
    const long long ldx = cls_only ? static_cast<long long>(T) * D : D;  // pitch of those rows in x
    if (cls_only) {
      KernelScope ks(PROF_ATTENTION, s);
      cls_attention_kernel<<<(B * v->H + 3) / 4, 128, 0, s>>>(v->qkv, v->att, B, T, v->H, 0.125f);
      EFFOCR_CUDA(cudaGetLastError());
    } else {
      EFFOCR_TRY(attention_f16(v->qkv, v->att, B, T, v->H, s, 0, (lnq_env && fused_tail) ? next_reverse() : 0));
    }
    if (fused_tail && !last_cls) {  // projection + residual + norm2 + MLP + residual in one kernel (blocktail_sm100.cuh)
      BlockTailArgs t;
      t.att = v->att; t.lda = D; t.wp = L.w_proj; t.bp = L.b_proj; t.gamma = L.ln2_w; t.beta = L.ln2_b; t.eps = v->eps;
      t.w1 = L.w_fc1; t.b1 = L.b_fc1; t.w2 = L.w_fc2; t.b2 = L.b_fc2; t.x = v->x; t.ldx = ldx; t.M = Mr; t.D = D; t.HID = v->mlp;
      t.reverse = lnq_env ? next_reverse() : 0;
      EFFOCR_TRY(block_tail_f16(t, s));
      continue;
    }
    if (fused_pln) {  // projection + residual + norm2 in one full-row kernel: x is read and written once, no L2 reductions
      ProjLnArgs p;
      p.att = v->att; p.lda = D; p.w = L.w_proj; p.bias = L.b_proj; p.x = v->x; p.ldx = ldx;
      p.gamma = L.ln2_w; p.beta = L.ln2_b; p.eps = v->eps; p.h = v->h16; p.ldh = D; p.M = Mr; p.D = D;
      EFFOCR_TRY(proj_ln_f16(p, s));
    } else {
      g = GemmArgs();
      g.A = v->att; g.lda = D; g.W = L.w_proj; g.ldw = D; g.M = Mr; g.N = D; g.K = D;
      g.out = v->x; g.ldo = ldx; g.out_f32 = 1; g.bias = L.b_proj; g.resid = v->x; g.ldr = ldx; g.prof_tag = PROF_GEMM_PROJ;
      EFFOCR_TRY(gemm_f16(g, s));
      EFFOCR_TRY(layernorm_f16(v->x, ldx, L.ln2_w, L.ln2_b, v->h16, D, Mr, D, v->eps, s, PROF_LAYERNORM));
    }
    if (fused_mlp) {  // fc1 + GELU + fc2 + residual in one kernel: the [M, mlp] hidden activations never reach HBM
      MlpArgs m;
      m.h = v->h16; m.ldh = D; m.w1 = L.w_fc1; m.b1 = L.b_fc1; m.w2 = L.w_fc2; m.b2 = L.b_fc2;
      m.x = v->x; m.ldx = ldx; m.M = Mr; m.D = D; m.HID = v->mlp;
      EFFOCR_TRY(mlp_fused_f16(m, s));
      continue;
    }
    g = GemmArgs();
    g.A = v->h16; g.lda = D; g.W = L.w_fc1; g.ldw = D; g.M = Mr; g.N = v->mlp; g.K = D;
    g.out = v->mid; g.ldo = v->mlp; g.bias = L.b_fc1; g.act = 1; g.prof_tag = PROF_GEMM_FC1;
    EFFOCR_TRY(gemm_f16(g, s));
    g = GemmArgs();
    g.A = v->mid; g.lda = v->mlp; g.W = L.w_fc2; g.ldw = v->mlp; g.M = Mr; g.N = D; g.K = v->mlp;
    g.out = v->x; g.ldo = ldx; g.out_f32 = 1; g.bias = L.b_fc2; g.resid = v->x; g.ldr = ldx; g.prof_tag = PROF_GEMM_FC2;
    EFFOCR_TRY(gemm_f16(g, s));
  }
  // final LayerNorm on the CLS rows only (row stride T * D)
  EFFOCR_TRY(layernorm_f32(v->x, static_cast<long long>(T) * D, v->lnf_w, v->lnf_b, emb, D, B, D, v->eps, s, PROF_FINAL_LN));
  return EFFOCR_OK;
}

}  // namespace effocr

using namespace effocr;

extern "C" int effocr_vit_create(int embed_dim, int num_heads, int depth, int mlp_dim, int max_batch, float ln_eps,
                                 const float* const* h_weights, int n_weights, effocr_vit_t* out) {
  if (!out) return fail(EFFOCR_ERR_INVALID, "vit_create: null out");
  *out = nullptr;
  EFFOCR_TRY(require_sm100());
  if (embed_dim != num_heads * 64) return fail(EFFOCR_ERR_INVALID, "vit_create: head dim must be 64");
  if (embed_dim != 192 && embed_dim != 384 && embed_dim != 768)
    return fail(EFFOCR_ERR_INVALID, "vit_create: embed dim must be 192, 384 or 768");
  if (n_weights != 4 + 12 * depth + 2) return fail(EFFOCR_ERR_INVALID, "vit_create: expected 4 + 12*depth + 2 weight tensors");
  if (max_batch <= 0 || mlp_dim % 8 != 0) return fail(EFFOCR_ERR_INVALID, "vit_create: bad max_batch / mlp_dim");
  VitHandle* v = new VitHandle();
  v->D = embed_dim; v->H = num_heads; v->depth = depth; v->mlp = mlp_dim; v->max_batch = max_batch; v->eps = ln_eps;
  const int D = embed_dim;
  int st = EFFOCR_OK;
  auto W = [&](int i) { return h_weights[i]; };
  do {
    if ((st = v->upload_f16(&v->w_patch, W(0), size_t(D) * 768))) break;
    {
      // the patch rows are white-centred (crop.cu): bias' = bias + sum_k fp16(W[n, k]) * white(channel of k)
      const float white[3] = {(1.0f - 0.485f) / 0.229f, (1.0f - 0.456f) / 0.224f, (1.0f - 0.406f) / 0.225f};
      std::vector<float> b(D);
      for (int n = 0; n < D; ++n) {
        double acc = W(1)[n];
        for (int k = 0; k < 768; ++k)
          acc += static_cast<double>(__half2float(__float2half_rn(W(0)[size_t(n) * 768 + k]))) * static_cast<double>(white[k / 256]);
        b[n] = static_cast<float>(acc);
      }
      if ((st = v->upload_f32(&v->b_patch, b.data(), D))) break;
    }
    if ((st = v->upload_f32(&v->cls, W(2), D))) break;
    if ((st = v->upload_f32(&v->pos, W(3), size_t(197) * D))) break;
    v->layers.resize(depth);
    for (int l = 0; l < depth && !st; ++l) {
      VitLayer& L = v->layers[l];
      const int o = 4 + 12 * l;
      if ((st = v->upload_f32(&L.ln1_w, W(o + 0), D))) break;
      if ((st = v->upload_f32(&L.ln1_b, W(o + 1), D))) break;
      if ((st = v->upload_f16(&L.w_qkv, W(o + 2), size_t(3) * D * D))) break;
      if ((st = v->upload_f32(&L.b_qkv, W(o + 3), 3 * D))) break;
      if ((st = v->upload_f16(&L.w_proj, W(o + 4), size_t(D) * D))) break;
      if ((st = v->upload_f32(&L.b_proj, W(o + 5), D))) break;
      if ((st = v->upload_f32(&L.ln2_w, W(o + 6), D))) break;
      if ((st = v->upload_f32(&L.ln2_b, W(o + 7), D))) break;
      if ((st = v->upload_f16(&L.w_fc1, W(o + 8), size_t(mlp_dim) * D))) break;
      if ((st = v->upload_f32(&L.b_fc1, W(o + 9), mlp_dim))) break;
      if ((st = v->upload_f16(&L.w_fc2, W(o + 10), size_t(D) * mlp_dim))) break;
      if ((st = v->upload_f32(&L.b_fc2, W(o + 11), D))) break;
    }
    if (st) break;
    if ((st = v->upload_f32(&v->lnf_w, W(4 + 12 * depth), D))) break;
    if ((st = v->upload_f32(&v->lnf_b, W(5 + 12 * depth), D))) break;
    const size_t M = size_t(max_batch) * 197;
    if ((st = v->alloc(&v->patches, size_t(max_batch) * 196 * 768))) break;
    if ((st = v->alloc(&v->x, M * D))) break;
    if ((st = v->alloc(&v->h16, M * D))) break;
    if ((st = v->alloc(&v->qkv, M * 3 * D))) break;
    if ((st = v->alloc(&v->att, M * D))) break;
    if ((st = v->alloc(&v->mid, M * mlp_dim))) break;
  } while (0);
  if (st) {
    delete v;
    return st;
  }
  *out = reinterpret_cast<effocr_vit_t>(v);
  return EFFOCR_OK;
}

extern "C" void effocr_vit_destroy(effocr_vit_t h) { delete reinterpret_cast<VitHandle*>(h); }

extern "C" int effocr_vit_embed_dim(effocr_vit_t h) { return h ? reinterpret_cast<VitHandle*>(h)->D : 0; }
extern "C" int effocr_vit_max_batch(effocr_vit_t h) { return h ? reinterpret_cast<VitHandle*>(h)->max_batch : 0; }
extern "C" void* effocr_vit_patch_buffer(effocr_vit_t h) { return h ? reinterpret_cast<VitHandle*>(h)->patches : nullptr; }

extern "C" int effocr_vit_forward(effocr_vit_t h, const void* d_input, int input_kind, int batch, float* d_emb,
                                  void* stream) {
  VitHandle* v = reinterpret_cast<VitHandle*>(h);
  if (!v || !d_emb || batch < 0) return fail(EFFOCR_ERR_INVALID, "vit_forward: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (input_kind == EFFOCR_INPUT_PATCH_BUFFER) {
    if (batch > v->max_batch) return fail(EFFOCR_ERR_INVALID, "vit_forward: batch exceeds max_batch for the in-place patch buffer");
    return batch == 0 ? EFFOCR_OK : vit_forward_chunk(v, batch, d_emb, s);
  }
  if (!d_input) return fail(EFFOCR_ERR_INVALID, "vit_forward: null input");
  for (int b0 = 0; b0 < batch; b0 += v->max_batch) {
    const int B = (batch - b0 < v->max_batch) ? batch - b0 : v->max_batch;
    if (input_kind == EFFOCR_INPUT_NCHW_F32) {
      const float* img = reinterpret_cast<const float*>(d_input) + size_t(b0) * 3 * 224 * 224;
      const long long total = static_cast<long long>(B) * 196 * 96;
      const int grid = static_cast<int>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
      {
        KernelScope ks(PROF_IM2PATCH, s);
        im2patch_kernel<<<grid, 256, 0, s>>>(img, v->patches, B);
      }
      EFFOCR_CUDA(cudaGetLastError());
    } else if (input_kind == EFFOCR_INPUT_PATCH_F16) {
      const __half* p = reinterpret_cast<const __half*>(d_input) + size_t(b0) * 196 * 768;
      EFFOCR_CUDA(cudaMemcpyAsync(v->patches, p, size_t(B) * 196 * 768 * 2, cudaMemcpyDeviceToDevice, s));
    } else {
      return fail(EFFOCR_ERR_INVALID, "vit_forward: unknown input_kind");
    }
    EFFOCR_TRY(vit_forward_chunk(v, B, d_emb + size_t(b0) * v->D, s));
  }
  return EFFOCR_OK;
}

// ---- building-block entry points (unit tests call these through the C ABI)
extern "C" int effocr_layernorm(const float* d_x, long long ldx, const float* d_gamma, const float* d_beta, void* d_out,
                                long long ldo, int rows, int dim, float eps, int out_f32, void* stream) {
  EFFOCR_TRY(require_sm100());
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (rows <= 0) return EFFOCR_OK;
  return out_f32 ? layernorm_f32(d_x, ldx, d_gamma, d_beta, reinterpret_cast<float*>(d_out), ldo, rows, dim, eps, s, PROF_LAYERNORM)
                 : layernorm_f16(d_x, ldx, d_gamma, d_beta, reinterpret_cast<__half*>(d_out), ldo, rows, dim, eps, s, PROF_LAYERNORM);
}

extern "C" int effocr_attention_f16(const void* d_qkv, void* d_out, int batch, int tokens, int heads, int impl,
                                    void* stream) {
  EFFOCR_TRY(require_sm100());
  if (batch <= 0) return EFFOCR_OK;
  return attention_f16(reinterpret_cast<const __half*>(d_qkv), reinterpret_cast<__half*>(d_out), batch, tokens, heads,
                       reinterpret_cast<cudaStream_t>(stream), impl);
}
