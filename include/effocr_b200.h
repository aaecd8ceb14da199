/* effocr_b200 -- C ABI of the B200-native EffOCR inference hot path.
 *
 * The reference (dell-research-harvard/effocr) is pure Python and reaches its arithmetic only
 * through third-party libraries (timm / onnxruntime / faiss / torchvision / OpenCV).  This
 * header is the boundary a maintainer would bind (ctypes stub in INTEGRATION.md) to replace
 * those library calls on the hot path  localize -> crop -> embed -> kNN  (SURVEY.md section 8a).
 * Each entry point names the reference call site it replaces.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory on the current CUDA device (sm_100a required);
 *     h_* is host memory; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - all calls are stream-ordered and never synchronise unless documented;
 *   - return value 0 = success, otherwise an EFFOCR_ERR_* code; effocr_last_error() returns a
 *     thread-local message; nothing throws, the caller owns every buffer it passes;
 *   - handles are immutable after create and may be used from several host threads as long as
 *     each thread passes its own workspace-owning handle or serialises calls on one handle
 *     (the Python shim holds a lock per handle, mirroring ORT's thread-safe session.run()).
 */
#ifndef EFFOCR_B200_H_
#define EFFOCR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFFOCR_B200_ABI_VERSION 1

#if defined(EFFOCR_BUILDING)
#define EFFOCR_API __attribute__((visibility("default")))
#else
#define EFFOCR_API
#endif

#define EFFOCR_OK 0
#define EFFOCR_ERR_INVALID 1
#define EFFOCR_ERR_CUDA 2
#define EFFOCR_ERR_NO_DEVICE 3
#define EFFOCR_ERR_NOMEM 4

/* ---- library ---------------------------------------------------------------------------- */
EFFOCR_API int effocr_abi_version(void);
/* bit 0: the library was built with -DEFFOCR_AB (A/B kernel variants kept for experiments: earlier attention kernels, the
 * 64-wide fused-MLP schedule at D = 384, direct-store GEMM epilogues, im2col convolutions).  The product build is 0. */
EFFOCR_API int effocr_build_flags(void);
EFFOCR_API const char* effocr_last_error(void);
/* 0 when the current CUDA device is an sm_100 part, EFFOCR_ERR_NO_DEVICE / _CUDA otherwise. */
EFFOCR_API int effocr_device_ok(void);

/* ---- launch accounting and per-kernel timing (what bench.py's gpu_launches / roofline report) ----
 * effocr_launch_count: kernels this library has launched since load.  With profiling enabled every
 * launch is bracketed by CUDA events on its own stream; effocr_profile_read synchronises the device
 * and returns launches / summed milliseconds for one kernel class (names: effocr_profile_tag_name). */
EFFOCR_API long long effocr_launch_count(void);
EFFOCR_API void effocr_profile_enable(int on);
EFFOCR_API void effocr_profile_reset(void);
EFFOCR_API int effocr_profile_num_tags(void);
EFFOCR_API const char* effocr_profile_tag_name(int tag);
EFFOCR_API int effocr_profile_read(int tag, long long* launches, double* total_ms);

/* ---- tcgen05 GEMM building block ----------------------------------------------------------
 * out[M,N] = act(A[M,K] . W[N,K]^T + bias) * gamma + resid      fp16 operands, fp32 accumulate.
 * Replaces the ATen/oneDNN/ORT matmul + elementwise calls inside timm Block.forward and yolov5
 * Conv.forward that the reference reaches through models/encoders.py:62-64 and
 * onnx_engines/recognizer_engine.py:27 / localizer_engine.py:54.
 * act: 0 none, 1 GELU(erf), 2 SiLU.  out_f32: element type of out and resid (0 fp16, 1 fp32).
 * block_n: 0 = auto, else 64/128/192/256 (| 0x10000 forces the direct-store epilogue instead of the
 * TMA-store / TMA-reduce one, | 0x20000 keeps the TMA epilogue but disables the A-stationary schedule).  bias/gamma fp32 [N] or NULL; resid may alias out (in-place update). */
EFFOCR_API int effocr_gemm_f16(const void* d_A, long long lda, const void* d_W, long long ldw, int M, int N, int K,
                    const float* d_bias, const float* d_gamma, const void* d_resid, long long ldr, void* d_out,
                    long long ldo, int act, int out_f32, int block_n, void* stream);

/* ---- fused transformer MLP block (tcgen05, CTA pairs) -------------------------------------
 * x[M,D] += GELU(h[M,D] . W1[HID,D]^T + b1) . W2[D,HID]^T + b2     fp16 operands, fp32 accumulate / residual.
 * The `x = x + mlp(norm2(x))` half of timm Block.forward (un-vendored; reached from models/encoders.py:58,62-64)
 * after the LayerNorm; the [M,HID] hidden activations stay on chip.  D in {192, 384}, HID % 128 == 0. */
EFFOCR_API int effocr_mlp_fused_f16(const void* d_h, long long ldh, const void* d_w1, const float* d_b1, const void* d_w2,
                         const float* d_b2, float* d_x, long long ldx, int M, int D, int HID, void* stream);

/* ---- fused attention projection + residual + LayerNorm (tcgen05, CTA pairs, full-row tiles) ------
 * x[M,D] += att[M,D] . Wp[D,D]^T + bp;   h[M,D] = LayerNorm(x) * gamma + beta (fp16).
 * The `x = x + attn(...)` tail and the `norm2` of timm Block.forward (un-vendored; models/encoders.py:58,62-64) in
 * one kernel: no L2 reductions, x is read and written once.  D = 384. */
EFFOCR_API int effocr_proj_ln_f16(const void* d_att, long long lda, const void* d_w, const float* d_bias, float* d_x,
                       long long ldx, const float* d_gamma, const float* d_beta, float eps, void* d_h, long long ldh,
                       int M, int D, void* stream);

/* ---- the second half of an encoder block in one kernel (tcgen05, CTA pairs, full-row tiles) -----
 * x[M,D] += att[M,D] . Wp[D,D]^T + bp;   x[M,D] += GELU(LayerNorm(x) * gamma + beta . W1[HID,D]^T + b1) . W2[D,HID]^T + b2.
 * `x = x + attn(...)` (projection part) and `x = x + mlp(norm2(x))` of timm Block.forward (un-vendored;
 * models/encoders.py:58,62-64): the residual stream is read once and written once, the normalised operand and the
 * hidden activations stay on chip.  D = 384, HID % 128 == 0. */
EFFOCR_API int effocr_block_tail_f16(const void* d_att, long long lda, const void* d_wp, const float* d_bp, const float* d_gamma,
                          const float* d_beta, float eps, const void* d_w1, const float* d_b1, const void* d_w2,
                          const float* d_b2, float* d_x, long long ldx, int M, int D, int HID, void* stream);

/* ---- K1: fused crop -> square white pad -> AA bilinear 224x224 -> normalise -------------------
 * Replaces the numpy slice + `create_paired_transform` per-crop CPU path
 * (infer_effocr.py:284-293, infer_effocr_onnx_multi.py:307-340, utils/datasets_utils.py:69-90,166-172).
 * `pixels` holds u8 RGB HWC line images; images[i] locates image i inside it; boxes[j] is the
 * Python slice im[y0:y1, x0:x1] (numpy semantics: negative indices wrap, ends clamp) of image
 * boxes[j].image.  An empty slice produces an all-zero output crop (the reference's ONNX path feeds
 * zeros for a failed transform, infer_effocr_onnx_multi.py:151-152,200-204). */
typedef struct {
  long long offset; /* byte offset of pixel (0,0) inside `pixels` */
  int height, width;
  int pitch;        /* bytes between rows (>= 3 * width) */
  int reserved;
} effocr_image_desc;
typedef struct {
  int image;
  int x0, y0, x1, y1;
} effocr_crop_box;
#define EFFOCR_CROP_NCHW_F16 0  /* out: fp16 [n, 3, 224, 224] */
#define EFFOCR_CROP_NCHW_F32 1  /* out: fp32 [n, 3, 224, 224]  (== create_paired_transform output) */
/* The two patch-major layouts are WHITE-CENTRED: they hold (normalised value - white level), white level =
 * ((1 - mean_c) / std_c) = (2.2489083, 2.4285715, 2.6400001): the white padding / background of a crop is then exactly 0
 * in fp16 instead of carrying the same rounding error on every pixel; the encoders fold W . white into their bias. */
#define EFFOCR_CROP_PATCH_F16 2 /* out: fp16 [n * 196, 768] patch-major (input of effocr_vit_forward) */
#define EFFOCR_CROP_PATCH4_F16 3 /* out: fp16 [n * 3136, 48] 4x4-patch-major (input of effocr_convnext_forward) */
EFFOCR_API int effocr_crop_resize(const uint8_t* d_pixels, const effocr_image_desc* d_images,
                                  const effocr_crop_box* d_boxes, int n_boxes, int layout, void* d_out, void* stream);

/* ---- a1: letterbox without resize (r == 1) on the device ------------------------------------------
 * Replaces EffLocalizer.load_localizer_img / letterbox (onnx_engines/localizer_engine.py:75-85,107-138) when the
 * line image already fits the model shape: pad to (height, width) with grey 114 (top = round(dh/2 - 0.1),
 * left = round(dw/2 - 0.1)), RGB order, / 255 -> fp32 [n, 3, height, width].  Images larger than the canvas are
 * not handled here (the resize path stays on the host with OpenCV, like the reference). */
EFFOCR_API int effocr_letterbox_pad(const uint8_t* d_pixels, const effocr_image_desc* d_images, int n_images, int height,
                                    int width, float* d_out, void* stream);

/* ---- a1: full letterbox (resize + pad) on the device ------------------------------------------------
 * Replaces EffLocalizer.load_localizer_img / letterbox (onnx_engines/localizer_engine.py:75-85,107-138) for any line
 * shape: cv2.resize(INTER_LINEAR) on u8 restated in its own fixed-point arithmetic (bit-exact), grey-114 padding,
 * RGB order, / 255 -> fp32 [n, 3, height, width].  One plan per image; taps = int32 quadruples
 * (source index 0, source index 1, coefficient 0, coefficient 1), coefficients scaled by 2048, built on the host by
 * effocr_b200.localizer_engine.letterbox_plan with OpenCV's float arithmetic. */
typedef struct effocr_letterbox_plan {
  int32_t new_width, new_height; /* size after the resize (cv2 `new_unpad`) */
  int32_t left, top;             /* padding before the content */
  int32_t xtap_offset, ytap_offset; /* first tap of this image's column / row table inside d_taps (in quadruples) */
} effocr_letterbox_plan;
EFFOCR_API int effocr_letterbox_resize(const uint8_t* d_pixels, const effocr_image_desc* d_images,
                                       const effocr_letterbox_plan* d_plans, const void* d_taps, int n_images, int height,
                                       int width, float* d_out, void* stream);

/* ---- recognizer encoder: timm vit_{tiny,small,base}_patch16_224, num_classes=0 ----------------
 * Replaces AutoEncoder.forward (models/encoders.py:62-64, called at infer_effocr.py:314) and
 * EffRecognizer.run (onnx_engines/recognizer_engine.py:23-27).
 * h_weights: HOST fp32 tensors in timm layout, in this order (n_weights = 4 + 12*depth + 2):
 *   patch_embed.proj.weight [D,3,16,16], patch_embed.proj.bias [D], cls_token [D], pos_embed [197,D],
 *   per block: norm1.weight, norm1.bias, attn.qkv.weight [3D,D], attn.qkv.bias, attn.proj.weight [D,D],
 *              attn.proj.bias, norm2.weight, norm2.bias, mlp.fc1.weight [4D,D], mlp.fc1.bias,
 *              mlp.fc2.weight [D,4D], mlp.fc2.bias,
 *   norm.weight, norm.bias.
 * ln_eps: LayerNorm epsilon (timm ViT: 1e-6).
 * The handle owns its weights and a workspace for max_batch crops (larger batches are chunked). */
typedef struct effocr_vit_s* effocr_vit_t;
#define EFFOCR_INPUT_NCHW_F32 0     /* d_input: fp32 [B,3,224,224] (the reference's tensor) */
#define EFFOCR_INPUT_PATCH_F16 1    /* d_input: fp16 [B*196,768] patch-major, white-centred (what EFFOCR_CROP_PATCH_F16 writes) */
#define EFFOCR_INPUT_PATCH_BUFFER 2 /* input already written into effocr_vit_patch_buffer(); B <= max_batch */
EFFOCR_API int effocr_vit_create(int embed_dim, int num_heads, int depth, int mlp_dim, int max_batch, float ln_eps,
                                 const float* const* h_weights, int n_weights, effocr_vit_t* out);
EFFOCR_API void effocr_vit_destroy(effocr_vit_t h);
EFFOCR_API int effocr_vit_embed_dim(effocr_vit_t h);
EFFOCR_API int effocr_vit_max_batch(effocr_vit_t h);
EFFOCR_API void* effocr_vit_patch_buffer(effocr_vit_t h);
/* d_emb: fp32 [B, D] pooled pre-logits (final-LayerNorm'd CLS token), NOT L2-normalised. */
EFFOCR_API int effocr_vit_forward(effocr_vit_t h, const void* d_input, int input_kind, int batch, float* d_emb,
                                  void* stream);

/* ---- recognizer encoder: timm convnext_tiny, num_classes=0 (BASELINE config 4) --------------------
 * Same call sites as the ViT handle (models/encoders.py:58,62-64).  h_weights: HOST fp32 tensors in timm order
 * (n_weights = 180): stem.0.weight [96,3,4,4], stem.0.bias, stem.1.weight, stem.1.bias; per stage i = 0..3:
 * (i > 0: downsample.0.weight, downsample.0.bias, downsample.1.weight [C,Cp,2,2], downsample.1.bias), per block:
 * conv_dw.weight [C,1,7,7], conv_dw.bias, norm.weight, norm.bias, mlp.fc1.weight [4C,C], mlp.fc1.bias,
 * mlp.fc2.weight [C,4C], mlp.fc2.bias, gamma [C]; then head.norm.weight, head.norm.bias.
 * Input kinds: EFFOCR_INPUT_NCHW_F32, or EFFOCR_INPUT_PATCH_BUFFER after effocr_crop_resize wrote
 * EFFOCR_CROP_PATCH4_F16 rows into effocr_convnext_patch_buffer().  d_emb: fp32 [B, 768]. */
typedef struct effocr_convnext_s* effocr_convnext_t;
EFFOCR_API int effocr_convnext_create(int max_batch, const float* const* h_weights, int n_weights, effocr_convnext_t* out);
EFFOCR_API void effocr_convnext_destroy(effocr_convnext_t h);
EFFOCR_API void* effocr_convnext_patch_buffer(effocr_convnext_t h);
EFFOCR_API int effocr_convnext_forward(effocr_convnext_t h, const void* d_input, int input_kind, int batch, float* d_emb,
                                       void* stream);

/* building blocks of the encoder, exported for the parity tests */
/* out fp16 [M, N] = (LayerNorm(x fp32 [M, 384]) * gamma + beta) . W[N, 384]^T + bias: timm Block's attn.qkv(norm1(x)) in one
 * kernel -- the normalised operand is produced on the SM and never reaches HBM.  D must be 384 (ViT-S). */
EFFOCR_API int effocr_ln_gemm_f16(const float* d_x, long long ldx, const float* d_gamma, const float* d_beta, float eps,
                                  const void* d_w, long long ldw, const float* d_bias, void* d_out, long long ldo, int M,
                                  int N, int D, void* stream);
EFFOCR_API int effocr_layernorm(const float* d_x, long long ldx, const float* d_gamma, const float* d_beta,
                                void* d_out, long long ldo, int rows, int dim, float eps, int out_f32, void* stream);
/* qkv fp16 [B*197, 3*H*64] (rows [q|k|v], each [H,64]) -> out fp16 [B*197, H*64].
 * impl 0: tcgen05 kernel (S and O in TMEM, V consumed in place as an MN-major operand); impl 1: mma.sync kernel. */
EFFOCR_API int effocr_attention_f16(const void* d_qkv, void* d_out, int batch, int tokens, int heads, int impl,
                                    void* stream);

/* ---- localizer: YOLOv5s forward (ultralytics yolov5s.yaml, nc classes) ---------------------------
 * Replaces the onnxruntime session behind EffLocalizer.run (onnx_engines/localizer_engine.py:25-29,49-55).
 * h_weights: HOST fp32 tensors in graph order (n_weights = 57 * 5 + 7): for every Conv block
 * (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var) in the order layer 0, 1,
 * C3@2 (cv1, cv2, cv3, m.0.cv1, m.0.cv2), 3, C3@4, 5, C3@6, 7, C3@8, SPPF (cv1, cv2), 10, C3@13, 14, C3@17,
 * 18, C3@20, 21, C3@23; then Detect m.0.weight, m.0.bias, m.1.*, m.2.*, anchors [3,3,2] (stride units).
 * BatchNorm (eps 1e-3) is folded at create.  d_images: fp32 [B,3,H,W] RGB in [0,1] (the letterboxed
 * tensor EffLocalizer.load_localizer_img builds), H and W multiples of 32 with H*W <= max_h*max_w.
 * d_pred: fp32 [B, effocr_yolo_num_predictions(H, W), 5+nc] = (cx, cy, w, h, obj, cls...) in input pixels. */
typedef struct effocr_yolo_s* effocr_yolo_t;
EFFOCR_API int effocr_yolo_create(int nc, int max_batch, int max_h, int max_w, const float* const* h_weights,
                                  int n_weights, effocr_yolo_t* out);
EFFOCR_API void effocr_yolo_destroy(effocr_yolo_t h);
/* Arithmetic of the forward pass.  The reference runs the detector in fp32 (onnxruntime CPU session,
 * localizer_engine.py:54).  EFFOCR_YOLO_SPLIT (default): every activation is a pair of fp16 planes (hi, lo) and every
 * folded weight [Whi | Wlo]; each convolution accumulates hi x Whi + lo x Whi + hi x Wlo in one fp32 TMEM accumulator,
 * i.e. fp32-accurate products on the fp16 tensor cores -- decoded boxes and confidences agree with the fp32 path to
 * ~1e-5, so thresholds, NMS decisions and crop-rectangle rounding come out the same.  EFFOCR_YOLO_FP16: single fp16
 * plane, ~2x faster, boxes within ~1 px / confidences within ~6e-3 of fp32 (a few per cent of lines then differ). */
#define EFFOCR_YOLO_SPLIT 0
#define EFFOCR_YOLO_FP16 1
EFFOCR_API int effocr_yolo_set_mode(effocr_yolo_t h, int mode);
EFFOCR_API int effocr_yolo_num_predictions(int height, int width);
EFFOCR_API int effocr_yolo_forward(effocr_yolo_t h, const float* d_images, int batch, int height, int width,
                                   float* d_pred, void* stream);

/* ---- K9: confidence filter + class-aware greedy NMS ---------------------------------------------
 * Replaces EffLocalizer.non_max_suppression incl. torchvision.ops.nms (localizer_engine.py:171-277):
 * obj > conf -> conf = obj * cls -> best class, conf > thr -> sort by conf desc -> boxes offset by cls*7680 ->
 * drop IoU > iou_thres -> first max_det.  d_out: fp32 [B, max_det, 6] (x1,y1,x2,y2,conf,cls), d_count: int [B]. */
EFFOCR_API int effocr_nms(const float* d_pred, int batch, int npred, int no, float conf_thres, float iou_thres,
                          int max_det, float* d_out, int* d_count, void* stream);

/* ---- a9: row-wise L2 normalisation, x / max(||x||, eps) -------------------------------------
 * Replaces torch.nn.functional.normalize at infer_effocr.py:316 / infer_effocr_onnx_multi.py:371. */
EFFOCR_API int effocr_l2_normalize(const float* d_x, float* d_out, int rows, int dim, float eps, void* stream);

/* ---- K10: exact inner-product top-k over a flat index ------------------------------------------
 * Replaces faiss.IndexFlatIP (add / search) behind pytorch_metric_learning.FaissKNN
 * (infer_effocr.py:184-187,317; infer_effocr_onnx_multi.py:496-500,372).
 * create copies d_vectors (fp32 [n, dim], device).  search writes, per query, the k best
 * (inner product desc, id asc) as fp32 distances and int64 ids; missing results are
 * (-3.4028235e38, -1) like faiss.  k <= 32.  A handle's search workspace is shared: serialise
 * concurrent searches on one handle. */
typedef struct effocr_knn_s* effocr_knn_t;
EFFOCR_API int effocr_knn_create(const float* d_vectors, int n, int dim, effocr_knn_t* out);
EFFOCR_API void effocr_knn_destroy(effocr_knn_t h);
EFFOCR_API int effocr_knn_ntotal(effocr_knn_t h);
EFFOCR_API int effocr_knn_dim(effocr_knn_t h);
EFFOCR_API const float* effocr_knn_vectors(effocr_knn_t h);
EFFOCR_API int effocr_knn_search(effocr_knn_t h, const float* d_queries, int nq, int k, float* d_dist,
                                 long long* d_idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EFFOCR_B200_H_ */
