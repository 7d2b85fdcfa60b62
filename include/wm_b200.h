/* wm_b200 -- C ABI of the B200-native WildlifeMapper tile-detection hot path (libwm_b200.so).
 *
 * The reference (lgemc/WildlifeMapper) has no FFI layer: every FLOP of its hot path runs inside PyTorch /
 * torchvision calls made from the Python modules under wildlifemapper/segment_anything/.  The boundary this
 * library sits behind is therefore that module surface; each entry point below names the
 * reference call site(s) it replaces.  The drop-in `segment_anything` package in this repository binds these
 * symbols with ctypes (wildlifemapper_b200/lib.py) and registers them as torch.library ops
 * (wildlifemapper_b200/ops.py).  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless noted; the caller owns all memory
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises, never
 *     allocates, and is CUDA-graph capturable
 *   - return value: 0 = ok, <0 = WM_ERR_*; wm_last_error() returns a thread-local message
 *   - bf16 = __nv_bfloat16 bits (uint16_t); leading dimensions are in ELEMENTS
 *   - requires compute capability 10.0 (sm_100a); anything else returns WM_ERR_ARCH -- there is no fallback
 */
#ifndef WM_B200_H
#define WM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WM_OK 0
#define WM_ERR_SHAPE (-1)
#define WM_ERR_ALIGN (-2)
#define WM_ERR_ARCH (-3)
#define WM_ERR_CUDA (-4)

#define WM_ACT_NONE 0
#define WM_ACT_GELU_ERF 1
#define WM_ACT_RELU 2
#define WM_ACT_SIGMOID 3

int wm_version(void);
const char* wm_last_error(void);
/* 0 if the current device is sm_100 and the driver exposes cuTensorMapEncodeTiled */
int wm_device_check(void);
/* Select the flash-attention kernel generation used by wm_attn_flash (A/B measurements; the superseded generations
 * 1-3, 5 and 6 of round 1 were measured slower and removed). */
int wm_set_flash_version(int version);
/* Measurement knobs (atomic, read once per call): "flash_version", "gemm_pairs" (0|1: large GEMMs on the CTA-pair kernel). */
int wm_set_option(const char* name, int value);
/* Diagnostics build only (csrc/build.sh with -DWM_F3_TRACE): copy the SM-clock event trace of CTA (0,0,0) of the last
 * flash-attention launch to host_out[3][64][8]; WM_ERR_ARCH in the product build. */
int wm_debug_flash_trace(uint64_t* host_out_3x64x4);
/* Same for CTA 0 of the last windowed-attention (v2) launch: host_out[3][64][8]. */
int wm_debug_window_trace(uint64_t* host_out_3x64x8);

/* C[M,N] = act(A[M,K] * W[N,K]^T + bias) + residual[(m % res_mod), :]      (tcgen05 / TMEM / TMA)
 * Replaces every nn.Linear / 1x1 Conv2d / patch-embed Conv2d(k16,s16) on the path:
 *   image_encoder.py:234-235,249,260 (qkv, proj)  common.py:21-26 (MLPBlock)  image_encoder.py:409-414,442-447
 *   (PatchEmbed / HfcEmbed after wm_patchify / wm_hfc_finalize)  :468-481,494-513 (HFC branch linears, proj_back)
 *   :105-111 (neck 1x1)  transformer.py:203-206 (decoder projections)  box_decoder.py:68-69,154-176 (heads).
 * A, W bf16; bias fp32 [N] or NULL; residual fp32 or NULL; out_bf16 and/or out_f32 may be NULL (not both).
 * bn_hint: 0 = auto, or 64 / 128 / 256 (N tile). */
int wm_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
                 int64_t ldr, int res_mod, void* out_bf16, int64_t ldc_bf16, float* out_f32, int64_t ldc_f32, int M,
                 int N, int K, int act, int bn_hint, void* stream);

/* SM count of the current device (> 0), or a negative WM_ERR_* code */
int wm_num_sms(void);

/* 3x3, pad 1, no-bias convolution on NHWC bf16 [B,64,64,C] as an implicit GEMM; W is [N, 9*C] bf16 with
 * k = (dy*3+dx)*C + c.  Replaces neck[2], image_encoder.py:113-119.  Output rows = pixels, [B*4096, N]. */
int wm_conv3x3_nhwc_bf16(const void* X, const void* W, void* out_bf16, float* out_f32, int B, int C, int N,
                         void* stream);

/* LayerNorm over the last dim of fp32 [rows, D] (D in {128,256,768,1024,1280}); fp32 statistics.
 * Outputs (each nullable): y_bf16, y_f32, y2_bf16 = bf16(y + add[(row % add_mod), :]).
 * Replaces nn.LayerNorm at image_encoder.py:173,183,476-477, transformer.py:58,135-145 and LayerNorm2d
 * (common.py:31-43) applied on NHWC rows. */
int wm_layernorm(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, const float* add,
                 int add_mod, void* y2_bf16, int rows, int D, float eps, void* stream);

/* NCHW fp32 tile batch [B,C,1024,1024] (C = 3 or 1) -> bf16 im2col rows [B*4096, C*256] (k = c*256 + ky*16 + kx)
 * for the patch-embed / hfc-embed GEMMs (image_encoder.py:409-417, 442-450) and, if C == 3 and gray != NULL, the
 * bf16 grayscale plane [B,1024,1024] (0.2989 R + 0.587 G + 0.114 B, network.py:41).  gray_split = 1: the plane is
 * written as rows of 3072 = [hi | lo | hi], hi = bf16(g), lo = bf16(g - hi): the A operand of the SPLIT low-pass GEMM
 * (weight [Lh | Lh | Ll]), which reproduces the reference's fp32 fft2 / ifft2 (network.py:43-55) to ~2e-6 instead of
 * the 1e-3 of single bf16 operands (x_hfc is a small difference of O(1) numbers on smooth imagery). */
int wm_patchify(const float* img, void* patches_bf16, void* gray_bf16, int gray_split, int B, int C, void* stream);

/* Batched 2-D transpose: in [batch, R, C] -> out [batch, C, R]; elt_bytes 2 or 4. */
int wm_transpose(const void* in, void* out, int batch, int R, int C, int elt_bytes, void* stream);

/* fp32 in [batch, R, C] -> bf16 hi / lo split of its transpose: out[b][c / 2][seg][c % 2][r], seg 0 and 2 = bf16(v),
 * seg 1 = bf16(v - bf16(v)); R, C multiples of 64.  With in = the first low-pass product [B, 1024 (y), 2048 (x', re/im)]
 * this is the A operand [B*1024, 6144] of the second split low-pass GEMM (network.py:43-55). */
int wm_transpose_split(const float* in, void* out_bf16, int batch, int R, int C, void* stream);

/* x_hfc = |gray(img) - lowpass| (network.py:53-55) where low_t [B,1024(x),1024(y)] fp32 is the TRANSPOSED low-pass
 * image from the two DFT-operator GEMMs; writes hfc_embed im2col rows [B*4096, 256] bf16 (image_encoder.py:442-450)
 * and, if hfc_img != NULL, the fp32 image [B,1024,1024]. */
int wm_hfc_finalize(const float* img, const float* low_t, void* patches_bf16, float* hfc_img, int B, void* stream);

/* out_bf16[row, :] = a[row, :] + b[(row % b_mod), :]   (b NULL: plain cast; a NULL: batch-broadcast cast of b; not both).
 * fp32 in, D % 4 == 0. */
int wm_add_cast(const float* a, const float* b, int b_mod, void* out_bf16, int rows, int D, void* stream);

/* Fused flash attention (tcgen05): out[b*Tq+t, h*hd+d] = softmax_k(scale * q.k [+ rel-pos]) v.
 * q/k/v are bf16 matrices with `*_rows` rows of `*_width` columns (leading dim ld*), head h of q at column
 * q_col0 + h*hd (same for k, v); rows of image b start at b*Tq (q) / b*Tk (k, v).  hd in {64,128}; Tq,Tk % 128 == 0.
 * rel_table: NULL, or bf16 [256, 64] (rows 0..126 = rel_pos_h, 128..254 = rel_pos_w, others 0) for the 64x64
 * global blocks -- the bias uses the UNSCALED q (image_encoder.py:253-256, 347-383).
 * Replaces Attention.forward core for global blocks (image_encoder.py:251-259) and nn.MultiheadAttention in
 * CrossAttentionHfcPatch (image_encoder.py:500-503). */
int wm_attn_flash(const void* q, int64_t q_rows, int64_t q_width, int64_t ldq, int q_col0, const void* k,
                  int64_t k_rows, int64_t k_width, int64_t ldk, int k_col0, const void* v, int64_t v_rows,
                  int64_t v_width, int64_t ldv, int v_col0, const void* rel_table, void* out_bf16, int64_t ldo, int B,
                  int H, int Tq, int Tk, int hd, float scale, void* stream);

/* Fused 14x14 windowed attention on qkv bf16 [B,64,64,3D] (q | k | v, head-major, hd = D / H = 64 or 80) written straight to
 * [B,64,64,D]; window partition, zero padding to 70x70, unpartition and crop happen in TMA coordinates.
 * The qkv GEMM must have been run with the k and v biases DROPPED (pad keys, see attn_window.cu).
 * rel_table: bf16 [64, hd]: rows 0..26 = rel_pos_h, 32..58 = rel_pos_w, others 0.
 * Replaces window_partition + Attention.forward + window_unpartition, image_encoder.py:192-199,246-311. */
int wm_attn_window(const void* qkv, const void* rel_table, void* out_bf16, int B, int H, int D, float scale,
                   void* stream);

/* Decoder attention on CUDA cores, hd in {16,32}: transformer.py:218-240. */
int wm_attn_small(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out_bf16,
                  int64_t ldo, int B, int H, int Tq, int Tk, int hd, float scale, void* stream);

/* PostProcess.forward (build_sam.py:219-258): logits fp32 [B,Q,C1] (C1 = classes + no-object), boxes fp32
 * [B,Q,4] cxcywh, sizes int64 [B,2].  packed fp32 [B,Q,6] = (x1,y1,x2,y2,score,label) compacted in query
 * order, query_idx int32 [B,Q], labels int64 [B,Q] (nullable), counts int32 [B].  from_prob != 0: `logits` already holds softmax
 * probabilities (integer-stage parity entry). */
int wm_postprocess(const float* logits, const float* boxes, const int64_t* sizes, float thr, int from_prob,
                   float* packed, int32_t* query_idx, int64_t* labels, int32_t* counts, int B, int Q, int C1,
                   void* stream);

/* sigmoid over the first C of C1 logits, stable top-K over Q*C (ties: lower flat index first).
 * prob_ws fp32 [B,Q*C] (input when from_prob != 0), order_ws int32 [B,Q*C].
 * Outputs scores fp32 [B,K], labels int32 [B,K], query int32 [B,K], out_boxes fp32 [B,K,4] (cxcywh). */
int wm_sigmoid_topk(const float* logits, const float* boxes, float* prob_ws, int32_t* order_ws, float* scores,
                    int32_t* labels, int32_t* query, float* out_boxes, int B, int Q, int C1, int C, int K,
                    int from_prob, void* stream);

/* Greedy NMS with torchvision.ops.nms semantics (visualize_prediction.py:150-154); labels != NULL -> per-class.
 * boxes fp32 [n,4] xyxy, scores fp32 [n], labels int64 [n].  order_ws int32 [n], mask_ws uint64
 * [n * ceil(n/64)].  keep int64 [n] (score-descending), num_keep int32 [1]. */
int wm_nms(const float* boxes, const float* scores, const int64_t* labels, int n, double iou_thr, int32_t* order_ws,
           uint64_t* mask_ws, int64_t* keep, int32_t* num_keep, void* stream);

/* Batched NMS over the packed PostProcess rows, one CTA per image, no host round trip (Q <= 1024):
 * candidates = rows [0, counts[b]) with score > score_thr (visualize_prediction.py:150), then the same greedy NMS
 * as wm_nms (per_class != 0: only same-label boxes suppress each other).  keep_idx int32 [B,Q]: kept row indices in
 * score order; keep_cnt int32 [B]. */
int wm_nms_batched(const float* packed, const int32_t* counts, int B, int Q, float score_thr, double iou_thr,
                   int per_class, int32_t* keep_idx, int32_t* keep_cnt, void* stream);

/* ---- either side of the path: uint8 tile front-end, image-level merge, COCO packing ---- */

/* Tiles cut out of ONE uint8 HWC survey image (device memory, `row_stride` bytes between rows) at origins int32 [T,2] =
 * (y0, x0) (device): out fp32 [T,3,1024,1024] = ((u8 / 255) - mean[c]) / std[c] for the first content_h x content_w
 * pixels of each tile that lie inside the image, 0 elsewhere.  Replaces torchvision to_tensor + normalize
 * (dataloader_coco.py:286-292, utils/augmentation.py:229-249) and the zero padding of nested_tensor_from_tensor_list
 * (utils/misc.py:46-67); bit-exact (IEEE division, same operation order).  mean/std: host pointers to 3 floats. */
int wm_tiles_from_u8(const uint8_t* img, int H, int W, int64_t row_stride, const int32_t* origins, int T, int content_h,
                     int content_w, const float* mean3, const float* std3, float* out, void* stream);

/* PIL-exact bilinear (antialiased) resize of T uint8 RGB tiles (tile_h x tile_w windows of ONE uint8 HWC image at origins
 * int32 [T,2] = (y0, x0), fully inside the image) to out_h x out_w: out uint8 [T,out_h,out_w,3]; tmp uint8 [T,tile_h,out_w,3]
 * is the intermediate of the horizontal pass.  Replaces RandomResize([768], max_size=768) of the reference's transforms
 * (dataloader_coco.py:275-292 -> utils/augmentation.py:77-107 -> torchvision F.resize -> PIL Image.resize(BILINEAR)).
 * xbounds / ybounds int32 [out,2] = (first input index, tap count), xk / yk int32 [out, ksize] = PIL's 22-bit fixed-point
 * coefficients (device pointers; the host computes them exactly as Pillow's precompute_coeffs + normalize_coeffs_8bpc do).
 * Bit-exact with Pillow: 8-bit intermediate, (2^21 + sum) >> 22, clipped. */
int wm_resize_tiles_u8(const uint8_t* img, int img_h, int img_w, int64_t row_stride, const int32_t* origins, int T, int tile_h,
                       int tile_w, uint8_t* tmp, uint8_t* out, int out_h, int out_w, const int32_t* xbounds, const int32_t* xk,
                       int xksize, const int32_t* ybounds, const int32_t* yk, int yksize, void* stream);

/* Image-level candidate list from the packed per-tile PostProcess rows (packed fp32 [T,Q,6], counts int32 [T], Q <= 1024):
 * rows with score > score_thr (visualize_prediction.py:150) in tile-major, query order; boxes moved by the tile origin
 * (fp32 add of (x0, y0)).  tile_n_ws int32 [T]; outputs boxes fp32 [T*Q,4], scores fp32 [T*Q], labels int64 [T*Q],
 * src int32 [T*Q,2] = (tile, row), total int32 [1].  Feed the first `total` entries to wm_nms (labels => per class). */
int wm_merge_detections(const float* packed, const int32_t* counts, const int32_t* origins, int T, int Q, float score_thr,
                        int32_t* tile_n_ws, float* boxes, float* scores, int64_t* labels, int32_t* src, int32_t* total,
                        void* stream);

/* COCO detection records for the kept detections (inference.py:149-171, convert_to_xywh :235-237):
 * out_xywh_score fp32 [n_keep,5] = (x, y, x1 - x0, y1 - y0, score), out_category int64 [n_keep]; keep int64 [n_keep]
 * indexes boxes/scores/labels (NULL = identity). */
int wm_pack_coco(const float* boxes, const float* scores, const int64_t* labels, const int64_t* keep, int n_keep,
                 float* out_xywh_score, int64_t* out_category, void* stream);

/* ---- evaluation-time criterion: inference.py:29-89 evaluate() calls criterion(outputs, targets) ---- */

/* Matching cost of HungarianMatcher.forward (modeling/matcher.py:58-73) for ALL targets of the batch at once:
 * cost[r, t] = w_bbox * L1(boxes[r], tgt_boxes[t]) + w_class * (-softmax(logits[r])[tgt_ids[t]])
 *              + w_giou * (-GIoU(xyxy(boxes[r]), xyxy(tgt_boxes[t]))),  r = b * Q + q in [0, rows), t in [0, T).
 * logits fp32 [rows, C1], boxes fp32 [rows, 4] (cxcywh), tgt_ids int64 [T], tgt_boxes fp32 [T, 4], cost fp32 [rows, T].
 * The per-image linear_sum_assignment on the diagonal blocks stays on the host (scipy), as in the reference. */
int wm_match_cost(const float* logits, const float* boxes, const int64_t* tgt_ids, const float* tgt_boxes, int rows, int T,
                  int C1, float w_class, float w_bbox, float w_giou, float* cost, void* stream);

/* SetCriterion losses (build_sam.py:95-150) given the matching: m_row int32 [n_match] = b * Q + q of every matched query,
 * m_label int64 / m_box fp32 [n_match(,4)] its target, tgt_len int32 [B] targets per image, empty_weight fp32 [C1]
 * (1, ..., 1, eos_coef), num_boxes the normaliser (clamped, all-reduced by the caller).  tcls_ws int32 [B*Q] scratch.
 * out5 fp32 [5] = loss_ce, class_error, cardinality_error, loss_bbox, loss_giou.  Forward values only. */
int wm_set_criterion(const float* logits, const float* boxes, const int32_t* m_row, const int64_t* m_label, const float* m_box,
                     int n_match, const int32_t* tgt_len, const float* empty_weight, int B, int Q, int C1, float num_boxes,
                     int32_t* tcls_ws, float* out5, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WM_B200_H */
