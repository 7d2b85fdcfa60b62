// Evaluation-time DETR criterion of the reference: `inference.py:29-89 evaluate()` calls
// `criterion(outputs, targets)` under torch.no_grad() for every batch, so the drop-in needs HungarianMatcher.forward
// (modeling/matcher.py:34-81) and SetCriterion.forward (build_sam.py:62-210) -- forward values only (the backward is the
// training path, section 8f row 1).  All latency-bound fp32 work on a handful of KB:
//
//   match_cost_kernel   C[b*Q + q, t] = w_bbox * L1(box_q, box_t) + w_class * (-softmax(logits_q)[label_t])
//                                       + w_giou * (-GIoU(xyxy(box_q), xyxy(box_t)))          (matcher.py:58-73)
//                       for ALL targets of the batch at once, like the reference ([B*Q, sum T]); the per-image linear
//                       assignment on the diagonal blocks stays scipy's on the host (as in the reference: C.cpu()).
//   criterion_kernel    given the matching: weighted cross entropy over all B*Q queries (no-object weight eos_coef),
//                       class_error = 100 - top-1 accuracy of the matched queries, cardinality error, L1 and GIoU box
//                       losses normalised by num_boxes                                         (build_sam.py:95-150)
//
// fp32 arithmetic in the reference's operation order (IEEE ops, no FMA contraction) for the per-element terms; the sums run
// in double (the reference sums in fp32: the two agree to ~1e-6 relative, the gate of the tests).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "wm_internal.h"

namespace wm {

__device__ __forceinline__ void cxcywh_to_xyxy(const float* b, float& x0, float& y0, float& x1, float& y1) {
  const float hw = __fmul_rn(0.5f, b[2]), hh = __fmul_rn(0.5f, b[3]);
  x0 = __fsub_rn(b[0], hw); y0 = __fsub_rn(b[1], hh);
  x1 = __fadd_rn(b[0], hw); y1 = __fadd_rn(b[1], hh);
}

// generalized_box_iou of one pair (utils/box_ops.py:24-62)
__device__ __forceinline__ float giou_pair(const float* a, const float* b) {
  float ax0, ay0, ax1, ay1, bx0, by0, bx1, by1;
  cxcywh_to_xyxy(a, ax0, ay0, ax1, ay1);
  cxcywh_to_xyxy(b, bx0, by0, bx1, by1);
  const float area_a = __fmul_rn(__fsub_rn(ax1, ax0), __fsub_rn(ay1, ay0));
  const float area_b = __fmul_rn(__fsub_rn(bx1, bx0), __fsub_rn(by1, by0));
  const float iw = fmaxf(__fsub_rn(fminf(ax1, bx1), fmaxf(ax0, bx0)), 0.0f);
  const float ih = fmaxf(__fsub_rn(fminf(ay1, by1), fmaxf(ay0, by0)), 0.0f);
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  const float iou = __fdiv_rn(inter, uni);
  const float ew = fmaxf(__fsub_rn(fmaxf(ax1, bx1), fminf(ax0, bx0)), 0.0f);
  const float eh = fmaxf(__fsub_rn(fmaxf(ay1, by1), fminf(ay0, by0)), 0.0f);
  const float earea = __fmul_rn(ew, eh);
  return __fsub_rn(iou, __fdiv_rn(__fsub_rn(earea, uni), earea));
}

__global__ void __launch_bounds__(256) match_cost_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                                         const long long* __restrict__ tgt_ids,
                                                         const float* __restrict__ tgt_boxes, int rows, int T, int C1,
                                                         float w_class, float w_bbox, float w_giou,
                                                         float* __restrict__ cost) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)rows * T) return;
  const int r = (int)(i / T), t = (int)(i % T);
  const float* x = logits + (size_t)r * C1;
  float mx = x[0];
  for (int c = 1; c < C1; ++c) mx = fmaxf(mx, x[c]);
  float den = 0.0f;
  for (int c = 0; c < C1; ++c) den = __fadd_rn(den, expf(__fsub_rn(x[c], mx)));
  const int lab = (int)tgt_ids[t];
  const float prob = (lab >= 0 && lab < C1) ? __fdiv_rn(expf(__fsub_rn(x[lab], mx)), den) : 0.0f;
  const float* a = boxes + (size_t)r * 4;
  const float* b = tgt_boxes + (size_t)t * 4;
  float l1 = 0.0f;
  for (int k = 0; k < 4; ++k) l1 = __fadd_rn(l1, fabsf(__fsub_rn(a[k], b[k])));
  const float g = giou_pair(a, b);
  // C = cost_bbox * cost_bbox_ + cost_class * (-prob) + cost_giou * (-giou), in the reference's order of additions
  cost[i] = __fadd_rn(__fadd_rn(__fmul_rn(w_bbox, l1), __fmul_rn(w_class, -prob)), __fmul_rn(w_giou, -g));
}

int match_cost_launch(const float* logits, const float* boxes, const long long* tgt_ids, const float* tgt_boxes, int rows,
                      int T, int C1, float w_class, float w_bbox, float w_giou, float* cost, cudaStream_t st) {
  const long long n = (long long)rows * T;
  if (n == 0) return WM_OK;
  match_cost_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logits, boxes, tgt_ids, tgt_boxes, rows, T, C1, w_class,
                                                                  w_bbox, w_giou, cost);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += scratch[w];
  return s;
}

// One block.  tcls_ws int32 [B*Q]: target class of every query (scratch).  matched: m_row[i] = b*Q + q of matched pair i,
// m_label / m_box its target.  out[5] = loss_ce, class_error, cardinality_error, loss_bbox, loss_giou.
__global__ void __launch_bounds__(1024) criterion_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                                         const int* __restrict__ m_row, const long long* __restrict__ m_label,
                                                         const float* __restrict__ m_box, int n_match,
                                                         const int* __restrict__ tgt_len, const float* __restrict__ empty_weight,
                                                         int B, int Q, int C1, float num_boxes, int* __restrict__ tcls_ws,
                                                         float* __restrict__ out) {
  __shared__ double scratch[32];
  const int rows = B * Q, no_obj = C1 - 1;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) tcls_ws[i] = no_obj;
  __syncthreads();
  for (int i = threadIdx.x; i < n_match; i += blockDim.x) tcls_ws[m_row[i]] = (int)m_label[i];
  __syncthreads();
  // ---- weighted cross entropy (F.cross_entropy(src_logits.transpose(1, 2), target_classes, empty_weight))
  double num = 0.0, den = 0.0;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const float* x = logits + (size_t)i * C1;
    float mx = x[0];
    for (int c = 1; c < C1; ++c) mx = fmaxf(mx, x[c]);
    float s = 0.0f;
    for (int c = 0; c < C1; ++c) s = __fadd_rn(s, expf(__fsub_rn(x[c], mx)));
    const int t = tcls_ws[i];
    const float logp = __fsub_rn(__fsub_rn(x[t], mx), logf(s));  // log_softmax
    const float w = empty_weight[t];
    num += (double)__fmul_rn(w, -logp);
    den += (double)w;
  }
  num = block_sum(num, scratch);
  den = block_sum(den, scratch);
  // ---- class_error (matched queries: argmax over the real classes == target label) and the box losses
  double correct = 0.0, l1 = 0.0, gl = 0.0;
  for (int i = threadIdx.x; i < n_match; i += blockDim.x) {
    const float* x = logits + (size_t)m_row[i] * C1;
    int am = 0;
    for (int c = 1; c < no_obj; ++c)
      if (x[c] > x[am]) am = c;  // first maximum (torch.topk k = 1 on distinct values)
    correct += (am == (int)m_label[i]) ? 1.0 : 0.0;
    const float* a = boxes + (size_t)m_row[i] * 4;
    const float* b = m_box + (size_t)i * 4;
    for (int k = 0; k < 4; ++k) l1 += (double)fabsf(__fsub_rn(a[k], b[k]));
    gl += (double)__fsub_rn(1.0f, giou_pair(a, b));
  }
  correct = block_sum(correct, scratch);
  l1 = block_sum(l1, scratch);
  gl = block_sum(gl, scratch);
  // ---- cardinality error: | #(argmax != no-object) - #targets | averaged over the images
  double card = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int cnt = 0;
    for (int q = 0; q < Q; ++q) {
      const float* x = logits + ((size_t)b * Q + q) * C1;
      int am = 0;
      for (int c = 1; c < C1; ++c)
        if (x[c] > x[am]) am = c;
      cnt += am != no_obj;
    }
    card += fabs((double)cnt - (double)tgt_len[b]);
  }
  card = block_sum(card, scratch);
  if (threadIdx.x == 0) {
    out[0] = (float)(num / den);
    out[1] = n_match > 0 ? (float)(100.0 - correct * (100.0 / n_match)) : 100.0f;  // accuracy() of nothing is 0
    out[2] = (float)(card / B);
    out[3] = (float)(l1 / (double)num_boxes);
    out[4] = (float)(gl / (double)num_boxes);
  }
}

int criterion_launch(const float* logits, const float* boxes, const int* m_row, const long long* m_label, const float* m_box,
                     int n_match, const int* tgt_len, const float* empty_weight, int B, int Q, int C1, float num_boxes,
                     int* tcls_ws, float* out, cudaStream_t st) {
  criterion_kernel<<<1, 1024, 0, st>>>(logits, boxes, m_row, m_label, m_box, n_match, tgt_len, empty_weight, B, Q, C1, num_boxes,
                                      tcls_ws, out);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

}  // namespace wm
