// Post-process kernels (sm_100a), bit-exact integer stages:
//   postprocess     PostProcess.forward, build_sam.py:219-258 (+ box_cxcywh_to_xyxy, utils/box_ops.py:9-13):
//                   softmax over 8 logits -> max over the 7 real classes (first index wins) -> keep score > fp32(thr)
//                   -> xyxy -> * [s0,s1,s0,s1]; compacted in query order into packed [B,Q,6] rows + counts
//   sigmoid_topk    north-star extension: sigmoid over the 7 class logits, stable top-K over Q*7 (ties: lower index)
//   nms             torchvision.ops.nms semantics as called at visualize_prediction.py:150-154 (+ per-class variant):
//                   stable score-descending order, fp32 IoU with IEEE ops and no FMA contraction,
//                   suppress iff (double)iou > (double)thr; 64x64 bitmask tiles + block-serial reduction.
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

// ------------------------------------------------------------------ PostProcess
// mode 0: logits -> softmax; mode 1: input already holds the class probabilities (integer-stage parity test)
__global__ void __launch_bounds__(1024) postprocess_kernel(const float* __restrict__ in, const float* __restrict__ boxes,
                                                           const long long* __restrict__ sizes, float thr, int mode,
                                                           float* __restrict__ packed, int* __restrict__ query_idx,
                                                           long long* __restrict__ labels_out, int* __restrict__ counts,
                                                           int Q, int C1) {
  __shared__ int warp_cnt[32];
  __shared__ int base;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) base = 0;
  const float s0 = (float)sizes[2 * b + 0], s1 = (float)sizes[2 * b + 1];  // img_w = sizes[0], img_h = sizes[1]
  __syncthreads();
  for (int q0 = 0; q0 < Q; q0 += 1024) {
    const int q = q0 + threadIdx.x;
    bool keep = false;
    float score = 0.f;
    int label = 0;
    if (q < Q) {
      const float* x = in + ((size_t)b * Q + q) * C1;
      float pr[16];
      if (mode == 0) {
        float mx = x[0];
        for (int c = 1; c < C1; ++c) mx = fmaxf(mx, x[c]);
        float sum = 0.f;
        for (int c = 0; c < C1; ++c) {
          pr[c] = expf(x[c] - mx);
          sum = __fadd_rn(sum, pr[c]);
        }
        for (int c = 0; c < C1; ++c) pr[c] = __fdiv_rn(pr[c], sum);
      } else {
        for (int c = 0; c < C1; ++c) pr[c] = x[c];
      }
      score = pr[0];
      for (int c = 1; c < C1 - 1; ++c)
        if (pr[c] > score) { score = pr[c]; label = c; }  // strict > : first maximum wins
      keep = score > thr;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    off += __popc(bal & ((1u << lane) - 1));
    if (keep) {
      const float* bx = boxes + ((size_t)b * Q + q) * 4;
      const float cx = bx[0], cy = bx[1], hw = __fmul_rn(0.5f, bx[2]), hh = __fmul_rn(0.5f, bx[3]);
      float* o = packed + ((size_t)b * Q + off) * 6;
      o[0] = __fmul_rn(__fsub_rn(cx, hw), s0);
      o[1] = __fmul_rn(__fsub_rn(cy, hh), s1);
      o[2] = __fmul_rn(__fadd_rn(cx, hw), s0);
      o[3] = __fmul_rn(__fadd_rn(cy, hh), s1);
      o[4] = score;
      o[5] = (float)label;
      query_idx[(size_t)b * Q + off] = q;
      if (labels_out) labels_out[(size_t)b * Q + off] = label;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += warp_cnt[w];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = base;
}

int postprocess_launch(const float* in, const float* boxes, const long long* sizes, float thr, int mode, float* packed,
                       int* query_idx, long long* labels_out, int* counts, int B, int Q, int C1, cudaStream_t st) {
  if (C1 < 2 || C1 > 16) return WM_ERR_SHAPE;
  if (B == 0) return WM_OK;
  postprocess_kernel<<<B, 1024, 0, st>>>(in, boxes, sizes, thr, mode, packed, query_idx, labels_out, counts, Q, C1);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------ stable descending rank (exact argsort, O(n^2))
// rank[i] = #{ j : s[j] > s[i]  or  (s[j] == s[i] and j < i) };  order[rank[i]] = i.
// The comparison runs on an integer key that orders exactly like the floats (-0 == +0) and puts every NaN first (the
// torch.sort(descending=True) convention), so the ranks are a permutation for ANY input and order[] is fully written.
__device__ __forceinline__ int rank_key(float f) {
  if (f != f) return 0x7fffffff;
  const int b = __float_as_int(f + 0.0f);  // -0 -> +0
  return b >= 0 ? b : (b ^ 0x7fffffff);
}
__global__ void __launch_bounds__(256) rank_desc_kernel(const float* __restrict__ s, int n, int* __restrict__ order,
                                                        size_t batch_stride_s, size_t batch_stride_o) {
  __shared__ int tile[1024];
  const float* sb = s + blockIdx.y * batch_stride_s;
  int* ob = order + blockIdx.y * batch_stride_o;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int si = i < n ? rank_key(sb[i]) : 0;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 1024) {
    __syncthreads();
    for (int t = threadIdx.x; t < 1024; t += 256) tile[t] = (j0 + t < n) ? rank_key(sb[j0 + t]) : (int)0x80000000;
    __syncthreads();
    const int lim = min(1024, n - j0);
    for (int t = 0; t < lim; ++t) {
      const int sj = tile[t];
      rank += (sj > si) || (sj == si && (j0 + t) < i);
    }
  }
  if (i < n) ob[rank] = i;
}

// ------------------------------------------------------------------ sigmoid + top-K
__global__ void __launch_bounds__(256) sigmoid_kernel(const float* __restrict__ logits, float* __restrict__ prob, int Q,
                                                      int C1, int C, size_t total) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const size_t bq = i / C;
  const int c = (int)(i % C);
  const float x = logits[bq * C1 + c];
  prob[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

__global__ void __launch_bounds__(256) topk_gather_kernel(const float* __restrict__ prob, const int* __restrict__ order,
                                                          const float* __restrict__ boxes, int n, int K, int C, int Q,
                                                          float* __restrict__ scores, int* __restrict__ labels,
                                                          int* __restrict__ query, float* __restrict__ out_boxes) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= K) return;
  const int idx = order[(size_t)b * n + r];
  const int q = idx / C;
  scores[(size_t)b * K + r] = prob[(size_t)b * n + idx];
  labels[(size_t)b * K + r] = idx % C;
  query[(size_t)b * K + r] = q;
  const float4 bx = *reinterpret_cast<const float4*>(boxes + ((size_t)b * Q + q) * 4);
  *reinterpret_cast<float4*>(out_boxes + ((size_t)b * K + r) * 4) = bx;
}

int sigmoid_topk_launch(const float* logits, const float* boxes, float* prob_ws, int* order_ws, float* scores,
                        int* labels, int* query, float* out_boxes, int B, int Q, int C1, int C, int K,
                        int from_prob, cudaStream_t st) {
  const int n = Q * C;
  if (K > n || K <= 0) return WM_ERR_SHAPE;
  if (B == 0) return WM_OK;
  const size_t total = (size_t)B * n;
  if (!from_prob) sigmoid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(logits, prob_ws, Q, C1, C, total);
  rank_desc_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(prob_ws, n, order_ws, n, n);
  topk_gather_kernel<<<dim3((K + 255) / 256, B), 256, 0, st>>>(prob_ws, order_ws, boxes, n, K, C, Q, scores, labels,
                                                              query, out_boxes);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------ NMS
__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float area_a, float area_b, double thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return (double)iou > thr;  // fp32 IoU against the DOUBLE threshold; NaN -> false
}

// mask[i][jb] bit t set  <=>  sorted box (jb*64+t) is suppressed by sorted box i   (only j > i matter)
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ boxes, const int* __restrict__ order,
                                                      const long long* __restrict__ labels, int n, double thr,
                                                      unsigned long long* __restrict__ mask) {
  const int row_blk = blockIdx.y, col_blk = blockIdx.x;
  if (col_blk < row_blk) return;
  __shared__ float4 cb[64];
  __shared__ float ca[64];
  __shared__ long long cl[64];
  const int nblk = (n + 63) / 64;
  const int cj = col_blk * 64 + threadIdx.x;
  if (cj < n) {
    const int oj = order[cj];
    const float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)oj * 4);
    cb[threadIdx.x] = bx;
    ca[threadIdx.x] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
    cl[threadIdx.x] = labels ? labels[oj] : 0;
  }
  __syncthreads();
  const int i = row_blk * 64 + threadIdx.x;
  if (i >= n) return;
  const int oi = order[i];
  const float4 a = *reinterpret_cast<const float4*>(boxes + (size_t)oi * 4);
  const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const long long la = labels ? labels[oi] : 0;
  const int cols = min(64, n - col_blk * 64);
  unsigned long long bits = 0;
  const int start = (row_blk == col_blk) ? threadIdx.x + 1 : 0;
  for (int t = start; t < cols; ++t)
    if (cl[t] == la && iou_gt(a, cb[t], area_a, ca[t], thr)) bits |= 1ull << t;
  mask[(size_t)i * nblk + col_blk] = bits;
}

// One block of 1024 threads walks the sorted boxes 64 at a time:
//   A  64 threads fetch the block's diagonal mask words into shared memory (one coalesced-latency round trip instead of
//      64 dependent global loads in the serial chain below: that chain was 5.4 ms of latency for 10 k boxes);
//   B  one thread resolves the in-block greedy chain from shared memory / registers and appends the survivors;
//   C  all threads OR the mask rows of the survivors into the running "removed" bitmap: thread = (16-row group, word),
//      16 independent predicated loads each, merged with a shared-memory atomicOr.
__global__ void __launch_bounds__(1024) nms_reduce_kernel(const unsigned long long* __restrict__ mask,
                                                          const int* __restrict__ order, int n,
                                                          long long* __restrict__ keep, int* __restrict__ num_keep) {
  extern __shared__ unsigned long long removed[];  // nblk words
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long alive_word;
  __shared__ int kept_total;
  const int nblk = (n + 63) / 64;
  for (int w = threadIdx.x; w < nblk; w += 1024) removed[w] = 0;
  if (threadIdx.x == 0) kept_total = 0;
  __syncthreads();
  for (int blk = 0; blk < nblk; ++blk) {
    const int cnt = min(64, n - blk * 64);
    if (threadIdx.x < 64) diag[threadIdx.x] = threadIdx.x < cnt ? mask[(size_t)(blk * 64 + threadIdx.x) * nblk + blk] : 0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long rem = removed[blk];
      unsigned long long alive = 0;
      int kt = kept_total;
      for (int t = 0; t < cnt; ++t) {
        if (!((rem >> t) & 1ull)) {
          alive |= 1ull << t;
          keep[kt++] = order[blk * 64 + t];
          rem |= diag[t];
        }
      }
      alive_word = alive;
      kept_total = kt;
    }
    __syncthreads();
    const unsigned long long alive = alive_word;
    const int nw = nblk - blk - 1;
    for (int idx = threadIdx.x; idx < 4 * nw; idx += 1024) {
      const int grp = idx / nw, w = blk + 1 + idx % nw;  // consecutive threads: consecutive words of the same rows
      const unsigned long long a16 = (alive >> (16 * grp)) & 0xffffull;
      if (a16 == 0) continue;
      const unsigned long long* mrow = mask + (size_t)(blk * 64 + 16 * grp) * nblk + w;
      unsigned long long acc = 0;
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const unsigned long long m = (16 * grp + t < cnt) ? mrow[(size_t)t * nblk] : 0ull;  // independent loads
        acc |= ((a16 >> t) & 1ull) ? m : 0ull;
      }
      if (acc) atomicOr(&removed[w], acc);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_keep = kept_total;
}

int nms_launch(const float* boxes, const float* scores, const long long* labels, int n, double thr, int* order_ws,
               unsigned long long* mask_ws, long long* keep, int* num_keep, cudaStream_t st) {
  if (n == 0) {
    cudaMemsetAsync(num_keep, 0, sizeof(int), st);
    return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
  }
  const int nblk = (n + 63) / 64;
  rank_desc_kernel<<<dim3((n + 255) / 256, 1), 256, 0, st>>>(scores, n, order_ws, 0, 0);
  nms_mask_kernel<<<dim3(nblk, nblk), 64, 0, st>>>(boxes, order_ws, labels, n, thr, mask_ws);
  nms_reduce_kernel<<<1, 1024, nblk * sizeof(unsigned long long), st>>>(mask_ws, order_ws, n, keep, num_keep);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------ batched small NMS (one block per image)
// Consumes the packed PostProcess rows directly: candidates = rows with score > score_thr (visualize_prediction.py:150),
// visited in stable score-descending order, greedy suppression with the same IoU arithmetic as above.
// keep_idx[b][0..keep_cnt[b]) = kept row indices (into the packed rows of image b) in score order.
constexpr int NMS_SMALL_MAX = 1024;
__global__ void __launch_bounds__(256) nms_batched_small_kernel(const float* __restrict__ packed,
                                                                const int* __restrict__ counts, int Q, float score_thr,
                                                                double iou_thr, int per_class,
                                                                int* __restrict__ keep_idx, int* __restrict__ keep_cnt) {
  __shared__ float4 sb[NMS_SMALL_MAX];
  __shared__ float sa[NMS_SMALL_MAX];
  __shared__ float ss[NMS_SMALL_MAX];
  __shared__ int sl[NMS_SMALL_MAX];
  __shared__ int order[NMS_SMALL_MAX];
  __shared__ unsigned char alive[NMS_SMALL_MAX];
  __shared__ int nv_s, nk_s;
  const int b = blockIdx.x;
  const int n = min(counts[b], Q);
  const float* rows = packed + (size_t)b * Q * 6;
  if (threadIdx.x == 0) { nv_s = 0; nk_s = 0; }
  for (int i = threadIdx.x; i < n; i += 256) {
    const float* r = rows + (size_t)i * 6;
    sb[i] = make_float4(r[0], r[1], r[2], r[3]);
    sa[i] = __fmul_rn(__fsub_rn(r[2], r[0]), __fsub_rn(r[3], r[1]));
    ss[i] = r[4];
    sl[i] = (int)r[5];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    const float si = ss[i];
    if (si > score_thr) {
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const float sj = ss[j];
        rank += (sj > score_thr) && ((sj > si) || (sj == si && j < i));
      }
      order[rank] = i;
      alive[rank] = 1;
      atomicAdd(&nv_s, 1);
    }
  }
  __syncthreads();
  const int nv = nv_s;
  for (int r = 0; r < nv; ++r) {
    if (alive[r]) {  // block-uniform: alive[r] is only written before the barrier that ends the previous round
      const int i = order[r];
      if (threadIdx.x == 0) keep_idx[(size_t)b * Q + nk_s++] = i;
      const float4 bi = sb[i];
      const float ai = sa[i];
      const int li = sl[i];
      for (int r2 = r + 1 + threadIdx.x; r2 < nv; r2 += 256) {
        if (alive[r2]) {
          const int j = order[r2];
          if ((!per_class || sl[j] == li) && iou_gt(bi, sb[j], ai, sa[j], iou_thr)) alive[r2] = 0;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) keep_cnt[b] = nk_s;
}

int nms_batched_small_launch(const float* packed, const int* counts, int B, int Q, float score_thr, double iou_thr,
                             int per_class, int* keep_idx, int* keep_cnt, cudaStream_t st) {
  if (Q > NMS_SMALL_MAX) return WM_ERR_SHAPE;
  if (B == 0) return WM_OK;
  nms_batched_small_kernel<<<B, 256, 0, st>>>(packed, counts, Q, score_thr, iou_thr, per_class, keep_idx, keep_cnt);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

}  // namespace wm
