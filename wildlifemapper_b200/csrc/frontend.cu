// Either side of the tile-detection path: the uint8 tile front-end and the
// image-level merge / COCO packing of the packed detections.  All three are bandwidth- or latency-bound byte / index
// work; integer results and the fp32 arithmetic are bit-exact restatements of the reference's host code.
//
//   tiles_from_u8_kernel     torchvision `to_tensor` + `normalize` (dataloader_coco.py:286-292 through
//                            utils/augmentation.py:229-249) + the zero padding to 1024 x 1024 of
//                            `nested_tensor_from_tensor_list` (utils/misc.py:46-67), for tiles cut out of one
//                            uint8 HWC survey image at given origins.
//   merge_count/compact      per-tile packed rows -> one image-level candidate list (score filter of
//                            visualize_prediction.py:150, boxes moved by the tile origin), tile-major, query order
//                            preserved (a stable compaction), ready for the per-class wm_nms.
//   resize_u8_pass_kernel    the PIL bilinear (antialiased) resize behind `RandomResize([768], max_size=768)`
//                            (dataloader_coco.py:275-292 -> utils/augmentation.py:77-107 -> torchvision F.resize ->
//                            PIL Image.resize(BILINEAR)) in PIL's own 8-bit fixed-point arithmetic: 22-bit integer
//                            coefficients, horizontal pass into a uint8 intermediate, then the vertical pass -- bit-exact.
//   coco_pack_kernel         `convert_to_xywh` + the per-detection record of `prepare_for_coco_detection`
//                            (inference.py:149-171, 235-237) for the kept detections.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wm_b200.h"

namespace wm {

constexpr int TILE = 1024;

// One CTA per (tile row, tile): 256 threads x 4 pixels.  The 3 x 256 possible results per channel are tabulated once
// per CTA with IEEE divisions ((u / 255 - mean) / std, the exact operation order of to_tensor + normalize), so the
// per-pixel work is three shared-memory lookups; 12 bytes read and 48 bytes written per thread, float4 stores.
__global__ void __launch_bounds__(256) tiles_from_u8_kernel(const uint8_t* __restrict__ img, int H, int W, long long row_stride,
                                                            const int* __restrict__ origins, int content_h, int content_w,
                                                            float m0, float m1, float m2, float s0, float s1, float s2,
                                                            float* __restrict__ out) {
  __shared__ float lut[3][256];
  {
    const float u = __fdiv_rn((float)threadIdx.x, 255.0f);
    lut[0][threadIdx.x] = __fdiv_rn(__fsub_rn(u, m0), s0);
    lut[1][threadIdx.x] = __fdiv_rn(__fsub_rn(u, m1), s1);
    lut[2][threadIdx.x] = __fdiv_rn(__fsub_rn(u, m2), s2);
  }
  __syncthreads();
  const int t = blockIdx.y, y = blockIdx.x, x = threadIdx.x * 4;
  const int oy = origins[2 * t], ox = origins[2 * t + 1];
  const int sy = oy + y;
  const bool row_in = y < content_h && sy >= 0 && sy < H;
  float r[3][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int sx = ox + x + i;
    const bool in = row_in && (x + i) < content_w && sx >= 0 && sx < W;
    if (in) {
      const uint8_t* p = img + (long long)sy * row_stride + (long long)sx * 3;
      r[0][i] = lut[0][__ldg(p)];
      r[1][i] = lut[1][__ldg(p + 1)];
      r[2][i] = lut[2][__ldg(p + 2)];
    } else {
      r[0][i] = r[1][i] = r[2][i] = 0.0f;  // padding is zero AFTER normalisation (misc.py:55)
    }
  }
  float* o = out + ((size_t)t * 3 * TILE + y) * TILE + x;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    *reinterpret_cast<float4*>(o + (size_t)c * TILE * TILE) = make_float4(r[c][0], r[c][1], r[c][2], r[c][3]);
}

int tiles_from_u8_launch(const uint8_t* img, int H, int W, long long row_stride, const int* origins, int T, int content_h,
                         int content_w, const float* mean, const float* stdv, float* out, cudaStream_t st) {
  if (T == 0) return WM_OK;
  tiles_from_u8_kernel<<<dim3(TILE, T), 256, 0, st>>>(img, H, W, row_stride, origins, content_h, content_w, mean[0], mean[1],
                                                      mean[2], stdv[0], stdv[1], stdv[2], out);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// merge: rows of tile t that pass `score > thr` (fp32 compare) go to [offset[t], offset[t] + n_t) in query order.
__device__ __forceinline__ bool merge_pass(const float* packed, const int* counts, int t, int q, int Q, float thr) {
  return q < Q && q < counts[t] && packed[((size_t)t * Q + q) * 6 + 4] > thr;
}

__global__ void __launch_bounds__(1024) merge_count_kernel(const float* __restrict__ packed, const int* __restrict__ counts,
                                                          int Q, float thr, int* __restrict__ tile_n) {
  const int t = blockIdx.x;
  const int n = __syncthreads_count(merge_pass(packed, counts, t, threadIdx.x, Q, thr));
  if (threadIdx.x == 0) tile_n[t] = n;
}

__global__ void __launch_bounds__(1024) merge_compact_kernel(const float* __restrict__ packed, const int* __restrict__ counts,
                                                            const int* __restrict__ origins, const int* __restrict__ tile_n,
                                                            int T, int Q, float thr, float* __restrict__ boxes,
                                                            float* __restrict__ scores, long long* __restrict__ labels,
                                                            int* __restrict__ src, int* __restrict__ total) {
  __shared__ int warp_sum[32];
  __shared__ int base_s;
  const int t = blockIdx.x, q = threadIdx.x, lane = q & 31, warp = q >> 5;
  // offset of this tile = number of candidates in the tiles before it (T is a few hundred at most)
  int part = 0;
  for (int i = q; i < T; i += blockDim.x) part += (i < t) ? tile_n[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) warp_sum[warp] = part;
  __syncthreads();
  if (warp == 0) {
    int v = warp_sum[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) base_s = v;
  }
  __syncthreads();
  const int base = base_s;
  if (t == T - 1 && q == 0) *total = base + tile_n[t];
  // stable in-tile position: ballot prefix inside the warp + prefix over the warps
  const bool pass = merge_pass(packed, counts, t, q, Q, thr);
  const unsigned bal = __ballot_sync(0xffffffffu, pass);
  __syncthreads();
  if (lane == 0) warp_sum[warp] = __popc(bal);
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += warp_sum[w];
  if (pass) {
    const int dst = base + before + __popc(bal & ((1u << lane) - 1u));
    const float* r = packed + ((size_t)t * Q + q) * 6;
    const float oy = (float)origins[2 * t], ox = (float)origins[2 * t + 1];
    boxes[4 * (size_t)dst + 0] = __fadd_rn(r[0], ox);
    boxes[4 * (size_t)dst + 1] = __fadd_rn(r[1], oy);
    boxes[4 * (size_t)dst + 2] = __fadd_rn(r[2], ox);
    boxes[4 * (size_t)dst + 3] = __fadd_rn(r[3], oy);
    scores[dst] = r[4];
    labels[dst] = (long long)r[5];
    src[2 * (size_t)dst] = t;
    src[2 * (size_t)dst + 1] = q;
  }
}

int merge_detections_launch(const float* packed, const int* counts, const int* origins, int T, int Q, float thr, int* tile_n_ws,
                            float* boxes, float* scores, long long* labels, int* src, int* total, cudaStream_t st) {
  if (Q > 1024) return WM_ERR_SHAPE;
  if (T == 0) return cudaMemsetAsync(total, 0, sizeof(int), st) == cudaSuccess ? WM_OK : WM_ERR_CUDA;
  merge_count_kernel<<<T, 1024, 0, st>>>(packed, counts, Q, thr, tile_n_ws);
  merge_compact_kernel<<<T, 1024, 0, st>>>(packed, counts, origins, tile_n_ws, T, Q, thr, boxes, scores, labels, src, total);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void coco_pack_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                 const long long* __restrict__ labels, const long long* __restrict__ keep, int n_keep,
                                 float* __restrict__ out_xywh_score, long long* __restrict__ out_cat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keep) return;
  const long long k = keep != nullptr ? keep[i] : i;
  const float x0 = boxes[4 * k], y0 = boxes[4 * k + 1], x1 = boxes[4 * k + 2], y1 = boxes[4 * k + 3];
  float* o = out_xywh_score + 5 * (size_t)i;
  o[0] = x0;
  o[1] = y0;
  o[2] = __fsub_rn(x1, x0);
  o[3] = __fsub_rn(y1, y0);
  o[4] = scores[k];
  out_cat[i] = labels[k];
}

int coco_pack_launch(const float* boxes, const float* scores, const long long* labels, const long long* keep, int n_keep,
                     float* out_xywh_score, long long* out_cat, cudaStream_t st) {
  if (n_keep == 0) return WM_OK;
  coco_pack_kernel<<<(n_keep + 255) / 256, 256, 0, st>>>(boxes, scores, labels, keep, n_keep, out_xywh_score, out_cat);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}


// ------------------------------------------------------------------ PIL-exact bilinear resize of uint8 RGB tiles
// One pass of ImagingResample for 8-bit pixels (Pillow src/libImaging/Resample.c, ImagingResampleHorizontal_8bpc /
// ImagingResampleVertical_8bpc): out = clip8((2^21 + sum_x in[xmin + x] * k[x]) >> 22), k = the precomputed 22-bit
// coefficients of the output position (computed on the host in double precision exactly as precompute_coeffs /
// normalize_coeffs_8bpc do).  Thread = one output pixel (3 channels); blockIdx.z = tile.
//   horizontal: in = tile window of the survey image (origin from `origins`), out[t][y][xo]
//   vertical:   in = the horizontal pass's output [t][y][xo], out[t][yo][xo]
__global__ void __launch_bounds__(256) resize_u8_pass_kernel(const uint8_t* __restrict__ in, long long in_row_stride,
                                                             long long in_tile_stride, const int* __restrict__ origins,
                                                             uint8_t* __restrict__ out, int out_w, int out_h,
                                                             const int* __restrict__ bounds, const int* __restrict__ kk,
                                                             int ksize, int vertical) {
  const int xo = blockIdx.x * 256 + threadIdx.x, yo = blockIdx.y, t = blockIdx.z;
  if (xo >= out_w) return;
  const uint8_t* base = in + (long long)t * in_tile_stride;
  if (origins) base += (long long)origins[2 * t] * in_row_stride + (long long)origins[2 * t + 1] * 3;
  const int o = vertical ? yo : xo;
  const int lo = bounds[2 * o], cnt = bounds[2 * o + 1];
  const int* k = kk + (size_t)o * ksize;
  int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
  for (int i = 0; i < cnt; ++i) {
    const uint8_t* px = vertical ? base + (long long)(lo + i) * in_row_stride + (long long)xo * 3
                                 : base + (long long)yo * in_row_stride + (long long)(lo + i) * 3;
    const int w = __ldg(k + i);
    s0 += (int)__ldg(px) * w;
    s1 += (int)__ldg(px + 1) * w;
    s2 += (int)__ldg(px + 2) * w;
  }
  uint8_t* q = out + (((size_t)t * out_h + yo) * out_w + xo) * 3;
  q[0] = (uint8_t)min(max(s0 >> 22, 0), 255);
  q[1] = (uint8_t)min(max(s1 >> 22, 0), 255);
  q[2] = (uint8_t)min(max(s2 >> 22, 0), 255);
}

int resize_u8_launch(const uint8_t* img, long long row_stride, const int* origins, int T, int H, int W, uint8_t* tmp,
                     uint8_t* out, int oh, int ow, const int* xbounds, const int* xk, int xks, const int* ybounds,
                     const int* yk, int yks, cudaStream_t st) {
  if (T == 0) return WM_OK;
  (void)W;
  // horizontal: [T, H, ow] from the image windows;  vertical: [T, oh, ow] from the intermediate
  resize_u8_pass_kernel<<<dim3((ow + 255) / 256, H, T), 256, 0, st>>>(img, row_stride, 0, origins, tmp, ow, H, xbounds, xk, xks, 0);
  resize_u8_pass_kernel<<<dim3((ow + 255) / 256, oh, T), 256, 0, st>>>(tmp, (long long)ow * 3, (long long)H * ow * 3, nullptr, out, ow,
                                                                       oh, ybounds, yk, yks, 1);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

}  // namespace wm
