// Bandwidth-bound kernels of the hot path (sm_100a): 128-bit coalesced global access, warp-shuffle reductions.
//   layernorm        nn.LayerNorm sites (image_encoder.py:173,183,476-477; transformer.py norms) and LayerNorm2d
//                    on NHWC rows (common.py:31-43) -- fp32 statistics, optional fused "+pos-embedding" copy
//   patchify         NCHW fp32 tile -> bf16 im2col rows for the 16x16/s16 patch-embed GEMM (image_encoder.py:409-417)
//                    + the grayscale plane of MedSAM.fft (network.py:41)
//   transpose        batched 2-D transpose (HFC low-pass second pass, the :512 raw-reshape operand, NHWC->NCHW)
//   hfc_finalize     |gray - lowpass| (network.py:53-55) written directly as hfc_embed im2col rows (:442-450)
//   add_cast         bf16(a + b[row % mod])   (decoder "+ positional encoding" operands, transformer.py:88-101,164-177)
//   attn_small       decoder attention, Tq x Tk x {16,32} (transformer.py:218-240) on CUDA cores (3.5 GFLOP/tile)
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

// ------------------------------------------------------------------ LayerNorm (one warp per row)
template <int VEC_PER_LANE>  // D = 128 * VEC_PER_LANE
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ y_bf16,
                                                        float* __restrict__ y_f32, const float* __restrict__ add,
                                                        int add_mod, __nv_bfloat16* __restrict__ y2_bf16, int rows,
                                                        float eps) {
  constexpr int D = 128 * VEC_PER_LANE;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  float4 v[VEC_PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    v[i] = __ldg(xr + lane + 32 * i);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const int col = (lane + 32 * i) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + bt.x;
    o.y = (v[i].y - mean) * rstd * g.y + bt.y;
    o.z = (v[i].z - mean) * rstd * g.z + bt.z;
    o.w = (v[i].w - mean) * rstd * g.w + bt.w;
    if (y_f32) *reinterpret_cast<float4*>(y_f32 + (size_t)row * D + col) = o;
    if (y_bf16)
      *reinterpret_cast<uint2*>(y_bf16 + (size_t)row * D + col) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    if (y2_bf16) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(add + (size_t)(row % add_mod) * D + col));
      *reinterpret_cast<uint2*>(y2_bf16 + (size_t)row * D + col) =
          make_uint2(pack_bf16(o.x + a.x, o.y + a.y), pack_bf16(o.z + a.z, o.w + a.w));
    }
  }
}

int layernorm_launch(const float* x, const float* gamma, const float* beta, __nv_bfloat16* y_bf16, float* y_f32,
                     const float* add, int add_mod, __nv_bfloat16* y2_bf16, int rows, int D, float eps,
                     cudaStream_t st) {
  if (rows <= 0) return WM_OK;
  const int grid = (rows + 7) / 8;
#define WM_LN_CASE(V)                                                                                              \
  case 128 * V:                                                                                                    \
    layernorm_kernel<V><<<grid, 256, 0, st>>>(x, gamma, beta, y_bf16, y_f32, add, add_mod, y2_bf16, rows, eps); \
    break;
  switch (D) {
    WM_LN_CASE(1)
    WM_LN_CASE(2)
    WM_LN_CASE(5)
    WM_LN_CASE(6)
    WM_LN_CASE(8)
    WM_LN_CASE(10)
    default:
      return WM_ERR_SHAPE;
  }
#undef WM_LN_CASE
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------ patchify (+ grayscale)
// One block per (image, patch row py): reads C x 16 x 1024 fp32 (coalesced rows), writes 64 patches x C*256 bf16.
template <int C>
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ patches,
                                                       __nv_bfloat16* __restrict__ gray) {
  const int b = blockIdx.y, py = blockIdx.x;
  const float* base = img + (size_t)b * C * 1024 * 1024 + (size_t)py * 16 * 1024;
  // each thread handles 4 consecutive pixels (x4 = 0..255) of each of the 16 rows
  const int x4 = threadIdx.x;  // 256 threads x 4 px = 1024
  const int px = x4 >> 2, kx = (x4 & 3) * 4;
  for (int ky = 0; ky < 16; ++ky) {
    float4 c[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch)
      c[ch] = __ldg(reinterpret_cast<const float4*>(base + (size_t)ch * 1024 * 1024 + ky * 1024) + x4);
    __nv_bfloat16* prow = patches + ((size_t)(b * 64 + py) * 64 + px) * (C * 256) + ky * 16 + kx;
#pragma unroll
    for (int ch = 0; ch < C; ++ch)
      *reinterpret_cast<uint2*>(prow + ch * 256) = make_uint2(pack_bf16(c[ch].x, c[ch].y), pack_bf16(c[ch].z, c[ch].w));
    if (C == 3 && gray) {
      float4 g;
      g.x = 0.2989f * c[0].x + 0.587f * c[C > 1 ? 1 : 0].x + 0.114f * c[C > 2 ? 2 : 0].x;
      g.y = 0.2989f * c[0].y + 0.587f * c[C > 1 ? 1 : 0].y + 0.114f * c[C > 2 ? 2 : 0].y;
      g.z = 0.2989f * c[0].z + 0.587f * c[C > 1 ? 1 : 0].z + 0.114f * c[C > 2 ? 2 : 0].z;
      g.w = 0.2989f * c[0].w + 0.587f * c[C > 1 ? 1 : 0].w + 0.114f * c[C > 2 ? 2 : 0].w;
      *reinterpret_cast<uint2*>(gray + ((size_t)b * 1024 + py * 16 + ky) * 1024 + x4 * 4) =
          make_uint2(pack_bf16(g.x, g.y), pack_bf16(g.z, g.w));
    }
  }
}

int patchify_launch(const float* img, __nv_bfloat16* patches, __nv_bfloat16* gray, int B, int C, cudaStream_t st) {
  if (C == 3)
    patchify_kernel<3><<<dim3(64, B), 256, 0, st>>>(img, patches, gray);
  else if (C == 1)
    patchify_kernel<1><<<dim3(64, B), 256, 0, st>>>(img, patches, nullptr);
  else
    return WM_ERR_SHAPE;
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------ batched transpose [R,C] -> [C,R]
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int C) {
  __shared__ T tile[32][33];
  const size_t boff = (size_t)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    if (r < R && c < C) tile[ty + 8 * i][tx] = in[boff + (size_t)r * C + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (r < R && c < C) out[boff + (size_t)c * R + r] = tile[tx][ty + 8 * i];
  }
}

// 2-byte elements, R and C even: 64 x 64 tiles moved as 4-byte words on both sides (a 32 x 32 tile of 2-byte elements reads
// and writes 64 bytes per warp access: half of a 128-byte line).  Word (r, c/2) of the input holds columns c, c+1 of row r;
// word (c, r/2) of the output holds rows r, r+1 of column c.
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int R, int C) {
  __shared__ uint16_t tile[64][66];  // [row][col], row stride 33 words: conflict-free on both passes
  const size_t boff = (size_t)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 words x 8 rows
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    if (r < R && c < C) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(in + boff + (size_t)r * C + c);
      *reinterpret_cast<uint32_t*>(&tile[ty + 8 * i][2 * tx]) = w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + 2 * tx;
    if (r < R && c < C) {
      const uint32_t w = (uint32_t)tile[2 * tx][ty + 8 * i] | ((uint32_t)tile[2 * tx + 1][ty + 8 * i] << 16);
      *reinterpret_cast<uint32_t*>(out + boff + (size_t)c * R + r) = w;
    }
  }
}

int transpose_launch(const void* in, void* out, int batch, int R, int C, int elt_bytes, cudaStream_t st) {
  dim3 grid((C + 31) / 32, (R + 31) / 32, batch);
  if (elt_bytes == 2 && R % 2 == 0 && C % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 3) == 0)
    transpose16_kernel<<<dim3((C + 63) / 64, (R + 63) / 64, batch), 256, 0, st>>>((const uint16_t*)in, (uint16_t*)out, R, C);
  else if (elt_bytes == 2)
    transpose_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)in, (uint16_t*)out, R, C);
  else if (elt_bytes == 4)
    transpose_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)in, (uint32_t*)out, R, C);
  else
    return WM_ERR_SHAPE;
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------ HFC finalize
// low_t[b][x][y] = lowpass transposed (fp32).  x_hfc[y][x] = | gray(img)[y][x] - low_t[x][y] |, written as the
// hfc_embed im2col row: patch (y/16, x/16), k = (y%16)*16 + x%16.  Optional fp32 image output for tests.
__global__ void __launch_bounds__(256) hfc_finalize_kernel(const float* __restrict__ img, const float* __restrict__ low_t,
                                                           __nv_bfloat16* __restrict__ patches, float* __restrict__ hfc_img) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* lt = low_t + (size_t)b * 1024 * 1024;
#pragma unroll
  for (int i = 0; i < 4; ++i) tile[ty + 8 * i][tx] = lt[(size_t)(x0 + ty + 8 * i) * 1024 + y0 + tx];  // [x][y]
  __syncthreads();
  const float* ib = img + (size_t)b * 3 * 1024 * 1024;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int y = y0 + ty + 8 * i, x = x0 + tx;
    const size_t pix = (size_t)y * 1024 + x;
    const float g = 0.2989f * ib[pix] + 0.587f * ib[pix + 1024 * 1024] + 0.114f * ib[pix + 2 * 1024 * 1024];
    const float v = fabsf(g - tile[tx][ty + 8 * i]);
    patches[((size_t)(b * 64 + (y >> 4)) * 64 + (x >> 4)) * 256 + (y & 15) * 16 + (x & 15)] = __float2bfloat16_rn(v);
    if (hfc_img) hfc_img[(size_t)b * 1024 * 1024 + pix] = v;
  }
}

int hfc_finalize_launch(const float* img, const float* low_t, __nv_bfloat16* patches, float* hfc_img, int B,
                        cudaStream_t st) {
  hfc_finalize_kernel<<<dim3(32, 32, B), 256, 0, st>>>(img, low_t, patches, hfc_img);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------ add + cast
__global__ void __launch_bounds__(256) add_cast_kernel(const float* __restrict__ a, const float* __restrict__ b, int b_mod,
                                                       __nv_bfloat16* __restrict__ out, int rows, int D4) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)rows * D4) return;
  const int row = (int)(i / D4), c4 = (int)(i % D4);
  float4 v = a ? __ldg(reinterpret_cast<const float4*>(a) + i) : make_float4(0.f, 0.f, 0.f, 0.f);  // a == null: broadcast-cast of b
  if (b) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(b) + (size_t)(row % b_mod) * D4 + c4);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}

int add_cast_launch(const float* a, const float* b, int b_mod, __nv_bfloat16* out, int rows, int D, cudaStream_t st) {
  if (D % 4 != 0) return WM_ERR_SHAPE;
  const size_t n = (size_t)rows * (D / 4);
  if (n == 0) return WM_OK;
  add_cast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b, b_mod, out, rows, D / 4);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// ------------------------------------------------------------------ decoder attention (small head dims)
// One THREAD per query row: q and the output accumulator live in registers, key/value tiles of 128 keys are staged
// in smem as fp32 and read by broadcast (all threads of a warp read the same key), online softmax in the exp2
// domain over chunks of 8 keys.  SPLIT = 4 partitions every key tile across 4 thread groups (for few queries,
// e.g. 51 tokens -> image) and merges the partial (m, l, acc) through smem.
template <int HD, int SPLIT>
__global__ void __launch_bounds__(256) attn_small_kernel(const __nv_bfloat16* __restrict__ q, int ldq,
                                                         const __nv_bfloat16* __restrict__ k, int ldk,
                                                         const __nv_bfloat16* __restrict__ v, int ldv,
                                                         __nv_bfloat16* __restrict__ out, int ldo, int Tq, int Tk,
                                                         float scale_log2e) {
  constexpr int QB = 256 / SPLIT;    // queries per block
  constexpr int KCH = 128 / SPLIT;   // keys of each tile handled by one group
  __shared__ __align__(16) float sK[128][HD];
  __shared__ __align__(16) float sV[128][HD];
  const int h = blockIdx.y, b = blockIdx.z;
  const int grp = threadIdx.x / QB, ql = threadIdx.x % QB;
  const int qi = blockIdx.x * QB + ql;
  const bool q_ok = qi < Tq;
  float qr[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) { qr[d] = 0.f; acc[d] = 0.f; }
  if (q_ok) {
    const uint4* qp = reinterpret_cast<const uint4*>(q + (size_t)(b * Tq + qi) * ldq + h * HD);
#pragma unroll
    for (int d8 = 0; d8 < HD / 8; ++d8) {
      const uint4 u = __ldg(qp + d8);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h2[e]);
        qr[d8 * 8 + 2 * e] = f.x * scale_log2e;
        qr[d8 * 8 + 2 * e + 1] = f.y * scale_log2e;
      }
    }
  }
  float m_run = -INFINITY, l_run = 0.f;
  for (int k0 = 0; k0 < Tk; k0 += 128) {
    __syncthreads();
    // stage 128 keys x HD of K and V (16-byte global loads, zero fill past Tk)
    for (int i = threadIdx.x; i < 128 * (HD / 8); i += 256) {
      const int kk = i / (HD / 8), d8 = i % (HD / 8);
      uint4 uk = make_uint4(0, 0, 0, 0), uv = make_uint4(0, 0, 0, 0);
      if (k0 + kk < Tk) {
        uk = __ldg(reinterpret_cast<const uint4*>(k + (size_t)(b * Tk + k0 + kk) * ldk + h * HD) + d8);
        uv = __ldg(reinterpret_cast<const uint4*>(v + (size_t)(b * Tk + k0 + kk) * ldv + h * HD) + d8);
      }
      const __nv_bfloat162* hk = reinterpret_cast<const __nv_bfloat162*>(&uk);
      const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&uv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fk = __bfloat1622float2(hk[e]), fv = __bfloat1622float2(hv[e]);
        sK[kk][d8 * 8 + 2 * e] = fk.x; sK[kk][d8 * 8 + 2 * e + 1] = fk.y;
        sV[kk][d8 * 8 + 2 * e] = fv.x; sV[kk][d8 * 8 + 2 * e + 1] = fv.y;
      }
    }
    __syncthreads();
    const int kbeg = grp * KCH;
    const int kend = min(kbeg + KCH, Tk - k0);  // may be <= kbeg
    for (int c0 = kbeg; c0 < kend; c0 += 8) {
      float sc[8];
      float mx = m_run;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float a = 0.f;
        const float4* kr = reinterpret_cast<const float4*>(&sK[c0 + u][0]);  // c0+u < 128 always
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
          const float4 kv4 = kr[d4];
          a = fmaf(qr[4 * d4], kv4.x, a); a = fmaf(qr[4 * d4 + 1], kv4.y, a);
          a = fmaf(qr[4 * d4 + 2], kv4.z, a); a = fmaf(qr[4 * d4 + 3], kv4.w, a);
        }
        sc[u] = (c0 + u < kend) ? a : -INFINITY;
        mx = fmaxf(mx, sc[u]);
      }
      const float alpha = ex2_approx(m_run - mx);  // 0 on the first chunk
      m_run = mx;
      l_run *= alpha;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] *= alpha;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float pe = ex2_approx(sc[u] - mx);  // 0 for masked keys
        l_run += pe;
        const float4* vr = reinterpret_cast<const float4*>(&sV[c0 + u][0]);
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
          const float4 vv = vr[d4];
          acc[4 * d4] = fmaf(pe, vv.x, acc[4 * d4]); acc[4 * d4 + 1] = fmaf(pe, vv.y, acc[4 * d4 + 1]);
          acc[4 * d4 + 2] = fmaf(pe, vv.z, acc[4 * d4 + 2]); acc[4 * d4 + 3] = fmaf(pe, vv.w, acc[4 * d4 + 3]);
        }
      }
    }
  }
  if (SPLIT > 1) {
    // merge the SPLIT partial softmax states of each query through smem (reuse the K/V staging buffers)
    __syncthreads();
    float* sm = &sK[0][0];               // [SPLIT][QB] m
    float* sl = sm + 256;                // [SPLIT][QB] l
    float* sa = &sV[0][0];               // [SPLIT][QB][HD] acc  (256*HD floats <= 128*HD*... guarded below)
    static_assert(SPLIT == 1 || 256 * HD <= 128 * HD * 2, "merge scratch");
    sm[grp * QB + ql] = m_run;
    sl[grp * QB + ql] = l_run;
    __syncthreads();
    float m_all = -INFINITY;
#pragma unroll
    for (int g = 0; g < SPLIT; ++g) m_all = fmaxf(m_all, sm[g * QB + ql]);
    const float w = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_all);
    // two rounds: acc of groups {0,1} then {2,3} would not fit at once for HD=32; accumulate via atomics-free tree
    for (int g = 0; g < SPLIT; ++g) {
      __syncthreads();
      if (grp == g) {
#pragma unroll
        for (int d = 0; d < HD; ++d) {
          float* cell = sa + (size_t)ql * HD + d;
          *cell = (g == 0 ? 0.f : *cell) + acc[d] * w;
        }
        float* lc = sl + SPLIT * QB + ql;  // merged l
        *lc = (g == 0 ? 0.f : *lc) + l_run * w;
      }
    }
    __syncthreads();
    if (grp == 0 && q_ok) {
      const float inv = 1.0f / sl[SPLIT * QB + ql];
      __nv_bfloat16* op = out + (size_t)(b * Tq + qi) * ldo + h * HD;
#pragma unroll
      for (int d = 0; d < HD; d += 2)
        *reinterpret_cast<uint32_t*>(op + d) = pack_bf16(sa[(size_t)ql * HD + d] * inv, sa[(size_t)ql * HD + d + 1] * inv);
    }
  } else if (q_ok) {
    const float inv = 1.0f / l_run;
    __nv_bfloat16* op = out + (size_t)(b * Tq + qi) * ldo + h * HD;
#pragma unroll
    for (int d8 = 0; d8 < HD / 8; ++d8)
      reinterpret_cast<uint4*>(op)[d8] =
          make_uint4(pack_bf16(acc[8 * d8] * inv, acc[8 * d8 + 1] * inv), pack_bf16(acc[8 * d8 + 2] * inv, acc[8 * d8 + 3] * inv),
                     pack_bf16(acc[8 * d8 + 4] * inv, acc[8 * d8 + 5] * inv), pack_bf16(acc[8 * d8 + 6] * inv, acc[8 * d8 + 7] * inv));
  }
}

int attn_small_launch(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                      __nv_bfloat16* out, int ldo, int B, int H, int Tq, int Tk, int hd, float scale, cudaStream_t st) {
  if (B * H * Tq == 0) return WM_OK;
  if ((ldq | ldk | ldv | ldo) % 8 != 0) return WM_ERR_SHAPE;
  const float sl2 = scale * 1.4426950408889634f;
  const bool split = Tq <= 1024;
  dim3 grid((Tq + (split ? 63 : 255)) / (split ? 64 : 256), H, B);
#define WM_AS(HD_, SP_) attn_small_kernel<HD_, SP_><<<grid, 256, 0, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, Tq, Tk, sl2)
  if (hd == 16) { if (split) WM_AS(16, 4); else WM_AS(16, 1); }
  else if (hd == 32) { if (split) WM_AS(32, 4); else WM_AS(32, 1); }
  else return WM_ERR_SHAPE;
#undef WM_AS
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

}  // namespace wm
