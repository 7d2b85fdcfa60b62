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
  // Rows are walked from the LAST one back: the GEMM in front of a LayerNorm writes the residual stream front to back and the
  // GEMM behind it reads the normalised rows front to back, so the rows this kernel touches first are the ones the producer
  // left in L2 and the ones it writes last are the first the consumer asks for (the tensors are 2-3 x the 126 MB L2; walked
  // forward, every byte came from and went to HBM).
#ifdef WM_LN_FORWARD
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
#else
  const int row = ((int)gridDim.x - 1 - (int)blockIdx.x) * 8 + (threadIdx.x >> 5);
#endif
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
// The grayscale plane feeds the low-pass DFT-operator GEMMs.  SPLIT: rows of 3072 = [hi | lo | hi] with hi = bf16(g),
// lo = bf16(g - hi): against the weight [Lh | Lh | Ll] one GEMM accumulates g_hi Lh + g_lo Lh + g_hi Ll in fp32, i.e. the
// product to ~2^-17 instead of 2^-9 (x_hfc = |g - low| is a small difference of O(1) numbers on smooth imagery).
template <int C, bool SPLIT>
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
      const uint2 hi = make_uint2(pack_bf16(g.x, g.y), pack_bf16(g.z, g.w));
      if (!SPLIT) {
        *reinterpret_cast<uint2*>(gray + ((size_t)b * 1024 + py * 16 + ky) * 1024 + x4 * 4) = hi;
      } else {
        const uint2 lo = make_uint2(pack_bf16(g.x - __uint_as_float(hi.x << 16), g.y - __uint_as_float(hi.x & 0xffff0000u)),
                                    pack_bf16(g.z - __uint_as_float(hi.y << 16), g.w - __uint_as_float(hi.y & 0xffff0000u)));
        __nv_bfloat16* grow = gray + ((size_t)b * 1024 + py * 16 + ky) * 3072 + x4 * 4;
        *reinterpret_cast<uint2*>(grow) = hi;
        *reinterpret_cast<uint2*>(grow + 1024) = lo;
        *reinterpret_cast<uint2*>(grow + 2048) = hi;
      }
    }
  }
}

int patchify_launch(const float* img, __nv_bfloat16* patches, __nv_bfloat16* gray, int gray_split, int B, int C, cudaStream_t st) {
  if (C == 3 && gray_split)
    patchify_kernel<3, true><<<dim3(64, B), 256, 0, st>>>(img, patches, gray);
  else if (C == 3)
    patchify_kernel<3, false><<<dim3(64, B), 256, 0, st>>>(img, patches, gray);
  else if (C == 1)
    patchify_kernel<1, false><<<dim3(64, B), 256, 0, st>>>(img, patches, nullptr);
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

// fp32 [batch, R, C] -> bf16 hi / lo split of the transpose, laid out as the A operand of the second low-pass GEMM:
// out[b][c / 2][seg][c % 2][r], seg 0 and 2 = hi = bf16(v), seg 1 = lo = bf16(v - hi)  (rows of 6 R elements, matching the
// weight [W2h | W2h | W2l]).  64 x 64 tiles; every warp store is 128 contiguous bytes (64 r values of one column).
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C) {
  __shared__ float tile[64][65];
  const size_t boff = (size_t)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i;
    const float2 v = *reinterpret_cast<const float2*>(in + boff + (size_t)r * C + c0 + 2 * tx);
    tile[ty + 8 * i][2 * tx] = v.x;
    tile[ty + 8 * i][2 * tx + 1] = v.y;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + 8 * i;
    const float a = tile[2 * tx][ty + 8 * i], b = tile[2 * tx + 1][ty + 8 * i];
    const uint32_t hi = pack_bf16(a, b);
    const uint32_t lo = pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    uint32_t* o = reinterpret_cast<uint32_t*>(out + ((size_t)blockIdx.z * (C / 2) + (c >> 1)) * (size_t)(6 * R) + (size_t)(c & 1) * R + r0) + tx;
    o[0] = hi;
    o[R] = lo;       // + 2 R elements = R words
    o[2 * R] = hi;
  }
}

int transpose_split_launch(const float* in, __nv_bfloat16* out, int batch, int R, int C, cudaStream_t st) {
  if (R % 64 != 0 || C % 64 != 0) return WM_ERR_SHAPE;
  transpose_split_kernel<<<dim3(C / 64, R / 64, batch), 256, 0, st>>>(in, out, R, C);
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
// (Round 1 ran one THREAD per query row on the CUDA cores, 253 us for 51 queries x 4096 keys at batch 32; superseded by the
// warp-level tensor-core kernel below and removed from the build.)
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
  return v;
}

// ------------------------------------------------------------------ decoder attention on warp-level tensor-core MMAs
// The three decoder shapes (tokens self-attention 51 x 51 x 32, tokens -> image 51 x 4096 x 16, image -> tokens
// 4096 x 51 x 16; transformer.py:218-240) are 3.5 GFLOP per tile batch: nothing for the tensor cores, but the
// thread-per-query kernel above needs one shared-memory broadcast load per 4 FMAs and ran them at 32 / 253 / 143 us
// (1.1 ms of the 42 ms step).  Head dims of 16 / 32 are one or two K = 16 steps: the warp-level mma.sync m16n8k16 form fits
// exactly (a tcgen05 tile would be 128 x N x 16 with 51 live rows and a TMEM round trip per 64 keys; this op is
// latency-bound, not throughput-bound).  One warp = 16 queries, FlashAttention-2 register layout: the S accumulators of two
// n8 tiles are the A fragment of the P V product, K fragments are plain 32-bit shared loads, V fragments come from
// ldmatrix.trans.  Many keys, few queries: the keys are split over a CLUSTER of KSPLIT CTAs and merged by rank 0 through
// distributed shared memory (no second kernel, no global scratch).
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_addr));
}

template <int HD, int KSPLIT>
__global__ void __launch_bounds__(128) attn_mma_kernel(const __nv_bfloat16* __restrict__ q, int ldq,
                                                       const __nv_bfloat16* __restrict__ k, int ldk,
                                                       const __nv_bfloat16* __restrict__ v, int ldv,
                                                       __nv_bfloat16* __restrict__ out, int ldo, int Tq, int Tk,
                                                       float scale_log2e) {
  constexpr int KT = 64;             // keys per staged tile
  constexpr int ROWB = HD * 2 + 16;  // smem row pitch in bytes (+16: ldmatrix / fragment loads without bank conflicts)
  __shared__ __align__(16) uint8_t sK[KT * ROWB];
  __shared__ __align__(16) uint8_t sV[KT * ROWB];
  __shared__ float sM[64], sL[64], sO[64][HD];  // published partial state (cluster merge)
  const int h = blockIdx.y, b = blockIdx.z;
  const int krank = KSPLIT > 1 ? (int)(blockIdx.x % KSPLIT) : 0;
  const int q0 = (int)(blockIdx.x / KSPLIT) * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  // ---- Q fragments (A operand): rows g and g + 8 of this warp's 16 queries, straight from global (zero beyond Tq)
  uint32_t qa[HD / 16][4];
  {
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    const __nv_bfloat16* p0 = q + (size_t)(b * Tq + r0) * ldq + h * HD;
    const __nv_bfloat16* p1 = q + (size_t)(b * Tq + r1) * ldq + h * HD;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      qa[ks][0] = r0 < Tq ? __ldg(reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 2 * t4)) : 0u;
      qa[ks][1] = r1 < Tq ? __ldg(reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 2 * t4)) : 0u;
      qa[ks][2] = r0 < Tq ? __ldg(reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 8 + 2 * t4)) : 0u;
      qa[ks][3] = r1 < Tq ? __ldg(reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 8 + 2 * t4)) : 0u;
    }
  }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  for (int k0 = krank * KT; k0 < Tk; k0 += KT * KSPLIT) {
    __syncthreads();
    // stage 64 keys x HD of K and V (16-byte loads, zero fill past Tk)
    for (int i = threadIdx.x; i < KT * (HD / 8); i += 128) {
      const int kk = i / (HD / 8), d8 = i % (HD / 8);
      uint4 uk = make_uint4(0, 0, 0, 0), uv = make_uint4(0, 0, 0, 0);
      if (k0 + kk < Tk) {
        uk = __ldg(reinterpret_cast<const uint4*>(k + (size_t)(b * Tk + k0 + kk) * ldk + h * HD) + d8);
        uv = __ldg(reinterpret_cast<const uint4*>(v + (size_t)(b * Tk + k0 + kk) * ldv + h * HD) + d8);
      }
      *reinterpret_cast<uint4*>(sK + kk * ROWB + d8 * 16) = uk;
      *reinterpret_cast<uint4*>(sV + kk * ROWB + d8 * 16) = uv;
    }
    __syncthreads();
    // ---- S = Q K^T for the 64 keys of the tile: 8 n8 tiles, accumulators s[j] = (row g: keys 8j + 2 t4, +1; row g + 8: same)
    float sc[KT / 8][4];
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint8_t* kr = sK + (8 * j + g) * ROWB + ks * 32 + 4 * t4;  // B[k = dim][n = key] = K[key][dim]
        mma_bf16_16816(sc[j], qa[ks], *reinterpret_cast<const uint32_t*>(kr), *reinterpret_cast<const uint32_t*>(kr + 16));
      }
    }
    // ---- online softmax (exp2 domain); rows g (c0, c1) and g + 8 (c2, c3); a row lives in the 4 lanes of a quad
    const int nvalid = Tk - k0;  // keys >= nvalid of this tile are padding
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = 8 * j + 2 * t4 + (e & 1);
        sc[j][e] = key < nvalid ? sc[j][e] * scale_log2e : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], sc[j][e]);
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
    }
    const float al[2] = {ex2_approx(m_run[0] - mx[0]), ex2_approx(m_run[1] - mx[1])};  // 0 on the first tile (every tile has a valid key)
    m_run[0] = mx[0];
    m_run[1] = mx[1];
    l_run[0] *= al[0];
    l_run[1] *= al[1];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= al[0]; o[i][1] *= al[0];
      o[i][2] *= al[1]; o[i][3] *= al[1];
    }
    // ---- P V: 16 keys per k-step; the accumulators of n8 tiles 2 kk and 2 kk + 1 ARE the A fragment of k-step kk
#pragma unroll
    for (int kk = 0; kk < KT / 16; ++kk) {
      // P is split into two bf16 terms (hi + lo, 16 mantissa bits): the thread-per-query kernel this one replaces kept P in
      // fp32, and a single bf16 rounding of P moved the full-size ViT-H box error to the edge of its gate (mean L1 1.02e-3
      // vs 1e-3).  Two MMAs per fragment instead of one: nothing for 3.5 GFLOP.
      uint32_t pa[4], pl[4];
      float e[8], el[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        e[u] = ex2_approx(sc[2 * kk + (u >> 2)][u & 3] - mx[(u >> 1) & 1]);  // masked keys: exp2(-inf) = 0
        l_run[(u >> 1) & 1] += e[u];
        el[u] = e[u] - __bfloat162float(__float2bfloat16_rn(e[u]));
      }
      pa[0] = pack_bf16(e[0], e[1]);  // row g,     keys 16 kk + 2 t4, +1
      pa[1] = pack_bf16(e[2], e[3]);  // row g + 8
      pa[2] = pack_bf16(e[4], e[5]);  // row g,     keys 16 kk + 8 + 2 t4, +1
      pa[3] = pack_bf16(e[6], e[7]);  // row g + 8
      pl[0] = pack_bf16(el[0], el[1]);
      pl[1] = pack_bf16(el[2], el[3]);
      pl[2] = pack_bf16(el[4], el[5]);
      pl[3] = pack_bf16(el[6], el[7]);
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        uint32_t b0, b1;  // B[k = key][n = dim] = V[key][dim]: two transposed 8 x 8 tiles (keys 16 kk .. +7, +8 .. +15)
        ldmatrix_x2_trans(b0, b1, smem_u32(sV + (16 * kk + (lane & 15)) * ROWB + i * 16));
        mma_bf16_16816(o[i], pa, b0, b1);
        mma_bf16_16816(o[i], pl, b0, b1);
      }
    }
  }
  // row sums: a row's l is spread over the 4 lanes of its quad
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
  }
  const int row0 = warp * 16 + g;  // local query rows row0, row0 + 8
  if (KSPLIT > 1) {
    if (t4 == 0) {
      sM[row0] = m_run[0]; sM[row0 + 8] = m_run[1];
      sL[row0] = l_run[0]; sL[row0 + 8] = l_run[1];
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      sO[row0][8 * i + 2 * t4] = o[i][0]; sO[row0][8 * i + 2 * t4 + 1] = o[i][1];
      sO[row0 + 8][8 * i + 2 * t4] = o[i][2]; sO[row0 + 8][8 * i + 2 * t4 + 1] = o[i][3];
    }
    cluster_sync();
    if (krank == 0 && threadIdx.x < 64) {  // thread = one query row
      const int rr = threadIdx.x;
      float mg = -INFINITY;
#pragma unroll
      for (int rk = 0; rk < KSPLIT; ++rk) mg = fmaxf(mg, ld_dsmem_f32(mapa_shared(smem_u32(&sM[rr]), rk)));
      float lg = 0.f, og[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) og[d] = 0.f;
#pragma unroll 1
      for (int rk = 0; rk < KSPLIT; ++rk) {
        const float mr = ld_dsmem_f32(mapa_shared(smem_u32(&sM[rr]), rk));
        const float wr = (mr == -INFINITY) ? 0.f : ex2_approx(mr - mg);
        lg = fmaf(ld_dsmem_f32(mapa_shared(smem_u32(&sL[rr]), rk)), wr, lg);
        const uint32_t ar = mapa_shared(smem_u32(&sO[rr][0]), rk);
#pragma unroll
        for (int d = 0; d < HD; ++d) og[d] = fmaf(ld_dsmem_f32(ar + 4 * d), wr, og[d]);
      }
      if (q0 + rr < Tq) {
        const float inv = 1.0f / lg;
        __nv_bfloat16* op = out + (size_t)(b * Tq + q0 + rr) * ldo + h * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 2) *reinterpret_cast<uint32_t*>(op + d) = pack_bf16(og[d] * inv, og[d + 1] * inv);
      }
    }
    cluster_sync();  // the peers' shared memory must stay alive until rank 0 has read it
  } else {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int qi = q0 + row0 + 8 * hh;
      if (qi < Tq) {
        const float inv = 1.0f / l_run[hh];
        __nv_bfloat16* op = out + (size_t)(b * Tq + qi) * ldo + h * HD;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i)
          *reinterpret_cast<uint32_t*>(op + 8 * i + 2 * t4) = pack_bf16(o[i][2 * hh] * inv, o[i][2 * hh + 1] * inv);
      }
    }
  }
}

template <int HD, int KS>
static int attn_mma_launch(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                           __nv_bfloat16* out, int ldo, int B, int H, int Tq, int Tk, float sl2, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((Tq + 63) / 64) * KS, H, B);
  cfg.blockDim = dim3(128, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = KS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = KS > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attn_mma_kernel<HD, KS>, q, ldq, k, ldk, v, ldv, out, ldo, Tq, Tk, sl2) == cudaSuccess
             ? WM_OK : WM_ERR_CUDA;
}

int attn_small_launch(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* v, int ldv,
                      __nv_bfloat16* out, int ldo, int B, int H, int Tq, int Tk, int hd, float scale, cudaStream_t st) {
  if (B * H * Tq == 0) return WM_OK;
  if ((ldq | ldk | ldv | ldo) % 8 != 0) return WM_ERR_SHAPE;
  const float sl2 = scale * 1.4426950408889634f;
  if ((ldq | ldk | ldv | ldo) % 8 == 0 && ((Tq + 63) / 64) * 8 <= 65535) {
    // tensor-core path (every decoder shape); few queries against many keys: keys split over a small cluster
    // Cluster size of the key split, measured at 51 x 4096 x 16, batch 32 (profiles/attn_small_scan.py): 1 CTA 60.6 us,
    // 2: 53.8, 4: 60.2, 8: 90.3 -- cluster launches carry a fixed cost that grows with the cluster size (~67 us at 8), the
    // per-tile cost is the same.  (Thread-per-query CUDA-core kernel: 253 us.)
#ifndef WM_ATTN_KS
#define WM_ATTN_KS 2
#endif
    const bool ks8 = Tq <= 256 && Tk >= 1024;
    if (hd == 16) return ks8 ? attn_mma_launch<16, WM_ATTN_KS>(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Tq, Tk, sl2, st)
                             : attn_mma_launch<16, 1>(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Tq, Tk, sl2, st);
    if (hd == 32) return ks8 ? attn_mma_launch<32, WM_ATTN_KS>(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Tq, Tk, sl2, st)
                             : attn_mma_launch<32, 1>(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Tq, Tk, sl2, st);
    return WM_ERR_SHAPE;
  }
  return WM_ERR_SHAPE;  // more than 65535 query blocks per (image, head)
}

}  // namespace wm
