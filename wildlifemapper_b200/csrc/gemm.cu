// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] * W[N,K]^T )
//
//   A : bf16, K contiguous (row-major activations)      -> K-major UMMA operand, TMA SWIZZLE_128B
//   W : bf16 [N,K], K contiguous (nn.Linear layout)     -> K-major UMMA operand
//   C : bf16 and/or fp32 row-major, fused epilogue: +bias[n] -> act (none | erf-GELU | ReLU) -> +residual
//
// Covers every dense linear site of the hot path: patch/hfc embed (after patchify),
// qkv/proj, MLP lin1/lin2, all HFC-branch linears (reference image_encoder.py:468-481, 494-513), neck 1x1
// and -- through a 4-D activation tensor map -- the neck 3x3 conv (image_encoder.py:113-119) as an
// implicit GEMM with zero padding supplied by TMA out-of-bounds fill; decoder projections and heads.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0    TMA producer        (kStages-deep smem ring, mbarrier full/empty)
//   warp 1    tcgen05.mma issuer  (UMMA 128 x BN x 16, fp32 accumulators in TMEM, 2 accumulator stages)
//   warp 2    TMEM allocator
//   warps 4+  epilogue            (tcgen05.ld 32x32b -> registers -> fused math -> 16-byte global stores)
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = (4 + kEpiWarps) * 32;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN >= 32 ? 2 * BN : 32;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = warp_idx_uniform();  // uniform role branches: see common.cuh
  const int lane = threadIdx.x & 31;
  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp runs the loop,
    // one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n_blk = t % num_n, m_blk = t / num_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (p.a_mode == 0) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
            // implicit 3x3 conv on NHWC [B,64,64,C]: m_blk -> (image, two rows); kb -> (tap, channel block)
            const int cblocks = p.conv_C / BK;
            const int tap = kb / cblocks, c0 = (kb % cblocks) * BK;
            const int img = m_blk >> 5, y0 = (m_blk & 31) * 2;
            tma_load_4d(sa, &tmap_a, &full_bar[stage], c0, tap % 3 - 1, y0 + tap / 3 - 1, img);
          }
          tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * BK, n_blk * BN);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane)
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    const bool leader = elect_one();  // the same lane issues every tcgen05.mma / tcgen05.commit
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = make_sdesc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_sdesc_sw128(sa + Cfg::A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr>>4) field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&tfull_bar[as]);  // accumulator complete
      __syncwarp();
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;      // column half of the tile
    constexpr int COLS_PER_WARP = BN / 2;
    int as = 0;
    uint32_t aphase = 0;
    const bool vec_ok = p.vec_ok != 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n_blk = t % num_n, m_blk = t / num_n;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float* res_row = nullptr;
      if (p.residual != nullptr && row_ok) res_row = p.residual + (size_t)(row % p.res_mod) * p.ldr;
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 32) {
        const int col0 = half * COLS_PER_WARP + c;
        const int n0 = n_blk * BN + col0;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + col0), v);
        tmem_ld_wait();
        if (!row_ok) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const bool full = vec_ok && (n0 + 32 <= p.N);
        if (full) {
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 b = __ldg(b4 + j);
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          } else if (p.act == 3) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 1.0f / (1.0f + expf(-f[j]));
          }
          if (res_row != nullptr) {
            const float4* r4 = reinterpret_cast<const float4*>(res_row + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 r = __ldg(r4 + j);
              f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
            }
          }
          if (p.out_f32 != nullptr) {
            float4* o4 = reinterpret_cast<float4*>(p.out_f32 + (size_t)row * p.ldc_f32 + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
          if (p.out_bf16 != nullptr) {
            uint4* o4 = reinterpret_cast<uint4*>(p.out_bf16 + (size_t)row * p.ldc_bf16 + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o4[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
          }
        } else {
          // ragged tail / unaligned output: scalar path (decoder heads N=8, N=4)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            if (n < p.N) {
              float x = f[j];
              if (p.bias != nullptr) x += __ldg(p.bias + n);
              if (p.act == 1) x = gelu_erf(x);
              else if (p.act == 2) x = fmaxf(x, 0.0f);
              else if (p.act == 3) x = 1.0f / (1.0f + expf(-x));
              if (res_row != nullptr) x += __ldg(res_row + n);
              if (p.out_f32 != nullptr) p.out_f32[(size_t)row * p.ldc_f32 + n] = x;
              if (p.out_bf16 != nullptr) p.out_bf16[(size_t)row * p.ldc_bf16 + n] = __float2bfloat16_rn(x);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, int num_sms, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) !=
        cudaSuccess)
      return WM_ERR_CUDA;
    attr_set = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_bf16_kernel<BN><<<grid, kGemmThreads, Cfg::SMEM_BYTES, st>>>(ta, tw, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

int gemm_dispatch(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, int bn, int num_sms,
                  cudaStream_t st) {
  if (bn == 256) return launch_gemm<256>(ta, tw, p, num_sms, st);
  if (bn == 128) return launch_gemm<128>(ta, tw, p, num_sms, st);
  if (bn == 64) return launch_gemm<64>(ta, tw, p, num_sms, st);
  return WM_ERR_SHAPE;
}

}  // namespace wm
