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

#ifdef WM_EPI_NO_SETMAXNREG
#define WM_SETMAXNREG_DEC() do { } while (0)
#define WM_SETMAXNREG_INC() do { } while (0)
#else
#define WM_SETMAXNREG_DEC() setmaxnreg_dec<40>()
#define WM_SETMAXNREG_INC() setmaxnreg_inc<232>()
#endif

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
  static constexpr int OFF_EPI = kStages * STAGE_BYTES;  // per-epilogue-warp 32 x 32 fp32 transpose buffers
  static constexpr int EPI_BYTES = kEpiWarps * 4096;
  static constexpr int OFF_BAR = OFF_EPI + EPI_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 /*barriers*/ + 1024 /*align slack*/;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static constexpr int TMEM_COLS = 2 * BN >= 32 ? 2 * BN : 32;
};

// Exact-erf GELU (nn.GELU(), M/common.py:13-26): x (0.5 + 0.5 erf(x / sqrt 2)) = x sat(0.5 + x Q(t)), t = min(x^2, 18), Q a
// degree-8 fit of 0.5 erf(x / sqrt 2) / x in x^2 on |x| <= 3 sqrt 2; beyond that x Q(18) leaves [-0.5, 0.5] and the
// saturating FMA (free on the FMA pipe) pins the factor to exactly 0 or 1.  11 FMA-pipe operations + 1 min and no MUFU:
// the GELU epilogue has to stay under the main-loop time of the next tile and is issue-bound (two epilogue warps per
// scheduler; erff() costs ~25 operations, a rcp/ex2 formulation 2 MUFU per element -- both measured slower;
// profiles/r01z_ncu_gemm_gelu_summary.txt).  |error| <= 4.8e-5 on GELU, far below the bf16 rounding of the stored activation.
#define WM_GELU_C0 0.3988664448261261f
#define WM_GELU_C1 -0.06624268740415573f
#define WM_GELU_C2 0.00973955076187849f
#define WM_GELU_C3 -0.0010831074323505163f
#define WM_GELU_C4 8.891491597751155e-05f
#define WM_GELU_C5 -5.17239732289454e-06f
#define WM_GELU_C6 1.99358936470162e-07f
#define WM_GELU_C7 -4.523354135699265e-09f
#define WM_GELU_C8 4.543853834859668e-11f
#define WM_GELU_LIM 4.242640687119285f
__device__ __forceinline__ float gelu_erf(float x) {
  const float t = fminf(x * x, WM_GELU_LIM * WM_GELU_LIM);
  float q = fmaf(WM_GELU_C8, t, WM_GELU_C7);
  q = fmaf(q, t, WM_GELU_C6);
  q = fmaf(q, t, WM_GELU_C5);
  q = fmaf(q, t, WM_GELU_C4);
  q = fmaf(q, t, WM_GELU_C3);
  q = fmaf(q, t, WM_GELU_C2);
  q = fmaf(q, t, WM_GELU_C1);
  q = fmaf(q, t, WM_GELU_C0);
  return x * __saturatef(fmaf(x, q, 0.5f));
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Two GELUs at once on the packed fp32x2 FMA path of sm_100 (FFMA2 / FMUL2; common.cuh)
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const uint64_t xp = pk2(x0, x1);
  float t0, t1;
  unpk2(mul2(xp, xp), t0, t1);
  const uint64_t t = pk2(fminf(t0, WM_GELU_LIM * WM_GELU_LIM), fminf(t1, WM_GELU_LIM * WM_GELU_LIM));
  uint64_t q = fma2(pk2(WM_GELU_C8, WM_GELU_C8), t, pk2(WM_GELU_C7, WM_GELU_C7));
  q = fma2(q, t, pk2(WM_GELU_C6, WM_GELU_C6));
  q = fma2(q, t, pk2(WM_GELU_C5, WM_GELU_C5));
  q = fma2(q, t, pk2(WM_GELU_C4, WM_GELU_C4));
  q = fma2(q, t, pk2(WM_GELU_C3, WM_GELU_C3));
  q = fma2(q, t, pk2(WM_GELU_C2, WM_GELU_C2));
  q = fma2(q, t, pk2(WM_GELU_C1, WM_GELU_C1));
  q = fma2(q, t, pk2(WM_GELU_C0, WM_GELU_C0));
  float q0, q1;
  unpk2(q, q0, q1);
  x0 *= __saturatef(fmaf(x0, q0, 0.5f));
  x1 *= __saturatef(fmaf(x1, q1, 0.5f));
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return gelu_erf(x);
  if (act == 2) return fmaxf(x, 0.0f);
  if (act == 3) return 1.0f / (1.0f + expf(-x));
  return x;
}

// Epilogue of one 128 x BN accumulator tile by the 8 epilogue warps (warp w: TMEM lane quarter w & 3, column half
// (w - 4) / 4), in 32-column chunks.  tcgen05.ld hands every lane one ROW of a chunk.  Bias, activation (and a
// residual that is not updated in place) are applied in that layout with broadcast loads; the result is written as
// 16-byte pieces into the warp's 4 KB staging buffer in the SWIZZLE_128B pattern (conflict-free) and leaves the SM as
// a TMA store: no shared-memory read-back, no per-lane global stores, no address arithmetic.  The in-place fp32
// residual update of the encoder (x += proj(...), x += lin2(...)) is a TMA REDUCE-ADD: the residual is never read by
// the SM.  (Per-lane row stores touch 32 cache lines per instruction; a shared-memory transpose with coalesced
// per-lane stores needed ~1500 instructions per warp and tile and left the MMA warp waiting for free accumulators
// more than half of the time -- profiles/r01c_gemm_shapes.jsonl, r01j_ncu_gemm_summary.txt.)
// `tempty` is arrived on (once per warp) as soon as this warp has read its share of the accumulator; `remote` = the
// barrier lives in the leader CTA of the pair (2-CTA kernel, non-leader CTA).
// STAGE_BUFS = 4 KB staging buffers per warp (2 in the CTA-pair kernel): with two, single-output stores alternate between
// them (`sbuf`, carried across tiles) and only wait for the store before the previous one.
// ACT >= 0 / FAST (1: bf16 output, 2: fp32 output): compile-time activation and "single output through the TMA path" (the CTA-pair kernel is instantiated
// for the hot combinations: the epilogue of the GELU GEMM is its bottleneck, and with every activation and output path
// inlined four times per tile the kernel was 155 KB of code with 8-12 % instruction-fetch stalls in the epilogue warps).
template <int BN, int STAGE_BUFS, int ACT = -1, int FAST = 0>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const CUtensorMap* tc16, const CUtensorMap* tc32,
                                              uint8_t* stage, uint32_t tmem_acc, int tile_row0, int tile_n0, int warp,
                                              int lane, uint64_t* tempty, bool remote, uint32_t& sbuf) {
  const int q = warp & 3;                // TMEM lane quarter this warp may access
  const int half = (warp - 4) >> 2;      // column half of the tile
  constexpr int NCH = BN / 64;           // 32-column chunks per warp
  const int row0 = tile_row0 + q * 32;
  const int row = row0 + lane;
  const int wcol0 = half * (BN / 2);     // first tile column of this warp
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)wcol0;
  // number of chunks of this warp that lie (at least partly) inside N: warp-uniform
  // (the CTA-pair kernel only runs with N % 256 == 0: no ragged right edge, and the dead edge code stays out of it)
  constexpr bool kWholeTiles = STAGE_BUFS == 2;
  int nact = NCH;
  if (!kWholeTiles) {
    nact = (p.N - (tile_n0 + wcol0) + 31) / 32;
    nact = nact < 0 ? 0 : (nact > NCH ? NCH : nact);
  }
  auto release = [&]() {  // this warp has read its share of the accumulator
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (remote) mbar_arrive_remote(tempty, 0);
      else mbar_arrive(tempty);
    }
  };
  if (nact == 0) {
    release();
    return;
  }
  const bool row_ok = row < p.M;
  const int act = ACT >= 0 ? ACT : p.act;
  if (FAST || p.tma_out) {
    const uint32_t st_row = smem_u32(stage) + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    const float* res_row = (p.residual != nullptr && !p.res_inplace) ? p.residual + (size_t)(row % p.res_mod) * p.ldr : nullptr;
    const bool alt = STAGE_BUFS == 2 && (FAST || p.tma_out == 1);  // single output: alternate the warp's two staging buffers
    uint32_t v[2][32];
    float4 bq[2][8];  // bias of a chunk (the same 32 values in every lane), fetched one chunk ahead of its use
    auto fetch_bias = [&](int c) {
      const int n0 = tile_n0 + wcol0 + c * 32;
      if (p.bias != nullptr && n0 + 32 <= p.N) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
        for (int j = 0; j < 8; ++j) bq[c & 1][j] = __ldg(b4 + j);
      }
    };
    tmem_ld32(taddr, v[0]);
    fetch_bias(0);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c < nact) {
        const int n0 = tile_n0 + wcol0 + c * 32;
        tmem_ld_wait();  // chunk c has landed
        if (c + 1 < NCH && c + 1 < nact) {
          tmem_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
          fetch_bias(c + 1);
        }
        if (c == nact - 1) release();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[c & 1][j]);
        const bool full = kWholeTiles || n0 + 32 <= p.N;  // (N % 8 == 0 on this path; partial chunks only at the right edge)
        if (p.bias != nullptr) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = bq[c & 1][j];
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) f[j] += __ldg(p.bias + n0 + j);
          }
        }
        if (act == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) gelu_erf2(f[2 * j], f[2 * j + 1]);
        } else if (act != 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], act);
        }
        if (res_row != nullptr && row_ok) {  // residual that is not updated in place (pos-embed broadcast, decoder): row-per-lane reads
          if (full) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 rq = __ldg(reinterpret_cast<const float4*>(res_row + n0) + j);
              f[4 * j] += rq.x; f[4 * j + 1] += rq.y; f[4 * j + 2] += rq.z; f[4 * j + 3] += rq.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) f[j] += __ldg(res_row + n0 + j);
          }
        }
        if (FAST != 1 && p.out_f32 != nullptr) {
          // ---- fp32 output: 32 columns = one 128-byte swizzled row per lane
          // the earlier stores of this warp have finished reading the staging buffer about to be rewritten
          const uint32_t off32 = alt ? (sbuf & 1u) * 4096u : 0u;
          if (lane == 0) {
            if (alt) tma_store_wait_read_but<1>();
            else tma_store_wait_read();
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            sts128(st_row + off32 + ((k ^ sw) << 4), __float_as_uint(f[4 * k]), __float_as_uint(f[4 * k + 1]),
                   __float_as_uint(f[4 * k + 2]), __float_as_uint(f[4 * k + 3]));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (p.res_inplace) tma_reduce_add_2d(tc32, stage + off32, n0, row0);
            else tma_store_2d(tc32, stage + off32, n0, row0);
            tma_store_commit();
          }
          ++sbuf;
        }
        if (FAST != 2 && p.out_bf16 != nullptr) {
          // ---- bf16 output: two chunks share one 128-byte row (64 columns): even chunk -> 16-byte pieces 0..3, odd -> 4..7.
          // An unpaired last chunk only happens at the right edge of the matrix, where TMA clips the unwritten half.
          // With both outputs (tma_out == 2, CTA-pair kernel) the bf16 rows use the warp's second staging buffer; the
          // wait above (every chunk) already covers its reuse.
          const uint32_t off16 = (p.out_f32 != nullptr) ? 4096u : (alt ? (sbuf & 1u) * 4096u : 0u);
          if ((c & 1) == 0 && p.out_f32 == nullptr) {
            if (lane == 0) {
              if (alt) tma_store_wait_read_but<1>();
              else tma_store_wait_read();
            }
            __syncwarp();
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            sts128(st_row + off16 + (((uint32_t)((c & 1) * 4 + k) ^ sw) << 4), pack_bf16(f[8 * k], f[8 * k + 1]),
                   pack_bf16(f[8 * k + 2], f[8 * k + 3]), pack_bf16(f[8 * k + 4], f[8 * k + 5]),
                   pack_bf16(f[8 * k + 6], f[8 * k + 7]));
          if ((c & 1) == 1 || c == nact - 1) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(tc16, stage + off16, n0 - (c & 1) * 32, row0);
              tma_store_commit();
            }
            if (p.out_f32 == nullptr) ++sbuf;
          }
        }
      }
    }
    return;
  }
  if (FAST) return;  // (compile time: the generic path below is not instantiated for the fast variants)
  // generic path (ragged or unaligned outputs -- decoder heads N = 8, N = 4 -- and dual fp32 + bf16 outputs): scalar
  // row-per-lane epilogue
  const float* res_row = (p.residual != nullptr && row_ok) ? p.residual + (size_t)(row % p.res_mod) * p.ldr : nullptr;
#pragma unroll 1
  for (int c = 0; c < nact; ++c) {
    const int n0 = tile_n0 + wcol0 + c * 32;
    uint32_t v[32];
    tmem_ld32(taddr + c * 32, v);
    tmem_ld_wait();
    if (c == nact - 1) release();
    if (row_ok) {
      const bool vec = p.vec_ok && (n0 + 32 <= p.N);
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr && n0 + j < p.N) f[j] += __ldg(p.bias + n0 + j);
        f[j] = apply_act(f[j], act);
      }
      if (vec) {
        if (res_row != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 rq = __ldg(reinterpret_cast<const float4*>(res_row + n0) + j);
            f[4 * j] += rq.x; f[4 * j + 1] += rq.y; f[4 * j + 2] += rq.z; f[4 * j + 3] += rq.w;
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
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + j;
          if (n < p.N) {
            float x = f[j];
            if (res_row != nullptr) x += __ldg(res_row + n);
            if (p.out_f32 != nullptr) p.out_f32[(size_t)row * p.ldc_f32 + n] = x;
            if (p.out_bf16 != nullptr) p.out_bf16[(size_t)row * p.ldc_bf16 + n] = __float2bfloat16_rn(x);
          }
        }
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_c16, const __grid_constant__ CUtensorMap tmap_c32,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
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

  // register split (setmaxnreg inside the role branches, never before a reconvergence point): the producer / MMA
  // warpgroup gives registers to the two epilogue warpgroups
  if (warp == 0) {
    WM_SETMAXNREG_DEC();
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
    WM_SETMAXNREG_DEC();
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
  } else if (warp < 4) {
    WM_SETMAXNREG_DEC();  // (all four warps of the warpgroup have to execute it)
  } else {
    WM_SETMAXNREG_INC();
    // ------------------------------------------------------------ epilogue (8 warps; see epilogue_tile)
    uint8_t* stage_buf = smem + Cfg::OFF_EPI + (warp - 4) * 4096;
    int as = 0;
    uint32_t aphase = 0, sbuf = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n_blk = t % num_n, m_blk = t / num_n;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      epilogue_tile<BN, 1>(p, &tmap_c16, &tmap_c32, stage_buf, tmem_base + as * BN, m_blk * BM, n_blk * BN, warp, lane,
                           &tempty_bar[as], false, sbuf);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait_read();  // the staging buffer must stay valid until the last store has read it
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each CTA loads
// its own 128 rows of A and HALF of the 256 W rows per k-block (32 KB instead of 48 KB per 128 x 256 of output), the
// leader CTA issues 256 x 256 x 16 MMAs that read both CTAs' shared memory and write both CTAs' tensor memory.  The
// single-CTA kernel is L2-bandwidth bound on the large encoder GEMMs (96 B/clk/SM of operand traffic).
//   full barriers   live in the LEADER: its producer arms them with both CTAs' bytes, both CTAs' TMA loads signal them
//   empty / tfull   are signalled in BOTH CTAs by multicast tcgen05.commit
//   tempty          lives in the leader: 2 x 8 epilogue warps arrive (the peer's remotely)
struct Gemm2Cfg {
  static constexpr int kStages = 5;
  static constexpr int A_BYTES = 128 * BK * 2;   // this CTA's 128 rows of A
  static constexpr int B_BYTES = 128 * BK * 2;   // this CTA's half of the 256 W rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_EPI = kStages * STAGE_BYTES;
  static constexpr int EPI_BYTES = kEpiWarps * 8192;  // two 4 KB staging buffers per epilogue warp (dual fp32 + bf16 output)
  static constexpr int OFF_BAR = OFF_EPI + EPI_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static constexpr int TMEM_COLS = 512;  // 2 accumulator stages x 256 columns
};

template <int ACT, int FAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_c16, const __grid_constant__ CUtensorMap tmap_c32,
                  const GemmParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_m = (p.M + 255) / 256;
  const int num_n = p.N / 256;  // host guarantees N % 256 == 0
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
      mbar_init(&tempty_bar[i], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync();  // barrier inits and the TMEM allocation of both CTAs are visible to the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    WM_SETMAXNREG_DEC();
    // ------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int n_blk = t % num_n, m_blk = t / num_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          tma_load_2d_2cta(sa, &tmap_a, leader_full, kb * BK, m_blk * 256 + (int)rank * 128);
          tma_load_2d_2cta(sb, &tmap_w, leader_full, kb * BK, n_blk * 256 + (int)rank * 128);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    WM_SETMAXNREG_DEC();
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, 256, 0, 0);
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
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
            for (int k = 0; k < BK / 16; ++k) umma_bf16_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2cta(&empty_bar[stage]);  // frees this smem stage in BOTH CTAs
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit_2cta(&tfull_bar[as]);  // accumulator halves complete in both CTAs
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp < 4) {
    WM_SETMAXNREG_DEC();
  } else {
    WM_SETMAXNREG_INC();
    // ------------------------------------------------------------ epilogue (both CTAs: own 128 rows of the tile)
    uint8_t* stage_buf = smem + Cfg::OFF_EPI + (warp - 4) * 8192;
    int as = 0;
    uint32_t aphase = 0, sbuf = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int n_blk = t % num_n, m_blk = t / num_n;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      epilogue_tile<BN, 2, ACT, FAST>(p, &tmap_c16, &tmap_c32, stage_buf, tmem_base + as * BN, m_blk * 256 + (int)rank * 128, n_blk * BN,
                           warp, lane, &tempty_bar[as], rank != 0, sbuf);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait_read();
  }

  tc_fence_before();
  cluster_sync();  // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int ACT, int FAST>
static int launch_gemm2_variant(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc16, const CUtensorMap& tc32,
                                const GemmParams& p, int num_sms, cudaStream_t st) {
  using Cfg = Gemm2Cfg;
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(gemm2_bf16_kernel<ACT, FAST>, Cfg::SMEM_BYTES, attr_done)) return rc;
  const int tiles = ((p.M + 255) / 256) * (p.N / 256);
  int clusters = num_sms / 2;
  if (tiles < clusters) clusters = tiles;
  gemm2_bf16_kernel<ACT, FAST><<<2 * clusters, kGemmThreads, Cfg::SMEM_BYTES, st>>>(ta, tw, tc16, tc32, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc16, const CUtensorMap& tc32,
                        const GemmParams& p, int num_sms, cudaStream_t st) {
  // lin1 + GELU (epilogue-bound) gets its own instantiation: -9 % (0.552 -> 0.501 ms at batch 32).  The epilogue-free
  // shapes are main-loop bound and measured 0-1 % slower when specialised the same way, so they stay on the general kernel.
  if (p.tma_out == 1 && p.act == 1 && p.out_bf16 != nullptr) return launch_gemm2_variant<1, 1>(ta, tw, tc16, tc32, p, num_sms, st);
  return launch_gemm2_variant<-1, 0>(ta, tw, tc16, tc32, p, num_sms, st);
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc16, const CUtensorMap& tc32,
                       const GemmParams& p, int num_sms, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(gemm_bf16_kernel<BN>, Cfg::SMEM_BYTES, attr_done)) return rc;
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_bf16_kernel<BN><<<grid, kGemmThreads, Cfg::SMEM_BYTES, st>>>(ta, tw, tc16, tc32, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

int gemm_dispatch(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc16, const CUtensorMap& tc32,
                  const GemmParams& p, int bn, int num_sms, cudaStream_t st) {
  if (bn == 512) return launch_gemm2(ta, tw, tc16, tc32, p, num_sms, st);  // CTA-pair kernel (A / W boxes of 128 rows)
  if (bn == 256) return launch_gemm<256>(ta, tw, tc16, tc32, p, num_sms, st);
  if (bn == 128) return launch_gemm<128>(ta, tw, tc16, tc32, p, num_sms, st);
  if (bn == 64) return launch_gemm<64>(ta, tw, tc16, tc32, p, num_sms, st);
  return WM_ERR_SHAPE;
}

}  // namespace wm
