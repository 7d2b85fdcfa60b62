// Flash attention v2 on tcgen05 / TMEM (sm_100a): global 64x64 attention with decomposed rel-pos bias (HD = 64)
// and the HFC cross-attention (HD = 128).  Same math and reference sites as attn_flash.cu
// (image_encoder.py:246-262, 347-383, 500-503); restructured around what the first profile showed
// (profiles/r01_*: one softmax warp per scheduler, MUFU- and issue-bound, tensor pipe 12 % active):
//
//   * one CTA = TWO 128-query tiles of one (image, head): 8 softmax warps (2 per scheduler) so that the exp /
//     FMA work of one tile overlaps the MMAs and TMEM latency of the other;
//   * key tiles of 64 (= one key row of the 64x64 grid): the row bias bias_h is a single scalar per query and
//     tile, read from TMEM (T_h = Q Rh_slice^T stays resident), and bias_w[64] lives in registers;
//   * O and the softmax denominators accumulate in TMEM across key tiles (P V and P 1 MMAs with accumulate),
//     with LAZY rescaling: the running reference maximum is only raised (and O, l rescaled through
//     tcgen05.ld / tcgen05.st) when a tile exceeds it by more than 2^8 -- the common path is a single pass
//     per score: FFMA, FADD, FMNMX, MUFU.EX2, pack, 16-byte swizzled store;
//   * row sums come from an extra N=16 MMA against a tile of ones, i.e. from the same bf16-rounded P that
//     multiplies V.
//
//   warp 0       TMA producer (Q tiles, tables, 3-stage K and V rings)
//   warp 1       tcgen05.mma issuer          warp 2  TMEM allocator
//   warps 4-7    softmax of query tile 0     warps 8-11  softmax of query tile 1 (thread = one query row)
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

constexpr int F2_THREADS = 384;
constexpr int F2_STAGES = 3;
constexpr float F2_LOG2E = 1.4426950408889634f;
constexpr float F2_TAU = 8.0f;  // lazy-rescale threshold (log2 units): p <= 2^8 between rescales

template <int HD, bool RELPOS>
struct Flash2Cfg {
  static constexpr int SUB = HD / 64;
  static constexpr int Q_BYTES = SUB * 16384;      // one 128-query tile
  static constexpr int KV_BYTES = SUB * 8192;      // one 64-key tile (K or V)
  static constexpr int OFF_Q = 0;                  // 2 query tiles
  static constexpr int OFF_K = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + F2_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + F2_STAGES * KV_BYTES;  // 2 x [128 x 64] bf16
  static constexpr int OFF_ONES = OFF_P + 2 * 16384;          // [16 x 64] bf16 of 1.0
  static constexpr int OFF_TAB = OFF_ONES + 2048;             // Rw [128 rows] + 2 x Rh slice [80 rows]; later fp32 scratch
  static constexpr int TAB_BYTES = RELPOS ? (16384 + 2 * 10240) : 0;
  static constexpr int OFF_BAR = OFF_TAB + TAB_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  // TMEM columns
  static constexpr int COL_S = 0;                  // S_t at COL_S + 64 t
  static constexpr int COL_O = 128;                // O_t at COL_O + HD t
  static constexpr int COL_L = COL_O + 2 * HD;     // L_t at COL_L + 16 t
  static constexpr int COL_TH = COL_L + 32;        // T_h_t at COL_TH + 80 t (RELPOS)
  static constexpr int TMEM_COLS = 512;
  static_assert(COL_TH + (RELPOS ? 160 : 0) <= 512, "TMEM budget");
};

template <int HD, bool RELPOS>
__global__ void __launch_bounds__(F2_THREADS, 1)
flash2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_rel,
              const FlashParams p) {
  using Cfg = Flash2Cfg<HD, RELPOS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [3]
  uint64_t* v_empty = bars + 10;  // [3]
  uint64_t* s_full = bars + 13;   // [2]
  uint64_t* p_full = bars + 15;   // [2]
  uint64_t* o_full = bars + 17;   // [2]
  uint64_t* t_full = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int nk = p.Tk / 64;

  // constant tile of ones (B operand of the row-sum MMA); any layout of all-ones is all-ones
  for (int i = threadIdx.x; i < 2048 / 4; i += F2_THREADS) reinterpret_cast<uint32_t*>(smem + Cfg::OFF_ONES)[i] = 0x3F803F80u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < F2_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
    mbar_init(t_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * Cfg::Q_BYTES + Cfg::TAB_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int s = 0; s < Cfg::SUB; ++s)
          tma_load_2d(smem + Cfg::OFF_Q + t * Cfg::Q_BYTES + s * 16384, &tmap_q, q_full, p.q_col0 + h * HD + s * 64,
                      b * p.Tq + m0 + t * 128);
      if (RELPOS) {
        // table tensor [256,64]: rows 0..126 rel_pos_h, 128..254 rel_pos_w.  Box = 16 rows.
        const int qi0 = m0 >> 6;  // first image row of this CTA (4 rows: 2 per query tile)
        for (int i = 0; i < 8; ++i) tma_load_2d(smem + Cfg::OFF_TAB + i * 2048, &tmap_rel, q_full, 0, 128 + 16 * i);
        for (int t = 0; t < 2; ++t)
          for (int i = 0; i < 5; ++i)
            tma_load_2d(smem + Cfg::OFF_TAB + 16384 + t * 10240 + i * 2048, &tmap_rel, q_full, 0, qi0 + 2 * t + 16 * i);
      }
      for (int j = 0; j < nk; ++j) {
        const int st = j % F2_STAGES;
        const uint32_t par = ((j / F2_STAGES) & 1) ^ 1;
        mbar_wait_parked(&k_empty[st], par);
        mbar_arrive_expect_tx(&k_full[st], Cfg::KV_BYTES);
#pragma unroll
        for (int s = 0; s < Cfg::SUB; ++s)
          tma_load_2d(smem + Cfg::OFF_K + st * Cfg::KV_BYTES + s * 8192, &tmap_k, &k_full[st], p.k_col0 + h * HD + s * 64,
                      b * p.Tk + j * 64);
        mbar_wait_parked(&v_empty[st], par);
        mbar_arrive_expect_tx(&v_full[st], Cfg::KV_BYTES);
#pragma unroll
        for (int s = 0; s < Cfg::SUB; ++s)
          tma_load_2d(smem + Cfg::OFF_V + st * Cfg::KV_BYTES + s * 8192, &tmap_v, &v_full[st], p.v_col0 + h * HD + s * 64,
                      b * p.Tk + j * 64);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);  // V is MN-major
      constexpr uint32_t idesc_l = make_idesc_bf16(128, 16, 0, 0);
      constexpr uint32_t idesc_tw = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_th = make_idesc_bf16(128, 80, 0, 0);
      constexpr uint32_t PB = RELPOS ? 1u : 0u;  // p_full phases consumed by the prologue
      const uint32_t sq = smem_u32(smem + Cfg::OFF_Q);
      const uint32_t sp = smem_u32(smem + Cfg::OFF_P);
      const uint32_t sones = smem_u32(smem + Cfg::OFF_ONES);
      const uint32_t stab = smem_u32(smem + Cfg::OFF_TAB);
      mbar_wait_parked(q_full, 0);
      tc_fence_after();
      if (RELPOS) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_sdesc_sw128(sq + t * Cfg::Q_BYTES + ks * 32, 16, 1024);
            umma_bf16(tmem_base + t * 128, ad, make_sdesc_sw128(stab + ks * 32, 16, 1024), idesc_tw, ks != 0);
            umma_bf16(tmem_base + Cfg::COL_TH + t * 80, ad, make_sdesc_sw128(stab + 16384 + t * 10240 + ks * 32, 16, 1024),
                      idesc_th, ks != 0);
          }
        }
        umma_commit(t_full);
      }
      auto issue_s = [&](int t, int j) {
        const int st = j % F2_STAGES;
        const uint32_t sk = smem_u32(smem + Cfg::OFF_K + st * Cfg::KV_BYTES);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          umma_bf16(tmem_base + Cfg::COL_S + t * 64,
                    make_sdesc_sw128(sq + t * Cfg::Q_BYTES + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                    make_sdesc_sw128(sk + (ks >> 2) * 8192 + (ks & 3) * 32, 16, 1024), idesc_s, ks != 0);
        }
        umma_commit(&s_full[t]);
      };
      // first score tiles (after the softmax warps have drained the rel-pos tables out of the S/O columns)
      mbar_wait_parked(&k_full[0], 0);
      for (int t = 0; t < 2; ++t) {
        if (RELPOS) mbar_wait_parked(&p_full[t], 0);
        tc_fence_after();
        issue_s(t, 0);
      }
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nk; ++j) {
        const int st = j % F2_STAGES;
        const uint32_t sv = smem_u32(smem + Cfg::OFF_V + st * Cfg::KV_BYTES);
        const bool more = j + 1 < nk;
        mbar_wait_parked(&v_full[st], (j / F2_STAGES) & 1);
        if (more) mbar_wait_parked(&k_full[(j + 1) % F2_STAGES], ((j + 1) / F2_STAGES) & 1);
        for (int t = 0; t < 2; ++t) {
          mbar_wait_parked(&p_full[t], (j + PB) & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // 64 keys, 16 per MMA
            const uint64_t ad = make_sdesc_sw128(sp + t * 16384 + ks * 32, 16, 1024);
            umma_bf16(tmem_base + Cfg::COL_O + t * HD, ad, make_sdesc_sw128(sv + ks * 2048, 8192, 1024), idesc_pv,
                      (j | ks) != 0);
            umma_bf16(tmem_base + Cfg::COL_L + t * 16, ad, make_sdesc_sw128(sones + ks * 32, 16, 1024), idesc_l,
                      (j | ks) != 0);
          }
          if (more) issue_s(t, j + 1);
          else umma_commit(&o_full[t]);
        }
        umma_commit(&v_empty[st]);
        if (more) umma_commit(&k_empty[(j + 1) % F2_STAGES]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax / correction / output
    const int t = (warp - 4) >> 2;  // query tile of this warpgroup
    const int q4 = warp & 3;        // TMEM lane quarter
    const int r = q4 * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const uint32_t s_addr = lane_addr + Cfg::COL_S + t * 64;
    const uint32_t o_addr = lane_addr + Cfg::COL_O + t * HD;
    const uint32_t l_addr = lane_addr + Cfg::COL_L + t * 16;
    const float c1 = p.scale * F2_LOG2E;
    uint8_t* sP = smem + Cfg::OFF_P + t * 16384;
    float tw[RELPOS ? 64 : 1];
    uint32_t th_col = 0;

    if (RELPOS) {
      const int tok = m0 + t * 128 + r;
      const int qj = tok & 63;
      // bias_h for key row kh sits in T_h column (qi - a_t) + 63 - kh, with (qi - a_t) = r / 64 (warp-uniform)
      th_col = lane_addr + Cfg::COL_TH + t * 80 + 63 + (r >> 6);
      mbar_wait(t_full, 0);
      tc_fence_after();
      // T_w (this tile's 128 table columns live in the not-yet-used S/O columns): per-thread scatter
      // bias_w[kw] = T_w[qj + 63 - kw] through a private, XOR-swizzled 32-float smem row, half at a time.
      float* scr = reinterpret_cast<float*>(smem + Cfg::OFF_TAB) + ((t * 128 + r) << 5);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(lane_addr + t * 128 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kw = qj + 63 - (c * 32 + i) - half * 32;
            if (kw >= 0 && kw < 32) scr[((((kw >> 2) ^ (r & 7)) << 2) | (kw & 3))] = __uint_as_float(v[i]) * F2_LOG2E;
          }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 w = *reinterpret_cast<const float4*>(scr + ((g ^ (r & 7)) << 2));
          tw[half * 32 + 4 * g] = w.x; tw[half * 32 + 4 * g + 1] = w.y;
          tw[half * 32 + 4 * g + 2] = w.z; tw[half * 32 + 4 * g + 3] = w.w;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);  // S/O columns of this tile may now be overwritten
    }

    float m_ref = -INFINITY;
    for (int j = 0; j < nk; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      float bh = 0.0f;
      if (RELPOS) {
        bh = __uint_as_float(tmem_ld1(th_col - j)) * F2_LOG2E;
        tmem_ld_wait();
      }
      // ---- optimistic single pass against the current reference maximum
      float d = bh - m_ref;  // +inf on the first tile: that tile always takes the exact path below
      float ymax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(s_addr + c * 32, v);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float y = RELPOS ? fmaf(__uint_as_float(v[i]), c1, tw[RELPOS ? c * 32 + i : 0]) : __uint_as_float(v[i]) * c1;
          ymax = fmaxf(ymax, y);
          e[i] = ex2_approx(y + d);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(sP + sw128_offset(r, c * 4 + g)) =
              make_uint4(pack_bf16(e[8 * g], e[8 * g + 1]), pack_bf16(e[8 * g + 2], e[8 * g + 3]),
                         pack_bf16(e[8 * g + 4], e[8 * g + 5]), pack_bf16(e[8 * g + 6], e[8 * g + 7]));
      }
      const float m_tile = ymax + bh;
      const bool need = m_tile > m_ref + F2_TAU;  // also true while m_ref == -inf
      if (__any_sync(0xffffffffu, need)) {
        // ---- exact path: raise the reference, rescale O and l (warp-collective TMEM traffic), redo this tile
        const float m_new = need ? m_tile : m_ref;
        if (j > 0) {
          const float alpha = ex2_approx(m_ref - m_new);  // 1 for lanes that did not need it
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(o_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(o_addr + c * 32, v);
          }
          const float l_old = __uint_as_float(tmem_ld1(l_addr));
          tmem_ld_wait();
          tmem_st1(l_addr, __float_as_uint(l_old * alpha));
          tmem_st_wait();
        }
        m_ref = m_new;
        d = bh - m_ref;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(s_addr + c * 32, v);
          tmem_ld_wait();
          float e[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float y = RELPOS ? fmaf(__uint_as_float(v[i]), c1, tw[RELPOS ? c * 32 + i : 0]) : __uint_as_float(v[i]) * c1;
            e[i] = ex2_approx(y + d);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(sP + sw128_offset(r, c * 4 + g)) =
                make_uint4(pack_bf16(e[8 * g], e[8 * g + 1]), pack_bf16(e[8 * g + 2], e[8 * g + 3]),
                           pack_bf16(e[8 * g + 4], e[8 * g + 5]), pack_bf16(e[8 * g + 6], e[8 * g + 7]));
        }
      }
      tc_fence_before();
      fence_proxy_async();  // generic-proxy P writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // ---- epilogue: O / l
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const float l = __uint_as_float(tmem_ld1(l_addr));
    tmem_ld_wait();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* dst = p.out + (size_t)(b * p.Tq + m0 + t * 128 + r) * p.ldo + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(o_addr + c * 32, v);
      tmem_ld_wait();
      uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        d4[g] = make_uint4(pack_bf16(__uint_as_float(v[8 * g]) * inv_l, __uint_as_float(v[8 * g + 1]) * inv_l),
                           pack_bf16(__uint_as_float(v[8 * g + 2]) * inv_l, __uint_as_float(v[8 * g + 3]) * inv_l),
                           pack_bf16(__uint_as_float(v[8 * g + 4]) * inv_l, __uint_as_float(v[8 * g + 5]) * inv_l),
                           pack_bf16(__uint_as_float(v[8 * g + 6]) * inv_l, __uint_as_float(v[8 * g + 7]) * inv_l));
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int HD, bool RELPOS>
static int launch_flash2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                         const FlashParams& p, cudaStream_t st) {
  using Cfg = Flash2Cfg<HD, RELPOS>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(flash2_kernel<HD, RELPOS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) !=
        cudaSuccess)
      return WM_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid(p.Tq / 256, p.H, p.B);
  flash2_kernel<HD, RELPOS><<<grid, F2_THREADS, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, trel, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// q tiles: box 128 rows; k/v tiles: box 64 rows; rel table [256,64]: box 16 rows
int flash2_dispatch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                    const FlashParams& p, int hd, cudaStream_t st) {
  if (p.Tq % 256 != 0 || p.Tk % 64 != 0 || p.Tk < 64) return WM_ERR_SHAPE;
  if (p.use_relpos) {
    if (hd != 64 || p.Tq != 4096 || p.Tk != 4096) return WM_ERR_SHAPE;
    return launch_flash2<64, true>(tq, tk, tv, trel, p, st);
  }
  if (hd == 64) return launch_flash2<64, false>(tq, tk, tv, trel, p, st);
  if (hd == 128) return launch_flash2<128, false>(tq, tk, tv, trel, p, st);
  return WM_ERR_SHAPE;
}

}  // namespace wm
