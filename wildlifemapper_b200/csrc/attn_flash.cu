// Fused flash-style attention on tcgen05 / TMEM for sm_100a (global 64x64 attention and the HFC cross-attention).
//
//   O = softmax( scale * Q K^T  [+ decomposed rel-pos bias from the UNSCALED q] ) V      per (image, head)
//
// Reference sites: encoder global blocks, image_encoder.py:246-262 with add_decomposed_rel_pos :347-383
// (bias[q,k] = q.Rh[qi-ki+63] + q.Rw[qj-kj+63], tables [127,hd]); HFC cross-attention :500-503
// (nn.MultiheadAttention, 8 heads x 128, q scaled by 1/sqrt(128), no bias).  The [h,4096,4096] score tensor
// the reference materialises never leaves the SM.
//
// One CTA = 128 query rows of one (image, head); key tiles of 128 stream through a TMA ring.
//   warp 0     TMA producer (Q once, K/V ring, rel-pos tables)
//   warp 1     tcgen05.mma issuer: S = Q K^T (128x128xHD) into a double-buffered TMEM tile, then
//              PV = P V (128xHDx128) with P staged as bf16 in swizzled smem and V as an MN-major operand
//   warps 2-5  softmax: thread = one query row (tcgen05.ld 32x32b), online softmax in fp32 (exp2 domain),
//              O accumulated in registers, rescaled by exp2(m_old - m_new)
// Rel-pos: T_h = Q Rh^T and T_w = Q Rw^T are two extra 128x128xHD MMAs per CTA; each thread scatters its row
// into smem as bias_h[key row], bias_w[key col] (pre-multiplied by log2 e) and adds them inside the softmax.
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

constexpr int FA_THREADS = 192;
constexpr int FA_KV_STAGES = 2;
constexpr int FA_BIAS_LD = 68;  // padded fp32 row (conflict-free float4 reads, one row per thread)
constexpr float LOG2E = 1.4426950408889634f;

template <int HD, bool RELPOS>
struct FlashCfg {
  static constexpr int SUB = HD / 64;                 // 64-column (128-byte) sub-tiles per operand tile
  static constexpr int TILE_BYTES = SUB * 128 * 128;  // 128 rows x HD bf16
  static constexpr int P_BYTES = 2 * 128 * 128;       // 128 x 128 bf16 (two sub-tiles)
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + TILE_BYTES;
  static constexpr int OFF_V = OFF_K + FA_KV_STAGES * TILE_BYTES;
  static constexpr int OFF_P = OFF_V + FA_KV_STAGES * TILE_BYTES;  // rel-pos tables alias P during the prologue
  static constexpr int OFF_BIAS = OFF_P + P_BYTES;
  static constexpr int BIAS_BYTES = RELPOS ? 2 * 128 * FA_BIAS_LD * 4 : 0;
  static constexpr int OFF_BAR = OFF_BIAS + BIAS_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int COL_S = 0, COL_PV = 256;
};

template <int HD, bool RELPOS>
__global__ void __launch_bounds__(FA_THREADS, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_rel,
                  const FlashParams p) {
  using Cfg = FlashCfg<HD, RELPOS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* k_full = bars + 1;                  // [2]
  uint64_t* k_empty = bars + 3;                 // [2]
  uint64_t* v_full = bars + 5;                  // [2]
  uint64_t* v_empty = bars + 7;                 // [2]
  uint64_t* s_full = bars + 9;                  // [2]
  uint64_t* s_empty = bars + 11;                // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* pv_full = bars + 14;
  uint64_t* t_full = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nk = p.Tk / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(pv_full, 1);
    mbar_init(t_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int rel_bytes = RELPOS ? 256 * 128 : 0;
      mbar_arrive_expect_tx(bar_q, Cfg::TILE_BYTES + rel_bytes);
#pragma unroll
      for (int s = 0; s < Cfg::SUB; ++s)
        tma_load_2d(smem + Cfg::OFF_Q + s * 16384, &tmap_q, bar_q, p.q_col0 + h * HD + s * 64, b * p.Tq + m0);
      if (RELPOS) tma_load_2d(smem + Cfg::OFF_P, &tmap_rel, bar_q, 0, 0);
      for (int j = 0; j < nk; ++j) {
        const int st = j & 1;
        const uint32_t par = ((j >> 1) & 1) ^ 1;
        mbar_wait(&k_empty[st], par);
        mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int s = 0; s < Cfg::SUB; ++s)
          tma_load_2d(smem + Cfg::OFF_K + st * Cfg::TILE_BYTES + s * 16384, &tmap_k, &k_full[st],
                      p.k_col0 + h * HD + s * 64, b * p.Tk + j * 128);
        mbar_wait(&v_empty[st], par);
        mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
#pragma unroll
        for (int s = 0; s < Cfg::SUB; ++s)
          tma_load_2d(smem + Cfg::OFF_V + st * Cfg::TILE_BYTES + s * 16384, &tmap_v, &v_full[st],
                      p.v_col0 + h * HD + s * 64, b * p.Tk + j * 128);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);  // B (=V) is MN-major
      const uint32_t sq = smem_u32(smem + Cfg::OFF_Q);
      const uint32_t sp = smem_u32(smem + Cfg::OFF_P);
      uint32_t s_uses[2] = {0, 0};
      mbar_wait(bar_q, 0);
      tc_fence_after();
      if (RELPOS) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint64_t ad = make_sdesc_sw128(sq + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
            const uint64_t bd = make_sdesc_sw128(sp + t * 16384 + (ks & 3) * 32, 16, 1024);
            umma_bf16(tmem_base + Cfg::COL_S + t * 128, ad, bd, idesc_s, ks != 0);
          }
          s_uses[t] = 1;
        }
        umma_commit(t_full);
      }
      auto issue_s = [&](int j) {
        const int st = j & 1, bsel = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        mbar_wait(&s_empty[bsel], (s_uses[bsel] & 1) ^ 1);
        s_uses[bsel]++;
        tc_fence_after();
        const uint32_t sk = smem_u32(smem + Cfg::OFF_K + st * Cfg::TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_bf16(tmem_base + Cfg::COL_S + bsel * 128, make_sdesc_sw128(sq + off, 16, 1024),
                    make_sdesc_sw128(sk + off, 16, 1024), idesc_s, ks != 0);
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[bsel]);
      };
      issue_s(0);
      for (int j = 0; j < nk; ++j) {
        if (j + 1 < nk) issue_s(j + 1);
        const int st = j & 1;
        mbar_wait(&v_full[st], (j >> 1) & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t sv = smem_u32(smem + Cfg::OFF_V + st * Cfg::TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // 128 keys, 16 per MMA
          const uint64_t ad = make_sdesc_sw128(sp + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
          // V tile: rows = keys (128 B each = 64 head dims); 16 keys = 2048 B; next 64 head dims at +16384
          const uint64_t bd = make_sdesc_sw128(sv + ks * 2048, 16384, 1024);
          umma_bf16(tmem_base + Cfg::COL_PV, ad, bd, idesc_pv, ks != 0);
        }
        umma_commit(&v_empty[st]);
        umma_commit(pv_full);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / output (one query row per thread)
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const float c1 = p.scale * LOG2E;
    float* sTH = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS) + r * FA_BIAS_LD;
    float* sTW = sTH + 128 * FA_BIAS_LD;
    uint8_t* sP = smem + Cfg::OFF_P;

    if (RELPOS) {
      const int tok = m0 + r;
      const int qi = tok >> 6, qj = tok & 63;
      mbar_wait(t_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        float* dst = t == 0 ? sTH : sTW;
        const int qc = (t == 0 ? qi : qj) + 63;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(lane_addr + Cfg::COL_S + t * 128 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kk = qc - (c * 32 + i);  // key row (t=0) / key column (t=1) this table entry belongs to
            if (kk >= 0 && kk < 64) dst[kk] = __uint_as_float(v[i]) * LOG2E;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[0]);
        mbar_arrive(&s_empty[1]);
      }
    }

    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f;

    for (int j = 0; j < nk; ++j) {
      const int bsel = j & 1;
      const uint32_t s_addr = lane_addr + Cfg::COL_S + bsel * 128;
      mbar_wait(&s_full[bsel], (j >> 1) & 1);
      tc_fence_after();
      // ---- pass 1: tile row maximum (log2 domain)
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(s_addr + c * 32, v);
        tmem_ld_wait();
        float mx = -INFINITY;
        if (RELPOS) {
          const float4* tw = reinterpret_cast<const float4*>(sTW + (c & 1) * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w = tw[i];
            mx = fmaxf(mx, fmaf(__uint_as_float(v[4 * i]), c1, w.x));
            mx = fmaxf(mx, fmaf(__uint_as_float(v[4 * i + 1]), c1, w.y));
            mx = fmaxf(mx, fmaf(__uint_as_float(v[4 * i + 2]), c1, w.z));
            mx = fmaxf(mx, fmaf(__uint_as_float(v[4 * i + 3]), c1, w.w));
          }
          mx += sTH[2 * j + (c >> 1)];
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]) * c1);
        }
        m_tile = fmaxf(m_tile, mx);
      }
      const float m_new = fmaxf(m_run, m_tile);
      const float alpha = exp2f(m_run - m_new);  // 0 on the first tile (m_run = -inf)
      m_run = m_new;
      // ---- fold in the previous tile's P V (its MMA ran while pass 1 executed), then rescale
      if (j > 0) {
        mbar_wait(pv_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(lane_addr + Cfg::COL_PV + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(v[i]);
        }
        tc_fence_before();
      }
#pragma unroll
      for (int i = 0; i < HD; ++i) o[i] *= alpha;
      l_run *= alpha;
      // ---- pass 2: probabilities -> bf16 P tile in swizzled smem (A operand of the P V MMA)
      float l_tile = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(s_addr + c * 32, v);
        tmem_ld_wait();
        float pr[32];
        if (RELPOS) {
          const float4* tw = reinterpret_cast<const float4*>(sTW + (c & 1) * 32);
          const float sh = sTH[2 * j + (c >> 1)] - m_new;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w = tw[i];
            pr[4 * i] = exp2f(fmaf(__uint_as_float(v[4 * i]), c1, w.x) + sh);
            pr[4 * i + 1] = exp2f(fmaf(__uint_as_float(v[4 * i + 1]), c1, w.y) + sh);
            pr[4 * i + 2] = exp2f(fmaf(__uint_as_float(v[4 * i + 2]), c1, w.z) + sh);
            pr[4 * i + 3] = exp2f(fmaf(__uint_as_float(v[4 * i + 3]), c1, w.w) + sh);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) pr[i] = exp2f(fmaf(__uint_as_float(v[i]), c1, -m_new));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) l_tile += pr[i];
        uint8_t* sub = sP + (c >> 1) * 16384;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 pk = make_uint4(pack_bf16(pr[8 * g], pr[8 * g + 1]), pack_bf16(pr[8 * g + 2], pr[8 * g + 3]),
                                      pack_bf16(pr[8 * g + 4], pr[8 * g + 5]), pack_bf16(pr[8 * g + 6], pr[8 * g + 7]));
          *reinterpret_cast<uint4*>(sub + sw128_offset(r, (c & 1) * 4 + g)) = pk;
        }
      }
      l_run += l_tile;
      tc_fence_before();
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(p_full);
        mbar_arrive(&s_empty[bsel]);
      }
    }
    // ---- last tile's P V, normalise, store
    mbar_wait(pv_full, (nk - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* dst = p.out + (size_t)(b * p.Tq + m0 + r) * p.ldo + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(lane_addr + Cfg::COL_PV + c * 32, v);
      tmem_ld_wait();
      uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = (o[c * 32 + g * 8 + i] + __uint_as_float(v[g * 8 + i])) * inv_l;
        d4[g] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int HD, bool RELPOS>
static int launch_flash(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                        const FlashParams& p, cudaStream_t st) {
  using Cfg = FlashCfg<HD, RELPOS>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(flash_attn_kernel<HD, RELPOS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::SMEM_BYTES) != cudaSuccess)
      return WM_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid(p.Tq / 128, p.H, p.B);
  flash_attn_kernel<HD, RELPOS><<<grid, FA_THREADS, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, trel, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

int flash_dispatch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                   const FlashParams& p, int hd, cudaStream_t st) {
  if (p.Tq % 128 != 0 || p.Tk % 128 != 0 || p.Tk < 128) return WM_ERR_SHAPE;
  if (p.use_relpos) {
    if (hd != 64 || p.Tq != 4096 || p.Tk != 4096) return WM_ERR_SHAPE;
    return launch_flash<64, true>(tq, tk, tv, trel, p, st);
  }
  if (hd == 64) return launch_flash<64, false>(tq, tk, tv, trel, p, st);
  if (hd == 128) return launch_flash<128, false>(tq, tk, tv, trel, p, st);
  return WM_ERR_SHAPE;
}

}  // namespace wm
