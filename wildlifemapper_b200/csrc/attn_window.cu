// Fused 14x14 windowed attention with decomposed rel-pos bias on tcgen05 / TMEM (sm_100a).
//
// Reference: Block.forward window path, image_encoder.py:188-204 -- norm1, window_partition (zero-pad 64->70,
// :265-286), Attention.forward (:246-262) with add_decomposed_rel_pos (:347-383, tables [27,hd]),
// window_unpartition + crop (:289-311).  Partition / pad / unpartition / crop never touch HBM here: the
// window is a 4-D TMA box over the [B,64,64,3D] qkv tensor, the zero padding is TMA out-of-bounds fill,
// and only in-image query rows are written back, straight into [B,64,64,D].
//
// Pad keys: pad tokens are zeros after the norm, so in the reference their k and v
// equal the qkv bias and they DO take part in the softmax.  The host computes the qkv GEMM with the k and v
// biases dropped (a per-query constant shift of the logits cancels in softmax; sum(p)=1 moves b_v to the
// output, where it is folded into the proj bias), which makes the pad keys exactly the zero rows TMA fills
// in.  Their rel-pos bias is still added, as in the reference.
//
// One CTA = 7 query rows (112 queries, padded to the 128-row MMA) of one (image, window, head):
//   S[128 x 224] = Q K^T  (keys laid out 14 rows x 16 columns; columns 14,15 are masked)
//   T[128 x 64]  = Q [Rh;Rw]^T   -> per-thread bias_h[14], bias_w[14] registers
//   softmax in registers (exp2 domain), P -> swizzled smem (bf16), O[128 x 64] = P V (V MN-major).
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

// (First-generation kernel: kept as the A/B reference for head dim 64 -- "window_version" 1 -- and as the path for head
// dim 80 (ViT-H), whose operand tiles do not fit the double-buffered layout of attn_window2.cu.  The probabilities are
// written back into the consumed score columns of tensor memory and P V reads its A operand from TMEM.)
constexpr int WA_THREADS = 192;
constexpr int WA_T_LD = 65;
constexpr float WA_LOG2E = 1.4426950408889634f;

template <int HD>
struct WinCfg {
  static_assert(HD == 64 || HD == 80, "head dim");
  static constexpr int SUB = (HD + 63) / 64;        // 64-column sub-tiles (a partial second one is loaded 64 wide)
  static constexpr int KSTEPS = HD / 16;
  static constexpr int Q_SUB = 16384;                // 128 rows x 128 B (112 loaded)
  static constexpr int KV_SUB = 28672;               // 224 rows x 128 B
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + SUB * Q_SUB;
  static constexpr int OFF_V = OFF_K + SUB * KV_SUB;
  static constexpr int OFF_REL = OFF_V + SUB * KV_SUB;  // 64 table rows x 128 B per sub-tile
  static constexpr int OFF_T = OFF_REL + SUB * 8192;    // fp32 [128][65] scratch
  static constexpr int OFF_BAR = OFF_T + 128 * WA_T_LD * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static constexpr int COL_S = 0, COL_T = 256, COL_O = 320;  // P: bf16 pairs over S columns [0, 112)
};

template <int HD>
__global__ void __launch_bounds__(WA_THREADS, 1)
window_attn_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   const __grid_constant__ CUtensorMap tmap_rel, const WindowParams p) {
  using Cfg = WinCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int mt = blockIdx.x;                 // which 7-row half of the window
  const int win = blockIdx.y / p.H, h = blockIdx.y % p.H;
  const int wy = win / 5, wx = win % 5;
  const int b = blockIdx.z;

  // rows 112..127 of the Q sub-tiles are never loaded: zero them so the unused MMA rows stay finite
  for (int i = threadIdx.x; i < Cfg::SUB * ((16384 - 14336) / 16); i += WA_THREADS)
    reinterpret_cast<uint4*>(smem + Cfg::OFF_Q + (i / 128) * Cfg::Q_SUB + 14336)[i % 128] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_rel);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_qk, Cfg::SUB * (14336 + 28672 + 8192));
      for (int s = 0; s < Cfg::SUB; ++s) {
        tma_load_4d(smem + Cfg::OFF_Q + s * Cfg::Q_SUB, &tmap_q, bar_qk, h * HD + s * 64, wx * 14, wy * 14 + mt * 7, b);
        tma_load_4d(smem + Cfg::OFF_K + s * Cfg::KV_SUB, &tmap_kv, bar_qk, p.D + h * HD + s * 64, wx * 14, wy * 14, b);
        tma_load_2d(smem + Cfg::OFF_REL + s * 8192, &tmap_rel, bar_qk, s * 64, 0);
      }
      mbar_arrive_expect_tx(bar_v, Cfg::SUB * 28672);
      for (int s = 0; s < Cfg::SUB; ++s)
        tma_load_4d(smem + Cfg::OFF_V + s * Cfg::KV_SUB, &tmap_kv, bar_v, 2 * p.D + h * HD + s * 64, wx * 14, wy * 14, b);
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 224, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);  // A = P from TMEM, V MN-major
    const bool leader = elect_one();
    const uint32_t sq = smem_u32(smem + Cfg::OFF_Q), sk = smem_u32(smem + Cfg::OFF_K);
    const uint32_t sv = smem_u32(smem + Cfg::OFF_V), sr = smem_u32(smem + Cfg::OFF_REL);
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    if (leader) {
#pragma unroll
      for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
        const uint32_t ko = (ks & 3) * 32;
        umma_bf16(tmem_base + Cfg::COL_T, make_sdesc_sw128(sq + (ks >> 2) * Cfg::Q_SUB + ko, 16, 1024),
                  make_sdesc_sw128(sr + (ks >> 2) * 8192 + ko, 16, 1024), idesc_t, ks != 0);
      }
#pragma unroll
      for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
        const uint32_t ko = (ks & 3) * 32;
        umma_bf16(tmem_base + Cfg::COL_S, make_sdesc_sw128(sq + (ks >> 2) * Cfg::Q_SUB + ko, 16, 1024),
                  make_sdesc_sw128(sk + (ks >> 2) * Cfg::KV_SUB + ko, 16, 1024), idesc_s, ks != 0);
      }
      umma_commit(s_full);
    }
    __syncwarp();
    mbar_wait(bar_v, 0);
    mbar_wait(p_full, 0);
    tc_fence_after();
    if (leader) {
#pragma unroll
      for (int ks = 0; ks < 14; ++ks)  // 224 keys, 16 per MMA; P: 8 TMEM columns per step
        umma_bf16_ts(tmem_base + Cfg::COL_O, tmem_base + Cfg::COL_S + ks * 8,
                     make_sdesc_sw128(sv + ks * 2048, Cfg::KV_SUB, 1024), idesc_o, ks != 0);
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int yy = r >> 4, x = r & 15;
    const int y = mt * 7 + yy;
    const float c1 = p.scale * WA_LOG2E;
    float* sT = reinterpret_cast<float*>(smem + Cfg::OFF_T) + r * WA_T_LD;

    mbar_wait(s_full, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(lane_addr + Cfg::COL_T + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) sT[c * 32 + i] = __uint_as_float(v[i]) * WA_LOG2E;
    }
    float bh[14], bw[16];
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      bh[k] = sT[y - k + 13];       // Rh[(q row) - (key row) + 13]
      bw[k] = sT[32 + x - k + 13];  // Rw[(q col) - (key col) + 13]   (x<=15 -> index <= 28)
    }
    bw[14] = -INFINITY;  // key columns 14,15 of the 16-wide box belong to the neighbouring window: masked
    bw[15] = -INFINITY;

    float m_row = -INFINITY;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      uint32_t v[32];
      tmem_ld32(lane_addr + Cfg::COL_S + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        m_row = fmaxf(m_row, fmaf(__uint_as_float(v[i]), c1, bw[i & 15]) + bh[2 * c + (i >> 4)]);
    }
    float l_row = 0.0f;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      uint32_t v[32];
      tmem_ld32(lane_addr + Cfg::COL_S + c * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float e0 = exp2f(fmaf(__uint_as_float(v[2 * i]), c1, bw[(2 * i) & 15]) + (bh[2 * c + ((2 * i) >> 4)] - m_row));
        const float e1 = exp2f(fmaf(__uint_as_float(v[2 * i + 1]), c1, bw[(2 * i + 1) & 15]) + (bh[2 * c + ((2 * i + 1) >> 4)] - m_row));
        l_row += e0 + e1;
        pk[i] = pack_bf16(e0, e1);
      }
      // P chunk c overwrites S columns [16c, 16c+16): already consumed (chunk c/2 <= c, chunk 0 is in registers)
      tmem_st16(lane_addr + Cfg::COL_S + c * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_full);

    mbar_wait(o_full, 0);
    tc_fence_after();
    const int gy = wy * 14 + y, gx = wx * 14 + x;
    const bool valid = (r < 112) && (x < 14) && (gy < 64) && (gx < 64);
    const float inv_l = 1.0f / l_row;
    __nv_bfloat16* dst = p.out + ((size_t)(b * 64 + gy) * 64 + gx) * p.D + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t v[16];
      tmem_ld16(lane_addr + Cfg::COL_O + c * 16, v);
      tmem_ld_wait();
      if (valid) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + c * 16);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]) * inv_l;
          d4[g] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HD>
static int launch_window(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& trel, const WindowParams& p,
                         cudaStream_t st) {
  using Cfg = WinCfg<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(window_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) !=
        cudaSuccess)
      return WM_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid(2, 25 * p.H, p.B);
  window_attn_kernel<HD><<<grid, WA_THREADS, Cfg::SMEM_BYTES, st>>>(tq, tkv, trel, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// tq: box (64, 16, 7, 1) over qkv [B,64,64,3D]; tkv: box (64, 16, 14, 1); trel: box (64, 64) over the [64, hd] table
int window_dispatch(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& trel, const WindowParams& p, int hd,
                    cudaStream_t st) {
  if (hd == 64) return launch_window<64>(tq, tkv, trel, p, st);
  if (hd == 80) return launch_window<80>(tq, tkv, trel, p, st);
  return WM_ERR_SHAPE;
}

}  // namespace wm
