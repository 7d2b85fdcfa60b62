// Flash attention v4 on tcgen05 / TMEM (sm_100a): global 64x64 attention with decomposed rel-pos bias and the HFC
// cross-attention.  Same math and reference sites as attn_flash.cu (image_encoder.py:246-262, 347-383, 500-503) and the
// same building blocks as v3 (attn_flash3.cu): two 128-query tiles per CTA, P kept in tensor memory (tcgen05.mma "ts"
// form), O accumulated in TMEM with lazy rescaling, bias_w in registers and T_h resident in TMEM, setmaxnreg split,
// warp-uniform TMA / MMA issue.  What changed, from the v3 event traces (profiles/r01g_flash3_trace_*.txt): in v3 the
// next score tile S_t(j+1) aliases P_t(j) and can only be issued after P_t(j) V, so every softmax warpgroup sat idle
// for ~1500 of every ~4500 cycles waiting for its next scores, and the MUFU ran at ~45 %.  Here
//
//   * the softmax works in steps of 64 keys (ONE key row of the 64x64 grid: bias_h is a single scalar per step) and
//     each query tile owns TWO 64-column score buffers: the MMA warp runs two steps ahead (P_t(j) V, then S_t(j+2)
//     into the buffer step j just released), so in steady state the scores of the next step are already waiting
//     when a warpgroup finishes a step, and the MUFU always has two warps per scheduler to choose from;
//   * K and V still arrive as 128-key TMA tiles; the two halves of a stage are addressed by descriptor offset.
//
//   warp 0       TMA producer (Q tiles, tables, K and V rings)
//   warp 1       tcgen05.mma issuer          warp 2  TMEM allocator
//   warps 4-7    softmax of query tile 0     warps 8-11  softmax of query tile 1 (thread = one query row)
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

#ifdef WM_F3_TRACE
// diagnostics build: SM-clock event trace of CTA (0,0,0) ([role][step][event]; see profiles/flash4_trace.py)
__device__ unsigned long long g_f4_trace[3][64][8];
#define F4_TRACE(role, j, ev)                                                                   \
  do {                                                                                          \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 64) g_f4_trace[role][j][ev] = clock64(); \
  } while (0)
#else
#define F4_TRACE(role, j, ev) do { } while (0)
#endif

// ablation builds (profiles/run_flash_ablation.sh; results are wrong on purpose): -DWM_F4_NO_EX2 replaces the exp2 of the
// hot loop by one FMA-pipe operation, -DWM_F4_ONE_MMA issues one MMA per group, -DWM_F4_SKELETON keeps only the barrier protocol
#ifdef WM_F4_ONE_MMA   // one tcgen05.mma per score tile / P V group instead of four
#define F4_MMA_STEPS 1
#else
#define F4_MMA_STEPS 4
#endif
#ifdef WM_F4_NO_EX2
#define F4_EX2(x) ((x) * 0.0009765625f + 1.0f)
#else
#define F4_EX2(x) ex2_approx(x)
#endif

// exp2 on the FMA / ALU pipes for arguments x <= TAU: 2^x = 2^round(x) * p(x - round(x)), cubic minimax on [-0.5, 0.5]
// (relative error ~1e-4, below the bf16 rounding of P).  WM_F4_POLY = n sends every n-th PAIR of a 32-score chunk this way
// (FA4-style MUFU relief).  Measured per launch at batch 32, hd 64 + rel-pos / hd 128 (profiles/flash_time.py, in-run A/B):
// off 2.827 / 1.728 ms, n = 8: 2.770 / 1.664, n = 6: 2.730 / 1.656, n = 5: 2.929 / 1.717, n = 4: 2.775 / 1.688,
// n = 3: 2.984 / 1.788, n = 2: 2.987 / 1.826 -- the schedule ptxas finds matters as much as the fraction.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -120.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: round to nearest integer in the mantissa
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#ifndef WM_F4_POLY
#define WM_F4_POLY 6
#endif

// one stage of the lane-dependent down-shift of x[0 .. 64 + SH - 1): lanes with bit SH set take x[i + SH]
template <int SH>
__device__ __forceinline__ void f4_barrel_stage(uint32_t (&x)[96], int lane) {
  const bool on = (lane & SH) != 0;
#pragma unroll
  for (int i = 0; i < 64 + SH - 1; ++i) x[i] = on ? x[i + SH] : x[i];  // ascending i: x[i + SH] is still the old value
}

constexpr int F4_THREADS = 384;
constexpr float F4_LOG2E = 1.4426950408889634f;
constexpr float F4_TAU = 16.0f;  // lazy-rescale threshold (log2 units): p <= 2^16 between rescales (bf16 P and the fp32 accumulators have the range)

template <int HD, bool RELPOS>
struct Flash4Cfg {
  static_assert(HD % 16 == 0 && HD >= 64 && HD <= 128, "head dim");
  static_assert(!RELPOS || HD <= 96, "rel-pos variant: tensor-memory budget");
  static constexpr int SUB = (HD + 63) / 64;       // 64-column (128-byte) sub-tiles per operand row (a partial second
                                                   // sub-tile -- head dim 80 -- is loaded 64 wide; only HD columns are used)
  static constexpr int KSTEPS = HD / 16;           // K = 16 MMA steps over the head dim
  static constexpr int STAGES = (SUB == 1) ? 4 : 2;
  static constexpr int TILE_BYTES = SUB * 16384;   // 128 rows (queries or keys) x SUB x 128 B
  static constexpr int OFF_Q = 0;                  // 2 query tiles
  static constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;
  static constexpr int OFF_V = OFF_K + STAGES * TILE_BYTES;
  // RELPOS prologue: Rw [128 rows] + 2 x Rh slice [80 rows] (x SUB sub-tiles).  Head dim 64: own region.  Larger head dims: the tables ALIAS K stage 1 and the V stages, whose
  // first loads wait for the prologue (t_done).
  static constexpr bool TAB_ALIAS = RELPOS && SUB > 1;
  static constexpr int TAB_BYTES = RELPOS ? SUB * (16384 + 2 * 10240) : 0;
  static constexpr int OFF_TAB = TAB_ALIAS ? OFF_K + TILE_BYTES : OFF_V + STAGES * TILE_BYTES;
  static_assert(!TAB_ALIAS || TAB_BYTES <= (2 * STAGES - 1) * TILE_BYTES, "aliased table region");
  static constexpr int OFF_BAR = OFF_V + STAGES * TILE_BYTES + (TAB_ALIAS ? 0 : TAB_BYTES);
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  // TMEM columns: S_t buffer b at 128 t + 64 b (P = bf16 pairs in its first 32 columns), O_t at 256 + HD t, then T_h:
  //   head dim 64: fp32, 64 columns per tile (+ column 64 in a register);  larger: fp16 pairs, 33 of 40 columns per tile
  static constexpr int COL_O = 256;
  static constexpr int COL_TH = COL_O + 2 * HD;
  static constexpr bool TH_PACKED = HD > 64;
  static constexpr int TH_STRIDE = TH_PACKED ? 40 : 64;
  static constexpr int TMEM_COLS = 512;
  static_assert(COL_TH + (RELPOS ? 2 * TH_STRIDE : 0) <= 512, "TMEM budget");
};

template <int HD, bool RELPOS>
__global__ void __launch_bounds__(F4_THREADS, 1)
flash4_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_rel,
              const FlashParams p) {
  using Cfg = Flash4Cfg<HD, RELPOS>;
  constexpr int NS = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [4]
  uint64_t* k_empty = bars + 5;   // [4]
  uint64_t* v_full = bars + 9;    // [4]
  uint64_t* v_empty = bars + 13;  // [4]
  uint64_t* s_full = bars + 17;   // [2 tiles][2 buffers]  S_t(j) complete
  uint64_t* p_full = bars + 21;   // [2 tiles][2 buffers]  P_t(j) stored to TMEM by the 4 warps of tile t
  uint64_t* pv_done = bars + 25;  // [2]  P_t(ns-2) V, P_t(ns-1) V complete (only waited on by the rare rescale path)
  uint64_t* o_full = bars + 27;   // [2]
  uint64_t* t_full = bars + 29;   // rel-pos table products complete
  uint64_t* t_done = bars + 30;   // ... and drained out of the S / O columns by the 8 softmax warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  // Tq and Tk need not be multiples of the tile sizes (decoder attention, 900 queries): query rows >= Tq are computed
  // from whatever the TMA box finds there and never stored; key columns >= Tk are masked to -inf in the softmax (their
  // V rows are multiplied by exact zeros).
  const int nk = (p.Tk + 127) / 128;  // 128-key TMA tiles
  const int ns = (p.Tk + 63) / 64;    // 64-key softmax / MMA steps

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);  // released by both MMA warps
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1);
    }
    mbar_init(t_full, 1);
    mbar_init(t_done, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<40>();
    // Producer and MMA roles: the WHOLE warp runs the control flow (uniform branches, every lane polls the
    // barriers); one elected lane issues the TMA / tcgen05 instructions.
    // Main loop of one MMA issuer warp for query tile t (warp 1: tile 0, warp 3: tile 1).  One issuer per tile: a single
    // warp walking both tiles spent ~2000 of the 2700 cycles of a 64-key step in its own control flow -- every mbarrier
    // wait costs ~100 cycles even when the phase has long completed, every tcgen05.mma / commit is issued by one lane
    // that shares its scheduler with two softmax warps -- and saw each tile's P 3000 cycles after it had been stored,
    // while the tensor pipe needs 52 cycles per 128 x 64 x 16 MMA (profiles/r01z_flash4_trace_before.txt,
    // profiles/micro/umma_chain_bench.cu).
    auto mma_main = [&](const int t, const bool leader) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);  // A = P from TMEM (K-major), V is MN-major
      const uint32_t sq = smem_u32(smem + Cfg::OFF_Q);
      // S_t(step) = Q_t K(step)^T (128 x 64 x HD) into score buffer (step & 1) of tile t; K rows of step: half (step & 1)
      // of stage (step / 2) % NS.
      auto issue_s = [&](int step) {
        if (leader) {
          const int kst = (step >> 1) % NS;
          const uint64_t qd = make_sdesc_sw128(sq + t * Cfg::TILE_BYTES, 16, 1024);
          const uint64_t kd = make_sdesc_sw128(smem_u32(smem + Cfg::OFF_K + kst * Cfg::TILE_BYTES) + (step & 1) * 8192, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < (F4_MMA_STEPS == 4 ? Cfg::KSTEPS : 1); ++ks) {
            const uint32_t off = (uint32_t)((ks >> 2) * (16384 >> 4) + (ks & 3) * 2);  // address field is bytes >> 4
            umma_bf16(tmem_base + t * 128 + (step & 1) * 64, qd + off, kd + off, idesc_s, ks != 0);
          }
          umma_commit(&s_full[t * 2 + (step & 1)]);
        }
        __syncwarp();
      };
      // two steps ahead of the softmax (both from K tile 0)
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      issue_s(1);
      if (leader) umma_commit(&k_empty[0]);
      __syncwarp();
      for (int j = 0; j < ns; ++j) {
        const int vst = (j >> 1) % NS;
        const uint32_t vph = (uint32_t)((j >> 1) / NS) & 1u;
        const bool more = j + 2 < ns;
        const uint32_t sv = smem_u32(smem + Cfg::OFF_V + vst * Cfg::TILE_BYTES) + (j & 1) * 8192;
        if (leader) F4_TRACE(2, j, 4 * t);
        mbar_wait(&p_full[t * 2 + (j & 1)], (uint32_t)(j >> 1) & 1u);
        if (leader) F4_TRACE(2, j, 4 * t + 1);
        if ((j & 1) == 0) {
          mbar_wait(&v_full[vst], vph);
          if (more) {  // step j + 2 opens K tile (j + 2) / 2
            const int kt = (j + 2) >> 1;
            mbar_wait(&k_full[kt % NS], (uint32_t)(kt / NS) & 1u);
          }
        }
        tc_fence_after();
        if (leader) {
          const uint64_t vd = make_sdesc_sw128(sv, 16384, 1024);
#pragma unroll
          for (int ks = 0; ks < F4_MMA_STEPS; ++ks)  // 64 keys, 16 per MMA; P: 8 TMEM columns per step; V: 2048 B per step
            umma_bf16_ts(tmem_base + Cfg::COL_O + t * HD, tmem_base + t * 128 + (j & 1) * 64 + ks * 8,
                         vd + (uint32_t)(ks * (2048 >> 4)), idesc_pv, (j | ks) != 0);
          // P_t(j) V complete: only the rare rescale path of step j + 1 needs it; while score tiles are still being
          // issued behind the P V groups, the commit of S_t(j + 2) covers it (the pipe completes in issue order), so the
          // explicit commit (~44 cycles of tensor-pipe issue time, profiles/micro/umma_chain_bench.cu) is only paid by the
          // last two steps.
          if (!more) umma_commit(&pv_done[t]);
          if (j == ns - 1) umma_commit(&o_full[t]);
        }
        __syncwarp();
        if (leader) F4_TRACE(2, j, 4 * t + 2);
        if (more) issue_s(j + 2);  // reuses the score buffer step j just released (behind its P V in the pipe)
        if (leader) F4_TRACE(2, j, 4 * t + 3);
        if (leader) {
          if (j & 1) umma_commit(&v_empty[vst]);                            // both halves of the V tile consumed by this tile
          if (more && (j & 1)) umma_commit(&k_empty[((j + 2) >> 1) % NS]);  // both halves of K tile (j + 2) / 2 issued
        }
        __syncwarp();
      }
    };
    if (warp == 0) {
      // ------------------------------------------------------------ TMA producer
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * Cfg::TILE_BYTES + Cfg::TAB_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int s = 0; s < Cfg::SUB; ++s)
            tma_load_2d(smem + Cfg::OFF_Q + t * Cfg::TILE_BYTES + s * 16384, &tmap_q, q_full, p.q_col0 + h * HD + s * 64,
                        b * p.Tq + m0 + t * 128);
        if (RELPOS) {
          // table tensor [256,HD]: rows 0..126 rel_pos_h, 128..254 rel_pos_w.  Box = 64 columns x 16 rows.
          const int qi0 = m0 >> 6;  // first image row of this CTA (4 rows: 2 per query tile)
          for (int sb = 0; sb < Cfg::SUB; ++sb) {
            for (int i = 0; i < 8; ++i)
              tma_load_2d(smem + Cfg::OFF_TAB + sb * 16384 + i * 2048, &tmap_rel, q_full, sb * 64, 128 + 16 * i);
            for (int t = 0; t < 2; ++t)
              for (int i = 0; i < 5; ++i)
                tma_load_2d(smem + Cfg::OFF_TAB + Cfg::SUB * 16384 + (t * Cfg::SUB + sb) * 10240 + i * 2048, &tmap_rel, q_full,
                            sb * 64, qi0 + 2 * t + 16 * i);
          }
        }
      }
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nk; ++j) {
        if (Cfg::TAB_ALIAS && j == 1) mbar_wait(t_done, 0);  // K stage 1 and the V stages hold the tables until then
        mbar_wait(&k_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SUB; ++s)
            tma_load_2d(smem + Cfg::OFF_K + st * Cfg::TILE_BYTES + s * 16384, &tmap_k, &k_full[st],
                        p.k_col0 + h * HD + s * 64, b * p.Tk + j * 128);
        }
        if (Cfg::TAB_ALIAS && j == 0) mbar_wait(t_done, 0);
        mbar_wait(&v_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SUB; ++s)
            tma_load_2d(smem + Cfg::OFF_V + st * Cfg::TILE_BYTES + s * 16384, &tmap_v, &v_full[st],
                        p.v_col0 + h * HD + s * 64, b * p.Tk + j * 128);
        }
        __syncwarp();
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------ MMA issuer
      const bool leader = elect_one();  // the same lane issues every tcgen05.mma / tcgen05.commit
      const uint32_t sq = smem_u32(smem + Cfg::OFF_Q);
      mbar_wait(q_full, 0);
      tc_fence_after();
      if (RELPOS) {
        constexpr uint32_t idesc_tw = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_th = make_idesc_bf16(128, Cfg::TH_PACKED ? 80 : 64, 0, 0);
        constexpr uint32_t idesc_tx = make_idesc_bf16(128, 16, 0, 0);
        const uint32_t stab = smem_u32(smem + Cfg::OFF_TAB);
        // T_w(t) -> S_t columns.  Head dim 64: T_h(t)[0..63] -> resident columns, T_h(t)[64..79] -> scratch in the O
        // columns.  Larger head dims: T_h(t)[0..79] -> O_t columns (scratch; the softmax warps repack it as fp16 pairs).
        if (leader) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
#pragma unroll
            for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
              const uint32_t koff = (ks & 3) * 32;
              const uint32_t srh = stab + Cfg::SUB * 16384 + (t * Cfg::SUB + (ks >> 2)) * 10240 + koff;
              const uint64_t ad = make_sdesc_sw128(sq + t * Cfg::TILE_BYTES + (ks >> 2) * 16384 + koff, 16, 1024);
              umma_bf16(tmem_base + t * 128, ad, make_sdesc_sw128(stab + (ks >> 2) * 16384 + koff, 16, 1024), idesc_tw, ks != 0);
              if (Cfg::TH_PACKED) {
                umma_bf16(tmem_base + Cfg::COL_O + t * HD, ad, make_sdesc_sw128(srh, 16, 1024), idesc_th, ks != 0);
              } else {
                umma_bf16(tmem_base + Cfg::COL_TH + t * 64, ad, make_sdesc_sw128(srh, 16, 1024), idesc_th, ks != 0);
                umma_bf16(tmem_base + Cfg::COL_O + t * 16, ad, make_sdesc_sw128(srh + 8192, 16, 1024), idesc_tx, ks != 0);
              }
            }
          }
          umma_commit(t_full);
        }
        __syncwarp();
        mbar_wait(t_done, 0);  // both warpgroups have drained the scratch out of the S / O columns
        tc_fence_after();
      }
      mma_main(0, leader);
    } else if (warp == 3) {
      // ------------------------------------------------------------ second MMA issuer (query tile 1)
      const bool leader = elect_one();
      mbar_wait(q_full, 0);
      if (RELPOS) mbar_wait(t_done, 0);  // table products (issued by warp 1) drained out of the S / O columns
      tc_fence_after();
      mma_main(1, leader);
    }
  } else {
    // ------------------------------------------------------------ softmax / correction / output
    setmaxnreg_inc<232>();
    const int t = (warp - 4) >> 2;  // query tile of this warpgroup
    const int q4 = warp & 3;        // TMEM lane quarter
    const int r = q4 * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const uint32_t s_addr = lane_addr + t * 128;
    const uint32_t o_addr = lane_addr + Cfg::COL_O + t * HD;
    const float c1 = p.scale * F4_LOG2E;
    float tw[RELPOS ? 64 : 1];
    float bh64 = 0.0f;                                            // T_h column 64 (only needed by key row 0)
    const int hi = r >> 6;                                        // image row of this query inside the tile (warp-uniform)
    const uint32_t th_addr = lane_addr + Cfg::COL_TH + t * Cfg::TH_STRIDE;  // bias_h[kh] = T_h[hi + 63 - kh]

    if (RELPOS) {
      const int qj = (m0 + t * 128 + r) & 63;
      mbar_wait(t_full, 0);
      tc_fence_after();
      if (Cfg::TH_PACKED) {
        // T_h(t)[0..79] sits in the O_t columns as fp32: keep [0..64] as fp16 pairs (x log2 e) in the resident columns
        uint32_t a[32], bq[32], cq[16];
        tmem_ld32(o_addr, a);
        tmem_ld32(o_addr + 32, bq);
        tmem_ld16(o_addr + 64, cq);
        tmem_ld_wait();
        uint32_t pkh[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pkh[i] = pack_f16(__uint_as_float(a[2 * i]) * F4_LOG2E, __uint_as_float(a[2 * i + 1]) * F4_LOG2E);
          pkh[16 + i] = pack_f16(__uint_as_float(bq[2 * i]) * F4_LOG2E, __uint_as_float(bq[2 * i + 1]) * F4_LOG2E);
        }
        tmem_st32(th_addr, pkh);
        tmem_st1(th_addr + 32, pack_f16(__uint_as_float(cq[0]) * F4_LOG2E, 0.0f));
        tmem_st_wait();
      } else {
        const uint32_t x = tmem_ld1(lane_addr + Cfg::COL_O + t * 16);
        tmem_ld_wait();
        bh64 = __uint_as_float(x) * F4_LOG2E;
      }
      // bias_w[kw] = T_w[qj + 63 - kw], qj = (q4 & 1) * 32 + lane: every lane needs a 64-column window of its T_w row that
      // starts at a lane-dependent column.  Load the 96 columns [base, base + 96) (base = 32 (q4 & 1), warp-uniform) and
      // shift them down by `lane` positions with a barrel of selects (5 stages, 346 SEL).  (Round 1 scattered through a
      // shared-memory scratch with ~10 instructions per examined column: 10 us of every CTA's 67, attn_flash7.cu.)
      {
        (void)qj;
        const uint32_t base = (uint32_t)(q4 & 1) * 32u;
        uint32_t x[96];
        {
          uint32_t x0[32], x1[32], x2[32];
          tmem_ld32(s_addr + base, x0);
          tmem_ld32(s_addr + base + 32, x1);
          tmem_ld32(s_addr + base + 64, x2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) { x[i] = x0[i]; x[32 + i] = x1[i]; x[64 + i] = x2[i]; }
        }
        f4_barrel_stage<16>(x, lane);
        f4_barrel_stage<8>(x, lane);
        f4_barrel_stage<4>(x, lane);
        f4_barrel_stage<2>(x, lane);
        f4_barrel_stage<1>(x, lane);
#pragma unroll
        for (int kw = 0; kw < 64; ++kw) tw[RELPOS ? kw : 0] = __uint_as_float(x[63 - kw]) * F4_LOG2E;  // x[i] = T_w[qj + i]
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_done);
    }

    float m_ref = -INFINITY, l_run = 0.0f;
    for (int j = 0; j < ns; ++j) {  // one step = 64 keys = key row j of the 64x64 grid
      const uint32_t sj = s_addr + (j & 1) * 64;  // this step's score buffer; P = bf16 pairs over its first 32 columns
      if (q4 == 0 && lane == 0) F4_TRACE(t, j, 0);
      mbar_wait(&s_full[t * 2 + (j & 1)], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      if (q4 == 0 && lane == 0) F4_TRACE(t, j, 1);
      // One pass over the 64 scores of this row in two 32-column chunks; the TMEM load of chunk 1 is in flight while
      // chunk 0 is processed.  Each chunk is first evaluated OPTIMISTICALLY against the current reference maximum;
      // only if some row of the warp exceeds it by more than 2^TAU is the reference raised (O, l and the chunk of P
      // already written are rescaled) and the chunk recomputed from the registers that still hold it.
#ifdef WM_F4_SKELETON  // ablation: barrier / MMA skeleton only -- no score loads, no softmax arithmetic, no P stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t * 2 + (j & 1)]);
      continue;
#endif
      uint32_t v[2][32];
      tmem_ld32(sj, v[0]);
      tmem_ld32(sj + 32, v[1]);
      float bh = 0.0f;
      if (RELPOS) {
        const int c0 = hi + 63 - j;  // T_h index of key row j
        if (Cfg::TH_PACKED) {
          const uint32_t a0 = tmem_ld1(th_addr + (c0 >> 1));
          tmem_ld_wait();
          bh = unpack_f16(a0, c0 & 1);
        } else {
          const uint32_t a0 = tmem_ld1(th_addr + (c0 > 63 ? 63 : c0));
          tmem_ld_wait();
          bh = (c0 > 63) ? bh64 : __uint_as_float(a0) * F4_LOG2E;
        }
      } else {
        tmem_ld_wait();
      }
      if (!RELPOS && (j + 1) * 64 > p.Tk) {  // ragged last step: keys >= Tk get a score of -inf (exp2 -> 0)
        const int nvalid = p.Tk - j * 64;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= nvalid) v[0][i] = 0xff800000u;
          if (32 + i >= nvalid) v[1][i] = 0xff800000u;
        }
      }
      uint64_t ls2[2] = {0ull, 0ull};  // this step's row sum (relative to m_ref): 2 packed pairs = 4 independent chains
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t(&cur)[32] = v[c];
        float d = RELPOS ? bh - m_ref : -m_ref;  // +inf while m_ref = -inf: the first chunk always takes the exact path
        float ymax[2] = {-INFINITY, -INFINITY};
        uint64_t cs2[2] = {0ull, 0ull};  // 2 x 2 partial row sums (packed fp32x2: FFMA2 / FADD2 halve the FMA-pipe issue load)
        uint32_t pk[16];
        const uint64_t c1p = pk2(c1, c1);
        {
          const uint64_t dp = pk2(d, d);
          uint64_t yp[16];  // exp2 arguments (FFMA2 / FADD2: placed by the compiler where their exp2 needs them)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t vp = pk2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1]));
            if (RELPOS) {
              const uint64_t y2 = fma2(vp, c1p, pk2(tw[RELPOS ? c * 32 + 2 * i : 0], tw[RELPOS ? c * 32 + 2 * i + 1 : 0]));
              float y0, y1;
              unpk2(y2, y0, y1);
              ymax[0] = fmaxf(ymax[0], y0);
              ymax[1] = fmaxf(ymax[1], y1);
              yp[i] = add2(y2, dp);
            } else {
              ymax[0] = fmaxf(ymax[0], __uint_as_float(cur[2 * i]));  // raw scores: c1 > 0
              ymax[1] = fmaxf(ymax[1], __uint_as_float(cur[2 * i + 1]));
              yp[i] = fma2(vp, c1p, dp);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a0, a1;
            unpk2(yp[i], a0, a1);
            const bool poly = WM_F4_POLY > 0 && (i % (WM_F4_POLY > 0 ? WM_F4_POLY : 1)) == 0;
            const float e0 = poly ? ex2_poly(a0) : F4_EX2(a0), e1 = poly ? ex2_poly(a1) : F4_EX2(a1);
            cs2[i & 1] = add2(cs2[i & 1], pk2(e0, e1));
            pk[i] = pack_bf16(e0, e1);
          }
        }
        const float m_chunk = RELPOS ? fmaxf(ymax[0], ymax[1]) + bh : fmaxf(ymax[0], ymax[1]) * c1;
        const bool need = m_chunk > m_ref + F4_TAU;
        if (__any_sync(0xffffffffu, need)) {
          // ---- exact path (rare after the first chunk of a row): raise the reference maximum.  O must be stable: the
          // MMA warp runs ahead with SCORE tiles only; P_t(j-1) V is the last MMA that touches O_t before p_full(j).
          const float m_new = need ? m_chunk : m_ref;
          const float alpha = ex2_approx(m_ref - m_new);  // 1 for lanes that did not need it, 0 on the very first chunk
          if (j > 0) {
            // P_t(j-1) V must have landed in O.  It was issued ahead of S_t(j+1): wait for that score tile (a peek, the
            // barrier is waited on again at step j + 1); the last step has no such tile and uses the explicit commit.
            if (j + 1 < ns) mbar_wait(&s_full[t * 2 + ((j + 1) & 1)], (uint32_t)((j + 1) >> 1) & 1u);
            else mbar_wait(&pv_done[t], 0);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
              uint32_t o[16];
              tmem_ld16(o_addr + k * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(o_addr + k * 16, o);
            }
          }
          l_run *= alpha;
          if (c > 0) {  // P chunk 0 of this step was written against the old reference
            tmem_st_wait();
            uint32_t o[16];
            tmem_ld16(sj, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float lo = __uint_as_float(o[i] << 16) * alpha, hi2 = __uint_as_float(o[i] & 0xffff0000u) * alpha;
              o[i] = pack_bf16(lo, hi2);
            }
            tmem_st16(sj, o);
            const uint64_t ap = pk2(alpha, alpha);
            ls2[0] = mul2(ls2[0], ap);
            ls2[1] = mul2(ls2[1], ap);
          }
          m_ref = m_new;
          d = RELPOS ? bh - m_ref : -m_ref;
          cs2[0] = 0ull;
          cs2[1] = 0ull;
          const uint64_t dp = pk2(d, d);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t vp = pk2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1]));
            uint64_t yp;
            if (RELPOS)
              yp = add2(fma2(vp, c1p, pk2(tw[RELPOS ? c * 32 + 2 * i : 0], tw[RELPOS ? c * 32 + 2 * i + 1 : 0])), dp);
            else
              yp = fma2(vp, c1p, dp);
            float a0, a1;
            unpk2(yp, a0, a1);
            const bool poly = WM_F4_POLY > 0 && (i % (WM_F4_POLY > 0 ? WM_F4_POLY : 1)) == 0;
            const float e0 = poly ? ex2_poly(a0) : F4_EX2(a0), e1 = poly ? ex2_poly(a1) : F4_EX2(a1);
            cs2[i & 1] = add2(cs2[i & 1], pk2(e0, e1));
            pk[i] = pack_bf16(e0, e1);
          }
        }
        ls2[0] = add2(ls2[0], cs2[0]);
        ls2[1] = add2(ls2[1], cs2[1]);
        // P chunk c (16 columns of bf16 pairs) overwrites score columns [16c, 16c+16) of chunk 0, which is in registers
        tmem_st16(sj + c * 16, pk);
      }
      {
        float s0, s1, s2, s3;
        unpk2(ls2[0], s0, s1);
        unpk2(ls2[1], s2, s3);
        l_run += (s0 + s1) + (s2 + s3);
      }
      if (q4 == 0 && lane == 0) F4_TRACE(t, j, 2);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t * 2 + (j & 1)]);
      if (q4 == 0 && lane == 0) F4_TRACE(t, j, 3);
    }
    // ---- epilogue: O / l
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* dst = p.out + (size_t)(b * p.Tq + m0 + t * 128 + r) * p.ldo + h * HD;
    const bool row_ok = m0 + t * 128 + r < p.Tq;
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t o[16];
      tmem_ld16(o_addr + c * 16, o);
      tmem_ld_wait();
      uint4* d4 = reinterpret_cast<uint4*>(dst + c * 16);
#pragma unroll
      for (int g = 0; g < 2; ++g)
        if (row_ok) d4[g] = make_uint4(pack_bf16(__uint_as_float(o[8 * g]) * inv_l, __uint_as_float(o[8 * g + 1]) * inv_l),
                           pack_bf16(__uint_as_float(o[8 * g + 2]) * inv_l, __uint_as_float(o[8 * g + 3]) * inv_l),
                           pack_bf16(__uint_as_float(o[8 * g + 4]) * inv_l, __uint_as_float(o[8 * g + 5]) * inv_l),
                           pack_bf16(__uint_as_float(o[8 * g + 6]) * inv_l, __uint_as_float(o[8 * g + 7]) * inv_l));
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
static int launch_flash4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                         const FlashParams& p, cudaStream_t st) {
  using Cfg = Flash4Cfg<HD, RELPOS>;
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(flash4_kernel<HD, RELPOS>, Cfg::SMEM_BYTES, attr_done)) return rc;
  dim3 grid((p.Tq + 255) / 256, p.H, p.B);
  flash4_kernel<HD, RELPOS><<<grid, F4_THREADS, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, trel, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// q tiles: box 128 rows; k/v tiles: box 128 rows; rel table [256,64]: box 16 rows
int flash4_dispatch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                    const FlashParams& p, int hd, cudaStream_t st) {
  if (p.Tq < 1 || p.Tk < 1) return WM_ERR_SHAPE;
  if (p.use_relpos) {
    // queries: the whole 64x64 grid; keys: its first Tk / 64 rows (the encoder always passes all 4096; fewer are accepted for
    // measurements: time vs key count separates the per-step cost from the per-CTA prologue)
    if (p.Tq != 4096 || p.Tk > 4096 || p.Tk % 128 != 0) return WM_ERR_SHAPE;
    if (hd == 64) return launch_flash4<64, true>(tq, tk, tv, trel, p, st);
    if (hd == 80) return launch_flash4<80, true>(tq, tk, tv, trel, p, st);
    return WM_ERR_SHAPE;
  }
  if (hd == 64) return launch_flash4<64, false>(tq, tk, tv, trel, p, st);
  if (hd == 80) return launch_flash4<80, false>(tq, tk, tv, trel, p, st);
  if (hd == 128) return launch_flash4<128, false>(tq, tk, tv, trel, p, st);
  return WM_ERR_SHAPE;
}

#ifdef WM_F3_TRACE
int flash4_read_trace(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_f4_trace, sizeof(g_f4_trace)) == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}
#else
int flash4_read_trace(unsigned long long*) { return WM_ERR_ARCH; }
#endif

}  // namespace wm
