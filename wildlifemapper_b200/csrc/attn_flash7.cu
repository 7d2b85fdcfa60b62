// Flash attention v7 on tcgen05 / TMEM (sm_100a), head dim 64: the global 64x64 attention of the encoder blocks with the
// decomposed rel-pos bias (reference image_encoder.py:246-262, 347-383) and the head-padded decoder attention of the
// dense-herd configuration (transformer.py:218-240, ragged Tq / Tk, no bias).  Same math, operand layouts and building
// blocks as v4 (attn_flash4.cu): two 128-query tiles per CTA, one MMA issuer warp per tile, 64-key softmax steps, P kept
// in tensor memory (tcgen05.mma "ts" form), O accumulated in TMEM with lazy rescaling, bias_w in registers.
//
// What changed, and why (round-2 analysis of the v4 ablation): v4 is LATENCY bound on the
// round trip  P_t(j) stored -> issuer sees it -> P_t(j) V and S_t(j+2) issued -> executed behind the other tile's MMAs ->
// commit -> softmax sees S_t(j+2).  With two score buffers per tile that round trip (~1600-1800 cycles) has to fit into ONE
// softmax pass (~1100 cycles), so every step waits; the tensor pipe is 28 % busy and the MUFU 45 %.  Here
//
//   * each query tile owns a RING OF THREE 64-column score buffers: the issuer runs three steps ahead, the round trip
//     may take two passes.  The tensor-memory columns come from moving T_h = q.Rh out of TMEM: it lives TRANSPOSED in
//     shared memory (th[key row][query], 64 KB, conflict-free: one float per thread and step);
//   * the lazy-rescale test no longer needs the maximum of every chunk (32 FMNMX per chunk): the chunk's ROW SUM, which is
//     computed anyway, exceeds 2^TAU whenever a probability does -- the maximum is only evaluated on the rare exact path;
//   * P_t(j) V completion is published every step on its own barrier (the rescale path waits on it directly).
//
//   warp 0       TMA producer (Q tiles, tables, K and V rings)      warp 2  TMEM allocator
//   warp 1 / 3   tcgen05.mma issuer of query tile 0 / 1
//   warps 4-7    softmax of query tile 0     warps 8-11  softmax of query tile 1 (thread = one query row)
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

#ifdef WM_F3_TRACE
__device__ unsigned long long g_f7_trace[3][64][8];
#define F7_TRACE(role, j, ev)                                                                   \
  do {                                                                                          \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 64) g_f7_trace[role][j][ev] = clock64(); \
  } while (0)
#else
#define F7_TRACE(role, j, ev) do { } while (0)
#endif

// exp2 on the FMA / ALU pipes (cubic minimax after rounding to the nearest integer, relative error ~1e-4, below the bf16
// rounding of P): WM_F7_POLY = n sends every n-th PAIR of a 32-score chunk this way (MUFU relief, FA4 style)
__device__ __forceinline__ float f7_ex2_poly(float x) {
  x = fminf(fmaxf(x, -120.0f), 126.0f);  // (saturate: the lazy-rescale test reads the row SUM, a wrapped exponent would hide an overflow)
  const float t = x + 12582912.0f;  // 1.5 * 2^23: round to nearest integer in the mantissa
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// the same for a packed pair (FADD2 / FFMA2: 12 instructions per pair instead of 18)
__device__ __forceinline__ void f7_ex2_poly2(uint64_t yp, float& e0, float& e1) {
  float x0, x1;
  unpk2(yp, x0, x1);
  x0 = fminf(fmaxf(x0, -120.0f), 126.0f);
  x1 = fminf(fmaxf(x1, -120.0f), 126.0f);
  const uint64_t x = pk2(x0, x1);
  const uint64_t t = add2(x, pk2(12582912.0f, 12582912.0f));
  const uint64_t rr = add2(t, pk2(-12582912.0f, -12582912.0f));
  const uint64_t f = fma2(rr, pk2(-1.0f, -1.0f), x);  // x - round(x), exact
  uint64_t pp = fma2(f, pk2(0.0555041f, 0.0555041f), pk2(0.2402265f, 0.2402265f));
  pp = fma2(pp, f, pk2(0.6931472f, 0.6931472f));
  pp = fma2(pp, f, pk2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  unpk2(pp, p0, p1);
  unpk2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
#ifndef WM_F7_POLY
#define WM_F7_POLY 0
#endif

// one stage of the lane-dependent down-shift of x[0 .. 64 + SH - 1): lanes with bit SH set take x[i + SH]
template <int SH>
__device__ __forceinline__ void f7_barrel_stage(uint32_t (&x)[96], int lane) {
  const bool on = (lane & SH) != 0;
#pragma unroll
  for (int i = 0; i < 64 + SH - 1; ++i) x[i] = on ? x[i + SH] : x[i];  // ascending i: x[i + SH] is still the old value
}

// every lane of the waiting warp polls: measured FASTER than one polling lane + __syncwarp (707 vs 800 ns per step), and a
// suspend-time hint ("parked" waits) is slower still (1300 ns): the wake-up latency sits on the critical path
#define f7_wait mbar_wait
#define F7_ROLE_WAIT mbar_wait  // (a __nanosleep back-off between the role warps' polls, 32 - 300 ns, changed nothing: 2.29 - 2.36 ms)

constexpr int F7_THREADS = 384;
constexpr int F7_NBUF = 3;  // score buffers per query tile
constexpr float F7_LOG2E = 1.4426950408889634f;
constexpr float F7_TAU = 16.0f;          // lazy-rescale threshold (log2 units)
constexpr float F7_SUM_LIMIT = 65536.0f;  // 2^TAU: a chunk whose row sum stays below holds no probability above 2^TAU

template <bool RELPOS>
struct Flash7Cfg {
  static constexpr int HD = 64;
  static constexpr int STAGES = 4;
  static constexpr int TILE_BYTES = 16384;  // 128 rows (queries or keys) x 128 B
  static constexpr int OFF_Q = 0;           // 2 query tiles
  static constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;
  static constexpr int OFF_V = OFF_K + STAGES * TILE_BYTES;
  // RELPOS prologue: Rw [128 rows] + 2 x Rh slice [80 rows].  The region ALIASES the V stages, whose first loads wait
  // until the table products have been issued and completed (t_full).
  static constexpr int TAB_BYTES = RELPOS ? 16384 + 2 * 10240 : 0;
  static constexpr int OFF_TAB = OFF_V;
  static_assert(TAB_BYTES <= STAGES * TILE_BYTES, "aliased table region");
  // T_h transposed: th[kh][t * 128 + r] = log2e * q_r . Rh[h_r - kh + 63]  (fp32, 64 key rows x 256 queries)
  static constexpr int OFF_TH = OFF_V + STAGES * TILE_BYTES;
  static constexpr int TH_BYTES = RELPOS ? 64 * 256 * 4 : 0;
  static constexpr int OFF_BAR = OFF_TH + TH_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  // TMEM columns: score buffer b of tile t at 192 t + 64 b (P = bf16 pairs in its first 32 columns), O_t at 384 + 64 t
  static constexpr int COL_O = 2 * F7_NBUF * 64;
  static constexpr int TMEM_COLS = 512;
  static_assert(COL_O + 2 * HD <= 512, "TMEM budget");
};

template <bool RELPOS>
__global__ void __launch_bounds__(F7_THREADS, 1)
flash7_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_rel,
              const FlashParams p) {
  using Cfg = Flash7Cfg<RELPOS>;
  constexpr int HD = Cfg::HD;
  constexpr int NS = Cfg::STAGES;
  constexpr int NB = F7_NBUF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [4]
  uint64_t* k_empty = bars + 5;   // [4]
  uint64_t* v_full = bars + 9;    // [4]
  uint64_t* v_empty = bars + 13;  // [4]
  uint64_t* s_full = bars + 17;   // [2 tiles][3 buffers]  S_t(j) complete
  uint64_t* p_full = bars + 23;   // [2 tiles][3 buffers]  P_t(j) stored to TMEM by the 4 warps of tile t
  uint64_t* pv_tail = bars + 29;  // [2 tiles][2]  P_t(ns-3) V / P_t(ns-2) V complete (single use; only the rare rescale path of the last two steps waits)
  uint64_t* o_full = bars + 33;   // [2]
  uint64_t* t_full = bars + 35;   // rel-pos table products complete
  uint64_t* t_done = bars + 36;   // ... and drained out of the S / O columns by the 8 softmax warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);

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
    for (int i = 0; i < 2 * NB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&pv_tail[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&o_full[i], 1);
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
    // Producer and MMA roles: the WHOLE warp runs the control flow (uniform branches, every lane polls the barriers); one
    // elected lane issues the TMA / tcgen05 instructions.
    auto mma_main = [&](const int t, const bool leader) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);  // A = P from TMEM (K-major), V is MN-major
      const uint32_t sq = smem_u32(smem + Cfg::OFF_Q) + t * Cfg::TILE_BYTES;
      const uint32_t s_base = tmem_base + t * (NB * 64);
      const uint32_t o_col = tmem_base + Cfg::COL_O + t * HD;
      // S_t(step) = Q_t K(step)^T (128 x 64 x 64) into ring buffer `buf`; K rows of the step: half (step & 1) of K stage
      // (step / 2) % NS.  An even step opens a K tile (wait for its TMA), an odd one (or the last) releases it.
      auto issue_s = [&](int step, int buf) {
        const int kt = step >> 1, kst = kt % NS;
        if ((step & 1) == 0) {
          F7_ROLE_WAIT(&k_full[kst], (uint32_t)(kt / NS) & 1u);
          tc_fence_after();
        }
        if (leader) {
          const uint64_t qd = make_sdesc_sw128(sq, 16, 1024);
          const uint64_t kd = make_sdesc_sw128(smem_u32(smem + Cfg::OFF_K + kst * Cfg::TILE_BYTES) + (step & 1) * 8192, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)
            umma_bf16(s_base + buf * 64, qd + (uint32_t)(ks * 2), kd + (uint32_t)(ks * 2), idesc_s, ks != 0);  // address field is bytes >> 4
          umma_commit(&s_full[t * NB + buf]);
          if ((step & 1) || step == ns - 1) umma_commit(&k_empty[kst]);
        }
        __syncwarp();
      };
      // three steps ahead of the softmax
      for (int s0 = 0; s0 < NB && s0 < ns; ++s0) issue_s(s0, s0);
      int buf = 0;
      uint32_t ph = 0;
      for (int j = 0; j < ns; ++j) {
        const int vt = j >> 1, vst = vt % NS;
        if (leader) F7_TRACE(2, j, 4 * t);
        F7_ROLE_WAIT(&p_full[t * NB + buf], ph);
        if (leader) F7_TRACE(2, j, 4 * t + 1);
        if ((j & 1) == 0) F7_ROLE_WAIT(&v_full[vst], (uint32_t)(vt / NS) & 1u);
        tc_fence_after();
        if (leader) {
          const uint64_t vd = make_sdesc_sw128(smem_u32(smem + Cfg::OFF_V + vst * Cfg::TILE_BYTES) + (j & 1) * 8192, 16384, 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)  // 64 keys, 16 per MMA; P: 8 TMEM columns per step; V: 2048 B per step
            umma_bf16_ts(o_col, s_base + buf * 64 + ks * 8, vd + (uint32_t)(ks * (2048 >> 4)), idesc_pv, (j | ks) != 0);
          // P_t(j) V complete: only the rare rescale path of step j + 1 needs it.  While score tiles are still being issued
          // behind the P V groups the commit of S_t(j + 3) covers it (the pipe completes in issue order); the last steps
          // publish it on single-use barriers (a shared, re-used barrier would alias: the softmax may run two phases ahead
          // of a lagging issuer).
          if (j == ns - 3) umma_commit(&pv_tail[t * 2]);
          if (j == ns - 2) umma_commit(&pv_tail[t * 2 + 1]);
          if (j == ns - 1) umma_commit(&o_full[t]);
          if ((j & 1) || j == ns - 1) umma_commit(&v_empty[vst]);  // both halves of the V tile consumed by this tile
        }
        __syncwarp();
        if (leader) F7_TRACE(2, j, 4 * t + 2);
        if (j + NB < ns) issue_s(j + NB, buf);  // reuses the score buffer step j just released (behind its P V in the pipe)
        if (leader) F7_TRACE(2, j, 4 * t + 3);
        if (++buf == NB) { buf = 0; ph ^= 1u; }
      }
    };
    if (warp == 0) {
      // ------------------------------------------------------------ TMA producer
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * Cfg::TILE_BYTES + Cfg::TAB_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t)
          tma_load_2d(smem + Cfg::OFF_Q + t * Cfg::TILE_BYTES, &tmap_q, q_full, p.q_col0 + h * HD, b * p.Tq + m0 + t * 128);
        if (RELPOS) {
          // table tensor [256,64]: rows 0..126 rel_pos_h, 128..254 rel_pos_w.  Box = 64 columns x 16 rows.
          const int qi0 = m0 >> 6;  // first image row of this CTA (4 rows: 2 per query tile)
          for (int i = 0; i < 8; ++i) tma_load_2d(smem + Cfg::OFF_TAB + i * 2048, &tmap_rel, q_full, 0, 128 + 16 * i);
          for (int t = 0; t < 2; ++t)
            for (int i = 0; i < 5; ++i)
              tma_load_2d(smem + Cfg::OFF_TAB + 16384 + t * 10240 + i * 2048, &tmap_rel, q_full, 0, qi0 + 2 * t + 16 * i);
        }
      }
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nk; ++j) {
        F7_ROLE_WAIT(&k_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
          tma_load_2d(smem + Cfg::OFF_K + st * Cfg::TILE_BYTES, &tmap_k, &k_full[st], p.k_col0 + h * HD, b * p.Tk + j * 128);
        }
        if (RELPOS && j == 0) mbar_wait(t_full, 0);  // the V stages hold the rel-pos tables until their products are complete
        F7_ROLE_WAIT(&v_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
          tma_load_2d(smem + Cfg::OFF_V + st * Cfg::TILE_BYTES, &tmap_v, &v_full[st], p.v_col0 + h * HD, b * p.Tk + j * 128);
        }
        __syncwarp();
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------ MMA issuer of tile 0 (+ the table products of both tiles)
      const bool leader = elect_one();  // the same lane issues every tcgen05.mma / tcgen05.commit
      mbar_wait(q_full, 0);
      tc_fence_after();
      if (RELPOS) {
        constexpr uint32_t idesc_tw = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_th = make_idesc_bf16(128, 64, 0, 0);
        constexpr uint32_t idesc_tx = make_idesc_bf16(128, 16, 0, 0);
        const uint32_t stab = smem_u32(smem + Cfg::OFF_TAB);
        const uint32_t sq = smem_u32(smem + Cfg::OFF_Q);
        // T_w(t) -> score buffers 0, 1 of tile t (128 columns); T_h(t)[0..63] -> score buffer 2; T_h(t)[64..79] -> 16
        // columns of the O region.  All scratch: the softmax warps move them to registers / shared memory.
        if (leader) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks) {
              const uint32_t koff = ks * 32;
              const uint32_t srh = stab + 16384 + t * 10240 + koff;
              const uint64_t ad = make_sdesc_sw128(sq + t * Cfg::TILE_BYTES + koff, 16, 1024);
              umma_bf16(tmem_base + t * (NB * 64), ad, make_sdesc_sw128(stab + koff, 16, 1024), idesc_tw, ks != 0);
              umma_bf16(tmem_base + t * (NB * 64) + 128, ad, make_sdesc_sw128(srh, 16, 1024), idesc_th, ks != 0);
              umma_bf16(tmem_base + Cfg::COL_O + t * 16, ad, make_sdesc_sw128(srh + 8192, 16, 1024), idesc_tx, ks != 0);
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
      // ------------------------------------------------------------ MMA issuer of tile 1
      const bool leader = elect_one();
      mbar_wait(q_full, 0);
      if (RELPOS) mbar_wait(t_done, 0);
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
    const uint32_t s_addr = lane_addr + t * (NB * 64);
    const uint32_t o_addr = lane_addr + Cfg::COL_O + t * HD;
    const float c1 = p.scale * F7_LOG2E;
    float tw[RELPOS ? 64 : 1];
    const uint32_t th_s = smem_u32(smem + Cfg::OFF_TH) + (uint32_t)(t * 128 + r) * 4u;  // + 1024 * key row

    if (RELPOS) {
      const int qj = (m0 + t * 128 + r) & 63;
      const int hi = r >> 6;  // image row of this query inside the tile (warp-uniform)
      mbar_wait(t_full, 0);
      tc_fence_after();
      {
        // bias_h[kh] = T_h[hi + 63 - kh] -> th[kh][query] (x log2 e); every thread writes and later reads only its own column
        uint32_t a[32], a2[32];
        tmem_ld32(s_addr + 128, a);
        tmem_ld32(s_addr + 160, a2);
        const uint32_t x64 = tmem_ld1(lane_addr + Cfg::COL_O + t * 16);
        tmem_ld_wait();
        float* tho = reinterpret_cast<float*>(smem + Cfg::OFF_TH) + t * 128 + r;
        if (hi == 0) {
#pragma unroll
          for (int kh = 0; kh < 64; ++kh) {
            const int c = 63 - kh;
            tho[kh * 256] = __uint_as_float(c < 32 ? a[c & 31] : a2[c & 31]) * F7_LOG2E;
          }
        } else {
          tho[0] = __uint_as_float(x64) * F7_LOG2E;
#pragma unroll
          for (int kh = 1; kh < 64; ++kh) {
            const int c = 64 - kh;
            tho[kh * 256] = __uint_as_float(c < 32 ? a[c & 31] : a2[c & 31]) * F7_LOG2E;
          }
        }
      }
      // bias_w[kw] = T_w[qj + 63 - kw], qj = (q4 & 1) * 32 + lane: every lane needs a 64-column window of its T_w row that
      // starts at a lane-dependent column.  Load the 96 columns [base, base + 96) (base = 32 (q4 & 1), warp-uniform) and
      // shift them down by `lane` positions with a barrel of selects (5 stages, 346 SEL) -- the first version scattered
      // through shared memory with ~10 instructions per examined column (~3000 per thread: 10 us of every CTA's 67).
      {
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
        f7_barrel_stage<16>(x, lane);  // (template: every index is a compile-time constant, x stays in registers)
        f7_barrel_stage<8>(x, lane);
        f7_barrel_stage<4>(x, lane);
        f7_barrel_stage<2>(x, lane);
        f7_barrel_stage<1>(x, lane);
#pragma unroll
        for (int kw = 0; kw < 64; ++kw) tw[kw] = __uint_as_float(x[63 - kw]) * F7_LOG2E;  // x[i] = T_w[qj + i]
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_done);
    }

    float m_ref = -INFINITY, l_run = 0.0f;
    int buf = 0;
    uint32_t ph = 0;
    uint32_t sj = s_addr;  // this step's score buffer; P = bf16 pairs over its first 32 columns
    const uint64_t c1p = pk2(c1, c1);
    // Software pipeline over the steps (the ring of three buffers gives the slack for it): the publication of P(j)
    // (tcgen05.wait::st, fence, warp barrier, arrive) is DEFERRED until the score loads of step j+1 are in flight, and chunk 1
    // of a step is loaded while its chunk 0 is processed.  Measured per 64-key step with rel-pos (profiles/flash_time2.py,
    // batch 32): neither 853 ns, deferred publication 780, + chunk 0 of step j+1 requested during chunk 1 of step j
    // (-DWM_F7_PREFETCH) 787; FMA-pipe exp2 for 3 of 16 pairs (-DWM_F7_POLY=6) costs more issue slots than it saves MUFU
    // time here: 745 ns without it (v4: 800).
    uint32_t v[2][32];
    f7_wait(&s_full[t * NB], 0);
    tc_fence_after();
    tmem_ld32(sj, v[0]);
    for (int j = 0; j < ns; ++j) {  // one step = 64 keys = key row j of the 64x64 grid
      float bh = 0.0f;
      if (RELPOS) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bh) : "r"(th_s + (uint32_t)j * 1024u));
      if (q4 == 0 && lane == 0) F7_TRACE(t, j, 0);
#ifndef WM_F7_PREFETCH
      if (j > 0) {
        f7_wait(&s_full[t * NB + buf], ph);
        tc_fence_after();
        tmem_ld32(sj, v[0]);
      }
#endif
      tmem_ld_wait();                // chunk 0 (requested during the previous step) has landed
      tmem_ld32(sj + 32, v[1]);      // chunk 1: in flight while chunk 0 is processed
#ifndef WM_F7_NO_DEFER
      if (j > 0) {                   // publish P(j-1): its stores were issued at the end of the previous step
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t * NB + (buf == 0 ? NB - 1 : buf - 1)]);
      }
#endif
      if (q4 == 0 && lane == 0) F7_TRACE(t, j, 1);
      const int nvalid = (!RELPOS && (j + 1) * 64 > p.Tk) ? p.Tk - j * 64 : 64;  // ragged last step: keys >= Tk get -inf (exp2 -> 0)
      float l_step = 0.0f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t(&cur)[32] = v[c];
        if (c == 1) {
          tmem_ld_wait();  // chunk 1 has landed (requested before chunk 0 was processed)
#ifdef WM_F7_PREFETCH
          if (j + 1 < ns) {  // request chunk 0 of the next step into the registers chunk 0 just vacated
            const int nb = buf + 1 == NB ? 0 : buf + 1;
            f7_wait(&s_full[t * NB + nb], buf + 1 == NB ? ph ^ 1u : ph);
            tc_fence_after();
            tmem_ld32(buf + 1 == NB ? s_addr : sj + 64, v[0]);
          }
#endif
        }
        if (!RELPOS && nvalid < 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= nvalid) cur[i] = 0xff800000u;
        }
        float d = bh - m_ref;  // +inf while m_ref = -inf: the first chunk always takes the exact path
        uint64_t cs2[2] = {0ull, 0ull};  // 2 x 2 partial row sums (packed fp32x2)
        uint32_t pk[16];
        {
          const uint64_t dp = pk2(d, d);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t vp = pk2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1]));
            uint64_t yp;
            if (RELPOS) yp = add2(fma2(vp, c1p, pk2(tw[RELPOS ? c * 32 + 2 * i : 0], tw[RELPOS ? c * 32 + 2 * i + 1 : 0])), dp);
            else yp = fma2(vp, c1p, dp);
            float e0, e1;
            if (WM_F7_POLY > 0 && (i % (WM_F7_POLY > 0 ? WM_F7_POLY : 1)) == 0) {
              f7_ex2_poly2(yp, e0, e1);
            } else {
              float a0, a1;
              unpk2(yp, a0, a1);
              e0 = ex2_approx(a0);
              e1 = ex2_approx(a1);
            }
            cs2[i & 1] = add2(cs2[i & 1], pk2(e0, e1));
            pk[i] = pack_bf16(e0, e1);
          }
        }
        float s0, s1, s2, s3;
        unpk2(cs2[0], s0, s1);
        unpk2(cs2[1], s2, s3);
        float csum = (s0 + s1) + (s2 + s3);
        const bool need = !(csum <= F7_SUM_LIMIT);  // also true for inf / NaN (first chunk: d = +inf)
        if (__any_sync(0xffffffffu, need)) {
          // ---- exact path (rare after the first chunk of a row): raise the reference maximum to this chunk's maximum.
          float ymax = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float y = RELPOS ? fmaf(__uint_as_float(cur[i]), c1, tw[RELPOS ? c * 32 + i : 0]) : __uint_as_float(cur[i]) * c1;
            ymax = fmaxf(ymax, y);
          }
          const float m_chunk = ymax + bh;
          const float m_new = need ? fmaxf(m_chunk, m_ref) : m_ref;
          const float alpha = ex2_approx(m_ref - m_new);  // 1 for lanes that did not need it, 0 on the very first chunk
          if (j > 0) {
            // O must be stable: P_t(j-1) V is the last MMA that touches O_t before p_full(j).  It was issued ahead of
            // S_t(j+2): wait for that score tile (a peek -- the barrier is waited on again later); the last two steps
            // have no such tile and use the single-use barriers.
            if (j + 2 < ns) {
              int b2 = buf + 2;
              uint32_t ph2 = ph;
              if (b2 >= NB) { b2 -= NB; ph2 ^= 1u; }
              f7_wait(&s_full[t * NB + b2], ph2);
            } else {
              f7_wait(&pv_tail[t * 2 + (j + 2 - ns)], 0);
            }
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
            l_step *= alpha;
          }
          m_ref = m_new;
          d = bh - m_ref;
          cs2[0] = 0ull;
          cs2[1] = 0ull;
          const uint64_t dp = pk2(d, d);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t vp = pk2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1]));
            uint64_t yp;
            if (RELPOS) yp = add2(fma2(vp, c1p, pk2(tw[RELPOS ? c * 32 + 2 * i : 0], tw[RELPOS ? c * 32 + 2 * i + 1 : 0])), dp);
            else yp = fma2(vp, c1p, dp);
            float a0, a1;
            unpk2(yp, a0, a1);
            const float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
            cs2[i & 1] = add2(cs2[i & 1], pk2(e0, e1));
            pk[i] = pack_bf16(e0, e1);
          }
          unpk2(cs2[0], s0, s1);
          unpk2(cs2[1], s2, s3);
          csum = (s0 + s1) + (s2 + s3);
        }
        l_step += csum;
        // P chunk c (16 columns of bf16 pairs) overwrites score columns [16c, 16c+16): chunk 0's, long in registers
        tmem_st16(sj + c * 16, pk);
      }
      l_run += l_step;
      if (q4 == 0 && lane == 0) F7_TRACE(t, j, 2);
#ifdef WM_F7_NO_DEFER
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t * NB + buf]);
#endif
      sj += 64;
      if (++buf == NB) { buf = 0; ph ^= 1u; sj = s_addr; }
    }
#ifndef WM_F7_NO_DEFER
    {  // publish the last P
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t * NB + (buf == 0 ? NB - 1 : buf - 1)]);
    }
#endif
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

template <bool RELPOS>
static int launch_flash7(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                         const FlashParams& p, cudaStream_t st) {
  using Cfg = Flash7Cfg<RELPOS>;
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(flash7_kernel<RELPOS>, Cfg::SMEM_BYTES, attr_done)) return rc;
  dim3 grid((p.Tq + 255) / 256, p.H, p.B);
  flash7_kernel<RELPOS><<<grid, F7_THREADS, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, trel, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

// head dim 64 only (q / k / v tiles: box 128 rows; rel table [256,64]: box 16 rows)
int flash7_dispatch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                    const FlashParams& p, int hd, cudaStream_t st) {
  if (p.Tq < 1 || p.Tk < 1 || hd != 64) return WM_ERR_SHAPE;
  if (p.use_relpos) {
    // queries: the whole 64x64 grid; keys: its first Tk / 64 rows (the encoder always passes all 4096; fewer are accepted for
    // measurements: time vs key count separates the per-step cost from the per-CTA prologue)
    if (p.Tq != 4096 || p.Tk > 4096 || p.Tk % 128 != 0) return WM_ERR_SHAPE;
    return launch_flash7<true>(tq, tk, tv, trel, p, st);
  }
  return launch_flash7<false>(tq, tk, tv, trel, p, st);
}

#ifdef WM_F3_TRACE
int flash7_read_trace(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_f7_trace, sizeof(g_f7_trace)) == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}
#else
int flash7_read_trace(unsigned long long*) { return WM_ERR_ARCH; }
#endif

}  // namespace wm
