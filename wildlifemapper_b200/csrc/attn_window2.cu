// Fused 14x14 windowed attention with decomposed rel-pos bias, second generation (sm_100a, tcgen05 / TMEM / TMA).
// Same math and reference sites as attn_window.cu (image_encoder.py:188-204, 246-311, 347-383; pad keys: SURVEY.md
// section 0.2).  The first kernel ran one non-pipelined CTA per (half window, head): TMA -> MMA -> softmax -> MMA ->
// store strictly in sequence, ~7.4 us per CTA and 88 TFLOP/s (profiles/r01a_ncu_window_summary.txt).  This one is
// PERSISTENT and pipelined:
//
//   * work item = one (image, window, head): K and V are loaded once for both 98-query halves (7 window rows x 14,
//     dense 14-wide TMA boxes: no 16-wide padding columns, no masked keys);
//   * S' = Q [K ; Rh ; Rw]^T is ONE 128 x 256 x 64 MMA group: the 196 score columns and the 27 + 27 rel-pos table
//     products T_h = q.Rh, T_w = q.Rw come out of the same instruction (the table rows sit behind the key rows of the
//     K stage buffers, loaded once per CTA);
//   * the probabilities never touch shared memory: bf16 pairs are written back over the consumed score columns in
//     tensor memory and the P V MMA reads its A operand from TMEM; O lands in the (by then dead) table columns;
//   * the two query halves ping-pong: each has its own MMA issuer warp (P V of item n-1, then S' of item n), so one
//     tile's tensor-core round trips overlap the other tile's exp2 pass; K/V stages are double buffered, Q tiles are
//     reloaded as soon as their S' has been issued;
//   * the rel-pos bias of a row (T_h[y - kh + 13], T_w[x - kw + 13]) is picked out of the 27 + 27 table products in
//     registers (barrel of selects), not through shared memory;
//   * the output leaves as a 4-D TMA store (box 64 ch x 14 x 7): the crop of the padded 70x70 grid back to 64x64 is
//     the TMA bounds check.
//
//   warp 0       TMA producer          warps 1, 3  tcgen05.mma issuers (tile 0, tile 1)          warp 2  TMEM allocator
//   warps 4-7    softmax / output of query tile 0     warps 8-11  of query tile 1     (thread = one query row)
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

#ifdef WM_F3_TRACE
// diagnostics build: SM-clock timeline of CTA 0 (role 0 / 1 = softmax warpgroup of tile 0 / 1, role 2 = MMA warp)
__device__ unsigned long long g_w2_trace[3][64][8];
#define W2_TRACE(role, n, ev)                                                          \
  do {                                                                                 \
    if (blockIdx.x == 0 && (n) < 64) g_w2_trace[role][n][ev] = clock64();              \
  } while (0)
#else
#define W2_TRACE(role, n, ev) do { } while (0)
#endif
#ifdef WM_W2_TRACE_WARPS  // per-warp skew of tile 0 instead of the MMA events: role 2, ev = warp (pass start), 4 + warp (P arrive)
#define W2_TRACE_MMA(n, ev) do { } while (0)
#define W2_TRACE_WARP(n, ev) W2_TRACE(2, n, ev)
#else
#define W2_TRACE_MMA(n, ev) W2_TRACE(2, n, ev)
#define W2_TRACE_WARP(n, ev) do { } while (0)
#endif

constexpr int W2_THREADS = 384;
constexpr float W2_LOG2E = 1.4426950408889634f;
#ifndef WM_W2_TAU
#define WM_W2_TAU 16.0f
#endif
// P = 2^(y - m_ref) may exceed 1 by up to 2^TAU before the reference maximum is raised: bf16 P and the fp32 accumulators
// have the exponent range for it (relative precision is unchanged), and raises become rare even for peaky logits.
constexpr float W2_TAU = WM_W2_TAU;
constexpr int W2_Q_BYTES = 16384;                 // 128 rows x 128 B (98 loaded)
constexpr int W2_K_BYTES = 32768;                 // 256 rows: 196 keys, 27 Rh, 27 Rw, 6 zero
constexpr int W2_V_BYTES = 208 * 128;             // 196 keys + 12 zero rows (K dimension of P V = 13 x 16)
constexpr int W2_STG_BYTES = 13312;               // one output staging buffer: 98 rows x 128 B, rounded up to 1 KB
constexpr int W2_OFF_Q = 0;                                  // [2 tiles]
constexpr int W2_OFF_K = W2_OFF_Q + 2 * W2_Q_BYTES;          // [2 stages]
constexpr int W2_OFF_V = W2_OFF_K + 2 * W2_K_BYTES;          // [2 stages]
constexpr int W2_OFF_SCR = W2_OFF_V + 2 * W2_V_BYTES + 1024; // [2 tiles], 1024-aligned (W2_V_BYTES is a multiple of 1024)
constexpr int W2_SCR_STRIDE = 2 * W2_STG_BYTES;          // two staging buffers per tile (items alternate)
constexpr int W2_OFF_BAR = W2_OFF_SCR + 2 * W2_SCR_STRIDE;
constexpr int W2_SMEM_BYTES = W2_OFF_BAR + 256 + 1024;
static_assert(W2_SMEM_BYTES <= 232448, "shared memory budget");
static_assert(W2_OFF_SCR % 1024 == 0 && W2_SCR_STRIDE % 1024 == 0, "staging alignment");
constexpr int W2_COL_T = 196;   // T_h at S' columns 196..222, T_w at 223..249
constexpr int W2_COL_O = 192;   // O_t (64 columns) over the consumed score / table columns

// Rare path of the single-pass softmax, kept OUT OF LINE (it sat in the middle of every unrolled chunk before, tripled the
// size of the hot loop and cost ~17 % of the softmax warps' time in instruction-fetch stalls): the reference maximum
// of this row was raised by log2(1/alpha) -- rescale the P chunks already written for it.
__device__ __noinline__ void window2_rescale_p(uint32_t p_addr, int nchunks, float alpha) {
  tmem_st_wait();
#pragma unroll 1
  for (int kk = 0; kk < nchunks; ++kk) {
    uint32_t o[16];
    tmem_ld16(p_addr + kk * 16, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float lo = __uint_as_float(o[i] << 16) * alpha, hi = __uint_as_float(o[i] & 0xffff0000u) * alpha;
      o[i] = pack_bf16(lo, hi);
    }
    tmem_st16(p_addr + kk * 16, o);
  }
}

// out[k] = in[a - k + 13] * LOG2E for k = 0..13 (a in [0, 13], `in` = the 27 table products of this row): a register
// array cannot be indexed by a per-lane value, so the shift by `a` is a 4-stage barrel of selects (67 SEL) instead of a
// round trip through shared memory (which cost ~2300 cycles per item: 54 generic stores, a warp sync, 28 loads, and a
// wait for the previous item's TMA store, whose staging buffer the scratch shared).
__device__ __forceinline__ void window2_pick14(const float (&in)[27], int a, float (&out)[14]) {
  float s8[21], s4[17], s2[15];
  const bool b8 = a & 8, b4 = a & 4, b2 = a & 2, b1 = a & 1;
#pragma unroll
  for (int j = 0; j < 21; ++j) s8[j] = (b8 && j + 8 < 27) ? in[j + 8 < 27 ? j + 8 : 26] : in[j];
#pragma unroll
  for (int j = 0; j < 17; ++j) s4[j] = b4 ? s8[j + 4] : s8[j];
#pragma unroll
  for (int j = 0; j < 15; ++j) s2[j] = b2 ? s4[j + 2] : s4[j];
#pragma unroll
  for (int j = 0; j < 14; ++j) out[13 - j] = (b1 ? s2[j + 1] : s2[j]) * W2_LOG2E;
}

__global__ void __launch_bounds__(W2_THREADS, 1)
window2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
               const __grid_constant__ CUtensorMap tmap_rel, const __grid_constant__ CUtensorMap tmap_out,
               const WindowParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W2_OFF_BAR);
  uint64_t* k_full = bars + 0;    // [2 stages]
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 6;
  uint64_t* q_full = bars + 8;    // [2 tiles]
  uint64_t* q_empty = bars + 10;
  uint64_t* s_full = bars + 12;   // S'_t complete
  uint64_t* p_full = bars + 14;   // P_t stored (4 warps)
  uint64_t* o_full = bars + 16;   // O_t complete
  uint64_t* s_free = bars + 18;   // O_t read back: the TMEM columns of tile t may be overwritten (4 warps)
  uint64_t* tab_full = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int items_total = p.B * 25 * p.H;
  const int n_items = (items_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // this CTA: blockIdx.x + n * gridDim.x

  // zero rows that no TMA box ever writes: keys 196..207 of both V stages, rows 250..255 of both K stages
  for (int i = threadIdx.x; i < 2 * (1536 / 16); i += W2_THREADS)
    reinterpret_cast<uint4*>(smem + W2_OFF_V + (i / 96) * W2_V_BYTES + 196 * 128)[i % 96] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < 2 * (768 / 16); i += W2_THREADS)
    reinterpret_cast<uint4*>(smem + W2_OFF_K + (i / 48) * W2_K_BYTES + 250 * 128)[i % 48] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_out);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2);  // released by both MMA warps
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2);
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1); mbar_init(&s_free[i], 4);
    }
    mbar_init(tab_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int n, int& b, int& wy, int& wx, int& h) {  // item n of this CTA
    const int it = (int)blockIdx.x + n * (int)gridDim.x;
    h = it % p.H;
    const int bw = it / p.H;
    const int win = bw % 25;
    b = bw / 25;
    wy = win / 5;
    wx = win % 5;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    setmaxnreg_dec<40>();
    if (elect_one()) {  // rel-pos tables behind the key rows of both K stages (table tensor [64,64]: rows 0..26 Rh, 32..58 Rw)
      mbar_arrive_expect_tx(tab_full, 4 * 27 * 128);
      for (int s = 0; s < 2; ++s) {
        tma_load_2d(smem + W2_OFF_K + s * W2_K_BYTES + 196 * 128, &tmap_rel, tab_full, 0, 0);
        tma_load_2d(smem + W2_OFF_K + s * W2_K_BYTES + 223 * 128, &tmap_rel, tab_full, 0, 32);
      }
    }
    __syncwarp();
    for (int n = 0; n < n_items; ++n) {
      int b, wy, wx, h;
      decode(n, b, wy, wx, h);
      const int st = n & 1;
      const uint32_t ph = (uint32_t)(n >> 1) & 1u;
      mbar_wait(&k_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[st], 196 * 128);
        tma_load_4d(smem + W2_OFF_K + st * W2_K_BYTES, &tmap_kv, &k_full[st], p.D + h * 64, wx * 14, wy * 14, b);
      }
      __syncwarp();
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&q_empty[t], (uint32_t)(n & 1) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[t], 98 * 128);
          tma_load_4d(smem + W2_OFF_Q + t * W2_Q_BYTES, &tmap_q, &q_full[t], h * 64, wx * 14, wy * 14 + t * 7, b);
        }
        __syncwarp();
      }
      mbar_wait(&v_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[st], 196 * 128);
        tma_load_4d(smem + W2_OFF_V + st * W2_V_BYTES, &tmap_kv, &v_full[st], 2 * p.D + h * 64, wx * 14, wy * 14, b);
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------------------------------ MMA issuers: warp 1 serves query tile 0, warp 3 tile 1
    // per iteration n:  P_t V (item n-1), then S'_t (item n) as soon as O_t has been read back.  One issuer per tile:
    // with a single warp walking both tiles in a fixed order, every blocking wait for one tile (P V round trip + O
    // read-back: ~2000 cycles) also held back work that was ready for the other tile, and the two tiles ran their
    // exp2 passes at the same time instead of alternating on the MUFU (profiles/r01z_window2_trace_before.txt).
    setmaxnreg_dec<40>();
    const int t = warp == 3;
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 256, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // A = P from TMEM, V MN-major
    const bool leader = elect_one();
    const uint32_t s_col = tmem_base + t * 256;
    mbar_wait(tab_full, 0);
    for (int n = 0; n <= n_items; ++n) {
      const int st = n & 1;
      const uint32_t ph = (uint32_t)(n >> 1) & 1u;
      const int pst = (n - 1) & 1;
      const uint32_t pph = (uint32_t)((n - 1) >> 1) & 1u;
      if (n > 0) {
        mbar_wait(&p_full[t], (uint32_t)(n - 1) & 1u);
        mbar_wait(&v_full[pst], pph);
        tc_fence_after();
        if (leader) {
          W2_TRACE_MMA(n, 4 * t);
          const uint64_t vd = make_sdesc_sw128(smem_u32(smem + W2_OFF_V + pst * W2_V_BYTES), 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 13; ++ks)  // 208 keys, 16 per MMA; P_t: 8 TMEM columns per step; V: 2048 B per step
            umma_bf16_ts(s_col + W2_COL_O, s_col + ks * 8, vd + (uint32_t)(ks * (2048 >> 4)), idesc_pv, ks != 0);
          umma_commit(&o_full[t]);
          umma_commit(&v_empty[pst]);
          W2_TRACE_MMA(n, 4 * t + 1);
        }
        __syncwarp();
      }
      if (n < n_items) {
        if (n > 0) mbar_wait(&s_free[t], (uint32_t)(n - 1) & 1u);
        mbar_wait(&q_full[t], (uint32_t)n & 1u);
        mbar_wait(&k_full[st], ph);
        tc_fence_after();
        if (leader) {
          W2_TRACE_MMA(n, 4 * t + 2);
          const uint64_t qd = make_sdesc_sw128(smem_u32(smem + W2_OFF_Q + t * W2_Q_BYTES), 16, 1024);
          const uint64_t kd = make_sdesc_sw128(smem_u32(smem + W2_OFF_K + st * W2_K_BYTES), 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(s_col, qd + 2 * ks, kd + 2 * ks, idesc_s, ks != 0);
          umma_commit(&s_full[t]);
          umma_commit(&q_empty[t]);
          umma_commit(&k_empty[st]);
          W2_TRACE_MMA(n, 4 * t + 3);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    setmaxnreg_dec<40>();  // (all four non-softmax warps have to give their registers back for the increase below)
  } else {
    // ------------------------------------------------------------ softmax / output of query tile t
    setmaxnreg_inc<232>();
    const int t = (warp - 4) >> 2;
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;                 // query row of the tile == TMEM lane (rows >= 98 are unused)
    const int yy = r / 14, x = r - yy * 14;
    const int y = t * 7 + yy;                     // row inside the 14x14 window
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + t * 256;
    const float c1 = p.scale * W2_LOG2E;
    uint8_t* scr = smem + W2_OFF_SCR + t * W2_SCR_STRIDE;
    const int bar_id = 4 + t;                     // named barrier of this warpgroup
    (void)scr;
    (void)bar_id;

    for (int n = 0; n < n_items; ++n) {
      int b, wy, wx, h;
      decode(n, b, wy, wx, h);
      const bool tracer = q4 == 0 && lane == 0;
      if (tracer) W2_TRACE(t, n, 0);
      mbar_wait(&s_full[t], (uint32_t)n & 1u);
      tc_fence_after();
      if (tracer) W2_TRACE(t, n, 1);
      // ---- rel-pos bias of this row: T_h[y - kh + 13], T_w[x - kw + 13] (columns 196..249), via the per-thread scratch
      uint32_t v[2][32];
      tmem_ld32(lane_addr + 192, v[0]);
      tmem_ld32(lane_addr + 224, v[1]);
      tmem_ld_wait();
      float bh[14], bw[14];
      {
        float th[27], tw[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) th[i] = __uint_as_float(v[0][4 + i]);        // S' columns 196..222
        tw[0] = __uint_as_float(v[0][31]);                                         // column 223
#pragma unroll
        for (int i = 1; i < 27; ++i) tw[i] = __uint_as_float(v[1][i - 1]);         // columns 224..249
        window2_pick14(th, y, bh);
        window2_pick14(tw, x, bw);
      }
      // ---- one pass over the 196 scores in 32-column chunks (optimistic exp2 against the running reference maximum,
      // exact redo from registers when a chunk exceeds it by more than 2^TAU; see attn_flash4.cu).  P chunk c (16
      // columns of bf16 pairs) overwrites score columns [16c, 16c+16), which chunk c/2 has already consumed.
      // The pass is issue-bound (two softmax warps per scheduler next to 8 cycles of MUFU per exp2 instruction), so
      // every per-score operation is a packed fp32x2 one: y = s * c1 + (bh[kh] - m_ref) + bw[kw] for two neighbouring
      // keys (same window row: 14 is even) is one FADD2 + one FFMA2 on operands that already sit in register pairs.
      uint64_t bwp[7], bhh[14];
#pragma unroll
      for (int j = 0; j < 7; ++j) bwp[j] = pk2(bw[2 * j], bw[2 * j + 1]);
      const uint64_t c1p = pk2(c1, c1);
      if (tracer) W2_TRACE(t, n, 2);
      if (t == 0 && lane == 0) W2_TRACE_WARP(n, q4);
      tmem_ld32(lane_addr, v[0]);
      tmem_ld_wait();
      // reference maximum = exact maximum of the first 32 scores (cheap: no exp2), so the first chunk never needs a redo
      float m_ref;
      {
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k = 2 * i;
          const uint64_t y2 = fma2(pk2(__uint_as_float(v[0][k]), __uint_as_float(v[0][k + 1])), c1p,
                                   add2(pk2(bh[k / 14], bh[k / 14]), bwp[(k % 14) / 2]));
          float y0, y1;
          unpk2(y2, y0, y1);
          mx[i & 1] = fmaxf(fmaxf(mx[i & 1], y0), y1);
        }
        m_ref = fmaxf(mx[0], mx[1]);
      }
#pragma unroll
      for (int k = 0; k < 14; ++k) bhh[k] = pk2(bh[k] - m_ref, bh[k] - m_ref);
      const bool row_valid = r < 98;  // rows 98..127 of the tile hold whatever the Q buffer held: they must not trigger redos
      uint64_t ls2[2] = {0ull, 0ull};
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        uint32_t(&cur)[32] = v[c & 1];
        if (c < 6) tmem_ld32(lane_addr + (c + 1) * 32, v[(c + 1) & 1]);
        constexpr int kPairsFull = 16;
        const int np = (c < 6) ? kPairsFull : 2;  // valid key pairs in this chunk (keys 192..195 in the last one)
        uint32_t pk[16];
        uint64_t cs2[2];
#pragma unroll 1
        for (int pass = 0;; ++pass) {  // one iteration unless the reference maximum has to be raised (rare)
          float ymax[2] = {-INFINITY, -INFINITY};
          cs2[0] = cs2[1] = 0ull;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < np) {
              const int k = c * 32 + 2 * i;      // key index (compile time): window row k / 14, column pair (k % 14) / 2
              const uint64_t y2 = fma2(pk2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])), c1p,
                                       add2(bhh[k / 14], bwp[(k % 14) / 2]));
              float y0, y1;
              unpk2(y2, y0, y1);
              ymax[i & 1] = fmaxf(fmaxf(ymax[i & 1], y0), y1);
              const float e0 = ex2_approx(y0), e1 = ex2_approx(y1);
              cs2[i & 1] = add2(cs2[i & 1], pk2(e0, e1));
              pk[i] = pack_bf16(e0, e1);
            } else {
              pk[i] = 0u;                        // pad keys 196..207
            }
          }
          const float m_chunk = fmaxf(ymax[0], ymax[1]);  // relative to m_ref
          const bool need = row_valid && m_chunk > W2_TAU;
          if (pass == 0 && __any_sync(0xffffffffu, need)) {
            const float delta = need ? m_chunk : 0.0f;
            const float alpha = ex2_approx(-delta);
            if (c > 0) window2_rescale_p(lane_addr, c, alpha);
            const uint64_t ap = pk2(alpha, alpha), dp = pk2(-delta, -delta);
            ls2[0] = mul2(ls2[0], ap);
            ls2[1] = mul2(ls2[1], ap);
#pragma unroll
            for (int k = 0; k < 14; ++k) bhh[k] = add2(bhh[k], dp);
            continue;
          }
          break;
        }
        ls2[0] = add2(ls2[0], cs2[0]);
        ls2[1] = add2(ls2[1], cs2[1]);
        if (c < 6) {
          tmem_st16(lane_addr + c * 16, pk);
          tmem_ld_wait();  // chunk c + 1 has landed
        } else {
          uint32_t pk8[8];  // keys 192..207
#pragma unroll
          for (int i = 0; i < 8; ++i) pk8[i] = pk[i];
          tmem_st8(lane_addr + 96, pk8);
        }
      }
      float l_row;
      {
        float s0, s1, s2, s3;
        unpk2(ls2[0], s0, s1);
        unpk2(ls2[1], s2, s3);
        l_row = (s0 + s1) + (s2 + s3);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (tracer) W2_TRACE(t, n, 3);
      if (t == 0 && lane == 0) W2_TRACE_WARP(n, 4 + q4);

      // ---- output: O_t / l -> bf16 -> swizzled staging -> 4-D TMA store (the crop to the 64x64 image is the bounds check)
      mbar_wait(&o_full[t], (uint32_t)n & 1u);
      tc_fence_after();
      if (tracer) W2_TRACE(t, n, 4);
      tmem_ld32(lane_addr + W2_COL_O, v[0]);
      tmem_ld32(lane_addr + W2_COL_O + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);  // the MMA warp may overwrite this tile's TMEM columns
      if (tracer) W2_TRACE(t, n, 5);
      const float inv_l = 1.0f / l_row;
#ifdef WM_W2_TMA_STORE
      // (round-1 output path, kept for A/B: swizzled staging + named barrier + 4-D TMA store; ~1000 cycles per item on the
      // tile's serial chain)
      uint8_t* stg = scr + (n & 1) * W2_STG_BYTES;
      const uint32_t st_row = smem_u32(stg) + (uint32_t)r * 128u;
      if (warp == 4 + 4 * t && lane == 0) tma_store_wait_read();
      if (r < 98) {
#pragma unroll
        for (int hv = 0; hv < 2; ++hv)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t(&o)[32] = v[hv];
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + (((uint32_t)(hv * 4 + k) ^ (uint32_t)(r & 7)) << 4)),
                         "r"(pack_bf16(__uint_as_float(o[8 * k]) * inv_l, __uint_as_float(o[8 * k + 1]) * inv_l)),
                         "r"(pack_bf16(__uint_as_float(o[8 * k + 2]) * inv_l, __uint_as_float(o[8 * k + 3]) * inv_l)),
                         "r"(pack_bf16(__uint_as_float(o[8 * k + 4]) * inv_l, __uint_as_float(o[8 * k + 5]) * inv_l)),
                         "r"(pack_bf16(__uint_as_float(o[8 * k + 6]) * inv_l, __uint_as_float(o[8 * k + 7]) * inv_l))
                         : "memory");
          }
      }
      fence_proxy_async();
      named_bar_sync(bar_id, 128);
      if (warp == 4 + 4 * t && lane == 0) {
        tma_store_4d(&tmap_out, stg, h * 64, wx * 14, wy * 14 + t * 7, b);
        tma_store_commit();
      }
#else
      // Every thread writes its own 128-byte output row straight from registers (8 x 16-byte stores): no staging, no
      // proxy fence, no warpgroup barrier, no TMA store on the tile's serial chain (S' -> pass -> P V -> O -> next S').  The crop
      // of the padded 70 x 70 grid back to 64 x 64 is the bounds test.  Measured (batch 32, in-run A/B): 0.269 -> 0.248 ms per
      // launch.  (The same change in the head-dim-80 kernel, 160-byte rows, was SLOWER: 0.826 -> 0.983 ms; it keeps the TMA store.)
      {
        const int gy = wy * 14 + y, gx = wx * 14 + x;
        if (r < 98 && gy < 64 && gx < 64) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)(b * 64 + gy) * 64 + gx) * p.D + h * 64);
#pragma unroll
          for (int hv = 0; hv < 2; ++hv)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t(&o)[32] = v[hv];
              dst[hv * 4 + k] = make_uint4(pack_bf16(__uint_as_float(o[8 * k]) * inv_l, __uint_as_float(o[8 * k + 1]) * inv_l),
                                           pack_bf16(__uint_as_float(o[8 * k + 2]) * inv_l, __uint_as_float(o[8 * k + 3]) * inv_l),
                                           pack_bf16(__uint_as_float(o[8 * k + 4]) * inv_l, __uint_as_float(o[8 * k + 5]) * inv_l),
                                           pack_bf16(__uint_as_float(o[8 * k + 6]) * inv_l, __uint_as_float(o[8 * k + 7]) * inv_l));
            }
        }
      }
#endif
      if (tracer) W2_TRACE(t, n, 6);
    }
#ifdef WM_W2_TMA_STORE
    if (warp == 4 + 4 * t && lane == 0) tma_store_wait_read();
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// tq: box (64, 14, 7, 1) over qkv [B,64,64,3D]; tkv: box (64, 14, 14, 1); trel: box (64, 27) over the [64,64] table;
// tout: box (64, 14, 7, 1) over out [B,64,64,D]
int window2_dispatch(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& trel, const CUtensorMap& tout,
                     const WindowParams& p, int num_sms, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(window2_kernel, W2_SMEM_BYTES, attr_done)) return rc;
  const int items = p.B * 25 * p.H;
  const int grid = items < num_sms ? items : num_sms;
  window2_kernel<<<grid, W2_THREADS, W2_SMEM_BYTES, st>>>(tq, tkv, trel, tout, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

#ifdef WM_F3_TRACE
int window2_read_trace(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_w2_trace, sizeof(g_w2_trace)) == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}
#else
int window2_read_trace(unsigned long long*) { return WM_ERR_ARCH; }
#endif

}  // namespace wm
