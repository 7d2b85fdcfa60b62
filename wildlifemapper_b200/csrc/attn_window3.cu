// Fused 14x14 windowed attention with decomposed rel-pos bias for HEAD DIM 80 (ViT-H), the structure of attn_window2.cu
// (persistent, work item = (image, window, head), S' = Q [K ; Rh ; Rw]^T in one 128 x 256 MMA group per query half, P in
// tensor memory, one MMA issuer warp per half, bias picked in registers, packed fp32x2 softmax pass, TMA output store).
// What head dim 80 changes:
//   * operands are two 64-column (128-byte, SWIZZLE_128B) sub-tiles per row; the second one is loaded 64 wide and only its
//     first 16 columns are used (k-step 4 of the score product; columns 64..79 of V / O);
//   * shared memory then only holds ONE K' / V stage (Q 64 KB + K' 64 KB + V 52 KB + staging 26 KB): the next item's K / V
//     loads start when this item's MMAs have consumed the stage;
//   * O (80 columns) sits at tensor-memory columns 112..191 of the tile (P occupies 0..103; the rest is dead by then);
//   * the output leaves as a TMA store of channels 0..63 of the head plus 32-byte row stores for channels 64..79.
// Replaces the non-pipelined first-generation kernel (attn_window.cu) for ViT-H: 96 -> see DESIGN.md ms per step.
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

#ifdef WM_F3_TRACE
// diagnostics build: SM-clock timeline of CTA 0 (role 0 / 1 = softmax warpgroup of tile 0 / 1, role 2 = MMA warp)
__device__ unsigned long long g_w3_trace[3][64][8];
#define W3_TRACE(role, n, ev)                                                          \
  do {                                                                                 \
    if (blockIdx.x == 0 && (n) < 64) g_w3_trace[role][n][ev] = clock64();              \
  } while (0)
#else
#define W3_TRACE(role, n, ev) do { } while (0)
#endif
#ifdef WM_W3_TRACE_WARPS  // per-warp skew of tile 0 instead of the MMA events: role 2, ev = warp (pass start), 4 + warp (P arrive)
#define W3_TRACE_MMA(n, ev) do { } while (0)
#define W3_TRACE_WARP(n, ev) W3_TRACE(2, n, ev)
#else
#define W3_TRACE_MMA(n, ev) W3_TRACE(2, n, ev)
#define W3_TRACE_WARP(n, ev) do { } while (0)
#endif

constexpr int W3_THREADS = 384;
constexpr float W3_LOG2E = 1.4426950408889634f;
#ifndef WM_W3_TAU
#define WM_W3_TAU 16.0f
#endif
// P = 2^(y - m_ref) may exceed 1 by up to 2^TAU before the reference maximum is raised: bf16 P and the fp32 accumulators
// have the exponent range for it (relative precision is unchanged), and raises become rare even for peaky logits.
constexpr float W3_TAU = WM_W3_TAU;
constexpr int W3_HD = 80;
constexpr int W3_Q_SUB = 16384;                   // 128 rows x 128 B (98 loaded) per sub-tile
constexpr int W3_Q_BYTES = 2 * W3_Q_SUB;
constexpr int W3_K_SUB = 32768;                   // 256 rows: 196 keys, 27 Rh, 27 Rw, 6 zero
constexpr int W3_K_BYTES = 2 * W3_K_SUB;
constexpr int W3_V_SUB = 208 * 128;               // 196 keys + 12 zero rows (K dimension of P V = 13 x 16)
constexpr int W3_V_BYTES = 2 * W3_V_SUB;
constexpr int W3_STG_BYTES = 13312;               // output staging of a tile: 98 rows x 128 B (channels 0..63), rounded up to 1 KB
constexpr int W3_OFF_Q = 0;                                  // [2 tiles][2 sub-tiles]
constexpr int W3_OFF_K = W3_OFF_Q + 2 * W3_Q_BYTES;          // one stage
constexpr int W3_OFF_V = W3_OFF_K + W3_K_BYTES;              // one stage
constexpr int W3_OFF_SCR = W3_OFF_V + W3_V_BYTES;            // [2 tiles] staging, 1024-aligned
constexpr int W3_SCR_STRIDE = W3_STG_BYTES;
constexpr int W3_OFF_BAR = W3_OFF_SCR + 2 * W3_SCR_STRIDE;
constexpr int W3_SMEM_BYTES = W3_OFF_BAR + 256 + 1024;
static_assert(W3_SMEM_BYTES <= 232448, "shared memory budget");
static_assert(W3_OFF_SCR % 1024 == 0 && W3_SCR_STRIDE % 1024 == 0 && W3_V_SUB % 1024 == 0, "alignment");
constexpr int W3_COL_T = 196;   // T_h at S' columns 196..222, T_w at 223..249
constexpr int W3_COL_O = 112;   // O_t (80 columns) over consumed score columns (P occupies 0..103)

// Rare path of the single-pass softmax, kept OUT OF LINE (it sat in the middle of every unrolled chunk before, tripled the
// size of the hot loop and cost ~17 % of the softmax warps' time in instruction-fetch stalls): the reference maximum
// of this row was raised by log2(1/alpha) -- rescale the P chunks already written for it.
__device__ __noinline__ void window3_rescale_p(uint32_t p_addr, int nchunks, float alpha) {
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
__device__ __forceinline__ void window3_pick14(const float (&in)[27], int a, float (&out)[14]) {
  float s8[21], s4[17], s2[15];
  const bool b8 = a & 8, b4 = a & 4, b2 = a & 2, b1 = a & 1;
#pragma unroll
  for (int j = 0; j < 21; ++j) s8[j] = (b8 && j + 8 < 27) ? in[j + 8 < 27 ? j + 8 : 26] : in[j];
#pragma unroll
  for (int j = 0; j < 17; ++j) s4[j] = b4 ? s8[j + 4] : s8[j];
#pragma unroll
  for (int j = 0; j < 15; ++j) s2[j] = b2 ? s4[j + 2] : s4[j];
#pragma unroll
  for (int j = 0; j < 14; ++j) out[13 - j] = (b1 ? s2[j + 1] : s2[j]) * W3_LOG2E;
}

__global__ void __launch_bounds__(W3_THREADS, 1)
window3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
               const __grid_constant__ CUtensorMap tmap_rel, const __grid_constant__ CUtensorMap tmap_out,
               const WindowParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W3_OFF_BAR);
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

  // zero rows that no TMA box ever writes: keys 196..207 of both V sub-tiles, rows 250..255 of both K' sub-tiles
  for (int i = threadIdx.x; i < 2 * (1536 / 16); i += W3_THREADS)
    reinterpret_cast<uint4*>(smem + W3_OFF_V + (i / 96) * W3_V_SUB + 196 * 128)[i % 96] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < 2 * (768 / 16); i += W3_THREADS)
    reinterpret_cast<uint4*>(smem + W3_OFF_K + (i / 48) * W3_K_SUB + 250 * 128)[i % 48] = make_uint4(0, 0, 0, 0);
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
    if (elect_one()) {  // rel-pos tables behind the key rows of both K' sub-tiles (table tensor [64,80]: rows 0..26 Rh, 32..58 Rw;
                        // the second box covers columns 64..127: 64..79 are data, the rest is out of bounds = zero fill)
      mbar_arrive_expect_tx(tab_full, 4 * 27 * 128);
      for (int s = 0; s < 2; ++s) {
        tma_load_2d(smem + W3_OFF_K + s * W3_K_SUB + 196 * 128, &tmap_rel, tab_full, s * 64, 0);
        tma_load_2d(smem + W3_OFF_K + s * W3_K_SUB + 223 * 128, &tmap_rel, tab_full, s * 64, 32);
      }
    }
    __syncwarp();
    for (int n = 0; n < n_items; ++n) {
      int b, wy, wx, h;
      decode(n, b, wy, wx, h);
      const uint32_t ph = (uint32_t)n & 1u;  // one K' / V stage: phase n
      mbar_wait(&k_empty[0], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[0], 2 * 196 * 128);
        for (int s = 0; s < 2; ++s)
          tma_load_4d(smem + W3_OFF_K + s * W3_K_SUB, &tmap_kv, &k_full[0], p.D + h * W3_HD + s * 64, wx * 14, wy * 14, b);
      }
      __syncwarp();
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&q_empty[t], (uint32_t)(n & 1) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[t], 2 * 98 * 128);
          for (int s = 0; s < 2; ++s)
            tma_load_4d(smem + W3_OFF_Q + t * W3_Q_BYTES + s * W3_Q_SUB, &tmap_q, &q_full[t], h * W3_HD + s * 64, wx * 14,
                        wy * 14 + t * 7, b);
        }
        __syncwarp();
      }
      mbar_wait(&v_empty[0], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[0], 2 * 196 * 128);
        for (int s = 0; s < 2; ++s)
          tma_load_4d(smem + W3_OFF_V + s * W3_V_SUB, &tmap_kv, &v_full[0], 2 * p.D + h * W3_HD + s * 64, wx * 14, wy * 14, b);
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
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, W3_HD, 0, 1);  // A = P from TMEM, V MN-major (two 64-column sub-tiles)
    const bool leader = elect_one();
    const uint32_t s_col = tmem_base + t * 256;
    mbar_wait(tab_full, 0);
    for (int n = 0; n <= n_items; ++n) {
      const uint32_t ph = (uint32_t)n & 1u, pph = (uint32_t)(n - 1) & 1u;  // one K' / V stage
      if (n > 0) {
        mbar_wait(&p_full[t], (uint32_t)(n - 1) & 1u);
        mbar_wait(&v_full[0], pph);
        tc_fence_after();
        if (leader) {
          W3_TRACE_MMA(n, 4 * t);
          const uint64_t vd = make_sdesc_sw128(smem_u32(smem + W3_OFF_V), W3_V_SUB, 1024);  // leading-dimension offset = sub-tile stride
#pragma unroll
          for (int ks = 0; ks < 13; ++ks)  // 208 keys, 16 per MMA; P_t: 8 TMEM columns per step; V: 2048 B per step
            umma_bf16_ts(s_col + W3_COL_O, s_col + ks * 8, vd + (uint32_t)(ks * (2048 >> 4)), idesc_pv, ks != 0);
          umma_commit(&o_full[t]);
          umma_commit(&v_empty[0]);
          W3_TRACE_MMA(n, 4 * t + 1);
        }
        __syncwarp();
      }
      if (n < n_items) {
        if (n > 0) mbar_wait(&s_free[t], (uint32_t)(n - 1) & 1u);
        mbar_wait(&q_full[t], (uint32_t)n & 1u);
        mbar_wait(&k_full[0], ph);
        tc_fence_after();
        if (leader) {
          W3_TRACE_MMA(n, 4 * t + 2);
          const uint64_t qd = make_sdesc_sw128(smem_u32(smem + W3_OFF_Q + t * W3_Q_BYTES), 16, 1024);
          const uint64_t kd = make_sdesc_sw128(smem_u32(smem + W3_OFF_K), 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) {  // head dim 80: k-steps 0..3 in sub-tile 0, k-step 4 = first 16 columns of sub-tile 1
            const uint32_t qo = (ks < 4) ? 2 * ks : (W3_Q_SUB >> 4), ko = (ks < 4) ? 2 * ks : (W3_K_SUB >> 4);
            umma_bf16(s_col, qd + qo, kd + ko, idesc_s, ks != 0);
          }
          umma_commit(&s_full[t]);
          umma_commit(&q_empty[t]);
          umma_commit(&k_empty[0]);
          W3_TRACE_MMA(n, 4 * t + 3);
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
    const float c1 = p.scale * W3_LOG2E;
    uint8_t* scr = smem + W3_OFF_SCR + t * W3_SCR_STRIDE;
    const int bar_id = 4 + t;                     // named barrier of this warpgroup

    for (int n = 0; n < n_items; ++n) {
      int b, wy, wx, h;
      decode(n, b, wy, wx, h);
      const bool tracer = q4 == 0 && lane == 0;
      if (tracer) W3_TRACE(t, n, 0);
      mbar_wait(&s_full[t], (uint32_t)n & 1u);
      tc_fence_after();
      if (tracer) W3_TRACE(t, n, 1);
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
        window3_pick14(th, y, bh);
        window3_pick14(tw, x, bw);
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
      if (tracer) W3_TRACE(t, n, 2);
      if (t == 0 && lane == 0) W3_TRACE_WARP(n, q4);
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
          const bool need = row_valid && m_chunk > W3_TAU;
          if (pass == 0 && __any_sync(0xffffffffu, need)) {
            const float delta = need ? m_chunk : 0.0f;
            const float alpha = ex2_approx(-delta);
            if (c > 0) window3_rescale_p(lane_addr, c, alpha);
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
      if (tracer) W3_TRACE(t, n, 3);
      if (t == 0 && lane == 0) W3_TRACE_WARP(n, 4 + q4);

      // ---- output: O_t / l -> bf16 -> swizzled staging -> 4-D TMA store (the crop to the 64x64 image is the bounds check)
      mbar_wait(&o_full[t], (uint32_t)n & 1u);
      tc_fence_after();
      if (tracer) W3_TRACE(t, n, 4);
      uint32_t ox[16];  // channels 64..79
      tmem_ld32(lane_addr + W3_COL_O, v[0]);
      tmem_ld32(lane_addr + W3_COL_O + 32, v[1]);
      tmem_ld16(lane_addr + W3_COL_O + 64, ox);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);  // the MMA warp may overwrite this tile's TMEM columns
      if (tracer) W3_TRACE(t, n, 5);
      const float inv_l = 1.0f / l_row;
      // one staging buffer per tile: the previous item's store (issued a whole item ago) must have finished reading it
      uint8_t* stg = scr;
      const uint32_t st_row = smem_u32(stg) + (uint32_t)r * 128u;
      if (warp == 4 + 4 * t && lane == 0) tma_store_wait_read();
      named_bar_sync(bar_id, 128);
      if (r < 98) {
        // channels 64..79 of the head: two 16-byte row stores (pixels of the padded 70 x 70 grid beyond 64 are cropped)
        const int gy = wy * 14 + y, gx = wx * 14 + x;
        if (gy < 64 && gx < 64) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)(b * 64 + gy) * 64 + gx) * p.D + h * W3_HD + 64);
#pragma unroll
          for (int g = 0; g < 2; ++g)
            dst[g] = make_uint4(pack_bf16(__uint_as_float(ox[8 * g]) * inv_l, __uint_as_float(ox[8 * g + 1]) * inv_l),
                                pack_bf16(__uint_as_float(ox[8 * g + 2]) * inv_l, __uint_as_float(ox[8 * g + 3]) * inv_l),
                                pack_bf16(__uint_as_float(ox[8 * g + 4]) * inv_l, __uint_as_float(ox[8 * g + 5]) * inv_l),
                                pack_bf16(__uint_as_float(ox[8 * g + 6]) * inv_l, __uint_as_float(ox[8 * g + 7]) * inv_l));
        }
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
        tma_store_4d(&tmap_out, stg, h * W3_HD, wx * 14, wy * 14 + t * 7, b);
        tma_store_commit();
      }
      if (tracer) W3_TRACE(t, n, 6);
    }
    if (warp == 4 + 4 * t && lane == 0) tma_store_wait_read();
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
int window3_dispatch(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& trel, const CUtensorMap& tout,
                     const WindowParams& p, int num_sms, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(window3_kernel, W3_SMEM_BYTES, attr_done)) return rc;
  const int items = p.B * 25 * p.H;
  const int grid = items < num_sms ? items : num_sms;
  window3_kernel<<<grid, W3_THREADS, W3_SMEM_BYTES, st>>>(tq, tkv, trel, tout, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

#ifdef WM_F3_TRACE
int window3_read_trace(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_w3_trace, sizeof(g_w3_trace)) == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}
#else
int window3_read_trace(unsigned long long*) { return WM_ERR_ARCH; }
#endif

}  // namespace wm
