// Flash attention v6 on tcgen05 / TMEM (sm_100a): the global 64x64 attention with decomposed rel-pos bias at head dim 64
// (image_encoder.py:246-262, 347-383) with THREE 128-query tiles per CTA.  Same math as v4 (attn_flash4.cu: 64-key
// steps, single-pass lazy softmax against a running reference maximum, P in tensor memory, bias_w in registers, T_h in
// TMEM); what changes is the occupancy of the exp2 pass.  In v4 the two softmax warps of a scheduler issue ~19 % of the
// cycles each (fixed-latency dependency stalls; ptxas emits the 32 exp2 of a chunk as one burst) and the MUFU runs at
// ~50 % (profiles/r01u_ncu_flash4_*.txt); neither the tensor pipe (52 cycles per 128 x 64 x 16 MMA) nor tensor-memory
// bandwidth is the limit (profiles/micro/).  Here every scheduler has three softmax warps:
//
//   * tensor memory: 160 columns per tile = one 64-column score buffer + P (bf16 pairs, 32 columns, NOT aliased with the
//     scores) + O (64).  With P out of the way the tile's MMA issuer writes S_t(j+1) as soon as the softmax warps have READ
//     S_t(j) (mid-pass), i.e. the next scores are waiting when a step ends although there is only one score buffer; P_t(j) V
//     follows when P_t(j) has been stored.  T_h moves out of tensor memory: 64 fp16 values per query row in a private,
//     XOR-swizzled shared-memory row (one conflict-free LDS per step);
//   * registers: 512 threads, setmaxnreg 40 / 152: scores are processed in place in 32-column chunks;
//   * a CTA covers 384 queries (6 image rows); the last CTA of an image runs two tiles.
//
//   warp 0       TMA producer (Q tiles, tables, K and V rings), TMEM allocator
//   warps 1-3    tcgen05.mma issuers of query tiles 0-2 (warp 1 also issues the table products of the prologue)
//   warps 4-15   softmax of query tile (warp - 4) / 4 (thread = one query row)
#include "common.cuh"
#include "wm_internal.h"

namespace wm {

constexpr int F6_THREADS = 512;
constexpr int F6_HD = 64;
constexpr int F6_NS = 3;                 // K / V ring stages (128 keys each)
constexpr int F6_TILE_BYTES = 16384;     // 128 rows x 128 B
constexpr float F6_LOG2E = 1.4426950408889634f;
constexpr float F6_TAU = 16.0f;          // lazy-rescale threshold (log2 units); bf16 P and the fp32 accumulators have the range
constexpr int F6_OFF_Q = 0;                                   // 3 query tiles
constexpr int F6_OFF_K = F6_OFF_Q + 3 * F6_TILE_BYTES;
constexpr int F6_OFF_V = F6_OFF_K + F6_NS * F6_TILE_BYTES;
// prologue tables: Rw [128 rows] + 3 x Rh slice [80 rows]; afterwards fp32 scatter scratch (128 B per softmax thread)
constexpr int F6_OFF_TAB = F6_OFF_V + F6_NS * F6_TILE_BYTES;
constexpr int F6_TAB_BYTES = 49152;
constexpr int F6_OFF_BAR = F6_OFF_TAB + F6_TAB_BYTES;
constexpr int F6_SMEM_BYTES = F6_OFF_BAR + 256 + 1024;
static_assert(16384 + 3 * 10240 <= F6_TAB_BYTES && 384 * 128 <= F6_TAB_BYTES, "table / scratch region");
static_assert(F6_SMEM_BYTES <= 232448, "shared memory budget");
constexpr int F6_TCOLS = 160;            // tensor-memory columns per tile: S 0..63 | P 64..95 | O 96..159
constexpr int F6_COL_P = 64, F6_COL_O = 96;

__global__ void __launch_bounds__(F6_THREADS, 1)
flash6_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_rel,
              const FlashParams p) {
  constexpr int HD = F6_HD, NS = F6_NS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F6_OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]  released by every active tile's issuer
  uint64_t* v_full = bars + 7;    // [3]
  uint64_t* v_empty = bars + 10;  // [3]
  uint64_t* s_full = bars + 13;   // [3 tiles]  S_t(j) complete
  uint64_t* p_full = bars + 16;   // [3 tiles]  P_t(j) stored to TMEM by the 4 warps of tile t
  uint64_t* o_full = bars + 19;   // [3 tiles]
  uint64_t* ta_full = bars + 22;  // T_h products complete ...
  uint64_t* ta_done = bars + 23;  // ... and repacked by the softmax warps
  uint64_t* tb_full = bars + 24;  // T_w products complete ...
  uint64_t* tb_done = bars + 25;  // ... and scattered into registers
  uint64_t* s_read = bars + 26;   // [3 tiles]  S_t(j) loaded into registers by the 4 warps of tile t
  uint64_t* pv_done = bars + 29;  // [3 tiles]  P_t(j) V complete: P_t may be overwritten, O_t is stable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 384, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = min(3, (p.Tq - m0) / 128);  // Tq % 128 == 0 (dispatch)
  const int nk = p.Tk / 128;  // 128-key TMA tiles
  const int ns = p.Tk / 64;   // 64-key softmax / MMA steps

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_rel);
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], ntiles);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], ntiles);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_read[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(ta_full, 1);
    mbar_init(ta_done, 4 * ntiles);
    mbar_init(tb_full, 1);
    mbar_init(tb_done, 4 * ntiles);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      // ------------------------------------------------------------ TMA producer
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, ntiles * (F6_TILE_BYTES + 10240) + 16384);
        for (int t = 0; t < ntiles; ++t)
          tma_load_2d(smem + F6_OFF_Q + t * F6_TILE_BYTES, &tmap_q, q_full, p.q_col0 + h * HD, b * p.Tq + m0 + t * 128);
        // table tensor [256,64]: rows 0..126 rel_pos_h, 128..254 rel_pos_w.  Box = 64 columns x 16 rows.
        const int qi0 = m0 >> 6;  // first image row of this CTA (2 rows per query tile)
        for (int i = 0; i < 8; ++i) tma_load_2d(smem + F6_OFF_TAB + i * 2048, &tmap_rel, q_full, 0, 128 + 16 * i);
        for (int t = 0; t < ntiles; ++t)
          for (int i = 0; i < 5; ++i)
            tma_load_2d(smem + F6_OFF_TAB + 16384 + t * 10240 + i * 2048, &tmap_rel, q_full, 0, qi0 + 2 * t + 16 * i);
      }
      __syncwarp();
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nk; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], F6_TILE_BYTES);
          tma_load_2d(smem + F6_OFF_K + st * F6_TILE_BYTES, &tmap_k, &k_full[st], p.k_col0 + h * HD, b * p.Tk + j * 128);
        }
        __syncwarp();
        mbar_wait(&v_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[st], F6_TILE_BYTES);
          tma_load_2d(smem + F6_OFF_V + st * F6_TILE_BYTES, &tmap_v, &v_full[st], p.v_col0 + h * HD, b * p.Tk + j * 128);
        }
        __syncwarp();
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    } else if (warp - 1 < ntiles) {
      // ------------------------------------------------------------ MMA issuer of query tile t
      const int t = warp - 1;
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);  // A = P from TMEM (K-major), V is MN-major
      const bool leader = elect_one();  // the same lane issues every tcgen05.mma / tcgen05.commit
      const uint32_t sq = smem_u32(smem + F6_OFF_Q);
      const uint32_t tcol = tmem_base + t * F6_TCOLS;
      mbar_wait(q_full, 0);
      tc_fence_after();
      if (t == 0) {
        // Prologue, both phases for all tiles by this warp.  A: T_h(t) = Q_t Rh_slice(t)^T (128 x 80) into the S | O columns,
        // repacked by the softmax warps as fp16 pairs into the T_h columns.  B: T_w(t) = Q_t Rw^T (128 x 128) into the
        // same columns, scattered into registers.  The scatter scratch overlays the tables, hence A before B and one
        // commit per phase for all tiles.
        constexpr uint32_t idesc_th = make_idesc_bf16(128, 80, 0, 0);
        constexpr uint32_t idesc_tw = make_idesc_bf16(128, 128, 0, 0);
        const uint32_t stab = smem_u32(smem + F6_OFF_TAB);
        if (leader) {
          for (int tt = 0; tt < ntiles; ++tt)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tmem_base + tt * F6_TCOLS, make_sdesc_sw128(sq + tt * F6_TILE_BYTES + ks * 32, 16, 1024),
                        make_sdesc_sw128(stab + 16384 + tt * 10240 + ks * 32, 16, 1024), idesc_th, ks != 0);
          umma_commit(ta_full);
        }
        __syncwarp();
        mbar_wait(ta_done, 0);
        tc_fence_after();
        if (leader) {
          for (int tt = 0; tt < ntiles; ++tt)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tmem_base + tt * F6_TCOLS, make_sdesc_sw128(sq + tt * F6_TILE_BYTES + ks * 32, 16, 1024),
                        make_sdesc_sw128(stab + ks * 32, 16, 1024), idesc_tw, ks != 0);
          umma_commit(tb_full);
        }
        __syncwarp();
      }
      mbar_wait(tb_done, 0);  // all tiles have drained the table products out of the S | O columns
      tc_fence_after();
      // S_t(step) = Q_t K(step)^T (128 x 64 x 64) into the tile's score buffer; K rows of step: half (step & 1) of stage
      // (step / 2) % NS.  K tile kt is released once both of its halves have been issued.
      auto issue_s = [&](int step) {
        const int kt = step >> 1;
        if ((step & 1) == 0) {
          mbar_wait(&k_full[kt % NS], (uint32_t)(kt / NS) & 1u);
          tc_fence_after();
        }
        if (leader) {
          const uint64_t qd = make_sdesc_sw128(sq + t * F6_TILE_BYTES, 16, 1024);
          const uint64_t kd = make_sdesc_sw128(smem_u32(smem + F6_OFF_K + (kt % NS) * F6_TILE_BYTES) + (step & 1) * 8192, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tcol, qd + 2 * ks, kd + 2 * ks, idesc_s, ks != 0);
          umma_commit(&s_full[t]);
          if (step & 1) umma_commit(&k_empty[kt % NS]);
        }
        __syncwarp();
      };
      issue_s(0);
      for (int j = 0; j < ns; ++j) {
        if (j + 1 < ns) {  // next scores first: they only need S_t(j) to have been read
          mbar_wait(&s_read[t], (uint32_t)j & 1u);
          tc_fence_after();
          issue_s(j + 1);
        }
        const int vt = j >> 1;
        mbar_wait(&p_full[t], (uint32_t)j & 1u);
        if ((j & 1) == 0) mbar_wait(&v_full[vt % NS], (uint32_t)(vt / NS) & 1u);
        tc_fence_after();
        if (leader) {
          const uint64_t vd = make_sdesc_sw128(smem_u32(smem + F6_OFF_V + (vt % NS) * F6_TILE_BYTES) + (j & 1) * 8192, 16384, 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)  // 64 keys, 16 per MMA; P: 8 TMEM columns per step; V: 2048 B per step
            umma_bf16_ts(tcol + F6_COL_O, tcol + F6_COL_P + ks * 8, vd + (uint32_t)(ks * (2048 >> 4)), idesc_pv, (j | ks) != 0);
          umma_commit(&pv_done[t]);
          if (j == ns - 1) umma_commit(&o_full[t]);
          if (j & 1) umma_commit(&v_empty[vt % NS]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / correction / output
    setmaxnreg_inc<152>();
    const int t = (warp - 4) >> 2;  // query tile of this warpgroup
    if (t < ntiles) {
      const int q4 = warp & 3;        // TMEM lane quarter
      const int r = q4 * 32 + lane;   // query row inside the tile == TMEM lane
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + t * F6_TCOLS;
      const uint32_t s_addr = lane_addr, p_addr = lane_addr + F6_COL_P, o_addr = lane_addr + F6_COL_O;
      const float c1 = p.scale * F6_LOG2E;
      const uint64_t c1p = pk2(c1, c1);
      const int hi = r >> 6;  // image row of this query inside the tile (warp-uniform): bias_h[kh] = T_h[hi + 63 - kh]
      float tw[64];
      float bh64;             // T_h[64] (only needed by key row 0 of the lower image row)
      uint32_t* th_row = reinterpret_cast<uint32_t*>(smem + F6_OFF_TAB) + ((t * 128 + r) << 5);  // private 128-byte row
      {
        // ---- phase A: T_h(t)[0..64] sits in columns 0..64 as fp32: keep [0..63] as fp16 pairs (x log2 e), [64] in a register
        mbar_wait(ta_full, 0);
        tc_fence_after();
        uint32_t a[32], pkh[32];
        tmem_ld32(s_addr, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) pkh[i] = pack_f16(__uint_as_float(a[2 * i]) * F6_LOG2E, __uint_as_float(a[2 * i + 1]) * F6_LOG2E);
        tmem_ld32(s_addr + 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) pkh[16 + i] = pack_f16(__uint_as_float(a[2 * i]) * F6_LOG2E, __uint_as_float(a[2 * i + 1]) * F6_LOG2E);
        const uint32_t x = tmem_ld1(s_addr + 64);
        tmem_ld_wait();
        bh64 = __uint_as_float(x) * F6_LOG2E;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ta_done);
        // ---- phase B: bias_w[kw] = T_w[qj + 63 - kw]: per-thread scatter through the private (XOR-swizzled) smem row
        const int qj = (m0 + t * 128 + r) & 63;
        mbar_wait(tb_full, 0);
        tc_fence_after();
        float* scr = reinterpret_cast<float*>(th_row);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld32(s_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int kw = qj + 63 - (c * 32 + i) - half * 32;
              if (kw >= 0 && kw < 32) scr[((((kw >> 2) ^ (r & 7)) << 2) | (kw & 3))] = __uint_as_float(v[i]) * F6_LOG2E;
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
        if (lane == 0) mbar_arrive(tb_done);
        // the row now keeps T_h: word w (fp16 pair T_h[2w], T_h[2w+1]) at position w ^ lane -- the per-step lookup reads
        // the same w in every lane of a warp, i.e. 32 different banks
#pragma unroll
        for (int w = 0; w < 32; ++w) th_row[w ^ lane] = pkh[w];
        __syncwarp();
      }
      float m_ref = -INFINITY, l_run = 0.0f;
      for (int j = 0; j < ns; ++j) {  // one step = 64 keys = key row j of the 64x64 grid
        float bh;
        {
          const int c0 = hi + 63 - j;  // T_h index of key row j (warp-uniform)
          const uint32_t a0 = th_row[(c0 > 63 ? 31 : (c0 >> 1)) ^ lane];
          bh = (c0 > 63) ? bh64 : unpack_f16(a0, c0 & 1);
        }
        mbar_wait(&s_full[t], (uint32_t)j & 1u);
        tc_fence_after();
        uint64_t ls2[2] = {0ull, 0ull};  // this step's row sum (relative to m_ref)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // scores of this chunk, processed in place: v[2i], v[2i+1] -> exp2 arguments -> probabilities -> bf16 pairs
          uint32_t v[32];
          tmem_ld32(s_addr + c * 32, v);
          tmem_ld_wait();
          if (c == 1) {  // both halves of S_t(j) are in registers: the issuer may overwrite the buffer with S_t(j+1)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_read[t]);
          }
          float ymax[2] = {-INFINITY, -INFINITY};
          uint64_t y2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            y2[i] = fma2(pk2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), c1p, pk2(tw[c * 32 + 2 * i], tw[c * 32 + 2 * i + 1]));
            float y0, y1;
            unpk2(y2[i], y0, y1);
            ymax[i & 1] = fmaxf(fmaxf(ymax[i & 1], y0), y1);
          }
          const float m_chunk = fmaxf(ymax[0], ymax[1]) + bh;
          const bool need = m_chunk > m_ref + F6_TAU;  // always on the very first chunk (m_ref = -inf)
          if (__any_sync(0xffffffffu, need)) {
            // ---- raise the reference maximum (rare after the first chunk of a row).  O must be stable: P_t(j-1) V is the last
            // MMA that touches O_t before p_full(j); its completion is pv_done(j-1).
            const float m_new = need ? m_chunk : m_ref;
            const float alpha = ex2_approx(m_ref - m_new);  // 1 for lanes that did not need it, 0 on the very first chunk
            if (j > 0) {
              if (c == 0) {  // (chunk 1 comes after the wait below)
                mbar_wait(&pv_done[t], (uint32_t)(j - 1) & 1u);
                tc_fence_after();
              }
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
              tmem_ld16(p_addr, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float lo = __uint_as_float(o[i] << 16) * alpha, hi2 = __uint_as_float(o[i] & 0xffff0000u) * alpha;
                o[i] = pack_bf16(lo, hi2);
              }
              tmem_st16(p_addr, o);
              const uint64_t ap = pk2(alpha, alpha);
              ls2[0] = mul2(ls2[0], ap);
              ls2[1] = mul2(ls2[1], ap);
            }
            m_ref = m_new;
          }
          const float d = bh - m_ref;
          const uint64_t dp = pk2(d, d);
          uint64_t cs2[2] = {0ull, 0ull};
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a0, a1;
            unpk2(add2(y2[i], dp), a0, a1);
            const float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
            cs2[i & 1] = add2(cs2[i & 1], pk2(e0, e1));
            pk[i] = pack_bf16(e0, e1);
          }
          ls2[0] = add2(ls2[0], cs2[0]);
          ls2[1] = add2(ls2[1], cs2[1]);
          if (c == 0 && j > 0) {  // P_t(j-1) V must have consumed the P columns (issued ~a pass ago: normally complete)
            mbar_wait(&pv_done[t], (uint32_t)(j - 1) & 1u);
            tc_fence_after();
          }
          tmem_st16(p_addr + c * 16, pk);
        }
        {
          float s0, s1, s2, s3;
          unpk2(ls2[0], s0, s1);
          unpk2(ls2[1], s2, s3);
          l_run += (s0 + s1) + (s2 + s3);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // ---- epilogue: O / l
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      __nv_bfloat16* dst = p.out + (size_t)(b * p.Tq + m0 + t * 128 + r) * p.ldo + h * HD;
#pragma unroll
      for (int c = 0; c < HD / 16; ++c) {
        uint32_t o[16];
        tmem_ld16(o_addr + c * 16, o);
        tmem_ld_wait();
        uint4* d4 = reinterpret_cast<uint4*>(dst + c * 16);
#pragma unroll
        for (int g = 0; g < 2; ++g)
          d4[g] = make_uint4(pack_bf16(__uint_as_float(o[8 * g]) * inv_l, __uint_as_float(o[8 * g + 1]) * inv_l),
                             pack_bf16(__uint_as_float(o[8 * g + 2]) * inv_l, __uint_as_float(o[8 * g + 3]) * inv_l),
                             pack_bf16(__uint_as_float(o[8 * g + 4]) * inv_l, __uint_as_float(o[8 * g + 5]) * inv_l),
                             pack_bf16(__uint_as_float(o[8 * g + 6]) * inv_l, __uint_as_float(o[8 * g + 7]) * inv_l));
      }
      tc_fence_before();
    }
  }

  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// q / k / v tiles: box 128 rows x 64 columns; rel table [256,64]: box 16 rows
int flash6_dispatch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                    const FlashParams& p, int hd, cudaStream_t st) {
  if (hd != F6_HD || !p.use_relpos || p.Tq != 4096 || p.Tk != 4096) return WM_ERR_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(flash6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F6_SMEM_BYTES) != cudaSuccess)
      return WM_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid((p.Tq + 383) / 384, p.H, p.B);
  flash6_kernel<<<grid, F6_THREADS, F6_SMEM_BYTES, st>>>(tq, tk, tv, trel, p);
  return cudaGetLastError() == cudaSuccess ? WM_OK : WM_ERR_CUDA;
}

}  // namespace wm
