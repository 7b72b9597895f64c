// attention_tc.cu — softmax(q kᵀ·scale + mask)·v forward AND backward on the Blackwell tensor cores, scores never in HBM.
//
// The reference materialises the [B,H,Lq,Lk] score tensor and ≥5 same-sized temporaries (llm/llama/model.py:112-121,
// examples/pydynet/transformer.py:93-104: 1.07 GB each at BASELINE config 4). Here one kernel template serves five passes
// that all have the same shape — "score MMA(s) into TMEM → element-wise in registers → bf16 hi/lo operand back INTO TMEM →
// accumulate MMA into TMEM":
//     LSE  rows = queries   S = Q·Kᵀ                      row-wise log-sum-exp (online), nothing accumulated
//     FWD  rows = queries   S = Q·Kᵀ                      P = exp(S − lse)              O  += P · V
//     DQ   rows = queries   S = Q·Kᵀ, dP = dO·Vᵀ          dS = P∘(dP − Δ)·scale         dQ += dS · K
//     DV   rows = keys      Sᵀ = K·Qᵀ                     Pᵀ = exp(Sᵀ − lse)            dV += Pᵀ · dO
//     DK   rows = keys      Sᵀ = K·Qᵀ, dPᵀ = V·dOᵀ        dSᵀ = Pᵀ∘(dPᵀ − Δ)·scale      dK += dSᵀ · Q
// Because lse is known before P is formed (LSE pass first), no accumulator is ever rescaled; the price is recomputing S.
//
// Operand placement (what the PDN_TC_TRACE timelines led to):
//   * the resident row operand (Q rows; K rows in the key-row passes) is copied once from its TMA-filled shared-memory tile into
//     TMEM and read there by the score MMAs (TS-mode tcgen05.mma): an SS-mode 128x64x16 MMA re-reads 4 KB of A from shared memory
//     every 32 cycles, more than the 128 B/clk the SM has;
//   * P / dS are split into bf16 hi/lo on the integer pipe (the XU pipe is saturated by ex2) and written with tcgen05.st into two
//     TMEM operand buffers; the accumulate MMA reads them as its A operand (TS-mode), no shared-memory round trip;
//   * the accumulate MMA's B operand (V, K, Q or dO tile, [64 items][d]) is consumed MN-major exactly as TMA loads it from the
//     tensors' row-major planes: no transposed copies of V / K / Q / dO exist;
//   * score operands and accumulate operands have separate full/empty barrier rings over 3-4 stages, so the TMA producer runs
//     ahead of the softmax instead of waiting for the accumulate MMA of the previous tile.
// fp32 parity comes from the same BF16x3 split as the GEMM (hi·hi + hi·lo + lo·hi), applied to P/dS as well.
//
// CTA = one 128-row tile of one (batch, head); warp 0 TMA producer, warp 1 MMA issuer (both walk their loops on all lanes and
// predicate the issuing instructions with elect.sync: uniform-register operands), warp 2 TMEM allocator, warps 4-11 = 256 softmax
// threads (two per row: a 32-column half each, row max / sum combined through shared memory only in the LSE pass).
// TMEM columns: S1[2] 0/64, S2[2] 128/192, accumulator 256, P/dS buffers 320/384 (hi 32 + lo 32 each), row operand 448 (hi 32 + lo 32).
#include "common.cuh"
#include "gemm_tc.h"
#include "tc_ptx.cuh"
#include <cudaTypedefs.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>

namespace pdn {

enum { AT_LSE = 0, AT_FWD = 1, AT_DQ = 2, AT_DV = 3, AT_DK = 4 };
constexpr int AT_R = 128;  // row tile
constexpr int AT_C = 64;   // column tile
constexpr int AT_KA = 2 * AT_R * 64 * 2;  // resident row operand, hi + lo: 32 KB
constexpr int AT_KB = 2 * AT_C * 64 * 2;  // one column-tile operand, hi + lo: 16 KB
constexpr int AT_KP = 2 * AT_R * AT_C * 2;  // P / dS tile, hi + lo: 32 KB
constexpr int AT_STAT = 2048;               // query columns whose statistics fit the shared-memory staging of the key-row passes

struct AtArgs {
  int64_t rows, cols;  // (Lq, Lk) in the query-row passes, (Lk, Lq) in the key-row passes
  int64_t H;
  int     D;           // real head dim (<= 64)
  float   scale;
  const float* mask;   // additive, element (b, query, key) at b*mask_bs + query*mask_qs + key ; nullptr = none
  int64_t mask_bs, mask_qs;
  const float* lse;    // [BH][Lq] (input of FWD / DQ / DV / DK)
  const float* delta;  // [BH][Lq] Σ_d dO·O (DQ, DK)
  float* lse_out;      // [BH][Lq] (LSE)
  float* out;          // [B, rows, H, D] fp32 (FWD: O, DQ: dQ, DV: dV, DK: dK)
  int ncol_tiles;
  int prefetch_ahead;  // linear CTA distance of the L2 prefetch of the row operand (0 = off)
  long long* trace;  // debug (PDN_TC_TRACE): clock64 stamps of CTA (0,0)
};

template <int MODE>
struct AtCfg {
  static constexpr bool TWO = (MODE == AT_DQ || MODE == AT_DK);
  static constexpr bool ACC = (MODE != AT_LSE);
  static constexpr bool TRANS = (MODE == AT_DV || MODE == AT_DK);
  static constexpr int  kNst = TWO ? 3 : 4;  // operand stages (P / dS live in tensor memory, so shared memory holds operands only)
  static constexpr int  kStage = AT_KB * (1 + (TWO ? 1 : 0) + (ACC ? 1 : 0));
  static constexpr int  kOffStages = AT_KA * (TWO ? 2 : 1);
  static constexpr int  kOffBars = kOffStages + kNst * kStage;
  // key-row passes: lse (and delta) of ALL query columns of this (batch, head) staged once in shared memory
  static constexpr int  kOffStat = kOffBars + 512;
  static constexpr int  kStat = TRANS ? (TWO ? 2 : 1) * AT_STAT * 4 : 0;
  static constexpr int  kSmem = kOffStat + kStat + 1024;
};

// 16-byte chunk c (8 bf16) of row r in a K-major [rows x 64] bf16 tile with 128-byte swizzle (what TMA writes / UMMA reads)
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

// two fp32 -> packed bf16x2 (lo in the low half), round-to-nearest-even: one F2FP instruction
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
#ifndef PDN_AT_PACKED_CVT
#define PDN_AT_PACKED_CVT 1
#endif
constexpr bool kPackedCvt = PDN_AT_PACKED_CVT != 0;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void umma3(uint32_t d, uint64_t ahi, uint64_t alo, uint64_t bhi, uint64_t blo, uint32_t idesc, bool first_clears) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
    umma_bf16(d, alo + adv, bhi + adv, idesc, (first_clears && k == 0) ? 0u : 1u);
    umma_bf16(d, ahi + adv, blo + adv, idesc, 1u);
    umma_bf16(d, ahi + adv, bhi + adv, idesc, 1u);
  }
}

// accumulate MMA with the A operand (P / dS hi and lo planes, 32 columns each) in tensor memory
__device__ __forceinline__ void umma3_ts(uint32_t d, uint32_t ahi, uint32_t alo, uint64_t bhi, uint64_t blo, uint32_t idesc, bool first_clears) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t adv = (uint64_t)((k * 16 * 128) >> 4);  // MN-major B: 16 k rows of 128 B
    const uint32_t ac = (uint32_t)(k * 8);                 // 16 bf16 = 8 columns
    umma_bf16_ts(d, alo + ac, bhi + adv, idesc, (first_clears && k == 0) ? 0u : 1u);
    umma_bf16_ts(d, ahi + ac, blo + adv, idesc, 1u);
    umma_bf16_ts(d, ahi + ac, bhi + adv, idesc, 1u);
  }
}

// score MMA with the A operand (row operand hi / lo planes) in tensor memory and a K-major B tile in shared memory
__device__ __forceinline__ void umma3_ts_k(uint32_t d, uint32_t ahi, uint32_t alo, uint64_t bhi, uint64_t blo, uint32_t idesc, bool first_clears) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
    const uint32_t ac = (uint32_t)(k * 8);
    umma_bf16_ts(d, alo + ac, bhi + adv, idesc, (first_clears && k == 0) ? 0u : 1u);
    umma_bf16_ts(d, ahi + ac, blo + adv, idesc, 1u);
    umma_bf16_ts(d, ahi + ac, bhi + adv, idesc, 1u);
  }
}

template <int MODE>
__global__ void __launch_bounds__(384, 1)
k_attn_tc(const __grid_constant__ CUtensorMap mA1, const __grid_constant__ CUtensorMap mA2, const __grid_constant__ CUtensorMap mB1,
          const __grid_constant__ CUtensorMap mB2, const __grid_constant__ CUtensorMap mB3, AtArgs a) {
  using Cfg = AtCfg<MODE>;
  constexpr bool TWO = Cfg::TWO, ACC = Cfg::ACC, TRANS = Cfg::TRANS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t*  smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + Cfg::kOffBars);
  // Two operand rings share the stage memory but not the barriers: the score operands of tile j (K, and V for dP) are released by
  // the score MMAs, the accumulate operand (V / Kt / Qt / dOt) only by the accumulate MMA one softmax later. With one ring the TMA
  // load of tile j+1 could not start before accumulate(j-1) had finished and every iteration paid a full TMA round trip
  // (first PDN_TC_TRACE timeline: 3300-cycle period for 1500 cycles of softmax work).
  constexpr int NST = Cfg::kNst;
  uint64_t *a_full = bars, *acc_full = bars + 1, *s_full = bars + 2, *s_empty = bars + 4, *p_full = bars + 6, *p_empty = bars + 8,
           *sc_full = bars + 10, *sc_empty = sc_full + NST, *ac_full = sc_empty + NST, *ac_empty = ac_full + NST;
  uint64_t* a_tmem = ac_empty + NST;  // the resident row operand (Q / K rows) has been copied from shared memory to tensor memory
  uint32_t* tmem_slot = (uint32_t*)(a_tmem + 1);

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const bool tr = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
#define AT_STAMP(i) do { if (tr) a.trace[i] = clock64(); } while (0)
  if (warp == 0) AT_STAMP(0);
  const int bh = blockIdx.y;
  const int64_t row0 = (int64_t)blockIdx.x * AT_R;
  const int n = a.ncol_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mA1);
    tma_prefetch_desc(&mB1);
    if (TWO) { tma_prefetch_desc(&mA2); tma_prefetch_desc(&mB2); }
    if (ACC) tma_prefetch_desc(&mB3);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_tmem, 8);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&sc_full[i], 1);
      mbar_init(&sc_empty[i], 1);
      mbar_init(&ac_full[i], 1);
      mbar_init(&ac_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);  // one arrival per softmax warp
      mbar_init(&p_full[i], 8);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  constexpr uint32_t kTmemCols = ACC ? 512u : 256u;  // LSE: two score buffers + the row operand
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) AT_STAMP(1);
  // TMEM columns: S1[2] at 0/64, S2[2] at 128/192, accumulator at 256, P / dS operand buffers [2] at 320/384 (hi 32 columns, lo 32)
  auto tmS1 = [&](int i) { return tmem_base + (uint32_t)(i * 64); };
  auto tmS2 = [&](int i) { return tmem_base + (uint32_t)(128 + i * 64); };
  auto tmP = [&](int i) { return tmem_base + (uint32_t)(320 + i * 64); };
  const uint32_t tmACC = tmem_base + 256u;
  // the first row operand (Q rows, or K rows in the key-row passes) as bf16 hi / lo A planes of the score MMA: 32 + 32 columns.
  // A from tensor memory costs no shared-memory bandwidth (SS-mode re-reads the 128 x 16 A slice for every MMA: 6 KB per 32-cycle
  // MMA against 128 B/clk) — measured 65 vs 37 cycles per 128x64x16 MMA.
  const uint32_t tmA1 = tmem_base + (ACC ? 448u : 128u);

  if (warp == 0) {
    // ===== TMA producer (whole warp walks the loop, the elected lane issues) =====
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(a_full, AT_KA * (TWO ? 2 : 1));
      tma_load_4d(&mA1, a_full, smem, 0, (int)row0, 0, bh);
      tma_load_4d(&mA1, a_full, smem + AT_KA / 2, 0, (int)row0, 1, bh);
      if (TWO) {
        tma_load_4d(&mA2, a_full, smem + AT_KA, 0, (int)row0, 0, bh);
        tma_load_4d(&mA2, a_full, smem + AT_KA + AT_KA / 2, 0, (int)row0, 1, bh);
      }
    }
    // The row operand of a CTA is read by nobody else, so its TMA load misses L2 and the CTA (one per SM, nothing to overlap
    // with) waits 1.5-2.5 us of its ~10 for it. CTAs start in linear order as SMs free up: the CTA one wave ahead prefetches this
    // tile into L2 at ITS start.
    if (leader && a.prefetch_ahead > 0) {
      const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + a.prefetch_ahead;
      if (lin < (long long)gridDim.x * gridDim.y) {
        const int pbh = (int)(lin / gridDim.x), prow = (int)(lin % gridDim.x) * AT_R;
        tma_prefetch_l2_4d(&mA1, 0, prow, 0, pbh);
        tma_prefetch_l2_4d(&mA1, 0, prow, 1, pbh);
        if (TWO) {
          tma_prefetch_l2_4d(&mA2, 0, prow, 0, pbh);
          tma_prefetch_l2_4d(&mA2, 0, prow, 1, pbh);
        }
      }
    }
    constexpr int kScoreBytes = AT_KB * (TWO ? 2 : 1);
    auto load_score = [&](int j) {  // K (and V for dP) rows of column tile j
      const int st = j % NST;
      mbar_wait(&sc_empty[st], (uint32_t)(((j / NST) & 1) ^ 1));
      if (leader) {
        uint8_t* sb = smem + Cfg::kOffStages + st * Cfg::kStage;
        mbar_expect_tx(&sc_full[st], kScoreBytes);
        const int col0 = j * AT_C;
        tma_load_4d(&mB1, &sc_full[st], sb, 0, col0, 0, bh);
        tma_load_4d(&mB1, &sc_full[st], sb + AT_KB / 2, 0, col0, 1, bh);
        if (TWO) {
          tma_load_4d(&mB2, &sc_full[st], sb + AT_KB, 0, col0, 0, bh);
          tma_load_4d(&mB2, &sc_full[st], sb + AT_KB + AT_KB / 2, 0, col0, 1, bh);
        }
      }
    };
    auto load_acc = [&](int j) {  // [column items (k of the accumulate MMA)][d]: consumed as an MN-major B operand, no transposed copy
      const int st = j % NST;
      mbar_wait(&ac_empty[st], (uint32_t)(((j / NST) & 1) ^ 1));
      if (leader) {
        uint8_t* sb = smem + Cfg::kOffStages + st * Cfg::kStage + kScoreBytes;
        mbar_expect_tx(&ac_full[st], AT_KB);
        const int col0 = j * AT_C;
        tma_load_4d(&mB3, &ac_full[st], sb, 0, col0, 0, bh);
        tma_load_4d(&mB3, &ac_full[st], sb + AT_KB / 2, 0, col0, 1, bh);
      }
    };
    if (n > 0) load_score(0);
    for (int j = 0; j < n; ++j) {
      if (j + 1 < n) load_score(j + 1);
      if (ACC) load_acc(j);
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp walks the loop, the elected lane issues) =====
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(AT_R, AT_C);  // M = 128, N = 64 for the score and the accumulate MMAs alike
      mbar_wait(a_full, 0);
      mbar_wait(a_tmem, 0);
      tc_fence_after();
      AT_STAMP(2);
      const uint32_t sA = smem_u32(smem);
      const uint64_t a2hi = make_smem_desc_sw128(sA + AT_KA), a2lo = make_smem_desc_sw128(sA + AT_KA + AT_KA / 2);
      auto score = [&](int j) {
        const int st = j % NST, sb = j & 1;
        mbar_wait(&sc_full[st], (uint32_t)((j / NST) & 1));
        if (j == 4) AT_STAMP(25);
        mbar_wait(&s_empty[sb], (uint32_t)(((j >> 1) & 1) ^ 1));
        tc_fence_after();
        if (j == 4) AT_STAMP(26);
        const uint32_t sB = smem_u32(smem + Cfg::kOffStages + st * Cfg::kStage);
        if (leader) {
          umma3_ts_k(tmS1(sb), tmA1, tmA1 + 32u, make_smem_desc_sw128(sB), make_smem_desc_sw128(sB + AT_KB / 2), idesc, true);
          if (TWO) umma3(tmS2(sb), a2hi, a2lo, make_smem_desc_sw128(sB + AT_KB), make_smem_desc_sw128(sB + AT_KB + AT_KB / 2), idesc, true);
          umma_commit(&s_full[sb]);
          umma_commit(&sc_empty[st]);
        }
        if (j == 4) AT_STAMP(27);
      };
      auto accumulate = [&](int j) {
        const int st = j % NST, pb = j & 1;
        mbar_wait(&ac_full[st], (uint32_t)((j / NST) & 1));
        if (j == 2) AT_STAMP(28);
        mbar_wait(&p_full[pb], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        if (j == 2) AT_STAMP(23);
        const uint32_t sB3 = smem_u32(smem + Cfg::kOffStages + st * Cfg::kStage + AT_KB * (TWO ? 2 : 1));
        if (leader) {
          umma3_ts(tmACC, tmP(pb), tmP(pb) + 32u, make_smem_desc_sw128_mn(sB3), make_smem_desc_sw128_mn(sB3 + AT_KB / 2),
                   idesc | IDESC_B_MN_MAJOR, j == 0);
          umma_commit(&p_empty[pb]);
          umma_commit(&ac_empty[st]);
        }
        if (j == 2) AT_STAMP(24);
      };
      if (n > 0) score(0);
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) score(j + 1);
        if (ACC) accumulate(j);
      }
      if (ACC && leader) umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ===== softmax / epilogue: two threads per row (8 warps): warp w owns TMEM lane quarter w%4 and column half (w-4)/4 =====
    const int q = warp & 3, half = (warp - 4) >> 2, r = q * 32 + lane;
    const int64_t gr = row0 + r;
    const bool    row_ok = gr < a.rows;
    const int64_t b = bh / a.H, h = bh % a.H;
    const uint32_t lane_sel = ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
    const int64_t Lq = TRANS ? a.cols : a.rows;
    float lse_r = 0.f, delta_r = 0.f;
    if (!TRANS && MODE != AT_LSE && row_ok) {
      lse_r = a.lse[(int64_t)bh * Lq + gr];
      if (TWO) delta_r = a.delta[(int64_t)bh * Lq + gr];
    }
    {
      // copy this thread's row of the first row operand from (swizzled) shared memory into tensor memory: the warps of column half
      // 0 take the hi plane, those of half 1 the lo plane (32 columns = 64 bf16 each)
      mbar_wait(a_full, 0);
      const uint8_t* arow = smem + half * (AT_KA / 2);
      uint32_t w[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(arow + sw128_off(r, c));
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
      }
      const uint32_t adst = tmA1 + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
      tmem_st_32x16(adst, w);
      tmem_st_32x16(adst + 16u, w + 16);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_tmem);
    }
    // Key-row passes read lse / delta per COLUMN: staged once per CTA (zero-filled to the tile boundary) instead of 8-16 global
    // 128-bit loads per thread and tile in the middle of the softmax chain (DV / DK spent 1500-2400 cycles per tile in this phase
    // against 700-800 in the query-row passes, PDN_TC_TRACE).
    float*     stat = reinterpret_cast<float*>(smem + Cfg::kOffStat);
    const bool stat_smem = TRANS && (int64_t)n * AT_C <= AT_STAT;
    if (stat_smem) {
      const int tid = (int)threadIdx.x - 128, ncols = (int)a.cols;
      for (int i = tid; i < n * AT_C; i += 256) {
        stat[i] = i < ncols ? __ldg(a.lse + (int64_t)bh * Lq + i) : 0.f;
        if (TWO) stat[AT_STAT + i] = i < ncols ? __ldg(a.delta + (int64_t)bh * Lq + i) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      if (warp == 4 && j < 4) AT_STAMP(3 + 4 * j);
      float s[32], dp[TWO ? 32 : 1];
      tmem_ld_32x32(tmS1(sb) + lane_sel, s);
      if (TWO) tmem_ld_32x32(tmS2(sb) + lane_sel, dp);
      tmem_ld_wait();
      if (warp == 4 && j < 4) AT_STAMP(4 + 4 * j);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
      const int col0 = j * AT_C + half * 32;
      const int ncols = (int)a.cols;
      // Everything below works in the log2 domain: t = (s*scale + mask) * log2(e), so that exp(x - lse) is ONE ex2.approx of
      // one FFMA result. Column validity only matters in the last tile (uniform branch); masks take the slower path.
      const float c1 = a.scale * 1.4426950408889634f;
      // without a mask the scale is folded into the exponent FFMA of the P passes (cs = c1); the LSE pass and masked problems
      // scale here (cs = 1)
      const bool  prescale = a.mask != nullptr || MODE == AT_LSE;
      const float cs = prescale ? 1.f : c1;
      if (a.mask) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int c = col0 + e;
          float mk = 0.f;
          if (c < ncols && row_ok) {
            const int64_t qi = TRANS ? c : gr, ki = TRANS ? gr : c;
            mk = __ldg(a.mask + b * a.mask_bs + qi * a.mask_qs + ki);
          }
          s[e] = fmaf(s[e], c1, mk * 1.4426950408889634f);
        }
      } else if (MODE == AT_LSE) {
#pragma unroll
        for (int e = 0; e < 32; ++e) s[e] *= c1;
      }
      if (col0 + 32 > ncols) {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (col0 + e >= ncols) s[e] = -INFINITY;
      }
      if (MODE == AT_LSE) {
        float tm = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; ++e) tm = fmaxf(tm, s[e]);
        const float mn = fmaxf(m_run, tm);
        if (mn != -INFINITY) {
          float sum = 0.f;
#pragma unroll
          for (int e = 0; e < 32; ++e) sum += ex2f(s[e] - mn);
          l_run = l_run * ex2f(m_run - mn) + sum;
          m_run = mn;
        }
      } else {
        // value that becomes the A operand of the accumulate MMA: P = 2^(t - lse2) [ * (dP - delta) * scale ]
        if (TRANS) {  // per-column statistics of this tile half (L1-resident; 128-bit loads when aligned)
          const float* lp = a.lse + (int64_t)bh * Lq + col0;
          const float* dpn = TWO ? a.delta + (int64_t)bh * Lq + col0 : nullptr;
          float cl[32], cd[TWO ? 32 : 1];
          if (stat_smem) {
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const float4 t = reinterpret_cast<const float4*>(stat + col0)[e4];
              cl[e4 * 4] = t.x; cl[e4 * 4 + 1] = t.y; cl[e4 * 4 + 2] = t.z; cl[e4 * 4 + 3] = t.w;
              if (TWO) {
                const float4 u = reinterpret_cast<const float4*>(stat + AT_STAT + col0)[e4];
                cd[e4 * 4] = u.x; cd[e4 * 4 + 1] = u.y; cd[e4 * 4 + 2] = u.z; cd[e4 * 4 + 3] = u.w;
              }
            }
          } else if (col0 + 32 <= ncols && (Lq & 3) == 0) {
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(lp) + e4);
              cl[e4 * 4] = t.x; cl[e4 * 4 + 1] = t.y; cl[e4 * 4 + 2] = t.z; cl[e4 * 4 + 3] = t.w;
              if (TWO) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(dpn) + e4);
                cd[e4 * 4] = u.x; cd[e4 * 4 + 1] = u.y; cd[e4 * 4 + 2] = u.z; cd[e4 * 4 + 3] = u.w;
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const bool ok = col0 + e < ncols;
              cl[e] = ok ? __ldg(lp + e) : 0.f;
              if (TWO) cd[e] = ok ? __ldg(dpn + e) : 0.f;
            }
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float p = ex2f(fmaf(s[e], cs, cl[e] * -1.4426950408889634f));
            if (TWO) p = p * (dp[e] - cd[e]) * a.scale;
            s[e] = p;
          }
        } else {
          const float nlse2 = lse_r * -1.4426950408889634f;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float p = ex2f(fmaf(s[e], cs, nlse2));
            if (TWO) p = p * (dp[e] - delta_r) * a.scale;
            s[e] = p;
          }
        }
        const int pb = j & 1;
        if (warp == 4 && j < 4) AT_STAMP(5 + 4 * j);
        mbar_wait(&p_empty[pb], (uint32_t)(((j >> 1) & 1) ^ 1));
        // bf16 hi/lo split on the integer pipe (the XU pipe is saturated by ex2): hi = round-to-nearest of the upper 16 bits
        // (+0x8000 then mask), lo = the same rounding of the exact remainder v - hi. The packed pairs go straight to tensor memory
        // (row = lane, two keys per 32-bit column): the accumulate MMA reads its A operand there.
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float v0 = s[2 * t], v1 = s[2 * t + 1];
          if (kPackedCvt) {  // 6 instructions per pair: packed convert, two unpacks, two subtractions, packed convert
            const uint32_t hp = cvt_bf16x2(v0, v1);
            hw[t] = hp;
            lw[t] = cvt_bf16x2(v0 - __uint_as_float(hp << 16), v1 - __uint_as_float(hp & 0xffff0000u));
          } else {
            const uint32_t b0 = (__float_as_uint(v0) + 0x8000u) & 0xffff0000u, b1 = (__float_as_uint(v1) + 0x8000u) & 0xffff0000u;
            hw[t] = __byte_perm(b0, b1, 0x7632);
            const uint32_t r0 = __float_as_uint(v0 - __uint_as_float(b0)) + 0x8000u, r1 = __float_as_uint(v1 - __uint_as_float(b1)) + 0x8000u;
            lw[t] = __byte_perm(r0, r1, 0x7632);
          }
        }
        const uint32_t pdst = tmP(pb) + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 16);
        tmem_st_32x16(pdst, hw);
        tmem_st_32x16(pdst + 32u, lw);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[pb]);
        if (warp == 4 && j < 4) AT_STAMP(6 + 4 * j);
      }
    }
    if (warp == 4) AT_STAMP(19);
    if (MODE == AT_LSE) {
      // combine the two column halves of every row through shared memory (the operand stages are idle by now)
      float* comb = reinterpret_cast<float*>(smem + Cfg::kOffStages);
      if (half == 1) { comb[r] = m_run; comb[128 + r] = l_run; }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 softmax warps only
      if (half == 0 && row_ok) {
        const float m2 = comb[r], l2 = comb[128 + r];
        const float mn = fmaxf(m_run, m2);  // running maxima are in the log2 domain
        const float l = (mn == -INFINITY) ? 0.f : l_run * ex2f(m_run - mn) + l2 * ex2f(m2 - mn);
        a.lse_out[(int64_t)bh * a.rows + gr] = mn * 0.6931471805599453f + logf(l);
      }
    } else {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      if (warp == 4) AT_STAMP(20);
      float o[32];
      tmem_ld_32x32(tmACC + lane_sel, o);
      tmem_ld_wait();
      if ((a.D & 3) == 0 && ((((uintptr_t)a.out) & 15) == 0)) {
        // a thread owns a row of the accumulator: transpose the warp's 32x32 block through the (now idle) P buffer so that every
        // store is a 128-bit piece of a row's contiguous D floats (the scalar row-per-thread stores cost 8500 cycles per CTA)
        float* stg = reinterpret_cast<float*>(smem + Cfg::kOffStages) + (warp - 4) * 1024;  // operand stages are idle by now
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
        __syncwarp();
        const int c = lane & 7, rs = lane >> 3;
        if (half * 32 + 4 * c < a.D) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int     rr = it * 4 + rs;
            const int64_t grow = row0 + q * 32 + rr;
            if (grow < a.rows)
              *reinterpret_cast<float4*>(a.out + ((b * a.rows + grow) * a.H + h) * a.D + half * 32 + 4 * c) =
                  *reinterpret_cast<const float4*>(stg + rr * 32 + ((c ^ (rr & 7)) << 2));
          }
        }
      } else if (row_ok) {
        float* dst = a.out + ((b * a.rows + gr) * a.H + h) * a.D + half * 32;
#pragma unroll
        for (int d = 0; d < 32; ++d)
          if (half * 32 + d < a.D) dst[d] = o[d];
      }
    }
  }
  if (warp == 4) AT_STAMP(21);
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
  if (warp == 2) AT_STAMP(22);
#undef AT_STAMP
}

// Δ[bh][q] = Σ_d dO·O over [B, Lq, H, D] contiguous tensors; one warp per (b, q, h) row, lanes across d (coalesced)
__global__ void __launch_bounds__(256) k_attn_delta(const float* __restrict__ g, const float* __restrict__ o, float* __restrict__ delta, int64_t B,
                                                    int64_t Lq, int64_t H, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t total = B * Lq * H;
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < total; i += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t h = i % H, q = (i / H) % Lq, b = i / (H * Lq);
    const float *gp = g + i * D, *op = o + i * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += gp[d] * op[d];
    s = warp_sum(s);
    if (lane == 0) delta[(b * H + h) * Lq + q] = s;
  }
}
// D % 4 == 0, 16-byte aligned: G = D/4 rounded up to a power of two lanes per row with one 128-bit load each, 32/G rows per warp
template <int G>
__global__ void __launch_bounds__(256) k_attn_delta_v4(const float4* __restrict__ g, const float4* __restrict__ o, float* __restrict__ delta,
                                                       int64_t B, int64_t Lq, int64_t H, int d4) {
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), gr = lane / G;
  constexpr int RPW = 32 / G;
  const int64_t total = B * Lq * H;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i0 = wid * RPW; i0 < total; i0 += nw * RPW) {
    const int64_t i = i0 + gr;
    float s = 0.f;
    if (i < total && gl < d4) {
      const float4 a = __ldg(g + i * d4 + gl), c = __ldg(o + i * d4 + gl);
      s = a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (i < total && gl == 0) {
      const int64_t h = i % H, q = (i / H) % Lq, b = i / (H * Lq);
      delta[(b * H + h) * Lq + q] = s;
    }
  }
}

struct AtOperand {
  Scratch       buf;
  PackedOperand op;
  CUtensorMap   map;
};

// rows x K fp32 view of a [B, L, H, D]-style tensor given (batch, head, row) strides -> planes [B*H][2][R][Kp] + TMA map
static int at_pack(AtOperand* o, const float* src, int64_t B, int64_t H, int64_t R, int64_t K, int64_t r_stride, int64_t k_stride, int64_t bs,
                   int64_t hs, int box_rows, long long version = -1) {
  const int64_t nb[3] = {1, B, H}, st[3] = {0, bs, hs};
  // version >= 0: the planes come from / stay in the operand-plane cache (backward re-uses the forward's Q, K, V planes)
  PDN_TRY(planes_cached(src, R, K, r_stride, k_stride, nb, st, version, &o->buf, &o->op));
  // a degenerate batch dimension (B*H == 1) or broadcast strides would collapse the packed batch count; attention always
  // has distinct (b, h) slices, so the packed batch index is b*H + h whenever the strides are non-zero
  return tc_make_map(&o->map, o->op.planes, R, K, o->op.Kp, o->op.nbatch, box_rows);
}

template <int MODE>
static int at_launch(const CUtensorMap& A1, const CUtensorMap& A2, const CUtensorMap& B1, const CUtensorMap& B2, const CUtensorMap& B3,
                     const AtArgs& a, int64_t BH) {
  using Cfg = AtCfg<MODE>;
  static bool attr = false;
  if (!attr) {
    PDN_CUDA(cudaFuncSetAttribute(k_attn_tc<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr = true;
  }
  dim3 grd((unsigned)((a.rows + AT_R - 1) / AT_R), (unsigned)BH);
  static long long* trace_buf = nullptr;
  static const bool trace_on = getenv("PDN_TC_TRACE") != nullptr;
  AtArgs aa = a;
  aa.trace = nullptr;
  static const int ahead_env = getenv("PDN_AT_PREFETCH") ? atoi(getenv("PDN_AT_PREFETCH")) : -1;
  aa.prefetch_ahead = ahead_env >= 0 ? ahead_env : sm_count();
  if (trace_on) {
    if (!trace_buf) PDN_CUDA(cudaMalloc(&trace_buf, 32 * sizeof(long long)));
    PDN_CUDA(cudaMemsetAsync(trace_buf, 0, 32 * sizeof(long long), stream()));
    aa.trace = trace_buf;
  }
  k_attn_tc<MODE><<<grd, 384, Cfg::kSmem, stream()>>>(A1, A2, B1, B2, B3, aa);
  PDN_LAUNCHED("attn_tc");
  if (trace_on) {
    long long h[32];
    PDN_CUDA(cudaStreamSynchronize(stream()));
    PDN_CUDA(cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[at trace] mode=%d rows=%lld cols=%lld | setup %lld a_full %lld |", MODE, (long long)a.rows, (long long)a.cols, h[1] - h[0], h[2] - h[0]);
    for (int j = 0; j < 4; ++j)
      fprintf(stderr, " it%d: s_full %lld ld %lld math %lld pwr %lld |", j, h[3 + 4 * j] - h[0], h[4 + 4 * j] - h[0], h[5 + 4 * j] - h[0], h[6 + 4 * j] - h[0]);
    fprintf(stderr, " loop_end %lld acc_full %lld stored %lld dealloc %lld | mma: acfull2 %lld pfull2 %lld acc2_issued %lld scfull4 %lld sempty4 %lld score4_issued %lld\n", h[19] - h[0], h[20] - h[0], h[21] - h[0], h[22] - h[0], h[28] - h[0], h[23] - h[0], h[24] - h[0], h[25] - h[0], h[26] - h[0], h[27] - h[0]);
  }
  return 0;
}

static int at_check(int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const int64_t* q_str, const int64_t* k_str, const int64_t* v_str) {
  PDN_CHECK(D >= 1 && D <= 64, "attention_tc: head dim %lld > 64", (long long)D);
  PDN_CHECK(B * H >= 1 && B * H <= 65535, "attention_tc: batch*heads out of range");
  PDN_CHECK(Lq >= 1 && Lk >= 1 && Lq <= 0x7fffffff && Lk <= 0x7fffffff, "attention_tc: bad sequence lengths");
  for (int i = 0; i < 3; ++i) PDN_CHECK(q_str[i] != 0 || (i < 2 && (i == 0 ? B : H) == 1), "attention_tc: broadcast q strides are not supported");
  (void)k_str; (void)v_str;
  return 0;
}

}  // namespace pdn

using namespace pdn;

extern "C" {

int pdn_attention_tc_fwd(const float* q, const float* k, const float* v, const float* mask, float* out, float* lse, int64_t B, int64_t H, int64_t Lq,
                         int64_t Lk, int64_t D, const int64_t* q_str, const int64_t* k_str, const int64_t* v_str, const int64_t* mask_str,
                         float scale, const int64_t* versions) {
  const long long qv = versions ? versions[0] : -1, kv = versions ? versions[1] : -1, vv = versions ? versions[2] : -1;
  PDN_TRY(ensure_init());
  PDN_TRY(at_check(B, H, Lq, Lk, D, q_str, k_str, v_str));
  AtOperand Qp, Kp, Vp;
  PDN_TRY(at_pack(&Qp, q, B, H, Lq, D, q_str[2], 1, q_str[0], q_str[1], AT_R, qv));
  PDN_TRY(at_pack(&Kp, k, B, H, Lk, D, k_str[2], 1, k_str[0], k_str[1], AT_C, kv));
  PDN_TRY(at_pack(&Vp, v, B, H, Lk, D, v_str[2], 1, v_str[0], v_str[1], AT_C, vv));  // [keys][d]: MN-major B of O += P V
  AtArgs a;
  a.rows = Lq; a.cols = Lk; a.H = H; a.D = (int)D; a.scale = scale;
  a.mask = mask; a.mask_bs = mask && mask_str ? mask_str[0] : 0; a.mask_qs = mask && mask_str ? mask_str[1] : 0;
  a.lse = lse; a.delta = nullptr; a.lse_out = lse; a.out = out;
  a.ncol_tiles = (int)((Lk + AT_C - 1) / AT_C);
  PDN_TRY((at_launch<AT_LSE>(Qp.map, Qp.map, Kp.map, Kp.map, Kp.map, a, B * H)));
  PDN_TRY((at_launch<AT_FWD>(Qp.map, Qp.map, Kp.map, Kp.map, Vp.map, a, B * H)));
  return 0;
}

int pdn_attention_tc_bwd(const float* q, const float* k, const float* v, const float* mask, const float* out, const float* lse, const float* g_out,
                         float* dq, float* dk, float* dv, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const int64_t* q_str,
                         const int64_t* k_str, const int64_t* v_str, const int64_t* mask_str, float scale, const int64_t* versions) {
  const long long qv = versions ? versions[0] : -1, kv = versions ? versions[1] : -1, vv = versions ? versions[2] : -1;
  PDN_TRY(ensure_init());
  PDN_TRY(at_check(B, H, Lq, Lk, D, q_str, k_str, v_str));
  Scratch sdelta;
  PDN_TRY(sdelta.alloc((size_t)B * H * Lq * sizeof(float)));
  if ((D & 3) == 0 && ((((uintptr_t)g_out) | ((uintptr_t)out)) & 15) == 0) {
    const int d4 = (int)(D / 4);
    const int grd = grid_for(B * Lq * H, 8 * 4);
    if (d4 <= 4) k_attn_delta_v4<4><<<grd, 256, 0, stream()>>>((const float4*)g_out, (const float4*)out, (float*)sdelta.p, B, Lq, H, d4);
    else if (d4 <= 8) k_attn_delta_v4<8><<<grd, 256, 0, stream()>>>((const float4*)g_out, (const float4*)out, (float*)sdelta.p, B, Lq, H, d4);
    else k_attn_delta_v4<16><<<grd, 256, 0, stream()>>>((const float4*)g_out, (const float4*)out, (float*)sdelta.p, B, Lq, H, d4);
  } else {
    k_attn_delta<<<grid_for(B * Lq * H, 8), 256, 0, stream()>>>(g_out, out, (float*)sdelta.p, B, Lq, H, (int)D);
  }
  PDN_LAUNCHED("attn_delta");
  const int64_t g_str[3] = {Lq * H * D, D, H * D};  // g_out is [B, Lq, H, D] contiguous
  AtArgs a;
  a.H = H; a.D = (int)D; a.scale = scale;
  a.mask = mask; a.mask_bs = mask && mask_str ? mask_str[0] : 0; a.mask_qs = mask && mask_str ? mask_str[1] : 0;
  a.lse = lse; a.delta = (const float*)sdelta.p; a.lse_out = nullptr;
  // every tensor is packed ONCE per layout; the row-operand (box 128) and column-operand (box 64) TMA maps share the planes
  AtOperand Qp, Kp, Vp, dOp;
  CUtensorMap Qp64, Kp128, Vp128, dOp64;
  PDN_TRY(at_pack(&Qp, q, B, H, Lq, D, q_str[2], 1, q_str[0], q_str[1], AT_R, qv));
  PDN_TRY(at_pack(&Kp, k, B, H, Lk, D, k_str[2], 1, k_str[0], k_str[1], AT_C, kv));
  PDN_TRY(at_pack(&dOp, g_out, B, H, Lq, D, g_str[2], 1, g_str[0], g_str[1], AT_R));
  PDN_TRY(at_pack(&Vp, v, B, H, Lk, D, v_str[2], 1, v_str[0], v_str[1], AT_C, vv));
  PDN_TRY(tc_make_map(&Qp64, Qp.op.planes, Lq, D, Qp.op.Kp, Qp.op.nbatch, AT_C));
  PDN_TRY(tc_make_map(&dOp64, dOp.op.planes, Lq, D, dOp.op.Kp, dOp.op.nbatch, AT_C));
  PDN_TRY(tc_make_map(&Kp128, Kp.op.planes, Lk, D, Kp.op.Kp, Kp.op.nbatch, AT_R));
  PDN_TRY(tc_make_map(&Vp128, Vp.op.planes, Lk, D, Vp.op.Kp, Vp.op.nbatch, AT_R));
  if (dq) {
    a.rows = Lq; a.cols = Lk; a.out = dq; a.ncol_tiles = (int)((Lk + AT_C - 1) / AT_C);
    PDN_TRY((at_launch<AT_DQ>(Qp.map, dOp.map, Kp.map, Vp.map, Kp.map, a, B * H)));  // dQ += dS K: the K tile again, MN-major
  }
  if (dk || dv) {
    a.rows = Lk; a.cols = Lq; a.ncol_tiles = (int)((Lq + AT_C - 1) / AT_C);
    if (dv) {
      a.out = dv;
      PDN_TRY((at_launch<AT_DV>(Kp128, Kp128, Qp64, Qp64, dOp64, a, B * H)));  // dV += P^T dO
    }
    if (dk) {
      a.out = dk;
      PDN_TRY((at_launch<AT_DK>(Kp128, Vp128, Qp64, dOp64, Qp64, a, B * H)));  // dK += dS^T Q
    }
  }
  return 0;
}

}  // extern "C"
