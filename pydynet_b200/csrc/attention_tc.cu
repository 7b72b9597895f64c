// attention_tc.cu — softmax(q kᵀ·scale + mask)·v forward AND backward on the Blackwell tensor cores, scores never in HBM.
//
// The reference materialises the [B,H,Lq,Lk] score tensor and ≥5 same-sized temporaries (llm/llama/model.py:112-121,
// examples/pydynet/transformer.py:93-104: 1.07 GB each at BASELINE config 4). Here one kernel template serves five passes
// that all have the same shape — "score MMA(s) into TMEM → element-wise in registers → one bf16 hi/lo operand tile into
// swizzled shared memory → accumulate MMA into TMEM":
//     LSE  rows = queries   S = Q·Kᵀ                      row-wise log-sum-exp (online), nothing accumulated
//     FWD  rows = queries   S = Q·Kᵀ                      P = exp(S − lse)              O  += P · V
//     DQ   rows = queries   S = Q·Kᵀ, dP = dO·Vᵀ          dS = P∘(dP − Δ)·scale         dQ += dS · K
//     DV   rows = keys      Sᵀ = K·Qᵀ                     Pᵀ = exp(Sᵀ − lse)            dV += Pᵀ · dO
//     DK   rows = keys      Sᵀ = K·Qᵀ, dPᵀ = V·dOᵀ        dSᵀ = Pᵀ∘(dPᵀ − Δ)·scale      dK += dSᵀ · Q
// Because lse is known before P is formed (LSE pass first), no accumulator is ever rescaled; the price is recomputing S.
// Every MMA operand is K-major: the transposed operands (Vᵀ, Kᵀ, Qᵀ, dOᵀ as [d][seq]) come from the strided pack kernel,
// and P / dS are written by the softmax warps directly in the 128-byte-swizzled layout tcgen05 expects. fp32 parity comes
// from the same BF16x3 split as the GEMM (hi·hi + hi·lo + lo·hi), applied to P/dS as well.
//
// CTA = one 128-row tile of one (batch, head); warp 0 TMA producer (2-stage ring over 64-column tiles), warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-7 = 128 softmax threads (one row each: row max / sum need no shuffles at all).
#include "common.cuh"
#include "gemm_tc.h"
#include "tc_ptx.cuh"
#include <cudaTypedefs.h>
#include <math.h>

namespace pdn {

enum { AT_LSE = 0, AT_FWD = 1, AT_DQ = 2, AT_DV = 3, AT_DK = 4 };
constexpr int AT_R = 128;  // row tile
constexpr int AT_C = 64;   // column tile
constexpr int AT_KA = 2 * AT_R * 64 * 2;  // resident row operand, hi + lo: 32 KB
constexpr int AT_KB = 2 * AT_C * 64 * 2;  // one column-tile operand, hi + lo: 16 KB
constexpr int AT_KP = 2 * AT_R * AT_C * 2;  // P / dS tile, hi + lo: 32 KB

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
};

template <int MODE>
struct AtCfg {
  static constexpr bool TWO = (MODE == AT_DQ || MODE == AT_DK);
  static constexpr bool ACC = (MODE != AT_LSE);
  static constexpr bool TRANS = (MODE == AT_DV || MODE == AT_DK);
  static constexpr int  kStage = AT_KB * (1 + (TWO ? 1 : 0) + (ACC ? 1 : 0));
  static constexpr int  kOffStages = AT_KA * (TWO ? 2 : 1);
  static constexpr int  kOffP = kOffStages + 2 * kStage;
  static constexpr int  kOffBars = kOffP + (ACC ? 2 * AT_KP : 0);
  static constexpr int  kSmem = kOffBars + 256 + 1024;
};

// 16-byte chunk c (8 bf16) of row r in a K-major [rows x 64] bf16 tile with 128-byte swizzle (what TMA writes / UMMA reads)
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

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

template <int MODE>
__global__ void __launch_bounds__(384, 1)
k_attn_tc(const __grid_constant__ CUtensorMap mA1, const __grid_constant__ CUtensorMap mA2, const __grid_constant__ CUtensorMap mB1,
          const __grid_constant__ CUtensorMap mB2, const __grid_constant__ CUtensorMap mB3, AtArgs a) {
  using Cfg = AtCfg<MODE>;
  constexpr bool TWO = Cfg::TWO, ACC = Cfg::ACC, TRANS = Cfg::TRANS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t*  smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + Cfg::kOffBars);
  uint64_t *a_full = bars, *b_full = bars + 1, *b_empty = bars + 3, *s_full = bars + 5, *s_empty = bars + 7, *p_full = bars + 9,
           *p_empty = bars + 11, *acc_full = bars + 13;
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);  // one arrival per softmax warp
      mbar_init(&p_full[i], 8);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  constexpr uint32_t kTmemCols = ACC ? 512u : 128u;  // LSE needs only the two score buffers (lets several CTAs share an SM)
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S1[2] at 0/64, S2[2] at 128/192, accumulator at 256
  auto tmS1 = [&](int i) { return tmem_base + (uint32_t)(i * 64); };
  auto tmS2 = [&](int i) { return tmem_base + (uint32_t)(128 + i * 64); };
  const uint32_t tmACC = tmem_base + 256u;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(a_full, AT_KA * (TWO ? 2 : 1));
      tma_load_4d(&mA1, a_full, smem, 0, (int)row0, 0, bh);
      tma_load_4d(&mA1, a_full, smem + AT_KA / 2, 0, (int)row0, 1, bh);
      if (TWO) {
        tma_load_4d(&mA2, a_full, smem + AT_KA, 0, (int)row0, 0, bh);
        tma_load_4d(&mA2, a_full, smem + AT_KA + AT_KA / 2, 0, (int)row0, 1, bh);
      }
      for (int j = 0; j < n; ++j) {
        const int st = j & 1;
        mbar_wait(&b_empty[st], (uint32_t)(((j >> 1) & 1) ^ 1));
        uint8_t* sb = smem + Cfg::kOffStages + st * Cfg::kStage;
        mbar_expect_tx(&b_full[st], Cfg::kStage);
        const int col0 = j * AT_C;
        tma_load_4d(&mB1, &b_full[st], sb, 0, col0, 0, bh);
        tma_load_4d(&mB1, &b_full[st], sb + AT_KB / 2, 0, col0, 1, bh);
        int off = AT_KB;
        if (TWO) {
          tma_load_4d(&mB2, &b_full[st], sb + off, 0, col0, 0, bh);
          tma_load_4d(&mB2, &b_full[st], sb + off + AT_KB / 2, 0, col0, 1, bh);
          off += AT_KB;
        }
        if (ACC) {  // [d rows][column items along k]
          tma_load_4d(&mB3, &b_full[st], sb + off, col0, 0, 0, bh);
          tma_load_4d(&mB3, &b_full[st], sb + off + AT_KB / 2, col0, 0, 1, bh);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(AT_R, AT_C);  // M = 128, N = 64 for the score and the accumulate MMAs alike
      mbar_wait(a_full, 0);
      tc_fence_after();
      const uint32_t sA = smem_u32(smem);
      const uint64_t a1hi = make_smem_desc_sw128(sA), a1lo = make_smem_desc_sw128(sA + AT_KA / 2);
      const uint64_t a2hi = make_smem_desc_sw128(sA + AT_KA), a2lo = make_smem_desc_sw128(sA + AT_KA + AT_KA / 2);
      auto score = [&](int j) {
        const int st = j & 1, sb = j & 1;
        mbar_wait(&b_full[st], (uint32_t)((j >> 1) & 1));
        mbar_wait(&s_empty[sb], (uint32_t)(((j >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t sB = smem_u32(smem + Cfg::kOffStages + st * Cfg::kStage);
        umma3(tmS1(sb), a1hi, a1lo, make_smem_desc_sw128(sB), make_smem_desc_sw128(sB + AT_KB / 2), idesc, true);
        if (TWO) umma3(tmS2(sb), a2hi, a2lo, make_smem_desc_sw128(sB + AT_KB), make_smem_desc_sw128(sB + AT_KB + AT_KB / 2), idesc, true);
        umma_commit(&s_full[sb]);
        if (!ACC) umma_commit(&b_empty[st]);
      };
      auto accumulate = [&](int j) {
        const int st = j & 1, pb = j & 1;
        mbar_wait(&p_full[pb], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const uint32_t sP = smem_u32(smem + Cfg::kOffP + pb * AT_KP);
        const uint32_t sB3 = smem_u32(smem + Cfg::kOffStages + st * Cfg::kStage + AT_KB * (TWO ? 2 : 1));
        umma3(tmACC, make_smem_desc_sw128(sP), make_smem_desc_sw128(sP + AT_KP / 2), make_smem_desc_sw128(sB3),
              make_smem_desc_sw128(sB3 + AT_KB / 2), idesc, j == 0);
        umma_commit(&p_empty[pb]);
        umma_commit(&b_empty[st]);
      };
      if (n > 0) score(0);
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) score(j + 1);
        if (ACC) accumulate(j);
      }
      if (ACC) umma_commit(acc_full);
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
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      float s[32], dp[TWO ? 32 : 1];
      tmem_ld_32x32(tmS1(sb) + lane_sel, s);
      if (TWO) tmem_ld_32x32(tmS2(sb) + lane_sel, dp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
      const int col0 = j * AT_C + half * 32;
      const int ncols = (int)a.cols;
      // Everything below works in the log2 domain: t = (s*scale + mask) * log2(e), so that exp(x - lse) is ONE ex2.approx of
      // one FFMA result. Column validity only matters in the last tile (uniform branch); masks take the slower path.
      const float c1 = a.scale * 1.4426950408889634f;
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
      } else {
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
          if (col0 + 32 <= ncols && (Lq & 3) == 0) {
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
            float p = ex2f(fmaf(cl[e], -1.4426950408889634f, s[e]));
            if (TWO) p = p * (dp[e] - cd[e]) * a.scale;
            s[e] = p;
          }
        } else {
          const float lse2 = lse_r * 1.4426950408889634f;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float p = ex2f(s[e] - lse2);
            if (TWO) p = p * (dp[e] - delta_r) * a.scale;
            s[e] = p;
          }
        }
        const int pb = j & 1;
        mbar_wait(&p_empty[pb], (uint32_t)(((j >> 1) & 1) ^ 1));
        uint8_t* ph = smem + Cfg::kOffP + pb * AT_KP;
        uint8_t* pl = ph + AT_KP / 2;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float v0 = s[c8 * 8 + 2 * t], v1 = s[c8 * 8 + 2 * t + 1];
            const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);   // one packed convert for two elements
            const float2         hf = __bfloat1622float2(hh);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
            hw[t] = *reinterpret_cast<const uint32_t*>(&hh);
            lw[t] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          const uint32_t off = sw128_off(r, half * 4 + c8);
          *reinterpret_cast<uint4*>(ph + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(pl + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[pb]);
      }
    }
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
      float o[32];
      tmem_ld_32x32(tmACC + lane_sel, o);
      tmem_ld_wait();
      if (row_ok) {
        float* dst = a.out + ((b * a.rows + gr) * a.H + h) * a.D + half * 32;
#pragma unroll
        for (int d = 0; d < 32; ++d)
          if (half * 32 + d < a.D) dst[d] = o[d];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
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

struct AtOperand {
  Scratch       buf;
  PackedOperand op;
  CUtensorMap   map;
};

// rows x K fp32 view of a [B, L, H, D]-style tensor given (batch, head, row) strides -> planes [B*H][2][R][Kp] + TMA map
static int at_pack(AtOperand* o, const float* src, int64_t B, int64_t H, int64_t R, int64_t K, int64_t r_stride, int64_t k_stride, int64_t bs,
                   int64_t hs, int box_rows) {
  const int64_t nb[3] = {1, B, H}, st[3] = {0, bs, hs};
  PDN_TRY(pack_operand_ex(src, R, K, r_stride, k_stride, 0, 0, nb, st, &o->buf, &o->op));
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
  k_attn_tc<MODE><<<grd, 384, Cfg::kSmem, stream()>>>(A1, A2, B1, B2, B3, a);
  PDN_LAUNCHED("attn_tc");
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
                         float scale) {
  PDN_TRY(ensure_init());
  PDN_TRY(at_check(B, H, Lq, Lk, D, q_str, k_str, v_str));
  AtOperand Qp, Kp, Vt;
  PDN_TRY(at_pack(&Qp, q, B, H, Lq, D, q_str[2], 1, q_str[0], q_str[1], AT_R));
  PDN_TRY(at_pack(&Kp, k, B, H, Lk, D, k_str[2], 1, k_str[0], k_str[1], AT_C));
  PDN_TRY(at_pack(&Vt, v, B, H, D, Lk, 1, v_str[2], v_str[0], v_str[1], AT_C));
  AtArgs a;
  a.rows = Lq; a.cols = Lk; a.H = H; a.D = (int)D; a.scale = scale;
  a.mask = mask; a.mask_bs = mask && mask_str ? mask_str[0] : 0; a.mask_qs = mask && mask_str ? mask_str[1] : 0;
  a.lse = lse; a.delta = nullptr; a.lse_out = lse; a.out = out;
  a.ncol_tiles = (int)((Lk + AT_C - 1) / AT_C);
  PDN_TRY((at_launch<AT_LSE>(Qp.map, Qp.map, Kp.map, Kp.map, Kp.map, a, B * H)));
  PDN_TRY((at_launch<AT_FWD>(Qp.map, Qp.map, Kp.map, Kp.map, Vt.map, a, B * H)));
  return 0;
}

int pdn_attention_tc_bwd(const float* q, const float* k, const float* v, const float* mask, const float* out, const float* lse, const float* g_out,
                         float* dq, float* dk, float* dv, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const int64_t* q_str,
                         const int64_t* k_str, const int64_t* v_str, const int64_t* mask_str, float scale) {
  PDN_TRY(ensure_init());
  PDN_TRY(at_check(B, H, Lq, Lk, D, q_str, k_str, v_str));
  Scratch sdelta;
  PDN_TRY(sdelta.alloc((size_t)B * H * Lq * sizeof(float)));
  k_attn_delta<<<grid_for(B * Lq * H, 8), 256, 0, stream()>>>(g_out, out, (float*)sdelta.p, B, Lq, H, (int)D);
  PDN_LAUNCHED("attn_delta");
  const int64_t g_str[3] = {Lq * H * D, D, H * D};  // g_out is [B, Lq, H, D] contiguous
  AtArgs a;
  a.H = H; a.D = (int)D; a.scale = scale;
  a.mask = mask; a.mask_bs = mask && mask_str ? mask_str[0] : 0; a.mask_qs = mask && mask_str ? mask_str[1] : 0;
  a.lse = lse; a.delta = (const float*)sdelta.p; a.lse_out = nullptr;
  // every tensor is packed ONCE per layout; the row-operand (box 128) and column-operand (box 64) TMA maps share the planes
  AtOperand Qp, Kp, Vp, dOp, Kt, Qt, dOt;
  CUtensorMap Qp64, Kp128, Vp128, dOp64;
  PDN_TRY(at_pack(&Qp, q, B, H, Lq, D, q_str[2], 1, q_str[0], q_str[1], AT_R));
  PDN_TRY(at_pack(&Kp, k, B, H, Lk, D, k_str[2], 1, k_str[0], k_str[1], AT_C));
  PDN_TRY(at_pack(&dOp, g_out, B, H, Lq, D, g_str[2], 1, g_str[0], g_str[1], AT_R));
  PDN_TRY(at_pack(&Vp, v, B, H, Lk, D, v_str[2], 1, v_str[0], v_str[1], AT_C));
  PDN_TRY(tc_make_map(&Qp64, Qp.op.planes, Lq, D, Qp.op.Kp, Qp.op.nbatch, AT_C));
  PDN_TRY(tc_make_map(&dOp64, dOp.op.planes, Lq, D, dOp.op.Kp, dOp.op.nbatch, AT_C));
  PDN_TRY(tc_make_map(&Kp128, Kp.op.planes, Lk, D, Kp.op.Kp, Kp.op.nbatch, AT_R));
  PDN_TRY(tc_make_map(&Vp128, Vp.op.planes, Lk, D, Vp.op.Kp, Vp.op.nbatch, AT_R));
  if (dq) {
    PDN_TRY(at_pack(&Kt, k, B, H, D, Lk, 1, k_str[2], k_str[0], k_str[1], AT_C));
    a.rows = Lq; a.cols = Lk; a.out = dq; a.ncol_tiles = (int)((Lk + AT_C - 1) / AT_C);
    PDN_TRY((at_launch<AT_DQ>(Qp.map, dOp.map, Kp.map, Vp.map, Kt.map, a, B * H)));
  }
  if (dk || dv) {
    a.rows = Lk; a.cols = Lq; a.ncol_tiles = (int)((Lq + AT_C - 1) / AT_C);
    if (dv) {
      PDN_TRY(at_pack(&dOt, g_out, B, H, D, Lq, 1, g_str[2], g_str[0], g_str[1], AT_C));
      a.out = dv;
      PDN_TRY((at_launch<AT_DV>(Kp128, Kp128, Qp64, Qp64, dOt.map, a, B * H)));
    }
    if (dk) {
      PDN_TRY(at_pack(&Qt, q, B, H, D, Lq, 1, q_str[2], q_str[0], q_str[1], AT_C));
      a.out = dk;
      PDN_TRY((at_launch<AT_DK>(Kp128, Vp128, Qp64, dOp64, Qt.map, a, B * H)));
    }
  }
  return 0;
}

}  // extern "C"
