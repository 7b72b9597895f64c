// attention.cu — fused softmax(q kᵀ·scale + mask)·v and the Llama RoPE + KV-cache append.
//
// Replaces the 5-node score/softmax/PV chain the reference's models build by hand (llm/llama/model.py:112-121,
// examples/pydynet/transformer.py:93-104: the S×S score tensor plus ≥5 temporaries) with one kernel that never
// materialises scores: one warp per (batch, head, query row), keys in chunks of 32 (one key per lane for q·k, then
// the lanes switch to owning head-dim columns for p·v), online softmax with warp-shuffle max/sum reductions.
// fp32 FFMA throughout (1e-4 parity with the NumPy path); q/k/v are consumed through element strides so head-split
// views of [B,L,H*D] projections and [B,S,H,D] KV-cache slices are read in place. This kernel serves decode / prefill /
// small training shapes; large training shapes go through the tcgen05 GEMM path (nn/_fused.py picks).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace pdn {

constexpr int ATT_MAXD = 128;           // head dim limit (4 columns per lane)
constexpr int ATT_DPL = ATT_MAXD / 32;  // output columns owned per lane

struct AttArgs {
  const float *q, *k, *v, *mask;
  float *out, *lse;
  int64_t B, H, Lq, Lk, D;
  int64_t qs[3], ks[3], vs[3];  // element strides: batch, head, row (D axis is unit stride)
  int64_t mask_bs, mask_qs;     // mask element strides for batch and query row (0 = broadcast); key axis unit stride
  float scale;
  const int64_t* lk_dev;        // optional device scalar: Lk = *lk_dev + lk_add (CUDA-graph replay of the decode step)
  int64_t lk_add;
  __nv_bfloat16* out_planes;    // optional: emit [B*Lq, H*D] as bf16 hi/lo operand planes [2][B*Lq][planes_kp] instead of fp32 out
  int64_t planes_kp;
};

__global__ void __launch_bounds__(128) k_attention_fwd(AttArgs a) {
  extern __shared__ float sm[];  // per warp: q row [D]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t w = (int64_t)blockIdx.x * 4 + wib;
  const int64_t total = a.B * a.H * a.Lq;
  if (w >= total) return;
  const int64_t iq = w % a.Lq, h = (w / a.Lq) % a.H, b = w / (a.Lq * a.H);
  const int D = (int)a.D;
  if (a.lk_dev) a.Lk = *a.lk_dev + a.lk_add;
  float* qsm = sm + wib * ATT_MAXD;
  const float* qrow = a.q + b * a.qs[0] + h * a.qs[1] + iq * a.qs[2];
  for (int d = lane; d < D; d += 32) qsm[d] = qrow[d] * a.scale;
  __syncwarp();
  const float* kb = a.k + b * a.ks[0] + h * a.ks[1];
  const float* vb = a.v + b * a.vs[0] + h * a.vs[1];
  const float* mrow = a.mask ? a.mask + b * a.mask_bs + iq * a.mask_qs : nullptr;
  float m = -INFINITY, l = 0.f;
  float acc[ATT_DPL];
#pragma unroll
  for (int i = 0; i < ATT_DPL; ++i) acc[i] = 0.f;
  const bool vec = (D % 4 == 0) && ((a.ks[2] & 3) == 0) && ((((uintptr_t)kb) & 15) == 0);
  for (int64_t j0 = 0; j0 < a.Lk; j0 += 32) {
    const int64_t j = j0 + lane;
    float s = -INFINITY;
    if (j < a.Lk) {
      const float* kr = kb + j * a.ks[2];
      float dot = 0.f;
      if (vec) {
        for (int d = 0; d < D; d += 4) {
          float4 kk = *reinterpret_cast<const float4*>(kr + d);
          dot += qsm[d] * kk.x + qsm[d + 1] * kk.y + qsm[d + 2] * kk.z + qsm[d + 3] * kk.w;
        }
      } else {
        for (int d = 0; d < D; ++d) dot += qsm[d] * kr[d];
      }
      s = dot;
      if (mrow) s += mrow[j];
    }
    const float cm = warp_max(s);
    const float mn = fmaxf(m, cm);
    // mn == -inf only while every key so far is masked out: keep the state empty
    const float corr = (mn == -INFINITY) ? 1.f : __expf(m - mn);
    const float p = (s == -INFINITY) ? 0.f : __expf(s - mn);
    l = l * corr + warp_sum(p);
    m = mn;
#pragma unroll
    for (int i = 0; i < ATT_DPL; ++i) acc[i] *= corr;
    const int cnt = (int)((a.Lk - j0) < 32 ? (a.Lk - j0) : 32);
    for (int t = 0; t < cnt; ++t) {
      const float pt = __shfl_sync(0xffffffffu, p, t);
      const float* vr = vb + (j0 + t) * a.vs[2];
#pragma unroll
      for (int i = 0; i < ATT_DPL; ++i) {
        int d = lane + i * 32;
        if (d < D) acc[i] += pt * vr[d];
      }
    }
  }
  const float inv = 1.f / l;
  if (a.out_planes) {  // the O-projection GEMM consumes this directly (no fp32 round trip, no pack launch)
    const int64_t rows = a.B * a.Lq, r = b * a.Lq + iq;
    __nv_bfloat16 *hi = a.out_planes + r * a.planes_kp + h * a.D, *lo = hi + rows * a.planes_kp;
#pragma unroll
    for (int i = 0; i < ATT_DPL; ++i) {
      int d = lane + i * 32;
      if (d < D) {
        const float v = acc[i] * inv;
        const __nv_bfloat16 hv = __float2bfloat16_rn(v);
        hi[d] = hv;
        lo[d] = __float2bfloat16_rn(v - __bfloat162float(hv));
      }
    }
  } else {
    float* orow = a.out + ((b * a.Lq + iq) * a.H + h) * a.D;
#pragma unroll
    for (int i = 0; i < ATT_DPL; ++i) {
      int d = lane + i * 32;
      if (d < D) orow[d] = acc[i] * inv;
    }
  }
  if (lane == 0 && a.lse) a.lse[(b * a.H + h) * a.Lq + iq] = m + logf(l);
}

// Decode form (few query rows, long key range): one CTA per (batch, head, query row), its 4 warps take interleaved
// 32-key chunks (flash-decoding inside the block) and merge their (max, sum, partial output) through shared memory.
// 4x more warps in flight than the warp-per-row kernel at the same batch: the first ncu capture of the decode step showed
// that kernel latency-bound at 29 % of HBM peak with 31 % occupancy (profiles/r1_ncu_summary.md).
__global__ void __launch_bounds__(128) k_attention_decode(AttArgs a) {
  __shared__ float qsm[ATT_MAXD];
  __shared__ float red_m[4], red_l[4];
  __shared__ float red_acc[4][ATT_MAXD];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t w = blockIdx.x;
  const int64_t iq = w % a.Lq, h = (w / a.Lq) % a.H, b = w / (a.Lq * a.H);
  const int D = (int)a.D;
  if (a.lk_dev) a.Lk = *a.lk_dev + a.lk_add;
  const float* qrow = a.q + b * a.qs[0] + h * a.qs[1] + iq * a.qs[2];
  for (int d = threadIdx.x; d < D; d += 128) qsm[d] = qrow[d] * a.scale;
  __syncthreads();
  const float* kb = a.k + b * a.ks[0] + h * a.ks[1];
  const float* vb = a.v + b * a.vs[0] + h * a.vs[1];
  const float* mrow = a.mask ? a.mask + b * a.mask_bs + iq * a.mask_qs : nullptr;
  float m = -INFINITY, l = 0.f;
  float acc[ATT_DPL];
#pragma unroll
  for (int i = 0; i < ATT_DPL; ++i) acc[i] = 0.f;
  const bool vec = (D % 4 == 0) && ((a.ks[2] & 3) == 0) && ((((uintptr_t)kb) & 15) == 0);
  for (int64_t j0 = (int64_t)wib * 32; j0 < a.Lk; j0 += 128) {
    const int64_t j = j0 + lane;
    float s = -INFINITY;
    if (j < a.Lk) {
      const float* kr = kb + j * a.ks[2];
      float dot = 0.f;
      if (vec) {
        for (int d = 0; d < D; d += 4) {
          float4 kk = *reinterpret_cast<const float4*>(kr + d);
          dot += qsm[d] * kk.x + qsm[d + 1] * kk.y + qsm[d + 2] * kk.z + qsm[d + 3] * kk.w;
        }
      } else {
        for (int d = 0; d < D; ++d) dot += qsm[d] * kr[d];
      }
      s = dot;
      if (mrow) s += mrow[j];
    }
    const float cm = warp_max(s);
    const float mn = fmaxf(m, cm);
    const float corr = (mn == -INFINITY) ? 1.f : __expf(m - mn);
    const float p = (s == -INFINITY) ? 0.f : __expf(s - mn);
    l = l * corr + warp_sum(p);
    m = mn;
#pragma unroll
    for (int i = 0; i < ATT_DPL; ++i) acc[i] *= corr;
    const int cnt = (int)((a.Lk - j0) < 32 ? (a.Lk - j0) : 32);
    for (int t = 0; t < cnt; ++t) {
      const float pt = __shfl_sync(0xffffffffu, p, t);
      const float* vr = vb + (j0 + t) * a.vs[2];
#pragma unroll
      for (int i = 0; i < ATT_DPL; ++i) {
        int d = lane + i * 32;
        if (d < D) acc[i] += pt * vr[d];
      }
    }
  }
  // merge the 4 warps
  if (lane == 0) { red_m[wib] = m; red_l[wib] = l; }
#pragma unroll
  for (int i = 0; i < ATT_DPL; ++i) {
    int d = lane + i * 32;
    if (d < D) red_acc[wib][d] = acc[i];
  }
  __syncthreads();
  if (wib == 0) {
    const float mt = fmaxf(fmaxf(red_m[0], red_m[1]), fmaxf(red_m[2], red_m[3]));
    float lt = 0.f, sc[4];
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) {
      sc[ww] = (red_m[ww] == -INFINITY) ? 0.f : __expf(red_m[ww] - mt);
      lt += red_l[ww] * sc[ww];
    }
    const float inv = 1.f / lt;
    const int64_t rows = a.B * a.Lq, r = b * a.Lq + iq;
#pragma unroll
    for (int i = 0; i < ATT_DPL; ++i) {
      int d = lane + i * 32;
      if (d < D) {
        const float v = (red_acc[0][d] * sc[0] + red_acc[1][d] * sc[1] + red_acc[2][d] * sc[2] + red_acc[3][d] * sc[3]) * inv;
        if (a.out_planes) {
          __nv_bfloat16 *hi = a.out_planes + r * a.planes_kp + h * a.D, *lo = hi + rows * a.planes_kp;
          const __nv_bfloat16 hv = __float2bfloat16_rn(v);
          hi[d] = hv;
          lo[d] = __float2bfloat16_rn(v - __bfloat162float(hv));
        } else {
          a.out[((b * a.Lq + iq) * a.H + h) * a.D + d] = v;
        }
      }
    }
    if (lane == 0 && a.lse) a.lse[(b * a.H + h) * a.Lq + iq] = mt + logf(lt);
  }
}

// Coalesced row kernel (decode and every 16-byte-aligned shape): G lanes (G*4 >= D) cooperate on ONE key row with one float4
// each, so a warp instruction reads 32/G whole key rows — contiguous 4*D-byte segments instead of 32 rows 4*H*D bytes apart —
// and the same lane keeps its float4 slice of the output, so P·V needs no role switch and no per-key shuffle broadcast.
// A chunk = U warp iterations: all 2*U float4 loads (K and V) are issued before the first use (8 x 16 B in flight per lane),
// then ONE online-softmax rescale per chunk. The 32/G key groups of a warp keep separate (max, sum, acc) states that are merged
// by xor-shuffles at the end; WPR = 4 additionally splits the chunks of one row over the CTA's 4 warps. First ncu capture of the decode step at batch 1024 (profiles/r1d_launches_b1024.csv): the lane-per-key kernel
// above ran at 2.25 TB/s of K/V traffic (35 % of the measured HBM peak) and was 54 % of the step.
template <int G, int WPR, int U, int MINB>
__global__ void __launch_bounds__(128, MINB) k_attention_rows(AttArgs a) {
  constexpr int KPI = 32 / G;  // key rows per warp iteration; U = iterations per chunk
  __shared__ float red_m[4], red_l[4];
  __shared__ float4 red_acc[4][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gl = lane & (G - 1), gk = lane / G;
  const int64_t total = a.B * a.H * a.Lq;
  const int64_t w = (WPR == 1) ? (int64_t)blockIdx.x * 4 + wib : (int64_t)blockIdx.x;
  if (w >= total) return;  // WPR == 4: uniform per CTA; WPR == 1: no block-wide barrier below
  const int64_t iq = w % a.Lq, h = (w / a.Lq) % a.H, b = w / (a.Lq * a.H);
  const int D = (int)a.D;
  if (a.lk_dev) a.Lk = *a.lk_dev + a.lk_add;
  const bool act = gl * 4 < D;
  float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) {
    const float* qrow = a.q + b * a.qs[0] + h * a.qs[1] + iq * a.qs[2] + gl * 4;
    q4 = make_float4(qrow[0] * a.scale, qrow[1] * a.scale, qrow[2] * a.scale, qrow[3] * a.scale);
  }
  const float* kb = a.k + b * a.ks[0] + h * a.ks[1] + gl * 4;
  const float* vb = a.v + b * a.vs[0] + h * a.vs[1] + gl * 4;
  const float* mrow = a.mask ? a.mask + b * a.mask_bs + iq * a.mask_qs : nullptr;
  float  m = -INFINITY, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int CH = KPI * U;  // keys per chunk
  for (int64_t j0 = (WPR == 1) ? 0 : (int64_t)wib * CH; j0 < a.Lk; j0 += (int64_t)CH * WPR) {
    float4 kk[U], vv[U];
    float  s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = j0 + u * KPI + gk;
      const bool ok = act && j < a.Lk;
      kk[u] = ok ? __ldcs(reinterpret_cast<const float4*>(kb + j * a.ks[2])) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = j0 + u * KPI + gk;
      const bool ok = act && j < a.Lk;
      vv[u] = ok ? __ldcs(reinterpret_cast<const float4*>(vb + j * a.vs[2])) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float cm = -INFINITY;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float dot = q4.x * kk[u].x + q4.y * kk[u].y + q4.z * kk[u].z + q4.w * kk[u].w;
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const int64_t j = j0 + u * KPI + gk;
      s[u] = (j < a.Lk) ? (mrow ? dot + __ldg(mrow + j) : dot) : -INFINITY;
      cm = fmaxf(cm, s[u]);
    }
    const float mn = fmaxf(m, cm);
    if (mn != -INFINITY) {  // -inf only while every key so far is masked out: keep the state empty
      const float corr = __expf(m - mn);
      l *= corr;
      acc.x *= corr; acc.y *= corr; acc.z *= corr; acc.w *= corr;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float p = __expf(s[u] - mn);  // exp(-inf) = 0 for masked / out-of-range keys
        l += p;
        acc.x += p * vv[u].x; acc.y += p * vv[u].y; acc.z += p * vv[u].z; acc.w += p * vv[u].w;
      }
      m = mn;
    }
  }
  // merge the key groups of the warp
#pragma unroll
  for (int o = G; o < 32; o <<= 1) {
    const float  m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float4 a2 = make_float4(__shfl_xor_sync(0xffffffffu, acc.x, o), __shfl_xor_sync(0xffffffffu, acc.y, o),
                                  __shfl_xor_sync(0xffffffffu, acc.z, o), __shfl_xor_sync(0xffffffffu, acc.w, o));
    const float mn = fmaxf(m, m2);
    const float c1 = (m == -INFINITY) ? 0.f : __expf(m - mn), c2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
    l = l * c1 + l2 * c2;
    acc.x = acc.x * c1 + a2.x * c2; acc.y = acc.y * c1 + a2.y * c2; acc.z = acc.z * c1 + a2.z * c2; acc.w = acc.w * c1 + a2.w * c2;
    m = mn;
  }
  if (WPR == 4) {  // merge the 4 warps of the row
    if (lane == 0) { red_m[wib] = m; red_l[wib] = l; }
    if (lane < G) red_acc[wib][lane] = acc;
    __syncthreads();
    if (wib != 0) return;
    const float mt = fmaxf(fmaxf(red_m[0], red_m[1]), fmaxf(red_m[2], red_m[3]));
    float  lt = 0.f;
    float4 at = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) {
      const float  sc = (red_m[ww] == -INFINITY) ? 0.f : __expf(red_m[ww] - mt);
      const float4 r = red_acc[ww][gl];
      lt += red_l[ww] * sc;
      at.x += r.x * sc; at.y += r.y * sc; at.z += r.z * sc; at.w += r.w * sc;
    }
    m = mt; l = lt; acc = at;
  }
  if (gk == 0 && act) {
    const float inv = 1.f / l;
    const float4 o = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    if (a.out_planes) {  // the O-projection GEMM consumes this directly (no fp32 round trip, no pack launch)
      const int64_t rows = a.B * a.Lq, r = b * a.Lq + iq;
      __nv_bfloat16 *hi = a.out_planes + r * a.planes_kp + h * a.D + gl * 4, *lo = hi + rows * a.planes_kp;
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(o.x, o.y), h23 = __floats2bfloat162_rn(o.z, o.w);
      const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(o.x - f01.x, o.y - f01.y), l23 = __floats2bfloat162_rn(o.z - f23.x, o.w - f23.y);
      *reinterpret_cast<uint2*>(hi) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
      *reinterpret_cast<uint2*>(lo) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    } else {
      *reinterpret_cast<float4*>(a.out + ((b * a.Lq + iq) * a.H + h) * a.D + gl * 4) = o;
    }
  }
  if (lane == 0 && a.lse) a.lse[(b * a.H + h) * a.Lq + iq] = m + logf(l);
}

// true when q/k/v/out admit the float4 row kernel: D a multiple of 4, 16-byte aligned bases, strides multiples of 4 elements
static bool rows_kernel_ok(const AttArgs& a) {
  if (a.D % 4 != 0 || a.D > 128) return false;
  auto al = [](const void* p) { return (((uintptr_t)p) & 15) == 0; };
  if (!al(a.k) || !al(a.v)) return false;
  for (int i = 0; i < 3; ++i)
    if ((a.ks[i] & 3) || (a.vs[i] & 3)) return false;
  if (a.out_planes ? (a.planes_kp & 3) != 0 : !al(a.out)) return false;
  return true;
}

template <int G>
static int launch_rows_g(const AttArgs& a, int64_t total) {
  // the 4 warps of a CTA split the keys of one row: finer work items (no half-empty last wave at batch 1024) and 4x the
  // loads in flight per row; 64 registers -> 8 CTAs per SM (measured 76 % of HBM peak vs 60 % with one warp per row)
  k_attention_rows<G, 4, 4, 8><<<(unsigned)total, 128, 0, stream()>>>(a);
  PDN_LAUNCHED("attention_rows");
  return 0;
}

static int launch_rows(const AttArgs& a, int64_t total) {
  if (a.D <= 32) return launch_rows_g<8>(a, total);
  if (a.D <= 64) return launch_rows_g<16>(a, total);
  return launch_rows_g<32>(a, total);
}

struct AttBwdArgs {
  AttArgs f;
  const float* g;  // grad of out, [B, Lq, H, D] contiguous
  float *dq, *dk, *dv;  // dq [B,Lq,H,D]; dk/dv [B,Lk,H,D] contiguous, pre-zeroed (atomic accumulation)
};

// one warp per (b, h, query row): Δ = Σ_d g·o ; for each key: p = exp(s - lse), dv_j += p g, ds = p (g·v_j − Δ),
// dq += ds k_j scale, dk_j += ds q scale
__global__ void __launch_bounds__(128) k_attention_bwd(AttBwdArgs a) {
  extern __shared__ float sm[];  // per warp: q*scale [D], g [D]
  const AttArgs& f = a.f;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t w = (int64_t)blockIdx.x * 4 + wib;
  const int64_t total = f.B * f.H * f.Lq;
  if (w >= total) return;
  const int64_t iq = w % f.Lq, h = (w / f.Lq) % f.H, b = w / (f.Lq * f.H);
  const int D = (int)f.D;
  float* qsm = sm + wib * 2 * ATT_MAXD;
  float* gsm = qsm + ATT_MAXD;
  const float* qrow = f.q + b * f.qs[0] + h * f.qs[1] + iq * f.qs[2];
  const int64_t orow = ((b * f.Lq + iq) * f.H + h) * f.D;
  float delta = 0.f;
  for (int d = lane; d < D; d += 32) {
    qsm[d] = qrow[d] * f.scale;
    float gg = a.g[orow + d];
    gsm[d] = gg;
    delta += gg * f.out[orow + d];
  }
  delta = warp_sum(delta);
  __syncwarp();
  const float lse = f.lse[(b * f.H + h) * f.Lq + iq];
  const float* kb = f.k + b * f.ks[0] + h * f.ks[1];
  const float* vb = f.v + b * f.vs[0] + h * f.vs[1];
  const float* mrow = f.mask ? f.mask + b * f.mask_bs + iq * f.mask_qs : nullptr;
  float dq[ATT_DPL];
#pragma unroll
  for (int i = 0; i < ATT_DPL; ++i) dq[i] = 0.f;
  for (int64_t j0 = 0; j0 < f.Lk; j0 += 32) {
    const int64_t j = j0 + lane;
    float p = 0.f, ds = 0.f;
    if (j < f.Lk) {
      const float *kr = kb + j * f.ks[2], *vr = vb + j * f.vs[2];
      float dot = 0.f, gv = 0.f;
      for (int d = 0; d < D; ++d) {
        dot += qsm[d] * kr[d];
        gv += gsm[d] * vr[d];
      }
      float s = dot + (mrow ? mrow[j] : 0.f);
      p = (s == -INFINITY) ? 0.f : __expf(s - lse);
      ds = p * (gv - delta);
    }
    const int cnt = (int)((f.Lk - j0) < 32 ? (f.Lk - j0) : 32);
    for (int t = 0; t < cnt; ++t) {
      const float pt = __shfl_sync(0xffffffffu, p, t);
      const float dst = __shfl_sync(0xffffffffu, ds, t);
      if (pt == 0.f && dst == 0.f) continue;
      const float* kr = kb + (j0 + t) * f.ks[2];
      const int64_t kvrow = ((b * f.Lk + j0 + t) * f.H + h) * f.D;
#pragma unroll
      for (int i = 0; i < ATT_DPL; ++i) {
        int d = lane + i * 32;
        if (d < D) {
          dq[i] += dst * kr[d];
          if (a.dv) atomicAdd(a.dv + kvrow + d, pt * gsm[d]);
          if (a.dk) atomicAdd(a.dk + kvrow + d, dst * qsm[d]);  // qsm already carries the scale
        }
      }
    }
  }
  if (a.dq) {
#pragma unroll
    for (int i = 0; i < ATT_DPL; ++i) {
      int d = lane + i * 32;
      if (d < D) a.dq[orow + d] = dq[i] * f.scale;
    }
  }
}

// ---------------------------------------------------------------- RoPE + KV append -------------------------------
// q,k,v: rows r = b*L + l of [H, D] values, `ld` elements apart (position pos0 + l); pairs (2i, 2i+1) rotated by angle[pos][i].
// k (rotated) and v rows are also written into the caches [Bmax, S, H, D] at [b, pos0 + l].
__global__ void __launch_bounds__(256) k_rope_kv_append(float* __restrict__ q, float* __restrict__ k, const float* __restrict__ v,
                                                        const float* __restrict__ cosT, const float* __restrict__ sinT, float* __restrict__ ck,
                                                        float* __restrict__ cv, int64_t B, int64_t L, int64_t H, int64_t D, int64_t S, int64_t pos0,
                                                        const int64_t* __restrict__ pos_dev, int64_t ld) {
  if (pos_dev) pos0 = *pos_dev;
  if (pos0 < 0 || pos0 + L > S) return;  // out of the cache: the host-side check could not run for a device-side position
  const int64_t half = D / 2, total = B * L * H * half;
  if ((D & 3) == 0 && (ld & 3) == 0 && total < 0x7fffffff && B * L * ld < 0x7fffffff && B * S * H * D < 0x7fffffff &&
      (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)ck | (uintptr_t)cv | (uintptr_t)cosT | (uintptr_t)sinT) & 15) == 0) {
    // two rotation pairs per thread: 128-bit loads/stores everywhere, 32-bit index arithmetic
    const uint32_t q4 = (uint32_t)(D >> 2), hq = (uint32_t)H * q4, tot4 = (uint32_t)(B * L) * hq, Lu = (uint32_t)L;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < tot4; i += gridDim.x * blockDim.x) {
      const uint32_t r = i / hq, rem = i - r * hq, hh = rem / q4, pj = rem - hh * q4;
      const uint32_t b = r / Lu, l = r - b * Lu;
      const int64_t pos = pos0 + l;
      const float2 c = __ldg(reinterpret_cast<const float2*>(cosT + pos * half) + pj), sn = __ldg(reinterpret_cast<const float2*>(sinT + pos * half) + pj);
      const uint32_t e = r * (uint32_t)ld + hh * (uint32_t)D + 4 * pj;
      const float4 qq = *reinterpret_cast<const float4*>(q + e), kk = *reinterpret_cast<const float4*>(k + e);
      *reinterpret_cast<float4*>(q + e) = make_float4(qq.x * c.x - qq.y * sn.x, qq.x * sn.x + qq.y * c.x, qq.z * c.y - qq.w * sn.y, qq.z * sn.y + qq.w * c.y);
      const float4 kr = make_float4(kk.x * c.x - kk.y * sn.x, kk.x * sn.x + kk.y * c.x, kk.z * c.y - kk.w * sn.y, kk.z * sn.y + kk.w * c.y);
      *reinterpret_cast<float4*>(k + e) = kr;
      if (ck) {
        const int64_t ce = (((int64_t)b * S + pos) * H + hh) * D + 4 * pj;
        *reinterpret_cast<float4*>(ck + ce) = kr;
        *reinterpret_cast<float4*>(cv + ce) = *reinterpret_cast<const float4*>(v + e);
      }
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pi = i % half, hh = (i / half) % H, r = i / (half * H);
    const int64_t b = r / L, l = r % L, pos = pos0 + l;
    const float c = cosT[pos * half + pi], s = sinT[pos * half + pi];
    const int64_t e = r * ld + hh * D + 2 * pi;  // ld = elements between consecutive rows (H*D, or 3*H*D inside a fused QKV buffer)
    float2 qq = *reinterpret_cast<float2*>(q + e);
    *reinterpret_cast<float2*>(q + e) = make_float2(qq.x * c - qq.y * s, qq.x * s + qq.y * c);
    float2 kk = *reinterpret_cast<float2*>(k + e);
    float2 kr = make_float2(kk.x * c - kk.y * s, kk.x * s + kk.y * c);
    *reinterpret_cast<float2*>(k + e) = kr;
    if (ck) {
      const int64_t ce = ((b * S + pos) * H + hh) * D + 2 * pi;
      *reinterpret_cast<float2*>(ck + ce) = kr;
      *reinterpret_cast<float2*>(cv + ce) = *reinterpret_cast<const float2*>(v + e);
    }
  }
}

}  // namespace pdn

using namespace pdn;

static int fill_att(AttArgs& a, const float* q, const float* k, const float* v, const float* mask, float* out, float* lse, int64_t B, int64_t H,
                    int64_t Lq, int64_t Lk, int64_t D, const int64_t* q_str, const int64_t* k_str, const int64_t* v_str, const int64_t* mask_str,
                    float scale) {
  PDN_CHECK(D >= 1 && D <= ATT_MAXD, "attention: head dim %lld not in [1, %d]", (long long)D, ATT_MAXD);
  a.q = q; a.k = k; a.v = v; a.mask = mask; a.out = out; a.lse = lse;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.D = D;
  for (int i = 0; i < 3; ++i) { a.qs[i] = q_str[i]; a.ks[i] = k_str[i]; a.vs[i] = v_str[i]; }
  a.mask_bs = mask && mask_str ? mask_str[0] : 0;
  a.mask_qs = mask && mask_str ? mask_str[1] : 0;
  a.scale = scale;
  a.lk_dev = nullptr;
  a.lk_add = 0;
  a.out_planes = nullptr;
  a.planes_kp = 0;
  return 0;
}

extern "C" {

int pdn_attention_fwd(const float* q, const float* k, const float* v, const float* mask, float* out, float* lse, int64_t B, int64_t H, int64_t Lq,
                      int64_t Lk, int64_t D, const int64_t* q_str, const int64_t* k_str, const int64_t* v_str, const int64_t* mask_str,
                      float scale, void* out_planes, int64_t planes_kp) {
  PDN_TRY(ensure_init());
  AttArgs a;
  PDN_TRY(fill_att(a, q, k, v, mask, out, lse, B, H, Lq, Lk, D, q_str, k_str, v_str, mask_str, scale));
  PDN_CHECK(!out_planes || (planes_kp >= H * D && (planes_kp & 7) == 0), "attention: bad planes padding");
  a.out_planes = (__nv_bfloat16*)out_planes;
  a.planes_kp = planes_kp;
  const int64_t total = B * H * Lq;
  if (total == 0) return 0;
  PDN_CHECK(Lk > 0, "attention: no keys");
  PDN_CHECK((total + 3) / 4 <= 0x7fffffff, "attention: too many query rows");
  if (Lq <= 16 && rows_kernel_ok(a)) return launch_rows(a, total);  // decode / short prefill: coalesced float4 row kernel
  if (Lq <= 4 && Lk >= 64 && total < (int64_t)sm_count() * 16) {  // decode with few rows: split the keys of each row over a whole CTA
    k_attention_decode<<<(unsigned)total, 128, 0, stream()>>>(a);
    PDN_LAUNCHED("attention_decode");
    return 0;
  }
  k_attention_fwd<<<(unsigned)((total + 3) / 4), 128, 4 * ATT_MAXD * sizeof(float), stream()>>>(a);
  PDN_LAUNCHED("attention_fwd");
  return 0;
}

int pdn_attention_bwd(const float* q, const float* k, const float* v, const float* mask, const float* out, const float* lse, const float* g_out,
                      float* dq, float* dk, float* dv, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const int64_t* q_str,
                      const int64_t* k_str, const int64_t* v_str, const int64_t* mask_str, float scale) {
  PDN_TRY(ensure_init());
  AttBwdArgs a;
  PDN_TRY(fill_att(a.f, q, k, v, mask, const_cast<float*>(out), const_cast<float*>(lse), B, H, Lq, Lk, D, q_str, k_str, v_str, mask_str, scale));
  a.g = g_out; a.dq = dq; a.dk = dk; a.dv = dv;
  const size_t kv_bytes = (size_t)(B * Lk * H * D) * sizeof(float);
  if (dk) PDN_CUDA(cudaMemsetAsync(dk, 0, kv_bytes, stream()));
  if (dv) PDN_CUDA(cudaMemsetAsync(dv, 0, kv_bytes, stream()));
  const int64_t total = B * H * Lq;
  if (total == 0) return 0;
  k_attention_bwd<<<(unsigned)((total + 3) / 4), 128, 8 * ATT_MAXD * sizeof(float), stream()>>>(a);
  PDN_LAUNCHED("attention_bwd");
  return 0;
}

int pdn_rope_kv_append(float* q, float* k, const float* v, const float* cosT, const float* sinT, float* cache_k, float* cache_v, int64_t B,
                       int64_t L, int64_t H, int64_t D, int64_t S, int64_t pos0, int64_t ld) {
  PDN_TRY(ensure_init());
  PDN_CHECK(D % 2 == 0, "rope: head dim must be even");
  PDN_CHECK(!cache_k || pos0 + L <= S, "rope_kv_append: positions [%lld, %lld) exceed the cache length %lld", (long long)pos0, (long long)(pos0 + L),
            (long long)S);
  const int64_t total = B * L * H * (D / 2);
  if (total == 0) return 0;
  k_rope_kv_append<<<grid_for(total, 256), 256, 0, stream()>>>(q, k, v, cosT, sinT, cache_k, cache_v, B, L, H, D, S, pos0, nullptr, ld > 0 ? ld : H * D);
  PDN_LAUNCHED("rope_kv_append");
  return 0;
}

/* Device-scalar variants used when one decode step is captured as a CUDA graph: the position lives in device memory
 * (pos_dev), so the recorded launches stay valid while the sequence grows. */
int pdn_rope_kv_append_dev(float* q, float* k, const float* v, const float* cosT, const float* sinT, float* cache_k, float* cache_v, int64_t B,
                           int64_t L, int64_t H, int64_t D, int64_t S, const int64_t* pos_dev, int64_t ld) {
  PDN_TRY(ensure_init());
  PDN_CHECK(D % 2 == 0 && pos_dev != nullptr, "rope_dev: bad arguments");
  const int64_t total = B * L * H * (D / 2);
  if (total == 0) return 0;
  k_rope_kv_append<<<grid_for(total, 256), 256, 0, stream()>>>(q, k, v, cosT, sinT, cache_k, cache_v, B, L, H, D, S, 0, pos_dev, ld > 0 ? ld : H * D);
  PDN_LAUNCHED("rope_kv_append");
  return 0;
}

int pdn_attention_fwd_dev(const float* q, const float* k, const float* v, float* out, int64_t B, int64_t H, int64_t Lq, int64_t D,
                          const int64_t* q_str, const int64_t* k_str, const int64_t* v_str, float scale, const int64_t* pos_dev, int64_t lk_add,
                          void* out_planes, int64_t planes_kp) {
  PDN_TRY(ensure_init());
  PDN_CHECK(pos_dev != nullptr, "attention_dev: null position");
  AttArgs a;
  PDN_TRY(fill_att(a, q, k, v, nullptr, out, nullptr, B, H, Lq, 1, D, q_str, k_str, v_str, nullptr, scale));
  PDN_CHECK(!out_planes || (planes_kp >= H * D && (planes_kp & 7) == 0), "attention: bad planes padding");
  a.out_planes = (__nv_bfloat16*)out_planes;
  a.planes_kp = planes_kp;
  a.lk_dev = pos_dev;
  a.lk_add = lk_add;
  const int64_t total = B * H * Lq;
  if (total == 0) return 0;
  if (Lq <= 16 && rows_kernel_ok(a)) return launch_rows(a, total);
  if (Lq <= 4 && total < (int64_t)sm_count() * 16) {  // graph-replayed decode step with few rows (key count only known on the device)
    k_attention_decode<<<(unsigned)total, 128, 0, stream()>>>(a);
    PDN_LAUNCHED("attention_decode");
    return 0;
  }
  k_attention_fwd<<<(unsigned)((total + 3) / 4), 128, 4 * ATT_MAXD * sizeof(float), stream()>>>(a);
  PDN_LAUNCHED("attention_fwd");
  return 0;
}

}  // extern "C"
