// Persistent whole-model decode step for SMALL batches (rows < 32): the reference's own configuration, llm/llama/infer.py:19-37
// (max_batch_size = 1), runs one token at a time through ~490 eager array expressions. At one token the model is
// latency-bound, not bandwidth- or FLOP-bound: 61 MB of fp32 weights are L2-resident after the first token and a token needs
// 30 MFLOP, so what costs time is the NUMBER OF DEPENDENT STEPS. This kernel runs a full decode step — embedding row, per layer
// {RMSNorm, Q/K/V projection, RoPE, KV-cache append, attention over the cache, O projection + residual, RMSNorm, gate/up, SwiGLU,
// down projection + residual}, final RMSNorm, lm_head, greedy argmax (reference llm/llama/model.py:142-150, 95-121, 56-58,
// 192-207, 254-256, 268) — as ONE cooperative launch of one CTA per SM with five grid barriers per layer:
//
//   S1  every CTA normalises the B residual rows itself (288 floats: cheaper than a barrier), then one WARP per pair of
//       output columns of [Wq|Wk|Wv]ᵀ (rows are unit-stride: 128-bit loads), rotates the pair (RoPE) and writes q to scratch,
//       k/v straight into the KV cache at [b, pos]
//   S2  one warp per (sequence, head, key split): lane-per-key online softmax over the cached keys, partial (m, l, acc[hd])
//   S3  every CTA merges the partials of all heads (tiny), one warp per output column of Woᵀ, residual add in place
//   S4  RMSNorm again per CTA, one warp per hidden unit: gate and up rows interleaved in memory, SwiGLU applied in the epilogue
//   S5  one warp per output column of W_downᵀ, residual add in place
//   end final RMSNorm per CTA, one warp per vocabulary row of W_lmᵀ (+bias): logits written, per-CTA argmax partials, the last
//       CTA to finish (atomic ticket) reduces them to the token id (first occurrence wins, like NumPy's argmax)
//
// Weights are read through transposed copies ([out][in], made once per weight version by the host side) so that a warp streams
// one contiguous row per output; activations cross CTAs through L2 only (ld.global.cg / st + release-acquire barrier: L1 is never
// trusted for mutable data). Summation order is fixed (lane-strided partial sums, xor-shuffle tree, partials merged in index
// order): results are bit-reproducible run to run. fp32 FFMA throughout — the reference's arithmetic type.
#include "common.cuh"

#include <vector>

namespace pdn {

struct MegaLayer {
  const float* wqkv_t;  // [3*dim][dim]
  const float* wo_t;    // [dim][dim]
  const float* wgu_t;   // [2*FF][dim], row 2j = gate column j, row 2j+1 = up column j
  const float* wd_t;    // [dim][FF]
  const float* n1;      // [dim]
  const float* n2;      // [dim]
  float*       ck;      // [Bmax][S][H][hd]
  float*       cv;
  float        eps1, eps2;
};

struct MegaArgs {
  const MegaLayer* layers;
  int              n_layers, B, dim, H, hd, FF, V, S;
  const float *    emb, *cosT, *sinT, *norm_w, *wlm_t, *lm_bias;
  float            eps_f;
  const int64_t*   ids;
  int64_t          ids_stride;
  int              pos;
  float *          h, *q, *part, *hid, *logits;
  int64_t*         ids_out;
  float*           amax_val;
  int*             amax_idx;
  unsigned long long* bar;
  unsigned long long  bar_base;
  unsigned int*       ticket;
  int                 nsplit;
};

constexpr int MEGA_THREADS = 512;
constexpr int MEGA_WARPS = MEGA_THREADS / 32;
constexpr int MEGA_MAXK = 1024;    // widest contraction kept in shared memory per row (max(dim, FF))
constexpr int MEGA_MAXSPLIT = 32;  // key splits per (sequence, head)

__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
    unsigned long long v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
      if (v < target && clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a lost CTA must surface as an error, never as a hung GPU
    } while (v < target);
  }
  __syncthreads();
}

// xs[b][k] = x[b][k] * rsqrt(mean_k x[b]^2 + eps) * w[k]   (reference norm.py:245-248); x rows read through L2
template <int NB>
__device__ __forceinline__ void load_norm_rows(const float* const* rows, const float* __restrict__ w, float eps, int K, float* xs, float* red) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float ss[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) ss[b] = 0.f;
  for (int k = tid; k < K; k += MEGA_THREADS) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float v = __ldcg(rows[b] + k);
      xs[b * MEGA_MAXK + k] = v;
      ss[b] += v * v;
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const float s = warp_sum(ss[b]);
    if (lane == 0) red[b * MEGA_WARPS + wid] = s;
  }
  __syncthreads();
  float rstd[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < MEGA_WARPS; ++i) tot += red[b * MEGA_WARPS + i];
    rstd[b] = 1.0f / sqrtf(tot / (float)K + eps);
  }
  for (int k = tid; k < K; k += MEGA_THREADS) {
    const float wk = __ldg(w + k);
#pragma unroll
    for (int b = 0; b < NB; ++b) xs[b * MEGA_MAXK + k] = xs[b * MEGA_MAXK + k] * rstd[b] * wk;
  }
  __syncthreads();
}

// plain copy of NB rows of K floats into shared memory (through L2)
template <int NB>
__device__ __forceinline__ void load_rows(const float* src, int ld, int K, float* xs) {
  for (int k = threadIdx.x; k < K; k += MEGA_THREADS) {
#pragma unroll
    for (int b = 0; b < NB; ++b) xs[b * MEGA_MAXK + k] = __ldcg(src + (size_t)b * ld + k);
  }
  __syncthreads();
}

// acc[b] = sum_k xs[b][k] * wrow[k]: lane-strided float4 partial sums + xor-shuffle tree; result valid in every lane
template <int NB>
__device__ __forceinline__ void warp_dot(const float* __restrict__ wrow, int K4, const float* xs, float* acc) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = 0.f;
  const float4* w4 = reinterpret_cast<const float4*>(wrow);
  for (int k = lane; k < K4; k += 32) {
    const float4 w = __ldg(w4 + k);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float4 x = *reinterpret_cast<const float4*>(xs + b * MEGA_MAXK + 4 * k);
      acc[b] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, acc[b]))));
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = warp_sum(acc[b]);
}

// two rows at once (independent loads in flight together): Q/K/V rotation pairs, gate|up
template <int NB>
__device__ __forceinline__ void warp_dot2(const float* __restrict__ w0, const float* __restrict__ w1, int K4, const float* xs, float* acc0, float* acc1) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int b = 0; b < NB; ++b) acc0[b] = acc1[b] = 0.f;
  const float4 *p0 = reinterpret_cast<const float4*>(w0), *p1 = reinterpret_cast<const float4*>(w1);
#pragma unroll 2
  for (int k = lane; k < K4; k += 32) {
    const float4 a0 = __ldg(p0 + k), a1 = __ldg(p1 + k);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float4 x = *reinterpret_cast<const float4*>(xs + b * MEGA_MAXK + 4 * k);
      acc0[b] = fmaf(a0.x, x.x, fmaf(a0.y, x.y, fmaf(a0.z, x.z, fmaf(a0.w, x.w, acc0[b]))));
      acc1[b] = fmaf(a1.x, x.x, fmaf(a1.y, x.y, fmaf(a1.z, x.z, fmaf(a1.w, x.w, acc1[b]))));
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    acc0[b] = warp_sum(acc0[b]);
    acc1[b] = warp_sum(acc1[b]);
  }
}

template <int NB, int HD4>
__global__ void __launch_bounds__(MEGA_THREADS, 1) k_decode_mega(MegaArgs a) {
  __shared__ __align__(16) float xs[NB * MEGA_MAXK];
  __shared__ float               red[NB * MEGA_WARPS];
  __shared__ int                 redi[NB * MEGA_WARPS];
  __shared__ __align__(16) float qs[MEGA_WARPS][64];  // the query row of the (sequence, head) a warp is working on in S2
  constexpr int HD = HD4 * 4;
  const int     tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int     G = gridDim.x, cta = blockIdx.x;
  const int     gw = wid * G + cta, nw = G * MEGA_WARPS;  // interleaved so that a short task list still touches every SM
  const int     dim = a.dim, FF = a.FF, H = a.H, B = a.B, pos = a.pos;
  const int     Lk = pos + 1;
  unsigned long long bar_t = a.bar_base;
  const float   scale = rsqrtf((float)HD);

  const float* rows[NB];
  for (int l = 0; l < a.n_layers; ++l) {
    const MegaLayer& Lw = a.layers[l];
    // residual rows: the embedding rows of the incoming token ids for the first layer (reference model.py:194), h afterwards
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int bb = b < B ? b : B - 1;
      rows[b] = (l == 0) ? a.emb + (size_t)a.ids[(size_t)bb * a.ids_stride] * dim : a.h + (size_t)bb * dim;
    }
    // ---------------- S1: RMSNorm -> Q/K/V columns (pairs) -> RoPE -> q scratch / KV cache --------------------------------
    load_norm_rows<NB>(rows, Lw.n1, Lw.eps1, dim, xs, red);
    {
      const int npairs = 3 * dim / 2, K4 = dim / 4;
      for (int t = gw; t < npairs; t += nw) {
        float y0[NB], y1[NB];
        warp_dot2<NB>(Lw.wqkv_t + (size_t)(2 * t) * dim, Lw.wqkv_t + (size_t)(2 * t + 1) * dim, K4, xs, y0, y1);
        if (lane == 0) {
          const int col = 2 * t, which = col / dim, c = col - which * dim;  // 0: q, 1: k, 2: v
          const int hh = c / HD, d = c - hh * HD;
          const float cs = __ldg(a.cosT + (size_t)pos * (HD / 2) + d / 2), sn = __ldg(a.sinT + (size_t)pos * (HD / 2) + d / 2);
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            if (b < B) {
              float o0 = y0[b], o1 = y1[b];
              if (which < 2) {  // interleaved-pair rotation (reference model.py:23-44)
                o0 = y0[b] * cs - y1[b] * sn;
                o1 = y0[b] * sn + y1[b] * cs;
              }
              float* dst = which == 0 ? a.q + (size_t)b * dim + c
                                      : (which == 1 ? Lw.ck : Lw.cv) + (((size_t)b * a.S + pos) * H + hh) * HD + d;
              *reinterpret_cast<float2*>(dst) = make_float2(o0, o1);
            }
          }
        }
      }
    }
    bar_t += G;
    grid_barrier(a.bar, bar_t);
    // ---------------- S2: attention partials, one warp per (b, head, key split), one key per lane and step ---------------
    {
      const int ns = a.nsplit, units = B * H * ns;
      for (int u = gw; u < units; u += nw) {
        const int sp = u % ns, hh = (u / ns) % H, b = u / (ns * H);
        __syncwarp();
        if (lane < HD4) reinterpret_cast<float4*>(qs[wid])[lane] = __ldcg(reinterpret_cast<const float4*>(a.q + (size_t)b * dim + hh * HD) + lane);
        __syncwarp();
        const float4* qv = reinterpret_cast<const float4*>(qs[wid]);
        float m = -INFINITY, lsum = 0.f;
        float4 acc[HD4];
#pragma unroll
        for (int i = 0; i < HD4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = sp * 32 + lane; s < Lk; s += ns * 32) {
          const size_t off = (((size_t)b * a.S + s) * H + hh) * HD;
          const float4* kr = reinterpret_cast<const float4*>(Lw.ck + off);
          const float4* vr = reinterpret_cast<const float4*>(Lw.cv + off);
          float sc = 0.f;
#pragma unroll
          for (int i = 0; i < HD4; ++i) {
            const float4 kk = __ldcg(kr + i);
            sc = fmaf(kk.x, qv[i].x, fmaf(kk.y, qv[i].y, fmaf(kk.z, qv[i].z, fmaf(kk.w, qv[i].w, sc))));
          }
          sc *= scale;
          const float mn = fmaxf(m, sc), corr = __expf(m - mn), p = __expf(sc - mn);
          lsum = lsum * corr + p;
#pragma unroll
          for (int i = 0; i < HD4; ++i) {
            const float4 vv = __ldcg(vr + i);
            acc[i].x = fmaf(p, vv.x, acc[i].x * corr);
            acc[i].y = fmaf(p, vv.y, acc[i].y * corr);
            acc[i].z = fmaf(p, vv.z, acc[i].z * corr);
            acc[i].w = fmaf(p, vv.w, acc[i].w * corr);
          }
          m = mn;
        }
        const float mw = warp_max(m);
        const float f = (m == -INFINITY) ? 0.f : __expf(m - mw);
        lsum = warp_sum(lsum * f);
        float* dst = a.part + (size_t)u * (HD + 2);
#pragma unroll
        for (int i = 0; i < HD4; ++i) {
          const float x = warp_sum(acc[i].x * f), y = warp_sum(acc[i].y * f), z = warp_sum(acc[i].z * f), w = warp_sum(acc[i].w * f);
          if (lane == 0) {
            dst[2 + 4 * i] = x;
            dst[3 + 4 * i] = y;
            dst[4 + 4 * i] = z;
            dst[5 + 4 * i] = w;
          }
        }
        if (lane == 0) {
          dst[0] = mw;
          dst[1] = lsum;
        }
      }
    }
    bar_t += G;
    grid_barrier(a.bar, bar_t);
    // ---------------- S3: merge partials (every CTA), O projection + residual, in place on h -----------------------------
    {
      const int ns = a.nsplit;
      for (int e = tid; e < B * dim; e += MEGA_THREADS) {
        const int b = e / dim, c = e - b * dim, hh = c / HD, d = c - hh * HD;
        const float* p0 = a.part + (size_t)((b * H + hh) * ns) * (HD + 2);
        float M = -INFINITY;
        for (int s = 0; s < ns; ++s) M = fmaxf(M, __ldcg(p0 + (size_t)s * (HD + 2)));
        float den = 0.f, num = 0.f;
        for (int s = 0; s < ns; ++s) {
          const float ms = __ldcg(p0 + (size_t)s * (HD + 2));
          const float f = (ms == -INFINITY) ? 0.f : __expf(ms - M);
          den = fmaf(__ldcg(p0 + (size_t)s * (HD + 2) + 1), f, den);
          num = fmaf(__ldcg(p0 + (size_t)s * (HD + 2) + 2 + d), f, num);
        }
        xs[b * MEGA_MAXK + c] = num / den;
      }
      __syncthreads();
      const int K4 = dim / 4;
      for (int n = gw; n < dim; n += nw) {
        float y[NB];
        warp_dot<NB>(Lw.wo_t + (size_t)n * dim, K4, xs, y);
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < B) a.h[(size_t)b * dim + n] = __ldcg(rows[b] + n) + y[b];
        }
      }
    }
    bar_t += G;
    grid_barrier(a.bar, bar_t);
    // ---------------- S4: RMSNorm -> gate / up rows (interleaved) -> SwiGLU -> hid ------------------------------------------
#pragma unroll
    for (int b = 0; b < NB; ++b) rows[b] = a.h + (size_t)(b < B ? b : B - 1) * dim;
    load_norm_rows<NB>(rows, Lw.n2, Lw.eps2, dim, xs, red);
    {
      const int K4 = dim / 4;
      for (int j = gw; j < FF; j += nw) {
        float g[NB], u[NB];
        warp_dot2<NB>(Lw.wgu_t + (size_t)(2 * j) * dim, Lw.wgu_t + (size_t)(2 * j + 1) * dim, K4, xs, g, u);
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < B) a.hid[(size_t)b * FF + j] = g[b] / (1.f + __expf(-g[b])) * u[b];  // x / (1 + exp(-x)), functional.py:39-40
        }
      }
    }
    bar_t += G;
    grid_barrier(a.bar, bar_t);
    // ---------------- S5: down projection + residual, in place on h ------------------------------------------------------------
    __syncthreads();
    load_rows<NB>(a.hid, FF, FF, xs);
    {
      const int K4 = FF / 4;
      for (int n = gw; n < dim; n += nw) {
        float y[NB];
        warp_dot<NB>(Lw.wd_t + (size_t)n * FF, K4, xs, y);
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < B) a.h[(size_t)b * dim + n] = __ldcg(a.h + (size_t)b * dim + n) + y[b];
        }
      }
    }
    bar_t += G;
    grid_barrier(a.bar, bar_t);
  }
  if (a.logits == nullptr) return;  // prompt positions before the last one: only the KV cache matters (model.py:255 keeps [-1])
  // ---------------- final RMSNorm -> lm_head rows (+bias) -> logits, argmax partials ---------------------------------------------
#pragma unroll
  for (int b = 0; b < NB; ++b) rows[b] = a.h + (size_t)(b < B ? b : B - 1) * dim;
  load_norm_rows<NB>(rows, a.norm_w, a.eps_f, dim, xs, red);
  float best[NB];
  int   besti[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    best[b] = -INFINITY;
    besti[b] = 0x7fffffff;
  }
  {
    const int K4 = dim / 4;
    for (int n = gw; n < a.V; n += nw) {
      float y[NB];
      warp_dot<NB>(a.wlm_t + (size_t)n * dim, K4, xs, y);
      const float bias = a.lm_bias ? __ldg(a.lm_bias + n) : 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float v = y[b] + bias;
        if (lane == 0 && b < B) a.logits[(size_t)b * a.V + n] = v;
        if (v > best[b] || (v == best[b] && n < besti[b])) {  // rows arrive in increasing n per warp: ties keep the first
          best[b] = v;
          besti[b] = n;
        }
      }
    }
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      red[b * MEGA_WARPS + wid] = best[b];
      redi[b * MEGA_WARPS + wid] = besti[b];
    }
  }
  __syncthreads();
  if (tid < NB && tid < B) {
    float bv = -INFINITY;
    int   bi = 0x7fffffff;
    for (int i = 0; i < MEGA_WARPS; ++i) {
      const float v = red[tid * MEGA_WARPS + i];
      const int   n = redi[tid * MEGA_WARPS + i];
      if (v > bv || (v == bv && n < bi)) {
        bv = v;
        bi = n;
      }
    }
    a.amax_val[(size_t)cta * NB + tid] = bv;
    a.amax_idx[(size_t)cta * NB + tid] = bi;
  }
  __shared__ unsigned int last;
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(a.ticket, 1u);
    last = (t == (unsigned)G - 1);
    if (last) *a.ticket = 0;  // every CTA has taken its ticket: safe to re-arm for the next launch
  }
  __syncthreads();
  if (last) {
    __threadfence();
    if (tid < NB && tid < B) {
      float bv = -INFINITY;
      int   bi = 0x7fffffff;
      for (int c = 0; c < G; ++c) {
        const float v = __ldcg(a.amax_val + (size_t)c * NB + tid);
        const int   n = __ldcg(a.amax_idx + (size_t)c * NB + tid);
        if (v > bv || (v == bv && n < bi)) {
          bv = v;
          bi = n;
        }
      }
      a.ids_out[tid] = bi;
    }
  }
}

struct MegaHandle {
  MegaArgs   args;
  MegaLayer* d_layers = nullptr;
  void*      scratch = nullptr;
  int        nb = 1, hd4 = 0, grid = 0, device = 0;
  unsigned long long launches = 0;
  int        bars_per_launch = 0;
};

template <int NB>
static const void* pick_kernel(int hd4) {
  switch (hd4) {
    case 8: return (const void*)k_decode_mega<NB, 8>;
    case 12: return (const void*)k_decode_mega<NB, 12>;
    case 16: return (const void*)k_decode_mega<NB, 16>;
  }
  return nullptr;
}
static const void* pick_kernel(int nb, int hd4) {
  switch (nb) {
    case 1: return pick_kernel<1>(hd4);
    case 2: return pick_kernel<2>(hd4);
    case 4: return pick_kernel<4>(hd4);
    case 8: return pick_kernel<8>(hd4);
  }
  return nullptr;
}

}  // namespace pdn

using namespace pdn;

extern "C" {

int pdn_decoder_create(void** handle, int n_layers, int B, int dim, int H, int FF, int V, int S, const void* const* layer_ptrs,
                       const float* layer_eps, const float* emb, const float* cosT, const float* sinT, const float* norm_w, float eps_f,
                       const float* wlm_t, const float* lm_bias) {
  PDN_TRY(ensure_init());
  PDN_CHECK(handle && n_layers > 0 && B >= 1 && B <= 8 && H > 0 && dim % H == 0, "decoder_create: bad arguments");
  const int hd = dim / H;
  PDN_CHECK(hd == 32 || hd == 48 || hd == 64, "decoder_create: head dim %d not in {32, 48, 64}", hd);
  PDN_CHECK(dim % 4 == 0 && FF % 4 == 0 && dim <= MEGA_MAXK && FF <= MEGA_MAXK, "decoder_create: dim / ffn width must be multiples of 4 and <= %d",
            MEGA_MAXK);
  MegaHandle* h = new MegaHandle();
  h->nb = B <= 1 ? 1 : (B <= 2 ? 2 : (B <= 4 ? 4 : 8));
  h->hd4 = hd / 4;
  h->grid = sm_count();
  cudaGetDevice(&h->device);
  const void* fn = pick_kernel(h->nb, h->hd4);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, MEGA_THREADS, 0) != cudaSuccess || per_sm < 1) {
    delete h;
    set_error("decoder_create: the persistent decode kernel does not fit one CTA per SM");
    return PDN_ERR_CUDA;
  }
  std::vector<MegaLayer> hl(n_layers);
  for (int l = 0; l < n_layers; ++l) {
    const void* const* p = layer_ptrs + (size_t)l * 8;
    hl[l] = MegaLayer{(const float*)p[0], (const float*)p[1], (const float*)p[2], (const float*)p[3], (const float*)p[4], (const float*)p[5],
                      (float*)p[6],       (float*)p[7],       layer_eps[2 * l],   layer_eps[2 * l + 1]};
  }
  // scratch: layer table | h [8][dim] | q [8][dim] | hid [8][FF] | partials [8][H][MAXSPLIT][hd+2] | argmax partials | barrier, ticket
  const size_t o_layers = 0, n_layers_b = sizeof(MegaLayer) * n_layers;
  auto   up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o_h = up(o_layers + n_layers_b), o_q = up(o_h + 8 * dim * 4), o_hid = up(o_q + 8 * dim * 4), o_part = up(o_hid + 8 * FF * 4);
  size_t o_av = up(o_part + (size_t)8 * H * MEGA_MAXSPLIT * (hd + 2) * 4), o_ai = up(o_av + (size_t)h->grid * 8 * 4);
  size_t o_bar = up(o_ai + (size_t)h->grid * 8 * 4), total = o_bar + 256;
  if (int r = dev_alloc(&h->scratch, total)) {
    delete h;
    return r;
  }
  char* base = (char*)h->scratch;
  cudaStream_t st = stream();
  PDN_CUDA(cudaMemsetAsync(base, 0, total, st));
  PDN_CUDA(cudaMemcpyAsync(base + o_layers, hl.data(), n_layers_b, cudaMemcpyHostToDevice, st));
  PDN_CUDA(cudaStreamSynchronize(st));  // hl is a host temporary
  MegaArgs& a = h->args;
  a.layers = (const MegaLayer*)(base + o_layers);
  a.n_layers = n_layers, a.B = B, a.dim = dim, a.H = H, a.hd = hd, a.FF = FF, a.V = V, a.S = S;
  a.emb = emb, a.cosT = cosT, a.sinT = sinT, a.norm_w = norm_w, a.wlm_t = wlm_t, a.lm_bias = lm_bias, a.eps_f = eps_f;
  a.h = (float*)(base + o_h), a.q = (float*)(base + o_q), a.hid = (float*)(base + o_hid), a.part = (float*)(base + o_part);
  a.amax_val = (float*)(base + o_av), a.amax_idx = (int*)(base + o_ai);
  a.bar = (unsigned long long*)(base + o_bar), a.ticket = (unsigned int*)(base + o_bar + 64);
  h->bars_per_launch = 5 * n_layers;
  *handle = h;
  return 0;
}

int pdn_decoder_step(void* handle, const int64_t* ids, int64_t ids_stride, int64_t pos, float* logits, int64_t* ids_out) {
  PDN_TRY(ensure_init());
  MegaHandle* h = (MegaHandle*)handle;
  PDN_CHECK(h && ids && pos >= 0 && pos < h->args.S, "decoder_step: position %lld outside the cache [0, %d)", (long long)pos, h ? h->args.S : 0);
  PDN_CHECK((logits == nullptr) == (ids_out == nullptr), "decoder_step: logits and ids_out are produced together");
  MegaArgs a = h->args;
  a.ids = ids, a.ids_stride = ids_stride, a.pos = (int)pos, a.logits = logits, a.ids_out = ids_out;
  const int keys = (int)pos + 1;
  int ns = (keys + 31) / 32;
  a.nsplit = ns < 1 ? 1 : (ns > MEGA_MAXSPLIT ? MEGA_MAXSPLIT : ns);
  a.bar_base = h->launches * (unsigned long long)h->bars_per_launch * (unsigned long long)h->grid;
  void* params[] = {&a};
  PDN_CUDA(cudaLaunchCooperativeKernel(pick_kernel(h->nb, h->hd4), dim3(h->grid), dim3(MEGA_THREADS), params, 0, stream()));
  PDN_LAUNCHED("decode_mega");
  h->launches++;  // only a launch that was accepted advances the barrier epoch
  return 0;
}

int pdn_decoder_destroy(void* handle) {
  MegaHandle* h = (MegaHandle*)handle;
  if (!h) return 0;
  if (h->scratch) dev_free(h->scratch);
  delete h;
  return 0;
}

}  // extern "C"
