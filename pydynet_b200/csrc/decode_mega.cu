// Persistent whole-model decode step for SMALL batches (rows < 32): the reference's own configuration, llm/llama/infer.py:19-37
// (max_batch_size = 1), runs one token at a time through ~490 eager array expressions. At one token the model is
// latency-bound, not bandwidth- or FLOP-bound: a token needs 30 MFLOP and 61 MB of fp32 weights (L2-resident after the first
// token), so what costs time is the NUMBER OF DEPENDENT STEPS and what sits on the critical path between them. This kernel runs a
// full decode step — embedding row, per layer {RMSNorm, Q/K/V projection, RoPE, KV-cache append, attention over the cache, O
// projection + residual, RMSNorm, gate/up, SwiGLU, down projection + residual}, final RMSNorm, lm_head, greedy argmax (reference
// llm/llama/model.py:142-150, 95-121, 56-58, 192-207, 254-256, 268) — as ONE cooperative launch of one CTA per SM, four dependent
// phases per layer and NO grid barrier: activations travel between CTAs as 8-byte {value, epoch} words (the low-latency "flag in the
// data" protocol: one relaxed 64-bit store per value, consumers poll the word until its epoch is the current launch number — no
// fence, no atomic, no barrier; a grid barrier costs ~4.7 µs on the 148 SMs / two dies of this part, 25 of them were the whole
// 120 µs of the previous version of this kernel):
//
//   P1  every CTA normalises the B residual rows itself (cheaper than a barrier), one WARP per pair of output columns of
//       [Wq|Wk|Wv]ᵀ (rows unit-stride: 128-bit loads), rotates the pair (RoPE), writes q to scratch and k/v into the KV cache
//   P2  one CTA per (sequence, head, key split): scores with 4 threads per key, block softmax statistics, P·V with 128-bit
//       coalesced V rows, and — because the O projection is linear — the UNNORMALISED partial O projection of this split
//       (head slice of Woᵀ, 288 x 48) in the same phase: no barrier between attention and O projection
//   P3  every CTA merges the split partials (softmax weights e^{m-M}/den), adds the residual, normalises, then one warp per hidden
//       unit: gate and up rows interleaved in memory, SwiGLU in the epilogue
//   P4  one warp per output column of W_downᵀ, residual add, new residual row to global
//   end final RMSNorm per CTA, lm_head rows (+bias): logits written, per-CTA argmax partials, the last CTA to finish (atomic
//       ticket) reduces them to the token id (first occurrence wins, like NumPy's argmax)
//
// What keeps each phase short:
//   * every warp has at most ONE task per phase, statically assigned, so the weight rows of the NEXT phase are loaded into
//     registers BEFORE the barrier (they do not depend on activations): after the barrier only the activations (one L2 round trip)
//     and the FMAs are left on the critical path;
//   * the CTA's block of lm_head rows (216 x 288 fp32 = 249 KB) starts streaming into shared memory with cp.async at kernel entry
//     and lands while the layers run, so the vocabulary projection at the end reads shared memory;
//   * attention and O projection are one phase, V rows are requested before the scores are known;
//   * every exchanged location is written exactly once per launch (one buffer set per layer), so the epoch test is the only
//     synchronisation; the KV cache rows of earlier positions were written by earlier launches (plain loads), the row of the current
//     position travels through the exchange buffers as well.
//
// Activations cross CTAs through L2 only (ld.relaxed.gpu / st.relaxed.gpu: L1 is never trusted for mutable data).
// Summation order is fixed (lane-strided partial sums, xor-shuffle tree, partials merged in index order): results are
// bit-reproducible run to run. fp32 FFMA throughout — the reference's arithmetic type.
#include "common.cuh"

#include <cmath>
#include <vector>

namespace pdn {

struct MegaLayer {
  const float* wqkv_t;  // [3*dim][dim]
  const float* wo_h;    // [H][dim][hd]: wo_h[h][n][d] = Wo[h*hd + d][n]
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
  unsigned long long* ll;  // exchange words: per layer {q, knew, vnew: [8][dim]; part: [MAXUNITS][dim+2]; hid: [8][FF]; hout: [8][dim]}
  unsigned int     epoch;  // launch number (> 0): the flag half of every word written by this launch
  float*           logits;
  int64_t*         ids_out;
  float*           amax_val;
  int*             amax_idx;
  unsigned int*       ticket;
  int                 nsplit;        // key splits per (sequence, head) in P2
  int                 lm_rows_cta;   // lm_head rows per CTA (contiguous block)
  int                 lm_rows_smem;  // how many of them are staged in shared memory
  unsigned long long* trace;         // PDN_MEGA_TRACE=1: globaltimer stamps of CTA 0 (and of the last P2 unit) at every phase boundary
};

__device__ __forceinline__ void mega_stamp(const MegaArgs& a, int& slot) {
  if (a.trace && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(blockIdx.x == 0 ? 0 : 512) + slot] = t;
  }
  ++slot;
}

constexpr int MEGA_THREADS = 512;
constexpr int MEGA_WARPS = MEGA_THREADS / 32;
constexpr int MEGA_MAXK = 1024;     // widest contraction kept in shared memory per row (max(dim, FF))
constexpr int MEGA_MAXUNITS = 160;  // (sequence, head, split) units of P2: one CTA each
constexpr int MEGA_SC = 2048;       // keys one P2 unit can take
constexpr int MEGA_LMROWS = 1024;   // lm_head rows per CTA (vocabulary <= 1024 * SMs)
constexpr int MEGA_PV = 12;         // partial rows a P3 thread polls in its first batch (heads x splits at one sequence, typical)
constexpr int MEGA_NJ = 6;          // float4 chunks of one weight row a lane keeps prefetched (row widths up to 768 floats)

// PDN_MEGA_TRACE: stamps INSIDE the phases of layer 2 for CTA 0 (a row CTA with a task in every phase, slots 256..) and for the first
// attention-unit CTA (slots 288..)
__device__ __forceinline__ void mega_fine(const MegaArgs& a, int layer, bool unit, int id) {
  if (a.trace && layer == 2 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(unit ? 288 : 256) + id] = t;
  }
}

__device__ __forceinline__ void ll_store(unsigned long long* p, float v, unsigned ep) {
  const unsigned long long w = ((unsigned long long)ep << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}

__device__ __forceinline__ unsigned long long ll_peek(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}

// a producer that never shows up must surface as an error, never as a hung GPU (~2 s)
__device__ __forceinline__ void ll_watchdog(long long& t0, int& spins) {
  if (++spins == 4096) {
    spins = 0;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ float ll_load(const unsigned long long* p, unsigned ep) {
  unsigned long long w = ll_peek(p);
  long long          t0 = 0;
  int                spins = 0;
  while ((unsigned)(w >> 32) != ep) {
    ll_watchdog(t0, spins);
    w = ll_peek(p);
  }
  return __uint_as_float((unsigned)w);
}

// N words at p[i * stride]: all requests in flight together, repeated until every word carries the epoch
template <int N>
__device__ __forceinline__ void ll_load_n(const unsigned long long* p, size_t stride, int n, unsigned ep, float* out) {
  unsigned long long w[N];
  long long          t0 = 0;
  int                spins = 0;
  bool               ok;
  do {
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = i < n ? ll_peek(p + (size_t)i * stride) : ((unsigned long long)ep << 32);
    ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) ok &= (unsigned)(w[i] >> 32) == ep;
    if (!ok) ll_watchdog(t0, spins);
  } while (!ok);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = __uint_as_float((unsigned)w[i]);
}

// the lane's chunks k = lane + 32 j (j < MEGA_NJ) of one weight row, requested now, consumed after the next barrier
__device__ __forceinline__ void pf_row(const float* __restrict__ row, int K4, float4* w) {
  const float4* p = reinterpret_cast<const float4*>(row);
  const int     lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < MEGA_NJ; ++j) {
    if (32 * j >= K4) break;  // warp-uniform; chunks beyond the row are never read by dot_row
    const int k = lane + 32 * j;
    w[j] = k < K4 ? __ldg(p + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// acc[b] = sum_k xs[b][k] * row[k] with the prefetched chunks (+ the tail of rows wider than 32*MEGA_NJ float4, loaded here);
// lane-strided partial sums + xor-shuffle tree; result valid in every lane
template <int NB>
__device__ __forceinline__ void dot_row(const float* __restrict__ row, int K4, const float4* w, const float* xs, float* acc) {
  const int lane = threadIdx.x & 31;
  float odd[NB];  // chunks alternate between two accumulators: two independent FMA chains instead of one
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = odd[b] = 0.f;
#pragma unroll
  for (int j = 0; j < MEGA_NJ; ++j) {
    if (32 * j >= K4) break;  // warp-uniform: rows narrower than 32 * MEGA_NJ float4 skip the rest (the phases are issue-bound:
                              // 16 warps x ~120 instructions per task on 4 schedulers, tools/probe/dot_probe.cu)
    const int k = lane + 32 * j;
    if (k < K4) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 x = *reinterpret_cast<const float4*>(xs + b * MEGA_MAXK + 4 * k);
        float&       t = (j & 1) ? odd[b] : acc[b];
        t = fmaf(w[j].x, x.x, fmaf(w[j].y, x.y, fmaf(w[j].z, x.z, fmaf(w[j].w, x.w, t))));
      }
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] += odd[b];
  const float4* p = reinterpret_cast<const float4*>(row);
  for (int k = lane + 32 * MEGA_NJ; k < K4; k += 32) {
    const float4 wv = __ldg(p + k);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float4 x = *reinterpret_cast<const float4*>(xs + b * MEGA_MAXK + 4 * k);
      acc[b] = fmaf(wv.x, x.x, fmaf(wv.y, x.y, fmaf(wv.z, x.z, fmaf(wv.w, x.w, acc[b]))));
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = warp_sum(acc[b]);
}

// two rows at once (Q/K/V rotation pairs, gate | up): the two dependency chains (LDS -> 4 FMA per chunk, then the 5-step shuffle tree)
// are interleaved, so the warp - alone on its scheduler in these phases - hides one chain's latency behind the other
template <int NB>
__device__ __forceinline__ void dot_row2(const float* __restrict__ row0, const float* __restrict__ row1, int K4, const float4* w0, const float4* w1,
                                         const float* xs, float* acc0, float* acc1) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int b = 0; b < NB; ++b) acc0[b] = acc1[b] = 0.f;
#pragma unroll
  for (int j = 0; j < MEGA_NJ; ++j) {
    if (32 * j >= K4) break;  // warp-uniform
    const int k = lane + 32 * j;
    if (k < K4) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 x = *reinterpret_cast<const float4*>(xs + b * MEGA_MAXK + 4 * k);
        acc0[b] = fmaf(w0[j].x, x.x, fmaf(w0[j].y, x.y, fmaf(w0[j].z, x.z, fmaf(w0[j].w, x.w, acc0[b]))));
        acc1[b] = fmaf(w1[j].x, x.x, fmaf(w1[j].y, x.y, fmaf(w1[j].z, x.z, fmaf(w1[j].w, x.w, acc1[b]))));
      }
    }
  }
  const float4 *p0 = reinterpret_cast<const float4*>(row0), *p1 = reinterpret_cast<const float4*>(row1);
  for (int k = lane + 32 * MEGA_NJ; k < K4; k += 32) {
    const float4 a0 = __ldg(p0 + k), a1 = __ldg(p1 + k);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float4 x = *reinterpret_cast<const float4*>(xs + b * MEGA_MAXK + 4 * k);
      acc0[b] = fmaf(a0.x, x.x, fmaf(a0.y, x.y, fmaf(a0.z, x.z, fmaf(a0.w, x.w, acc0[b]))));
      acc1[b] = fmaf(a1.x, x.x, fmaf(a1.y, x.y, fmaf(a1.z, x.z, fmaf(a1.w, x.w, acc1[b]))));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      acc0[b] += __shfl_xor_sync(0xffffffffu, acc0[b], o);
      acc1[b] += __shfl_xor_sync(0xffffffffu, acc1[b], o);
    }
  }
}

// rstd[b] = 1 / sqrt(mean_k x[b]^2 + eps) from per-thread partial sums of squares (reference norm.py:245-248)
template <int NB>
__device__ __forceinline__ void block_rstd(float* ss, int K, float eps, float* red, float* rstd) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const float s = warp_sum(ss[b]);
    if (lane == 0) red[b * MEGA_WARPS + wid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < MEGA_WARPS; ++i) tot += red[b * MEGA_WARPS + i];
    rstd[b] = 1.0f / sqrtf(tot / (float)K + eps);
  }
  __syncthreads();  // red is reused by the next reduction
}

// this thread's elements (k = tid, tid + 512) of a norm weight vector: requested before the rows are polled
__device__ __forceinline__ void pf_norm(const float* __restrict__ w, int K, float* nw) {
  nw[0] = threadIdx.x < K ? __ldg(w + threadIdx.x) : 0.f;
  nw[1] = threadIdx.x + MEGA_THREADS < K ? __ldg(w + threadIdx.x + MEGA_THREADS) : 0.f;
}

// xs[b][k] = x[b][k] * rstd[b] * w[k]; the raw rows are kept in ``keep`` (the residual of the attention block). Rows come from the
// embedding table (plain loads) or from the previous layer's exchange words. K <= 2 * MEGA_THREADS.
template <int NB>
__device__ __forceinline__ void load_norm_rows(const float* const* emb_rows, const unsigned long long* ll_rows, unsigned ep, int B, const float* nw, float eps,
                                               int K, float* xs, float* keep, float* red) {
  float ss[NB], v[NB][2];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int bb = b < B ? b : B - 1;
    const int cnt = threadIdx.x < K ? (threadIdx.x + MEGA_THREADS < K ? 2 : 1) : 0;
    if (emb_rows) {
      v[b][0] = cnt > 0 ? __ldg(emb_rows[b] + threadIdx.x) : 0.f;
      v[b][1] = cnt > 1 ? __ldg(emb_rows[b] + threadIdx.x + MEGA_THREADS) : 0.f;
    } else {
      ll_load_n<2>(ll_rows + (size_t)bb * K + threadIdx.x, MEGA_THREADS, cnt, ep, v[b]);
    }
    ss[b] = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j < cnt) {
        if (keep) keep[b * MEGA_MAXK + threadIdx.x + j * MEGA_THREADS] = v[b][j];
        ss[b] += v[b][j] * v[b][j];
      }
    }
  }
  float rstd[NB];
  block_rstd<NB>(ss, K, eps, red, rstd);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = threadIdx.x + j * MEGA_THREADS;
    if (k < K) {
#pragma unroll
      for (int b = 0; b < NB; ++b) xs[b * MEGA_MAXK + k] = v[b][j] * rstd[b] * nw[j];
    }
  }
  __syncthreads();
}

template <int NB, int HD4>
__global__ void __launch_bounds__(MEGA_THREADS, 1) k_decode_mega(MegaArgs a) {
  extern __shared__ __align__(16) float dsm[];
  float* xs = dsm;                  // [NB][MAXK] activations of the running phase
  float* hn = dsm + NB * MEGA_MAXK;  // [NB][MAXK] residual rows: layer input (P1 -> P3), after the attention block (P3 -> P4)
  float* lmw = hn + NB * MEGA_MAXK;  // [lm_rows_smem][dim] this CTA's first lm_head rows
  __shared__ float               red[NB * MEGA_WARPS];
  __shared__ int                 redi[NB * MEGA_WARPS];
  __shared__ float               sc[MEGA_SC];           // P2: scores, then softmax numerators of this unit's keys
  __shared__ __align__(16) float pvs[2048];             // P2: per key-group partial P.V
  __shared__ __align__(16) float qs[64], accs[64], knew[64], vnew[64];  // P2: query row, unnormalised output, K / V row of this position
  __shared__ float               ml_m[MEGA_MAXUNITS], ml_l[MEGA_MAXUNITS], fac[MEGA_MAXUNITS];
  __shared__ float               lmb[MEGA_LMROWS];  // lm_head bias of this CTA's rows
  constexpr int HD = HD4 * 4;
  constexpr int WR = HD4 > 2 * MEGA_NJ ? HD4 : 2 * MEGA_NJ;
  const int     tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int     G = gridDim.x, cta = blockIdx.x;
  // The warp's task index in every phase is BLOCKED (CTA c owns tasks 16c .. 16c+15): a phase with T tasks runs on ceil(T/16) CTAs,
  // so only those CTAs poll its inputs (every polling CTA re-reads the whole activation row; measured on the real step: 48 polling
  // CTAs make P4 twice as fast as 148). The attention units of P2 live on the CTAs after them; the rest of the grid only takes part
  // in the vocabulary projection.
  const int     gw = cta * MEGA_WARPS + wid;
  const int     dim = a.dim, FF = a.FF, H = a.H, B = a.B, pos = a.pos;
  const int     Lk = pos + 1, ns = a.nsplit, units = B * H * ns;
  const int     dim4 = dim / 4, FF4 = FF / 4;
  const unsigned ep = a.epoch;
  const float   scale = rsqrtf((float)HD);
  float4        wreg[WR];  // weight chunks of the NEXT phase's task, requested before that phase's inputs are polled
  int           tslot = 0;
  mega_stamp(a, tslot);
  // exchange words of one layer
  const size_t o_q = 0, o_kn = o_q + (size_t)8 * dim, o_vn = o_kn + (size_t)8 * dim, o_part = o_vn + (size_t)8 * dim;
  const size_t o_hid = o_part + (size_t)MEGA_MAXUNITS * (dim + 2), o_hout = o_hid + (size_t)8 * FF, ll_layer = o_hout + (size_t)8 * dim;

  // ---- the CTA's lm_head rows start moving into shared memory now; they are needed ~4 phases x n_layers later ----------------------
  const int lm_r0 = cta * a.lm_rows_cta;
  const int lm_n = a.logits ? max(0, min(a.lm_rows_cta, a.V - lm_r0)) : 0;
  const int lm_ns = min(lm_n, a.lm_rows_smem);
  {
    const float4* src = reinterpret_cast<const float4*>(a.wlm_t + (size_t)lm_r0 * dim);
    const unsigned dst0 = (unsigned)__cvta_generic_to_shared(lmw);
    for (int i = tid; i < lm_ns * dim4; i += MEGA_THREADS)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 16u * i), "l"(src + i) : "memory");
    if (a.lm_bias) {  // the bias block as well (4-byte copies: the block start is not 16-byte aligned in general)
      const unsigned b0 = (unsigned)__cvta_generic_to_shared(lmb);
      for (int i = tid; i < lm_n; i += MEGA_THREADS)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(b0 + 4u * i), "l"(a.lm_bias + lm_r0 + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const int npairs = 3 * dim / 2;
  const int  CH = (max(npairs, max(FF, dim)) + MEGA_WARPS - 1) / MEGA_WARPS;  // CTAs that carry the row phases P1 / P3 / P4
  const int  u_id = (cta - CH % G + G) % G;                                   // P2 unit of this CTA (if < units)
  const bool rowcta = cta < CH, unitcta = u_id < units;
  float nw[2];  // this thread's norm weights of the next normalisation
  pf_norm(a.layers[0].n1, dim, nw);
  float rope_c = 1.f, rope_s = 0.f;  // rotation of this warp's P1 column pair (the same in every layer)
  if (gw < npairs) {
    pf_row(a.layers[0].wqkv_t + (size_t)(2 * gw) * dim, dim4, wreg);
    pf_row(a.layers[0].wqkv_t + (size_t)(2 * gw + 1) * dim, dim4, wreg + MEGA_NJ);
    const int c = (2 * gw) % dim, d = c % HD;
    rope_c = __ldg(a.cosT + (size_t)pos * (HD / 2) + d / 2);
    rope_s = __ldg(a.sinT + (size_t)pos * (HD / 2) + d / 2);
  }

  const float* rows[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) rows[b] = a.emb + (size_t)a.ids[(size_t)(b < B ? b : B - 1) * a.ids_stride] * dim;  // reference model.py:194
  for (int l = 0; l < a.n_layers; ++l) {
    const MegaLayer&    Lw = a.layers[l];
    unsigned long long* X = a.ll + (size_t)l * ll_layer;  // this layer's exchange words
    // ================= P1: RMSNorm -> Q/K/V columns (pairs) -> RoPE -> q / new K, V row (exchange) + KV cache ==========================
    const bool f0 = cta == 0, fu = unitcta && u_id == 0;
    if (f0) mega_fine(a, l, false, 0);
    if (rowcta) load_norm_rows<NB>(l == 0 ? rows : nullptr, l == 0 ? nullptr : (X - ll_layer) + o_hout, ep, B, nw, Lw.eps1, dim, xs, hn, red);
    if (f0) mega_fine(a, l, false, 1);
    pf_norm(Lw.n2, dim, nw);  // for P3
    if (gw < npairs) {
      float y0[NB], y1[NB];
      dot_row2<NB>(Lw.wqkv_t + (size_t)(2 * gw) * dim, Lw.wqkv_t + (size_t)(2 * gw + 1) * dim, dim4, wreg, wreg + MEGA_NJ, xs, y0, y1);
      if (lane == 0) {
        const int   col = 2 * gw, which = col / dim, c = col - which * dim;  // 0: q, 1: k, 2: v
        const int   hh = c / HD, d = c - hh * HD;
        const float cs = rope_c, sn = rope_s;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          if (b < B) {
            float o0 = y0[b], o1 = y1[b];
            if (which < 2) {  // interleaved-pair rotation (reference model.py:23-44)
              o0 = y0[b] * cs - y1[b] * sn;
              o1 = y0[b] * sn + y1[b] * cs;
            }
            unsigned long long* x = X + (which == 0 ? o_q : (which == 1 ? o_kn : o_vn)) + (size_t)b * dim + c;
            ll_store(x, o0, ep);
            ll_store(x + 1, o1, ep);
            if (which) *reinterpret_cast<float2*>((which == 1 ? Lw.ck : Lw.cv) + (((size_t)b * a.S + pos) * H + hh) * HD + d) = make_float2(o0, o1);
          }
        }
      }
    }
    mega_stamp(a, tslot);
    // ================= P2: one CTA per (sequence, head, key split): attention partial + its unnormalised O projection ================
    if (unitcta) {
      const int u_sp = u_id % ns, u_hh = (u_id / ns) % H, u_b = u_id / (ns * H);
      if (tid < dim) {  // the head slice of Woᵀ (thread n keeps row n: HD floats)
        const float4* p = reinterpret_cast<const float4*>(Lw.wo_h + ((size_t)u_hh * dim + tid) * HD);
#pragma unroll
        for (int i = 0; i < HD4; ++i) wreg[i] = __ldg(p + i);
      }
      const int     chunk = (Lk + ns - 1) / ns, k0 = u_sp * chunk;
      const int     nk = max(0, min(Lk, k0 + chunk) - k0);
      const int     knew_raw = pos - k0;
      const int     knew_i = (knew_raw >= 0 && knew_raw < nk) ? knew_raw : -1;  // the current position among this unit's keys, or -1
      const size_t  base = (((size_t)u_b * a.S + k0) * H + u_hh) * HD;  // key k0 of this head; consecutive keys are H*HD apart
      const size_t  kstr4 = (size_t)H * HD4;
      const float4* kc = reinterpret_cast<const float4*>(Lw.ck + base);
      const float4* vc = reinterpret_cast<const float4*>(Lw.cv + base);
      // V rows of the first keys of this thread's key group are requested before the scores exist
      constexpr int NKG = MEGA_THREADS / HD4;  // key groups; thread -> (key group, float4 column of the head)
      const int     kg = tid / HD4, d4 = tid - kg * HD4;
      float4        vpre[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = kg + i * NKG;
        vpre[i] = (kg < NKG && kk < nk && kk != knew_i) ? __ldg(vc + (size_t)kk * kstr4 + d4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // ... and so are the K rows of the first 128 keys (4 threads per key, HD/4 features each)
      constexpr int PER = HD4 / 4;
      const int     part = tid & 3;
      float4        kpre[PER];
      {
        const int kk = tid >> 2;
#pragma unroll
        for (int i = 0; i < PER; ++i)
          kpre[i] = (kk < nk && kk != knew_i) ? __ldg(kc + (size_t)kk * kstr4 + part * PER + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      {
        const size_t hoff = (size_t)u_b * dim + u_hh * HD;
        if (tid < HD) qs[tid] = ll_load(X + o_q + hoff + tid, ep);
        else if (knew_i >= 0 && tid < 2 * HD) knew[tid - HD] = ll_load(X + o_kn + hoff + tid - HD, ep);
        else if (knew_i >= 0 && tid < 3 * HD) vnew[tid - 2 * HD] = ll_load(X + o_vn + hoff + tid - 2 * HD, ep);
      }
      __syncthreads();
      if (fu) mega_fine(a, l, true, 1);  // q (and the new K / V row) arrived
      float         m_loc = -INFINITY;
      for (int kb = 0; kb < nk; kb += MEGA_THREADS / 4) {
        const int kk = kb + (tid >> 2);
        float     s = 0.f;
        if (kk < nk) {
          const float4* kr = kk == knew_i ? reinterpret_cast<const float4*>(knew) + part * PER : kc + (size_t)kk * kstr4 + part * PER;
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            const float4 kv = (kb == 0 && kk != knew_i) ? kpre[i] : kr[i];  // cache rows of earlier positions: written by earlier launches
            const float4 qv = reinterpret_cast<const float4*>(qs)[part * PER + i];
            s = fmaf(kv.x, qv.x, fmaf(kv.y, qv.y, fmaf(kv.z, qv.z, fmaf(kv.w, qv.w, s))));
          }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s *= scale;
        if (kk < nk) {
          if (part == 0) sc[kk] = s;
          m_loc = fmaxf(m_loc, s);
        }
      }
      m_loc = warp_max(m_loc);
      if (lane == 0) red[wid] = m_loc;
      __syncthreads();
      float M = -INFINITY;
#pragma unroll
      for (int i = 0; i < MEGA_WARPS; ++i) M = fmaxf(M, red[i]);
      __syncthreads();
      if (fu) mega_fine(a, l, true, 2);  // scores + block max
      float l_loc = 0.f;
      for (int kk = tid; kk < nk; kk += MEGA_THREADS) {
        const float p = __expf(sc[kk] - M);
        sc[kk] = p;
        l_loc += p;
      }
      l_loc = warp_sum(l_loc);
      if (lane == 0) red[wid] = l_loc;
      __syncthreads();
      float lsum = 0.f;
#pragma unroll
      for (int i = 0; i < MEGA_WARPS; ++i) lsum += red[i];
      if (fu) mega_fine(a, l, true, 3);  // softmax numerators + sum
      // P.V: thread (key group, float4 column) walks its keys
      if (kg < NKG) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kk = kg + i * NKG;
          if (kk < nk) {
            const float  p = sc[kk];
            const float4 v = kk == knew_i ? reinterpret_cast<const float4*>(vnew)[d4] : vpre[i];
            acc.x = fmaf(p, v.x, acc.x), acc.y = fmaf(p, v.y, acc.y), acc.z = fmaf(p, v.z, acc.z), acc.w = fmaf(p, v.w, acc.w);
          }
        }
        for (int k4 = kg + 4 * NKG; k4 < nk; k4 += 4 * NKG) {  // further keys: four rows in flight at a time
          float4 vv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kk = k4 + i * NKG;
            vv[i] = (kk < nk && kk != knew_i) ? __ldg(vc + (size_t)kk * kstr4 + d4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kk = k4 + i * NKG;
            if (kk < nk) {
              const float  p = sc[kk];
              const float4 v = kk == knew_i ? reinterpret_cast<const float4*>(vnew)[d4] : vv[i];
              acc.x = fmaf(p, v.x, acc.x), acc.y = fmaf(p, v.y, acc.y), acc.z = fmaf(p, v.z, acc.z), acc.w = fmaf(p, v.w, acc.w);
            }
          }
        }
        reinterpret_cast<float4*>(pvs)[kg * HD4 + d4] = acc;
      }
      __syncthreads();
      if (tid < HD * 8) {  // 8 lanes per feature: each sums every 8th key group, then a 3-step shuffle tree (fixed order)
        const int d = tid >> 3, j = tid & 7;
        float     s = 0.f;
        for (int g = j; g < NKG; g += 8) s += pvs[g * HD + d];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (j == 0) accs[d] = s;
      }
      __syncthreads();
      if (fu) mega_fine(a, l, true, 4);  // P.V reduced
      unsigned long long* dst = X + o_part + (size_t)u_id * (dim + 2);
      // a single split per head needs no merge: its partial is normalised here and P3 just adds the heads (no (max, sum) exchange,
      // no softmax-weight stage on P3's critical path: 1.35 us per layer in the timeline of the multi-split version)
      const float onorm = ns == 1 ? 1.f / lsum : 1.f;
      if (tid < dim) {
        float o = 0.f;
#pragma unroll
        for (int i = 0; i < HD4; ++i) {
          const float4 x = reinterpret_cast<const float4*>(accs)[i];
          o = fmaf(wreg[i].x, x.x, fmaf(wreg[i].y, x.y, fmaf(wreg[i].z, x.z, fmaf(wreg[i].w, x.w, o))));
        }
        ll_store(dst + tid, o * onorm, ep);
      }
      for (int n = tid + MEGA_THREADS; n < dim; n += MEGA_THREADS) {  // models wider than the CTA: rows loaded here
        const float4* p = reinterpret_cast<const float4*>(Lw.wo_h + ((size_t)u_hh * dim + n) * HD);
        float         o = 0.f;
#pragma unroll
        for (int i = 0; i < HD4; ++i) {
          const float4 wv = __ldg(p + i), x = reinterpret_cast<const float4*>(accs)[i];
          o = fmaf(wv.x, x.x, fmaf(wv.y, x.y, fmaf(wv.z, x.z, fmaf(wv.w, x.w, o))));
        }
        ll_store(dst + n, o * onorm, ep);
      }
      if (tid == 0) {
        ll_store(dst + dim, M, ep);  // -inf for an empty split
        ll_store(dst + dim + 1, lsum, ep);
      }
      if (fu) mega_fine(a, l, true, 5);  // partial O projection stored
    }
    // next: gate / up rows of hidden unit gw
    if (gw < FF) {
      pf_row(Lw.wgu_t + (size_t)(2 * gw) * dim, dim4, wreg);
      pf_row(Lw.wgu_t + (size_t)(2 * gw + 1) * dim, dim4, wreg + MEGA_NJ);
    }
    mega_stamp(a, tslot);
    // ================= P3: merge the split partials + residual -> hn, RMSNorm -> gate / up -> SwiGLU -> hid ===========================
    if (rowcta) {
      const unsigned long long* P = X + o_part;
      const int                 hs = H * ns;
      // the partial rows this thread combines are requested first (they do not depend on the softmax weights) ...
      if (f0) mega_fine(a, l, false, 3);
      float pv[MEGA_PV];  // (first sequence; further sequences are polled in the combine loop)
      if (tid < dim) ll_load_n<MEGA_PV>(P + tid, (size_t)(dim + 2), min(MEGA_PV, hs), ep, pv);
      if (f0) mega_fine(a, l, false, 4);  // first batch of partial rows arrived
      // ... while the (max, sum) pairs of the units are fetched and turned into e^{m-M} / den (several splits per head only)
      if (ns == 1) {
        if (tid < units) fac[tid] = 1.f;
      } else if (tid < units) {
        float ml[2];
        ll_load_n<2>(P + (size_t)tid * (dim + 2) + dim, 1, 2, ep, ml);
        ml_m[tid] = ml[0];
        ml_l[tid] = ml[1];
      }
      if (ns > 1) __syncthreads();
      if (ns > 1 && tid < units) {
        const int bh = tid / ns;
        float     M = -INFINITY;
        for (int s = 0; s < ns; ++s) M = fmaxf(M, ml_m[bh * ns + s]);
        float den = 0.f;
        for (int s = 0; s < ns; ++s) {
          const float ms = ml_m[bh * ns + s];
          den = fmaf(ms == -INFINITY ? 0.f : __expf(ms - M), ml_l[bh * ns + s], den);
        }
        const float mt = ml_m[tid];
        fac[tid] = (mt == -INFINITY) ? 0.f : __expf(mt - M) / den;
      }
      __syncthreads();
      if (f0) mega_fine(a, l, false, 5);  // softmax weights of the units ready
      float ss[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        ss[b] = 0.f;
        if (b < B) {
          for (int n = tid; n < dim; n += MEGA_THREADS) {
            const unsigned long long* p = P + (size_t)(b * hs) * (dim + 2) + n;
            float                     v = 0.f;
            const bool first = (b == 0 && n == tid);
            if (first) {
#pragma unroll
              for (int i = 0; i < MEGA_PV; ++i)
                if (i < hs) v = fmaf(fac[i], pv[i], v);
            }
            for (int i0 = (first ? MEGA_PV : 0); i0 < hs; i0 += 8) {  // more units than the first batch holds / rows wider than the CTA
              float q8[8];
              ll_load_n<8>(p + (size_t)i0 * (dim + 2), (size_t)(dim + 2), min(8, hs - i0), ep, q8);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i0 + i < hs) v = fmaf(fac[b * hs + i0 + i], q8[i], v);
            }
            const float hv = hn[b * MEGA_MAXK + n] + v;
            hn[b * MEGA_MAXK + n] = hv;
            ss[b] += hv * hv;
          }
        }
      }
      if (f0) mega_fine(a, l, false, 6);  // partials combined
      float rstd[NB];
      block_rstd<NB>(ss, dim, Lw.eps2, red, rstd);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = tid + j * MEGA_THREADS;
        if (k < dim) {
#pragma unroll
          for (int b = 0; b < NB; ++b) xs[b * MEGA_MAXK + k] = (b < B ? hn[b * MEGA_MAXK + k] : 0.f) * rstd[b] * nw[j];
        }
      }
      __syncthreads();
      if (f0) mega_fine(a, l, false, 7);  // normalised
    }
    pf_norm(l + 1 < a.n_layers ? a.layers[l + 1].n1 : a.norm_w, dim, nw);  // for the next P1 / the final normalisation
    if (gw < FF) {
      float g[NB], u[NB];
      dot_row2<NB>(Lw.wgu_t + (size_t)(2 * gw) * dim, Lw.wgu_t + (size_t)(2 * gw + 1) * dim, dim4, wreg, wreg + MEGA_NJ, xs, g, u);
      if (f0) mega_fine(a, l, false, 10);  // gate / up dot products done
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (b < B) ll_store(X + o_hid + (size_t)b * FF + gw, __fdividef(g[b], 1.f + __expf(-g[b])) * u[b], ep);  // x / (1 + exp(-x)), functional.py:39-40
      }
    }
    if (f0) mega_fine(a, l, false, 8);  // gate / up / SwiGLU stored
    if (gw < dim) pf_row(Lw.wd_t + (size_t)gw * FF, FF4, wreg);
    mega_stamp(a, tslot);
    // ================= P4: down projection + residual -> next layer's input ==============================================================
    if (cta * MEGA_WARPS < dim) {
      __syncthreads();  // every warp is done with xs
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float                     hv[2];
        const unsigned long long* src = X + o_hid + (size_t)(b < B ? b : B - 1) * FF + tid;
        const int                 cnt = tid < FF ? (tid + MEGA_THREADS < FF ? 2 : 1) : 0;
        ll_load_n<2>(src, MEGA_THREADS, cnt, ep, hv);
        if (cnt > 0) xs[b * MEGA_MAXK + tid] = hv[0];
        if (cnt > 1) xs[b * MEGA_MAXK + tid + MEGA_THREADS] = hv[1];
      }
      __syncthreads();
      if (f0) mega_fine(a, l, false, 9);  // hidden row arrived
      if (gw < dim) {
        float y[NB];
        dot_row<NB>(Lw.wd_t + (size_t)gw * FF, FF4, wreg, xs, y);
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < B) ll_store(X + o_hout + (size_t)b * dim + gw, hn[b * MEGA_MAXK + gw] + y[b], ep);
        }
      }
    }
    if (l + 1 < a.n_layers && gw < npairs) {
      pf_row(a.layers[l + 1].wqkv_t + (size_t)(2 * gw) * dim, dim4, wreg);
      pf_row(a.layers[l + 1].wqkv_t + (size_t)(2 * gw + 1) * dim, dim4, wreg + MEGA_NJ);
    }
    mega_stamp(a, tslot);
    __syncthreads();  // xs / hn are rewritten by the next phase
  }
  if (a.logits == nullptr) return;  // prompt positions before the last one: only the KV cache matters (model.py:255 keeps [-1])
  // ================= final RMSNorm -> lm_head rows (+bias) -> logits, argmax partials ====================================================
  // the rows of this CTA's block that did not fit in shared memory: the first three of this warp are requested now (row widths up to
  // 96 float4), so that the vocabulary projection never waits for a weight
  const bool lm_pre = dim4 <= 96;
  const int  lm_first = lm_ns + ((wid - lm_ns % MEGA_WARPS) + MEGA_WARPS) % MEGA_WARPS;  // this warp's first row past the staged ones
  if (lm_pre) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int     r = lm_first + MEGA_WARPS * j;
      const float4* w4 = reinterpret_cast<const float4*>(a.wlm_t + (size_t)(lm_r0 + r) * dim);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int k = lane + 32 * i;
        wreg[j * 3 + i] = (r < lm_n && k < dim4) ? __ldg(w4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  load_norm_rows<NB>(nullptr, a.ll + (size_t)(a.n_layers - 1) * ll_layer + o_hout, ep, B, nw, a.eps_f, dim, xs, nullptr, red);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  mega_stamp(a, tslot);
  float best[NB];
  int   besti[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    best[b] = -INFINITY;
    besti[b] = 0x7fffffff;
  }
  for (int r = wid; r < lm_n; r += MEGA_WARPS) {
    const int n = lm_r0 + r;
    float     y[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) y[b] = 0.f;
    if (r < lm_ns) {
      const float4* w4 = reinterpret_cast<const float4*>(lmw + (size_t)r * dim);
      for (int k = lane; k < dim4; k += 32) {
        const float4 w = w4[k];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float4 x = *reinterpret_cast<const float4*>(xs + b * MEGA_MAXK + 4 * k);
          y[b] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, y[b]))));
        }
      }
    } else if (lm_pre && r < lm_first + 3 * MEGA_WARPS) {
      const int j = (r - lm_first) / MEGA_WARPS;
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) {
        if (jj == j) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int k = lane + 32 * i;
            if (k < dim4) {
              const float4 w = wreg[jj * 3 + i];
#pragma unroll
              for (int b = 0; b < NB; ++b) {
                const float4 x = *reinterpret_cast<const float4*>(xs + b * MEGA_MAXK + 4 * k);
                y[b] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, y[b]))));
              }
            }
          }
        }
      }
    } else {
      const float4* w4 = reinterpret_cast<const float4*>(a.wlm_t + (size_t)n * dim);
      for (int k = lane; k < dim4; k += 32) {
        const float4 w = __ldg(w4 + k);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float4 x = *reinterpret_cast<const float4*>(xs + b * MEGA_MAXK + 4 * k);
          y[b] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, y[b]))));
        }
      }
    }
    const float bias = a.lm_bias ? lmb[r] : 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float v = warp_sum(y[b]) + bias;
      if (lane == 0 && b < B) a.logits[(size_t)b * a.V + n] = v;
      if (v > best[b] || (v == best[b] && n < besti[b])) {  // rows arrive in increasing n per warp: ties keep the first
        best[b] = v;
        besti[b] = n;
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
  mega_stamp(a, tslot);
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
  void*      scratch = nullptr;
  const void* fn = nullptr;
  int        nb = 1, hd4 = 0, grid = 0, device = 0;
  size_t     dyn_smem = 0;
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
  PDN_CHECK(S <= MEGA_SC, "decoder_create: max_seq_len %d > %d", S, MEGA_SC);
  PDN_CHECK((V + sm_count() - 1) / sm_count() <= MEGA_LMROWS, "decoder_create: vocabulary %d too large for %d rows per SM", V, MEGA_LMROWS);
  MegaHandle* h = new MegaHandle();
  h->nb = B <= 1 ? 1 : (B <= 2 ? 2 : (B <= 4 ? 4 : 8));
  h->hd4 = hd / 4;
  h->grid = sm_count();
  cudaGetDevice(&h->device);
  h->fn = pick_kernel(h->nb, h->hd4);
  // every warp owns at most one task per phase (that is what lets it request the task's weights before its inputs have arrived)
  const int warps = h->grid * MEGA_WARPS;
  if (warps < 3 * dim / 2 || warps < FF || B * H > h->grid || B * H > MEGA_MAXUNITS) {
    delete h;
    set_error("decoder_create: model too wide for one task per warp on %d SMs", sm_count());
    return PDN_ERR_INVALID;
  }
  cudaFuncAttributes fa;
  int                max_optin = 0;
  if (cudaFuncGetAttributes(&fa, h->fn) != cudaSuccess ||
      cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device) != cudaSuccess) {
    delete h;
    set_error("decoder_create: cannot query the decode kernel");
    return PDN_ERR_CUDA;
  }
  const size_t fixed = (size_t)2 * h->nb * MEGA_MAXK * 4;
  const long   avail = (long)max_optin - (long)fa.sharedSizeBytes - (long)fixed - 1024;
  const int    rows_cta = (V + h->grid - 1) / h->grid;
  int          rows_smem = avail > 0 ? (int)(avail / ((long)dim * 4)) : 0;
  if (rows_smem > rows_cta) rows_smem = rows_cta;
  h->dyn_smem = fixed + (size_t)rows_smem * dim * 4;
  int per_sm = 0;
  if (cudaFuncSetAttribute(h->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->dyn_smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, h->fn, MEGA_THREADS, h->dyn_smem) != cudaSuccess || per_sm < 1) {
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
  // scratch: layer table | exchange words (per layer: q, knew, vnew [8][dim]; partials [MAXUNITS][dim+2]; hid [8][FF]; hout [8][dim])
  //          | argmax partials | ticket
  const size_t o_layers = 0, n_layers_b = sizeof(MegaLayer) * n_layers;
  auto   up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t ll_layer = (size_t)8 * dim * 3 + (size_t)MEGA_MAXUNITS * (dim + 2) + (size_t)8 * FF + (size_t)8 * dim;
  size_t o_ll = up(o_layers + n_layers_b), o_av = up(o_ll + ll_layer * n_layers * 8), o_ai = up(o_av + (size_t)h->grid * 8 * 4);
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
  a.ll = (unsigned long long*)(base + o_ll);  // zero-filled: epoch 0 is never a launch number
  a.amax_val = (float*)(base + o_av), a.amax_idx = (int*)(base + o_ai);
  a.ticket = (unsigned int*)(base + o_bar + 64);
  a.lm_rows_cta = rows_cta, a.lm_rows_smem = rows_smem;
  h->bars_per_launch = 4 * n_layers;
  a.trace = nullptr;
  if (getenv("PDN_MEGA_TRACE")) {
    void* t = nullptr;
    if (dev_alloc(&t, 1024 * 8) == 0) a.trace = (unsigned long long*)t;
  }
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
  // key splits per (sequence, head): an attention unit streams keys*2*hd*4 bytes of cache, every CTA then reads one partial row
  // (dim + 2 floats) per unit in P3 — the split count balances the two
  const double keys = (double)pos + 1.0, bh = (double)a.B * a.H;
  int          ns = (int)std::lround(std::sqrt(keys * 2.0 * a.hd / (bh * (a.dim + 2))));
  const int    cap = h->grid / (a.B * a.H) < MEGA_MAXUNITS / (a.B * a.H) ? h->grid / (a.B * a.H) : MEGA_MAXUNITS / (a.B * a.H);
  if (keys <= 320.0) ns = 1;  // short contexts: one unit per head reads all keys (<= 120 KB of cache) and P3 needs no merge stage
  ns = ns < 1 ? 1 : (ns > cap ? cap : ns);
  while (((int)keys + ns - 1) / ns > MEGA_SC) ++ns;  // unreachable for S <= MEGA_SC, kept as a guard
  a.nsplit = ns;
  a.epoch = (unsigned int)((h->launches % 0xfffffffeull) + 1);
  void* params[] = {&a};
  // Cooperative launch: the runtime guarantees that all CTAs are co-resident, which is what makes the polling between them safe.
  // (PDN_MEGA_COOP=0 launches the same grid as an ordinary kernel - one CTA per SM by its shared-memory footprint, so co-resident
  // on an otherwise idle device - to measure what the cooperative launch path itself costs per token.)
  static const bool coop = !(getenv("PDN_MEGA_COOP") && getenv("PDN_MEGA_COOP")[0] == '0');
  if (coop) PDN_CUDA(cudaLaunchCooperativeKernel(h->fn, dim3(h->grid), dim3(MEGA_THREADS), params, h->dyn_smem, stream()));
  else PDN_CUDA(cudaLaunchKernel(h->fn, dim3(h->grid), dim3(MEGA_THREADS), params, h->dyn_smem, stream()));
  PDN_LAUNCHED("decode_mega");
  h->launches++;  // only a launch that was accepted advances the epoch
  if (a.trace && logits && (h->launches % 64) == 40) {  // debug timeline of one step (CTA 0 and the last CTA), in ns from kernel entry
    unsigned long long t[1024];
    cudaStreamSynchronize(stream());
    cudaMemcpy(t, a.trace, sizeof(t), cudaMemcpyDeviceToHost);
    const int n = 3 + h->bars_per_launch;  // entry | P1 P2 P3 P4 per layer | final norm | lm_head
    fprintf(stderr, "[mega trace] pos %d nsplit %d\n cta0:", (int)pos, a.nsplit);
    for (int i = 1; i < n; ++i) fprintf(stderr, "%s%llu", (i % 4) == 1 ? " | " : " ", t[i] - t[i - 1]);
    fprintf(stderr, "\n ctaL:");
    for (int i = 1; i < n; ++i) fprintf(stderr, "%s%llu", (i % 4) == 1 ? " | " : " ", t[512 + i] - t[512 + i - 1]);
    fprintf(stderr, "\n total cta0 %llu ns\n", t[n - 1] - t[0]);
    const char* rn[11] = {"P1 start", "h polled + normalised", "QKV stored", "P3 start", "partials arrived", "weights ready", "combined", "normalised",
                          "hidden stored", "P4 hidden arrived", "gate/up dots done"};
    fprintf(stderr, " layer 2, CTA 0 (ns after its P1 start):");
    for (int i = 0; i < 11; ++i) fprintf(stderr, " %s=%lld", rn[i], (long long)(t[256 + i] - t[256]));
    const char* un[6] = {"P2 start", "q arrived", "scores+max", "softmax", "P.V", "O-partial stored"};
    fprintf(stderr, "\n layer 2, first attention unit (ns after CTA 0's P1 start):");
    for (int i = 0; i < 6; ++i) fprintf(stderr, " %s=%lld", un[i], (long long)(t[288 + i] - t[256]));
    fprintf(stderr, "\n");
  }
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
