// fused_rows.cu — single-pass row kernels: softmax / log-softmax, RMSNorm, cross-entropy, SwiGLU, Adam.
// Each replaces a chain of 5-12 eager array expressions of the reference (softmax functional.py:43-58, RMSNorm
// norm.py:245-248, CE functional.py:364-381, silu*up llm/llama/model.py:56-58, Adam optimizer.py:185-196) with ONE
// HBM pass: 128-bit coalesced loads, warp-shuffle row reductions, values cached in registers between passes.
#include "common.cuh"
#include <math.h>

namespace pdn {

constexpr int ROW_CACHE = 8;  // float4 per lane cached in registers: rows up to 32*4*8 = 1024 floats stay on-chip

// ---------------------------------------------------------------- softmax ---------------------------------------
// one warp per row; n <= 1024 and n % 4 == 0 take the register-cached float4 path, everything else the strided loop.
template <bool LOG>
__global__ void __launch_bounds__(256) k_softmax_fwd(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * n;
  float*       yr = y + row * n;
  const bool vec = (n % 4 == 0) && n <= 32 * 4 * ROW_CACHE && ((((uintptr_t)xr) | ((uintptr_t)yr)) & 15) == 0;
  if (vec) {
    const int n4 = n >> 2;
    float4    c[ROW_CACHE];
    float     m = -INFINITY;
#pragma unroll
    for (int i = 0; i < ROW_CACHE; ++i) {
      int j = lane + i * 32;
      if (j < n4) {
        c[i] = __ldg(reinterpret_cast<const float4*>(xr) + j);
        m = fmaxf(m, fmaxf(fmaxf(c[i].x, c[i].y), fmaxf(c[i].z, c[i].w)));
      }
    }
    m = warp_max(m);
    if (m == -INFINITY) m = 0.f;  // fully masked row: exp(-inf - 0) = 0 everywhere, like NumPy's nan-free path is not needed
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_CACHE; ++i) {
      int j = lane + i * 32;
      if (j < n4) {
        c[i].x -= m; c[i].y -= m; c[i].z -= m; c[i].w -= m;
        float4 e = make_float4(__expf(c[i].x), __expf(c[i].y), __expf(c[i].z), __expf(c[i].w));
        s += (e.x + e.y) + (e.z + e.w);
        if (!LOG) c[i] = e;
      }
    }
    s = warp_sum(s);
    const float inv = 1.f / s, ls = logf(s);
#pragma unroll
    for (int i = 0; i < ROW_CACHE; ++i) {
      int j = lane + i * 32;
      if (j < n4) {
        float4 o = LOG ? make_float4(c[i].x - ls, c[i].y - ls, c[i].z - ls, c[i].w - ls)
                       : make_float4(c[i].x * inv, c[i].y * inv, c[i].z * inv, c[i].w * inv);
        reinterpret_cast<float4*>(yr)[j] = o;
      }
    }
    return;
  }
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, xr[j]);
  m = warp_max(m);
  if (m == -INFINITY) m = 0.f;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += __expf(xr[j] - m);
  s = warp_sum(s);
  const float inv = 1.f / s, ls = logf(s);
  for (int j = lane; j < n; j += 32) yr[j] = LOG ? (xr[j] - m - ls) : __expf(xr[j] - m) * inv;
}

// dx = y * (g - sum(g*y))            (softmax)
// dx = g - exp(y) * sum(g)           (log-softmax, y = log p)
template <bool LOG>
__global__ void __launch_bounds__(256) k_softmax_bwd(const float* __restrict__ y, const float* __restrict__ g, float* __restrict__ dx,
                                                     int64_t rows, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* yr = y + row * n;
  const float* gr = g + row * n;
  float*       dr = dx + row * n;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += LOG ? gr[j] : gr[j] * yr[j];
  s = warp_sum(s);
  for (int j = lane; j < n; j += 32) dr[j] = LOG ? (gr[j] - __expf(yr[j]) * s) : yr[j] * (gr[j] - s);
}

// generic-dtype fallback (fp64 / fp16 rows): block per row, three passes
template <typename T, bool LOG>
__global__ void __launch_bounds__(256) k_softmax_fwd_any(const T* x, T* y, int64_t rows, int64_t n) {
  using A = typename Acc<T>::type;
  __shared__ A red[32];
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = x + row * n;
    T*       yr = y + row * n;
    A m = -INFINITY;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) { A v = ld<T>(xr + j); m = v > m ? v : m; }
    m = block_max<A>(m, red, (A)-INFINITY);
    if (m == (A)-INFINITY) m = 0;
    A s = 0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) s += exp(ld<T>(xr + j) - m);
    s = block_sum<A>(s, red);
    A ls = log(s);
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      A v = ld<T>(xr + j) - m;
      st<T>(yr + j, LOG ? v - ls : exp(v) / s);
    }
  }
}
template <typename T, bool LOG>
__global__ void __launch_bounds__(256) k_softmax_bwd_any(const T* y, const T* g, T* dx, int64_t rows, int64_t n) {
  using A = typename Acc<T>::type;
  __shared__ A red[32];
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const T *yr = y + row * n, *gr = g + row * n;
    T* dr = dx + row * n;
    A s = 0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) s += LOG ? ld<T>(gr + j) : ld<T>(gr + j) * ld<T>(yr + j);
    s = block_sum<A>(s, red);
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      A yy = ld<T>(yr + j), gg = ld<T>(gr + j);
      st<T>(dr + j, LOG ? gg - exp(yy) * s : yy * (gg - s));
    }
  }
}

// ---------------------------------------------------------------- RMSNorm ---------------------------------------
__global__ void __launch_bounds__(256) k_rmsnorm_fwd(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                                     float* __restrict__ rstd, int64_t rows, int n, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * n;
  float*       yr = y + row * n;
  float ss = 0.f;
  for (int j = lane; j < n; j += 32) { float v = xr[j]; ss += v * v; }
  ss = warp_sum(ss);
  const float r = 1.f / sqrtf(ss / (float)n + eps);
  if (lane == 0 && rstd) rstd[row] = r;
  for (int j = lane; j < n; j += 32) yr[j] = xr[j] * r * __ldg(w + j);
}

// Same normalisation, but the result is emitted directly as the next GEMM's A operand: K-major bf16 hi/lo planes
// [2][rows][Kp] (x = hi + lo), skipping the fp32 round trip and the separate pack launch.
__device__ __forceinline__ void put_pair(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t idx, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__global__ void __launch_bounds__(256) k_rmsnorm_planes(const float* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ planes,
                                                        int64_t rows, int n, int64_t Kp, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * n;
  float ss = 0.f;
  for (int j = lane; j < n; j += 32) { float v = xr[j]; ss += v * v; }
  ss = warp_sum(ss);
  const float r = 1.f / sqrtf(ss / (float)n + eps);
  __nv_bfloat16 *hi = planes, *lo = planes + rows * Kp;
  for (int j = lane; j < Kp; j += 32) put_pair(hi, lo, row * Kp + j, j < n ? xr[j] * r * __ldg(w + j) : 0.f);
}
// four consecutive values -> 8 bytes of the hi plane + 8 bytes of the lo plane
__device__ __forceinline__ void put_quad(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t idx, float4 v) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
  *reinterpret_cast<uint2*>(hi + idx) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  *reinterpret_cast<uint2*>(lo + idx) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}
// n % 4 == 0, n <= 128 * NV, 16-byte aligned x / w: the row is read ONCE with 128-bit loads and stays in registers
template <int NV>
__global__ void __launch_bounds__(256) k_rmsnorm_planes_v4(const float* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ planes,
                                                           int64_t rows, int n, int64_t Kp, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * n);
  const int n4 = n >> 2;
  float4 v[NV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + i * 32;
    v[i] = j < n4 ? __ldg(xr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  ss = warp_sum(ss);
  const float r = 1.f / sqrtf(ss / (float)n + eps);
  __nv_bfloat16 *hi = planes, *lo = planes + rows * Kp;
  const int kp4 = (int)(Kp >> 2);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + i * 32;
    if (j < kp4) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < n4) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + j);
        o = make_float4(v[i].x * r * ww.x, v[i].y * r * ww.y, v[i].z * r * ww.z, v[i].w * r * ww.w);
      }
      put_quad(hi, lo, row * Kp + 4 * j, o);
    }
  }
}
__global__ void __launch_bounds__(256) k_swiglu_rows_planes(const float* __restrict__ gu, __nv_bfloat16* __restrict__ planes, int64_t rows,
                                                            int64_t F, int64_t Kp);

// dx = r*w*g - x * r^3 * mean(g*w*x) ; dw[j] += sum_rows g*x*r   (per-warp register partials, then atomics)
template <int MAXJ>
__global__ void __launch_bounds__(256) k_rmsnorm_bwd(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ rstd,
                                                     const float* __restrict__ g, float* __restrict__ dx, float* __restrict__ dw,
                                                     int64_t rows, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float acc[MAXJ];
#pragma unroll
  for (int i = 0; i < MAXJ; ++i) acc[i] = 0.f;
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const float *xr = x + row * n, *gr = g + row * n;
    float*      dr = dx + row * n;
    const float r = rstd[row];
    float dot = 0.f;
    for (int j = lane; j < n; j += 32) dot += gr[j] * __ldg(w + j) * xr[j];
    dot = warp_sum(dot);
    const float c = dot * r * r * r / (float)n;
#pragma unroll
    for (int i = 0; i < MAXJ; ++i) {
      int j = lane + i * 32;
      if (j < n) {
        float xv = xr[j], gv = gr[j];
        if (dx) dr[j] = r * __ldg(w + j) * gv - xv * c;
        acc[i] += gv * xv * r;
      }
    }
  }
  if (dw) {
#pragma unroll
    for (int i = 0; i < MAXJ; ++i) {
      int j = lane + i * 32;
      if (j < n) atomicAdd(dw + j, acc[i]);
    }
  }
}

// ---------------------------------------------------------------- cross entropy ---------------------------------
// loss = reduce_rows(lse(row) - logit[row, target]);  one warp per row; the scalar is accumulated with one atomic per warp.
__global__ void __launch_bounds__(256) k_ce_fwd(const float* __restrict__ logits, const long long* __restrict__ target, float* __restrict__ loss,
                                                float* __restrict__ lse, int64_t N, int C, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* xr = logits + row * C;
  float m = -INFINITY;
  for (int j = lane; j < C; j += 32) m = fmaxf(m, xr[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < C; j += 32) s += __expf(xr[j] - m);
  s = warp_sum(s);
  if (lane == 0) {
    float l = m + logf(s);
    lse[row] = l;
    long long t = target[row];
    if (t < 0) t += C;
    atomicAdd(loss, (l - xr[t]) * scale);
  }
}
__global__ void __launch_bounds__(256) k_ce_bwd(const float* __restrict__ logits, const long long* __restrict__ target,
                                                const float* __restrict__ lse, const float* __restrict__ gloss, float* __restrict__ dlogits,
                                                int64_t N, int C, float scale) {
  const int64_t total = N * (int64_t)C;
  const float gs = gloss[0] * scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / C;
    int     c = (int)(i - row * C);
    long long t = target[row];
    if (t < 0) t += C;
    float p = __expf(logits[i] - lse[row]);
    dlogits[i] = (p - (c == t ? 1.f : 0.f)) * gs;
  }
}

// ---------------------------------------------------------------- SwiGLU ----------------------------------------
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__global__ void __launch_bounds__(256) k_swiglu(const float4* gate, const float4* up, float4* out, int64_t n4,
                                                const float* gate_s, const float* up_s, float* out_s, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = __ldg(gate + i), b = __ldg(up + i);
    out[i] = make_float4(silu_f(a.x) * b.x, silu_f(a.y) * b.y, silu_f(a.z) * b.z, silu_f(a.w) * b.w);
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) out_s[i] = silu_f(gate_s[i]) * up_s[i];
}
// fused gate|up projection output [rows, 2F]: out[r, j] = silu(gu[r, j]) * gu[r, F + j]
__global__ void __launch_bounds__(256) k_swiglu_rows(const float* __restrict__ gu, float* __restrict__ out, int64_t rows, int64_t F) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * F; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F, j = i - r * F;
    out[i] = silu_f(gu[r * 2 * F + j]) * gu[r * 2 * F + F + j];
  }
}
__global__ void __launch_bounds__(256) k_swiglu_rows_planes(const float* __restrict__ gu, __nv_bfloat16* __restrict__ planes, int64_t rows,
                                                            int64_t F, int64_t Kp) {
  __nv_bfloat16 *hi = planes, *lo = planes + rows * Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * Kp; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Kp, j = i - r * Kp;
    put_pair(hi, lo, i, j < F ? silu_f(gu[r * 2 * F + j]) * gu[r * 2 * F + F + j] : 0.f);
  }
}
// F % 4 == 0, 16-byte aligned gu, rows * Kp / 4 < 2^31: one thread = four consecutive outputs of one row
__global__ void __launch_bounds__(256) k_swiglu_rows_planes_v4(const float* __restrict__ gu, __nv_bfloat16* __restrict__ planes, int rows, int F,
                                                               int Kp) {
  __nv_bfloat16 *hi = planes, *lo = planes + (int64_t)rows * Kp;
  const uint32_t kp4 = (uint32_t)Kp >> 2, f4 = (uint32_t)F >> 2, total = (uint32_t)rows * kp4;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / kp4, j = i - r * kp4;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < f4) {
      const float4* row = reinterpret_cast<const float4*>(gu + (int64_t)r * 2 * F);
      const float4 a = __ldg(row + j), b = __ldg(row + f4 + j);
      o = make_float4(silu_f(a.x) * b.x, silu_f(a.y) * b.y, silu_f(a.z) * b.z, silu_f(a.w) * b.w);
    }
    put_quad(hi, lo, (int64_t)r * Kp + 4 * j, o);
  }
}
// dgate = g * up * silu'(gate), dup = g * silu(gate); silu'(x) = s + x*s*(1-s), s = sigmoid(x)
__global__ void __launch_bounds__(256) k_swiglu_bwd(const float* __restrict__ gate, const float* __restrict__ up, const float* __restrict__ g,
                                                    float* __restrict__ dgate, float* __restrict__ dup, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float x = gate[i], u = up[i], gg = g[i];
    float s = 1.f / (1.f + __expf(-x));
    dgate[i] = gg * u * (s + x * s * (1.f - s));
    dup[i] = gg * x * s;
  }
}

// ---------------------------------------------------------------- Adam ------------------------------------------
__global__ void __launch_bounds__(256) k_adam(float4* p, const float4* grad, float4* m, float4* v,
                                              int64_t n4, float step_size, float b1, float b2, float eps, float wd, float gscale,
                                              float* ps, const float* gs, float* ms, float* vs, int64_t n, const float* step_dev) {
  if (step_dev) step_size = __ldg(step_dev);  // recorded steps (CUDA graph): the bias-corrected step size lives in device memory
  auto upd = [&](float& pp, float g, float& mm, float& vv) {
    g = g * gscale + wd * pp;
    mm = b1 * mm + (1.f - b1) * g;
    vv = b2 * vv + (1.f - b2) * g * g;
    pp -= step_size * mm / (sqrtf(vv) + eps);
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 P = p[i], G = __ldg(grad + i), M = m[i], V = v[i];
    upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
    p[i] = P; m[i] = M; v[i] = V;
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) upd(ps[i], gs[i], ms[i], vs[i]);
}

// state = {t (as float bits of an int), lr}: writes lr * sqrt(1 - b2^t) / (1 - b1^t) and advances t — the host arithmetic of
// pdn_adam_step (reference optimizer.py:185-196), done on the device so that a recorded step can be replayed
__global__ void k_adam_prepare(int* t, const float* lr, float* step_size, float b1, float b2) {
  const int    tt = *t;
  const double a_t = sqrt(1.0 - pow((double)b2, (double)tt)) / (1.0 - pow((double)b1, (double)tt));
  *step_size = (float)((double)*lr * a_t);
  *t = tt + 1;
}

static inline int rows_grid(int64_t rows, int warps_per_block) { return (int)((rows + warps_per_block - 1) / warps_per_block); }

}  // namespace pdn

using namespace pdn;

extern "C" {

int pdn_softmax_fwd(int dtype, const void* x, void* y, int64_t rows, int64_t n, int log_mode) {
  PDN_TRY(ensure_init());
  if (rows == 0 || n == 0) return 0;
  if (dtype == PDN_F32 && n <= 0x7fffffff) {
    int grd = rows_grid(rows, 8);
    if (log_mode) k_softmax_fwd<true><<<grd, 256, 0, stream()>>>((const float*)x, (float*)y, rows, (int)n);
    else k_softmax_fwd<false><<<grd, 256, 0, stream()>>>((const float*)x, (float*)y, rows, (int)n);
    PDN_LAUNCHED("softmax_fwd");
    return 0;
  }
  int grd = (int)(rows < sm_count() * 8 ? rows : sm_count() * 8);
  if (dtype == PDN_F64) {
    if (log_mode) k_softmax_fwd_any<double, true><<<grd, 256, 0, stream()>>>((const double*)x, (double*)y, rows, n);
    else k_softmax_fwd_any<double, false><<<grd, 256, 0, stream()>>>((const double*)x, (double*)y, rows, n);
  } else if (dtype == PDN_F16) {
    if (log_mode) k_softmax_fwd_any<__half, true><<<grd, 256, 0, stream()>>>((const __half*)x, (__half*)y, rows, n);
    else k_softmax_fwd_any<__half, false><<<grd, 256, 0, stream()>>>((const __half*)x, (__half*)y, rows, n);
  } else {
    set_error("softmax: unsupported dtype %d", dtype);
    return PDN_ERR_UNSUPPORTED;
  }
  PDN_LAUNCHED("softmax_fwd_any");
  return 0;
}

int pdn_softmax_bwd(int dtype, const void* y, const void* g, void* dx, int64_t rows, int64_t n, int log_mode) {
  PDN_TRY(ensure_init());
  if (rows == 0 || n == 0) return 0;
  if (dtype == PDN_F32 && n <= 0x7fffffff) {
    int grd = rows_grid(rows, 8);
    if (log_mode) k_softmax_bwd<true><<<grd, 256, 0, stream()>>>((const float*)y, (const float*)g, (float*)dx, rows, (int)n);
    else k_softmax_bwd<false><<<grd, 256, 0, stream()>>>((const float*)y, (const float*)g, (float*)dx, rows, (int)n);
    PDN_LAUNCHED("softmax_bwd");
    return 0;
  }
  int grd = (int)(rows < sm_count() * 8 ? rows : sm_count() * 8);
  if (dtype == PDN_F64) {
    if (log_mode) k_softmax_bwd_any<double, true><<<grd, 256, 0, stream()>>>((const double*)y, (const double*)g, (double*)dx, rows, n);
    else k_softmax_bwd_any<double, false><<<grd, 256, 0, stream()>>>((const double*)y, (const double*)g, (double*)dx, rows, n);
  } else {
    set_error("softmax_bwd: unsupported dtype %d", dtype);
    return PDN_ERR_UNSUPPORTED;
  }
  PDN_LAUNCHED("softmax_bwd_any");
  return 0;
}

int pdn_rmsnorm_fwd(const float* x, const float* w, float* y, float* rstd, int64_t rows, int64_t n, float eps) {
  PDN_TRY(ensure_init());
  if (rows == 0 || n == 0) return 0;
  PDN_CHECK(n <= 0x7fffffff, "rmsnorm: row too long");
  k_rmsnorm_fwd<<<rows_grid(rows, 8), 256, 0, stream()>>>(x, w, y, rstd, rows, (int)n, eps);
  PDN_LAUNCHED("rmsnorm_fwd");
  return 0;
}

int pdn_rmsnorm_planes(const float* x, const float* w, void* planes, int64_t rows, int64_t n, int64_t Kp, float eps) {
  PDN_TRY(ensure_init());
  if (rows == 0 || n == 0) return 0;
  PDN_CHECK(n <= 0x7fffffff && Kp >= n && (Kp & 7) == 0, "rmsnorm_planes: bad K padding");
  const bool v4 = (n % 4 == 0) && n <= 1024 && ((((uintptr_t)x) | ((uintptr_t)w) | ((uintptr_t)planes)) & 15) == 0;
  if (v4 && n <= 384) k_rmsnorm_planes_v4<3><<<rows_grid(rows, 8), 256, 0, stream()>>>(x, w, (__nv_bfloat16*)planes, rows, (int)n, Kp, eps);
  else if (v4) k_rmsnorm_planes_v4<8><<<rows_grid(rows, 8), 256, 0, stream()>>>(x, w, (__nv_bfloat16*)planes, rows, (int)n, Kp, eps);
  else k_rmsnorm_planes<<<rows_grid(rows, 8), 256, 0, stream()>>>(x, w, (__nv_bfloat16*)planes, rows, (int)n, Kp, eps);
  PDN_LAUNCHED("rmsnorm_planes");
  return 0;
}

int pdn_swiglu_rows_planes(const float* gu, void* planes, int64_t rows, int64_t F, int64_t Kp) {
  PDN_TRY(ensure_init());
  if (rows * F == 0) return 0;
  PDN_CHECK(Kp >= F && (Kp & 7) == 0, "swiglu_rows_planes: bad K padding");
  if ((F % 4 == 0) && ((((uintptr_t)gu) | ((uintptr_t)planes)) & 15) == 0 && rows * (Kp / 4) < 0x7fffffff && rows < 0x7fffffff) {
    k_swiglu_rows_planes_v4<<<grid_for(rows * (Kp / 4), 256), 256, 0, stream()>>>(gu, (__nv_bfloat16*)planes, (int)rows, (int)F, (int)Kp);
  } else {
    k_swiglu_rows_planes<<<grid_for(rows * Kp, 256), 256, 0, stream()>>>(gu, (__nv_bfloat16*)planes, rows, F, Kp);
  }
  PDN_LAUNCHED("swiglu_rows_planes");
  return 0;
}

int pdn_rmsnorm_bwd(const float* x, const float* w, const float* rstd, const float* g, float* dx, float* dw, int64_t rows, int64_t n) {
  PDN_TRY(ensure_init());
  if (dw) PDN_CUDA(cudaMemsetAsync(dw, 0, (size_t)n * sizeof(float), stream()));
  if (rows == 0 || n == 0) return 0;
  PDN_CHECK(n <= 32 * 64, "rmsnorm_bwd: normalized size %lld > 2048 not supported by the fused kernel", (long long)n);
  int64_t want = (rows + 7) / 8;
  int     grd = (int)(want < sm_count() * 4 ? want : sm_count() * 4);
  if (n <= 32 * 16) k_rmsnorm_bwd<16><<<grd, 256, 0, stream()>>>(x, w, rstd, g, dx, dw, rows, (int)n);
  else k_rmsnorm_bwd<64><<<grd, 256, 0, stream()>>>(x, w, rstd, g, dx, dw, rows, (int)n);
  PDN_LAUNCHED("rmsnorm_bwd");
  return 0;
}

int pdn_ce_loss_fwd(const float* logits, const int64_t* target, float* loss, float* lse, int64_t N, int64_t C, int mean) {
  PDN_TRY(ensure_init());
  PDN_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream()));
  if (N == 0) return 0;
  PDN_CHECK(C > 0 && C <= 0x7fffffff, "cross entropy: bad class count");
  k_ce_fwd<<<rows_grid(N, 8), 256, 0, stream()>>>(logits, (const long long*)target, loss, lse, N, (int)C, mean ? 1.f / (float)N : 1.f);
  PDN_LAUNCHED("ce_fwd");
  return 0;
}

int pdn_ce_loss_bwd(const float* logits, const int64_t* target, const float* lse, const float* gloss, float* dlogits, int64_t N, int64_t C,
                    int mean) {
  PDN_TRY(ensure_init());
  if (N == 0) return 0;
  k_ce_bwd<<<grid_for(N * C, 256), 256, 0, stream()>>>(logits, (const long long*)target, lse, gloss, dlogits, N, (int)C,
                                                       mean ? 1.f / (float)N : 1.f);
  PDN_LAUNCHED("ce_bwd");
  return 0;
}

int pdn_swiglu(const float* gate, const float* up, float* out, int64_t n) {
  PDN_TRY(ensure_init());
  if (n == 0) return 0;
  bool    al = (((uintptr_t)gate | (uintptr_t)up | (uintptr_t)out) & 15) == 0;
  int64_t n4 = al ? n / 4 : 0;
  k_swiglu<<<grid_for(n4 > 0 ? n4 : 1, 256), 256, 0, stream()>>>((const float4*)gate, (const float4*)up, (float4*)out, n4, gate, up, out, n);
  PDN_LAUNCHED("swiglu");
  return 0;
}

int pdn_swiglu_rows(const float* gu, float* out, int64_t rows, int64_t F) {
  PDN_TRY(ensure_init());
  if (rows * F == 0) return 0;
  k_swiglu_rows<<<grid_for(rows * F, 256), 256, 0, stream()>>>(gu, out, rows, F);
  PDN_LAUNCHED("swiglu_rows");
  return 0;
}

int pdn_swiglu_bwd(const float* gate, const float* up, const float* g, float* dgate, float* dup, int64_t n) {
  PDN_TRY(ensure_init());
  if (n == 0) return 0;
  k_swiglu_bwd<<<grid_for(n, 256), 256, 0, stream()>>>(gate, up, g, dgate, dup, n);
  PDN_LAUNCHED("swiglu_bwd");
  return 0;
}

int pdn_adam_step(float* p, const float* grad, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float wd, int t,
                  float grad_scale) {
  PDN_TRY(ensure_init());
  if (n == 0) return 0;
  PDN_CHECK(t >= 1, "adam: step counter starts at 1");
  double a_t = sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t));
  float  step_size = (float)((double)lr * a_t);
  bool    al = (((uintptr_t)p | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  int64_t n4 = al ? n / 4 : 0;
  k_adam<<<grid_for(n4 > 0 ? n4 : 1, 256), 256, 0, stream()>>>((float4*)p, (const float4*)grad, (float4*)m, (float4*)v, n4, step_size, b1, b2, eps,
                                                                wd, grad_scale, p, grad, m, v, n, nullptr);
  PDN_LAUNCHED("adam");
  return 0;
}

int pdn_adam_step_dev(float* p, const float* grad, float* m, float* v, int64_t n, float b1, float b2, float eps, float wd, float grad_scale,
                      int* t_dev, const float* lr_dev, float* step_dev) {
  PDN_TRY(ensure_init());
  if (n == 0) return 0;
  k_adam_prepare<<<1, 1, 0, stream()>>>(t_dev, lr_dev, step_dev, b1, b2);
  PDN_LAUNCHED("adam_prepare");
  bool    al = (((uintptr_t)p | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  int64_t n4 = al ? n / 4 : 0;
  k_adam<<<grid_for(n4 > 0 ? n4 : 1, 256), 256, 0, stream()>>>((float4*)p, (const float4*)grad, (float4*)m, (float4*)v, n4, 0.f, b1, b2, eps, wd,
                                                                grad_scale, p, grad, m, v, n, step_dev);
  PDN_LAUNCHED("adam");
  return 0;
}

int pdn_adam_multi(int n_tensors, float* const* p, const float* const* grad, float* const* m, float* const* v, const int64_t* sizes, float lr,
                   float b1, float b2, float eps, float wd, int t, float grad_scale) {
  for (int i = 0; i < n_tensors; ++i) PDN_TRY(pdn_adam_step(p[i], grad[i], m[i], v[i], sizes[i], lr, b1, b2, eps, wd, t, grad_scale));
  return 0;
}

}  // extern "C"
