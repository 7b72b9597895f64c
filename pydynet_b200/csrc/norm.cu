// norm.cu — per-feature batch-statistic normalisation shared by BatchNorm1d/2d and the reference's "LayerNorm"
// (norm.py:58-73, 132-147, 203-218: mean / biased variance per feature over every other axis). x is viewed as
// [outer, C, inner]; the reference runs ~12 eager array expressions (each a full HBM pass); here: two reduction passes
// for the statistics (mean, then centred squares — same two-pass definition as the reference, not E[x^2]-E[x]^2), one
// pass to apply, and for backward one dual-reduction pass + one elementwise pass.
#include "common.cuh"

namespace pdn {

// MODE 0: s1 += x            MODE 1: s1 += (x - mean)^2        MODE 2: s1 += g, s2 += g * (x - mean) * rstd
// inner == 1 : block = 32 channels x 8 row-lanes, rows strided by gridDim.y chunks (coalesced along C)
template <int MODE>
__global__ void __launch_bounds__(256) k_feat_reduce_rows(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mean,
                                                          const float* __restrict__ var, float eps, float* __restrict__ s1, float* __restrict__ s2,
                                                          int64_t rows, int64_t C, float inv_m) {
  __shared__ float r1[8][33], r2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  float a1 = 0.f, a2 = 0.f;
  if (c < C) {
    const float mu = MODE >= 1 ? mean[c] : 0.f;
    const float rs = MODE == 2 ? rsqrtf(var[c] + eps) : 0.f;
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < rows; r += (int64_t)gridDim.y * 8) {
      const float xv = x[r * C + c];
      if (MODE == 0) a1 += xv;
      else if (MODE == 1) { float d = xv - mu; a1 += d * d; }
      else { float gv = g[r * C + c]; a1 += gv; a2 += gv * (xv - mu) * rs; }
    }
  }
  r1[ty][tx] = a1; r2[ty][tx] = a2;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { a1 += r1[i][tx]; a2 += r2[i][tx]; }
    atomicAdd(s1 + c, a1 * inv_m);
    if (MODE == 2) atomicAdd(s2 + c, a2 * inv_m);
  }
}

// inner == 1, C % 4 == 0: same reduction with 128-bit loads — block = 32 column-quads (128 channels) x 8 row-lanes
template <int MODE>
__global__ void __launch_bounds__(256) k_feat_reduce_rows4(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mean,
                                                           const float* __restrict__ var, float eps, float* __restrict__ s1, float* __restrict__ s2,
                                                           int64_t rows, int64_t C, float inv_m) {
  __shared__ float4 r1[8][33], r2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = ((int64_t)blockIdx.x * 32 + tx) * 4;
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  if (c < C) {
    float4 mu = a1, rs = a1;
    if (MODE >= 1) mu = *reinterpret_cast<const float4*>(mean + c);
    if (MODE == 2) {
      const float4 vv = *reinterpret_cast<const float4*>(var + c);
      rs = make_float4(rsqrtf(vv.x + eps), rsqrtf(vv.y + eps), rsqrtf(vv.z + eps), rsqrtf(vv.w + eps));
    }
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < rows; r += (int64_t)gridDim.y * 8) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + r * C + c));
      if (MODE == 0) { a1.x += xv.x; a1.y += xv.y; a1.z += xv.z; a1.w += xv.w; }
      else if (MODE == 1) {
        const float dx = xv.x - mu.x, dy = xv.y - mu.y, dz = xv.z - mu.z, dw = xv.w - mu.w;
        a1.x += dx * dx; a1.y += dy * dy; a1.z += dz * dz; a1.w += dw * dw;
      } else {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + r * C + c));
        a1.x += gv.x; a1.y += gv.y; a1.z += gv.z; a1.w += gv.w;
        a2.x += gv.x * (xv.x - mu.x) * rs.x; a2.y += gv.y * (xv.y - mu.y) * rs.y;
        a2.z += gv.z * (xv.z - mu.z) * rs.z; a2.w += gv.w * (xv.w - mu.w) * rs.w;
      }
    }
  }
  r1[ty][tx] = a1; r2[ty][tx] = a2;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float4 b1 = r1[i][tx], b2 = r2[i][tx];
      a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
      a2.x += b2.x; a2.y += b2.y; a2.z += b2.z; a2.w += b2.w;
    }
    atomicAdd(s1 + c, a1.x * inv_m); atomicAdd(s1 + c + 1, a1.y * inv_m); atomicAdd(s1 + c + 2, a1.z * inv_m); atomicAdd(s1 + c + 3, a1.w * inv_m);
    if (MODE == 2) {
      atomicAdd(s2 + c, a2.x * inv_m); atomicAdd(s2 + c + 1, a2.y * inv_m); atomicAdd(s2 + c + 2, a2.z * inv_m); atomicAdd(s2 + c + 3, a2.w * inv_m);
    }
  }
}

// ONE pass for mean and variance (inner == 1, C % 4 == 0): sums of (x - K) and (x - K)^2 around a pilot value K[c] = mean of the first
// <= 32 rows (recomputed by every thread from L2-resident rows, bit-identical everywhere), so that var = S2/m - (S1/m)^2 cancels
// against a shift that is within ~sigma/6 of the mean — the accuracy of the two-pass form without reading x twice (the second
// full pass was 35 of the 106 us a forward at [65536, 512] took).
__device__ __forceinline__ float4 feat_pilot(const float* __restrict__ x, int64_t rows, int64_t C, int64_t c) {
  const int n = rows < 32 ? (int)rows : 32;
  float4    k = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < n; ++r) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)r * C + c));
    k.x += v.x; k.y += v.y; k.z += v.z; k.w += v.w;
  }
  const float inv = 1.f / (float)n;
  return make_float4(k.x * inv, k.y * inv, k.z * inv, k.w * inv);
}
__global__ void __launch_bounds__(256) k_feat_stats1_rows4(const float* __restrict__ x, float* __restrict__ s1, float* __restrict__ s2, int64_t rows,
                                                           int64_t C, float inv_m) {
  __shared__ float4 r1[8][33], r2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = ((int64_t)blockIdx.x * 32 + tx) * 4;
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  if (c < C) {
    const float4 k = feat_pilot(x, rows, C, c);
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < rows; r += (int64_t)gridDim.y * 8) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + r * C + c));
      const float dx = xv.x - k.x, dy = xv.y - k.y, dz = xv.z - k.z, dw = xv.w - k.w;
      a1.x += dx; a1.y += dy; a1.z += dz; a1.w += dw;
      a2.x += dx * dx; a2.y += dy * dy; a2.z += dz * dz; a2.w += dw * dw;
    }
  }
  r1[ty][tx] = a1; r2[ty][tx] = a2;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float4 b1 = r1[i][tx], b2 = r2[i][tx];
      a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
      a2.x += b2.x; a2.y += b2.y; a2.z += b2.z; a2.w += b2.w;
    }
    atomicAdd(s1 + c, a1.x * inv_m); atomicAdd(s1 + c + 1, a1.y * inv_m); atomicAdd(s1 + c + 2, a1.z * inv_m); atomicAdd(s1 + c + 3, a1.w * inv_m);
    atomicAdd(s2 + c, a2.x * inv_m); atomicAdd(s2 + c + 1, a2.y * inv_m); atomicAdd(s2 + c + 2, a2.z * inv_m); atomicAdd(s2 + c + 3, a2.w * inv_m);
  }
}
// mean = K + S1/m, var = S2/m - (S1/m)^2, in place over (s1 -> mean, s2 -> var)
__global__ void __launch_bounds__(128) k_feat_stats1_finish(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ var,
                                                            int64_t rows, int64_t C) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= C) return;
  const float4 k = feat_pilot(x, rows, C, c);
  const float4 d = *reinterpret_cast<const float4*>(mean + c), q = *reinterpret_cast<const float4*>(var + c);
  *reinterpret_cast<float4*>(mean + c) = make_float4(k.x + d.x, k.y + d.y, k.z + d.z, k.w + d.w);
  *reinterpret_cast<float4*>(var + c) = make_float4(fmaxf(q.x - d.x * d.x, 0.f), fmaxf(q.y - d.y * d.y, 0.f), fmaxf(q.z - d.z * d.z, 0.f),
                                                    fmaxf(q.w - d.w * d.w, 0.f));
}
// running statistics in place: r = (1 - momentum) r + momentum stat (norm.py:66-69), both vectors in one launch
__global__ void __launch_bounds__(256) k_feat_running(float* __restrict__ rm, float* __restrict__ rv, const float* __restrict__ mean,
                                                      const float* __restrict__ var, float momentum, int64_t C) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    // the same two roundings as the array expressions `r *= (1 - m); r += stat * m`
    rm[c] = rm[c] * (1.f - momentum) + mean[c] * momentum;
    rv[c] = rv[c] * (1.f - momentum) + var[c] * momentum;
  }
}

// inner > 1 : block = (channel, chunk); threads walk j = (o, i) with i fastest (coalesced along inner)
template <int MODE>
__global__ void __launch_bounds__(256) k_feat_reduce_chan(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mean,
                                                          const float* __restrict__ var, float eps, float* __restrict__ s1, float* __restrict__ s2,
                                                          int64_t outer, int64_t C, int64_t inner, float inv_m) {
  __shared__ float red[32];
  const int64_t c = blockIdx.x, m = outer * inner;
  const float mu = MODE >= 1 ? mean[c] : 0.f;
  const float rs = MODE == 2 ? rsqrtf(var[c] + eps) : 0.f;
  float a1 = 0.f, a2 = 0.f;
  for (int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.y * blockDim.x) {
    const int64_t o = j / inner, i = j - o * inner;
    const int64_t e = (o * C + c) * inner + i;
    const float xv = x[e];
    if (MODE == 0) a1 += xv;
    else if (MODE == 1) { float d = xv - mu; a1 += d * d; }
    else { float gv = g[e]; a1 += gv; a2 += gv * (xv - mu) * rs; }
  }
  a1 = block_sum<float>(a1, red);
  if (MODE == 2) a2 = block_sum<float>(a2, red);
  if (threadIdx.x == 0) {
    atomicAdd(s1 + c, a1 * inv_m);
    if (MODE == 2) atomicAdd(s2 + c, a2 * inv_m);
  }
}

// Element-wise passes. The channel of a flat element index needs a division; it is done once per FOUR elements (float4,
// C*inner and inner are multiples of 4 on this path) and in 32-bit arithmetic (64-bit div/mod made these kernels ALU-bound:
// 23 % of HBM peak in the first ncu capture, profiles/r1_ncu_summary.md).
__device__ __forceinline__ uint32_t feat_channel(uint32_t e, uint32_t C, uint32_t inner) { return inner == 1 ? e % C : (e / inner) % C; }

template <bool VEC>
__global__ void __launch_bounds__(256) k_feat_apply(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                                                    const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y,
                                                    int64_t total, uint32_t C, uint32_t inner, float eps) {
  if (VEC) {
    const int64_t n4 = total >> 2;
    // offset inside one [C, inner] slab, advanced incrementally (no division in the loop; slab < 2^31 checked on the host)
    const uint32_t slab = C * inner;
    const int64_t  i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, di = (int64_t)gridDim.x * blockDim.x;
    uint32_t       e = (uint32_t)((i0 << 2) % slab);
    const uint32_t de = (uint32_t)((di << 2) % slab);
    for (int64_t i = i0; i < n4; i += di, e = (e + de >= slab) ? e + de - slab : e + de) {
      const float4   xv = __ldg(reinterpret_cast<const float4*>(x) + i);
      float4 o;
      if (inner == 1) {
        const uint32_t c = e;  // 4 consecutive channels (per-channel vectors are L1-resident: scalar loads, no alignment demand)
        o.x = (xv.x - mean[c]) * rsqrtf(var[c] + eps) * scale[c] + shift[c];
        o.y = (xv.y - mean[c + 1]) * rsqrtf(var[c + 1] + eps) * scale[c + 1] + shift[c + 1];
        o.z = (xv.z - mean[c + 2]) * rsqrtf(var[c + 2] + eps) * scale[c + 2] + shift[c + 2];
        o.w = (xv.w - mean[c + 3]) * rsqrtf(var[c + 3] + eps) * scale[c + 3] + shift[c + 3];
      } else {
        const uint32_t c = e / inner;  // all four elements share a channel (inner % 4 == 0)
        const float mu = mean[c], rs = rsqrtf(var[c] + eps) * scale[c], sh = shift[c];
        o.x = (xv.x - mu) * rs + sh; o.y = (xv.y - mu) * rs + sh; o.z = (xv.z - mu) * rs + sh; o.w = (xv.w - mu) * rs + sh;
      }
      reinterpret_cast<float4*>(y)[i] = o;
    }
  } else {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t c = (e / inner) % C;
      y[e] = (x[e] - mean[c]) * rsqrtf(var[c] + eps) * scale[c] + shift[c];
    }
  }
}

// dx = scale * rstd * (g - mean(g) - xhat * mean(g * xhat)); mg / mgx arrive already divided by m
template <bool VEC>
__global__ void __launch_bounds__(256) k_feat_bwd_dx(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mean,
                                                     const float* __restrict__ var, const float* __restrict__ scale, const float* __restrict__ mg,
                                                     const float* __restrict__ mgx, float* __restrict__ dx, int64_t total, uint32_t C, uint32_t inner,
                                                     float eps) {
  if (VEC) {
    const int64_t n4 = total >> 2;
    const uint32_t slab = C * inner;
    const int64_t  i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, di = (int64_t)gridDim.x * blockDim.x;
    uint32_t       e = (uint32_t)((i0 << 2) % slab);
    const uint32_t de = (uint32_t)((di << 2) % slab);
    for (int64_t i = i0; i < n4; i += di, e = (e + de >= slab) ? e + de - slab : e + de) {
      const float4   xv = __ldg(reinterpret_cast<const float4*>(x) + i), gv = __ldg(reinterpret_cast<const float4*>(g) + i);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
      float o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t c = inner == 1 ? e + t : e / inner;
        const float rs = rsqrtf(var[c] + eps);
        const float xhat = (xs[t] - mean[c]) * rs;
        o[t] = scale[c] * rs * (gs[t] - mg[c] - xhat * mgx[c]);
      }
      reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
  } else {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t c = (e / inner) % C;
      const float rs = rsqrtf(var[c] + eps);
      const float xhat = (x[e] - mean[c]) * rs;
      dx[e] = scale[c] * rs * (g[e] - mg[c] - xhat * mgx[c]);
    }
  }
}

// inner == 1 and C/4 a power of two <= 256 (the encoder's d_model 512): a thread keeps ONE float4 column group for the whole kernel,
// its per-channel constants in registers (the generic kernels above re-read 16-20 scalars of the statistics vectors and take four
// rsqrt per 16 bytes of data: load/store-unit-bound at ~55 % of HBM peak), rows strided, four rows in flight per thread.
__global__ void __launch_bounds__(256) k_feat_apply_cols(const float4* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                                                         const float* __restrict__ scale, const float* __restrict__ shift, float4* __restrict__ y,
                                                         int64_t rows, uint32_t C4, float eps) {
  const uint32_t c4 = threadIdx.x & (C4 - 1), rpb = 256 / C4, rin = threadIdx.x / C4, c = 4 * c4;
  float a[4], b[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    a[t] = rsqrtf(var[c + t] + eps) * scale[c + t];
    b[t] = shift[c + t];
  }
  const float   m0 = mean[c], m1 = mean[c + 1], m2 = mean[c + 2], m3 = mean[c + 3];
  const int64_t step = (int64_t)gridDim.x * rpb;
  for (int64_t r = (int64_t)blockIdx.x * rpb + rin; r < rows; r += 4 * step) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * step < rows) v[u] = __ldcs(x + (r + u * step) * C4 + c4);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * step < rows)
        __stcs(y + (r + u * step) * C4 + c4,
               make_float4((v[u].x - m0) * a[0] + b[0], (v[u].y - m1) * a[1] + b[1], (v[u].z - m2) * a[2] + b[2], (v[u].w - m3) * a[3] + b[3]));
  }
}
__global__ void __launch_bounds__(256) k_feat_bwd_dx_cols(const float4* __restrict__ x, const float4* __restrict__ g, const float* __restrict__ mean,
                                                          const float* __restrict__ var, const float* __restrict__ scale, const float* __restrict__ mg,
                                                          const float* __restrict__ mgx, float4* __restrict__ dx, int64_t rows, uint32_t C4, float eps) {
  const uint32_t c4 = threadIdx.x & (C4 - 1), rpb = 256 / C4, rin = threadIdx.x / C4, c = 4 * c4;
  float rs[4], a[4], mu[4], g0[4], gx[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    rs[t] = rsqrtf(var[c + t] + eps);
    a[t] = scale[c + t] * rs[t];
    mu[t] = mean[c + t], g0[t] = mg[c + t], gx[t] = mgx[c + t];
  }
  const int64_t step = (int64_t)gridDim.x * rpb;
  for (int64_t r = (int64_t)blockIdx.x * rpb + rin; r < rows; r += 2 * step) {
    float4 xv[2], gv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (r + u * step < rows) {
        xv[u] = __ldcs(x + (r + u * step) * C4 + c4);
        gv[u] = __ldcs(g + (r + u * step) * C4 + c4);
      }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (r + u * step < rows)
        __stcs(dx + (r + u * step) * C4 + c4, make_float4(a[0] * (gv[u].x - g0[0] - (xv[u].x - mu[0]) * rs[0] * gx[0]),
                                                           a[1] * (gv[u].y - g0[1] - (xv[u].y - mu[1]) * rs[1] * gx[1]),
                                                           a[2] * (gv[u].z - g0[2] - (xv[u].z - mu[2]) * rs[2] * gx[2]),
                                                           a[3] * (gv[u].w - g0[3] - (xv[u].w - mu[3]) * rs[3] * gx[3])));
  }
}
static bool feat_cols_ok(int64_t C, int64_t inner) {
  const int64_t C4 = C / 4;
  return inner == 1 && C % 4 == 0 && C4 >= 1 && C4 <= 256 && (C4 & (C4 - 1)) == 0;
}

__global__ void k_scale_vec(float* a, float* b, float s, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] *= s; if (b) b[i] *= s; }
}

// float4 path: 16-byte aligned tensors, 4 elements never straddle a channel boundary (inner == 1: C % 4 == 0 and the stats
// vectors are read as float4; inner > 1: inner % 4 == 0), one [C, inner] slab addressable in 32 bits
static bool feat_vec_ok(const float* a, const float* b, int64_t C, int64_t inner) {
  if ((((uintptr_t)a) | ((uintptr_t)b)) & 15) return false;
  if (C * inner >= 0x7fffffffLL) return false;
  return inner == 1 ? (C % 4 == 0) : (inner % 4 == 0);
}

template <int MODE>
static int feat_reduce(const float* x, const float* g, const float* mean, const float* var, float eps, float* s1, float* s2, int64_t outer,
                       int64_t C, int64_t inner, float inv_m) {
  PDN_CUDA(cudaMemsetAsync(s1, 0, (size_t)C * sizeof(float), stream()));
  if (s2) PDN_CUDA(cudaMemsetAsync(s2, 0, (size_t)C * sizeof(float), stream()));
  const int sms = sm_count();
  const bool al16 = ((((uintptr_t)x) | ((uintptr_t)g) | ((uintptr_t)mean) | ((uintptr_t)var)) & 15) == 0;
  if (inner == 1 && (C & 3) == 0 && al16) {
    int64_t gx = (C / 4 + 31) / 32;
    int64_t gy = (sms * 8 + gx - 1) / gx;
    if (gy > (outer + 7) / 8) gy = (outer + 7) / 8;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    k_feat_reduce_rows4<MODE><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream()>>>(x, g, mean, var, eps, s1, s2, outer, C, inv_m);
  } else if (inner == 1) {
    int64_t gx = (C + 31) / 32;
    int64_t gy = (sms * 4 + gx - 1) / gx;
    if (gy > (outer + 7) / 8) gy = (outer + 7) / 8;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    k_feat_reduce_rows<MODE><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream()>>>(x, g, mean, var, eps, s1, s2, outer, C, inv_m);
  } else {
    int64_t m = outer * inner;
    int64_t gy = (sms * 4 + C - 1) / C;
    if (gy > (m + 255) / 256) gy = (m + 255) / 256;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    PDN_CHECK(C <= 0x7fffffff, "feature norm: too many channels");
    k_feat_reduce_chan<MODE><<<dim3((unsigned)C, (unsigned)gy), 256, 0, stream()>>>(x, g, mean, var, eps, s1, s2, outer, C, inner, inv_m);
  }
  PDN_LAUNCHED("feat_reduce");
  return 0;
}

}  // namespace pdn

using namespace pdn;

extern "C" {

int pdn_bnorm_stats(const float* x, float* mean, float* var, int64_t outer, int64_t C, int64_t inner) {
  PDN_TRY(ensure_init());
  if (C == 0) return 0;
  PDN_CHECK(outer * inner > 0, "feature norm: empty reduction");
  const float inv_m = 1.f / (float)(outer * inner);
  static const bool two_pass = getenv("PDN_NORM_TWO_PASS") != nullptr;
  if (!two_pass && inner == 1 && (C & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)mean) | ((uintptr_t)var)) & 15) == 0) {
    PDN_CUDA(cudaMemsetAsync(mean, 0, (size_t)C * sizeof(float), stream()));
    PDN_CUDA(cudaMemsetAsync(var, 0, (size_t)C * sizeof(float), stream()));
    int64_t gx = (C / 4 + 31) / 32, gy = (sm_count() * 8 + gx - 1) / gx;
    if (gy > (outer + 7) / 8) gy = (outer + 7) / 8;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    k_feat_stats1_rows4<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream()>>>(x, mean, var, outer, C, inv_m);
    PDN_LAUNCHED("feat_stats1");
    k_feat_stats1_finish<<<(unsigned)((C / 4 + 127) / 128), 128, 0, stream()>>>(x, mean, var, outer, C);
    PDN_LAUNCHED("feat_stats1_finish");
    return 0;
  }
  PDN_TRY((feat_reduce<0>(x, nullptr, nullptr, nullptr, 0.f, mean, nullptr, outer, C, inner, inv_m)));
  PDN_TRY((feat_reduce<1>(x, nullptr, mean, nullptr, 0.f, var, nullptr, outer, C, inner, inv_m)));
  return 0;
}

/* running statistics of the batch-coupled norms, in place: r <- (1 - momentum) r + momentum stat (norm.py:66-69, 140-143, 211-214) */
int pdn_bnorm_running(float* running_mean, float* running_var, const float* mean, const float* var, float momentum, int64_t C) {
  PDN_TRY(ensure_init());
  if (C == 0) return 0;
  k_feat_running<<<(unsigned)((C + 255) / 256), 256, 0, stream()>>>(running_mean, running_var, mean, var, momentum, C);
  PDN_LAUNCHED("feat_running");
  return 0;
}

/* data-parallel building blocks: local partial statistics scaled by 1/m_global so that a cross-rank SUM gives the global
 * mean / variance / gradient means (equal shards). which = 0: mean partial, 1: centred-square partial (needs the global mean) */
int pdn_bnorm_partial(const float* x, const float* mean, float* out, int64_t outer, int64_t C, int64_t inner, int which, float inv_m_global) {
  PDN_TRY(ensure_init());
  if (C == 0) return 0;
  if (which == 0) return feat_reduce<0>(x, nullptr, nullptr, nullptr, 0.f, out, nullptr, outer, C, inner, inv_m_global);
  return feat_reduce<1>(x, nullptr, mean, nullptr, 0.f, out, nullptr, outer, C, inner, inv_m_global);
}

/* backward phase 1: mg[c] = Σ_local g / m_global, mgx[c] = Σ_local g·xhat / m_global (all-reduce these, then phase 2) */
int pdn_bnorm_bwd_reduce(const float* x, const float* mean, const float* var, const float* g, float* mg, float* mgx, int64_t outer, int64_t C,
                         int64_t inner, float eps, float inv_m_global) {
  PDN_TRY(ensure_init());
  if (C == 0) return 0;
  return feat_reduce<2>(x, g, mean, var, eps, mg, mgx, outer, C, inner, inv_m_global);
}

/* backward phase 2: dx from the (global) means; then mg, mgx are rescaled by m_local_scale to become this rank's share of
 * dshift / dscale (their cross-rank sum — done by the gradient all-reduce — is the global gradient) */
int pdn_bnorm_bwd_dx(const float* x, const float* mean, const float* var, const float* scale, const float* g, const float* mg, const float* mgx,
                     float* dx, int64_t outer, int64_t C, int64_t inner, float eps) {
  PDN_TRY(ensure_init());
  const int64_t total = outer * C * inner;
  if (total == 0 || !dx) return 0;
  if (feat_vec_ok(x, dx, C, inner) && (((uintptr_t)g) & 15) == 0 && feat_cols_ok(C, inner))
    k_feat_bwd_dx_cols<<<sm_count() * 8, 256, 0, stream()>>>((const float4*)x, (const float4*)g, mean, var, scale, mg, mgx, (float4*)dx, outer,
                                                             (uint32_t)(C / 4), eps);
  else if (feat_vec_ok(x, dx, C, inner) && (((uintptr_t)g) & 15) == 0)
    k_feat_bwd_dx<true><<<grid_for(total / 4, 256), 256, 0, stream()>>>(x, g, mean, var, scale, mg, mgx, dx, total, (uint32_t)C, (uint32_t)inner, eps);
  else
    k_feat_bwd_dx<false><<<grid_for(total, 256), 256, 0, stream()>>>(x, g, mean, var, scale, mg, mgx, dx, total, (uint32_t)C, (uint32_t)inner, eps);
  PDN_LAUNCHED("feat_bwd_dx");
  return 0;
}

int pdn_bnorm_apply(const float* x, const float* mean, const float* var, const float* scale, const float* shift, float* y, int64_t outer,
                    int64_t C, int64_t inner, float eps) {
  PDN_TRY(ensure_init());
  const int64_t total = outer * C * inner;
  if (total == 0) return 0;
  if (feat_vec_ok(x, y, C, inner) && feat_cols_ok(C, inner))
    k_feat_apply_cols<<<sm_count() * 8, 256, 0, stream()>>>((const float4*)x, mean, var, scale, shift, (float4*)y, outer, (uint32_t)(C / 4), eps);
  else if (feat_vec_ok(x, y, C, inner)) k_feat_apply<true><<<grid_for(total / 4, 256), 256, 0, stream()>>>(x, mean, var, scale, shift, y, total, (uint32_t)C, (uint32_t)inner, eps);
  else k_feat_apply<false><<<grid_for(total, 256), 256, 0, stream()>>>(x, mean, var, scale, shift, y, total, (uint32_t)C, (uint32_t)inner, eps);
  PDN_LAUNCHED("feat_apply");
  return 0;
}

int pdn_bnorm_bwd(const float* x, const float* mean, const float* var, const float* scale, const float* g, float* dx, float* dscale,
                  float* dshift, int64_t outer, int64_t C, int64_t inner, float eps) {
  PDN_TRY(ensure_init());
  const int64_t total = outer * C * inner;
  if (total == 0) return 0;
  const float m = (float)(outer * inner);
  // dshift <- mean(g), dscale <- mean(g * xhat) (scaled by 1/m inside the reduction), used by dx, then rescaled to sums
  PDN_TRY((feat_reduce<2>(x, g, mean, var, eps, dshift, dscale, outer, C, inner, 1.f / m)));
  if (dx) {
    if (feat_vec_ok(x, dx, C, inner) && (((uintptr_t)g) & 15) == 0 && feat_cols_ok(C, inner))
      k_feat_bwd_dx_cols<<<sm_count() * 8, 256, 0, stream()>>>((const float4*)x, (const float4*)g, mean, var, scale, dshift, dscale, (float4*)dx, outer,
                                                               (uint32_t)(C / 4), eps);
    else if (feat_vec_ok(x, dx, C, inner) && (((uintptr_t)g) & 15) == 0)
      k_feat_bwd_dx<true><<<grid_for(total / 4, 256), 256, 0, stream()>>>(x, g, mean, var, scale, dshift, dscale, dx, total, (uint32_t)C, (uint32_t)inner, eps);
    else
      k_feat_bwd_dx<false><<<grid_for(total, 256), 256, 0, stream()>>>(x, g, mean, var, scale, dshift, dscale, dx, total, (uint32_t)C, (uint32_t)inner, eps);
    PDN_LAUNCHED("feat_bwd_dx");
  }
  k_scale_vec<<<(unsigned)((C + 255) / 256), 256, 0, stream()>>>(dshift, dscale, m, C);
  PDN_LAUNCHED("scale_vec");
  return 0;
}

}  // extern "C"
