// conv.cu — Conv2d forward / backward-data / backward-weight as tcgen05 GEMMs over gathered operands, and Pool2d.
//
// The reference convolves by materialising an fp32 im2col matrix (zero-pad -> 6-D strided view -> .copy() -> reshape ->
// sgemm -> NCHW transpose view; backward = np.add.at scatter — functional.py:194-281). Here the im2col gather writes the
// GEMM's operand format directly (K-major bf16 hi/lo planes of the BF16x3 split, gemm_tc.cu), so there is no fp32 column
// matrix, no pad copy and no scatter-add:
//   forward        y[n,o,pix]   = Σ_ckk  col(x)[(n,pix), ckk] · W[o, ckk]          epilogue writes NCHW (+ bias[o])
//   backward-data  dx[n,c,pix'] = Σ_okk  gat(g)[(n,pix'), okk] · W[o, c, kk]        (transposed-conv gather: no col2im atomics)
//   backward-weight dW[o, ckk]  = Σ_m    gᵀ[o, m] · colᵀ(x)[ckk, m]                 split-K over m = (n, pix)
// Pooling is a direct window kernel (the reference runs it through the same im2col + max/mean + add.at).
#include "common.cuh"
#include "gemm_tc.h"
#include "conv_gather.cuh"
#include <stdlib.h>

namespace pdn {

// Writes planes [2][R][Kp]: ROWS_M ? (R = Mtot, k index = kk) : (R = Ktot, k index = m). 32(m) x 64(kk) tile through smem so
// that both the gather (along m = consecutive pixels) and the plane stores (along the plane's k axis) are coalesced.
template <int MODE, bool ROWS_M>
__global__ void __launch_bounds__(256) k_conv_pack(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, ConvGeom g, int64_t Mtot,
                                                   int64_t Ktot, int64_t Kp) {
  __shared__ float tile[64][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * 32;
  const int kk0 = blockIdx.y * 64;
  const RowPos rp = conv_row<MODE>(g, m0 + tx, Mtot);
#pragma unroll
  for (int i = 0; i < 8; ++i) tile[ty + i * 8][tx] = conv_fetch<MODE>(src, g, rp, kk0 + ty + i * 8, (int)Ktot);
  __syncthreads();
  const int64_t R = ROWS_M ? Mtot : Ktot;
  __nv_bfloat16* hi = dst;
  __nv_bfloat16* lo = dst + (size_t)R * Kp;
  if (ROWS_M) {  // rows = m: each warp writes 64 consecutive k (one bf16x2 per lane) of one row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int     rl = ty + i * 8;
      const int64_t row = m0 + rl, col = kk0 + 2 * tx;
      if (row < R && col < Kp) {  // Kp is even
        const float v0 = tile[2 * tx][rl], v1 = tile[2 * tx + 1][rl];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        __nv_bfloat162 H2, L2;
        H2.x = h0; H2.y = h1;
        L2.x = __float2bfloat16_rn(v0 - __bfloat162float(h0));
        L2.y = __float2bfloat16_rn(v1 - __bfloat162float(h1));
        *reinterpret_cast<__nv_bfloat162*>(hi + row * Kp + col) = H2;
        *reinterpret_cast<__nv_bfloat162*>(lo + row * Kp + col) = L2;
      }
    }
  } else {  // rows = kk, k axis = m
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t row = kk0 + ty + i * 8, col = m0 + tx;
      if (row < R && col < Kp) {
        const float v = tile[ty + i * 8][tx];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[row * Kp + col] = h;
        lo[row * Kp + col] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
  }
}

template <int MODE, bool ROWS_M>
static int conv_pack(const float* src, const ConvGeom& g, int64_t Mtot, int64_t Ktot, Scratch* buf, PackedOperand* out) {
  const int64_t R = ROWS_M ? Mtot : Ktot, K = ROWS_M ? Ktot : Mtot;
  const int64_t Kp = (K + 7) & ~(int64_t)7;
  PDN_TRY(buf->alloc((size_t)2 * R * Kp * sizeof(__nv_bfloat16)));
  // the grid walks (m tiles, kk tiles); pad columns [K, Kp) are covered because conv_fetch returns 0 out of range
  const int64_t m_ext = ROWS_M ? Mtot : Kp, kk_ext = ROWS_M ? Kp : Ktot;
  dim3 grd((unsigned)((m_ext + 31) / 32), (unsigned)((kk_ext + 63) / 64));
  PDN_CHECK(grd.y <= 65535, "conv: C*k*k too large for the pack grid");
  k_conv_pack<MODE, ROWS_M><<<grd, 256, 0, stream()>>>(src, (__nv_bfloat16*)buf->p, g, Mtot, Ktot, Kp);
  PDN_LAUNCHED("conv_pack");
  out->planes = buf->p; out->R = R; out->K = K; out->Kp = Kp; out->nbatch = 1;
  out->pbs[0] = out->pbs[1] = out->pbs[2] = 0;
  return 0;
}

static int make_geom(ConvGeom& g, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int k, int stride, int pad) {
  PDN_CHECK(k >= 1 && stride >= 1 && pad >= 0, "conv: bad kernel/stride/pad");
  PDN_CHECK(H + 2 * pad >= k && W + 2 * pad >= k, "conv: kernel larger than the padded input");
  PDN_CHECK(C * k * k <= 0x7fffffff && O * k * k <= 0x7fffffff && H * W <= 0x7fffffff, "conv: dimension too large");
  g.N = N; g.C = C; g.H = H; g.W = W; g.O = O; g.k = k; g.stride = stride; g.pad = pad;
  g.oh = (H + 2 * pad - k) / stride + 1;
  g.ow = (W + 2 * pad - k) / stride + 1;
  return 0;
}

static void tc_defaults(TcArgs& t) {
  for (int i = 0; i < 3; ++i) { t.nb[i] = 1; t.c_bs[i] = 0; t.a_pbs[i] = 0; t.b_pbs[i] = 0; }
  t.bias = nullptr; t.accumulate = 0; t.splits = 1; t.nchw_hw = 0; t.c_clear_bytes = 0; t.amax_val = nullptr; t.amax_idx = nullptr;
}

// per-output-channel sum of an NCHW tensor: dbias[o] = Σ_{n,pix} g[n,o,pix]; grid (O, image chunks), atomics into a zeroed vector
__global__ void __launch_bounds__(256) k_channel_sum(const float* __restrict__ g, float* __restrict__ out, int64_t N, int64_t O, int64_t hw) {
  __shared__ float red[32];
  const int64_t o = blockIdx.x;
  float s = 0.f;
  for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
    const float* p = g + (n * O + o) * hw;
    for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) s += p[i];
  }
  s = block_sum<float>(s, red);
  if (threadIdx.x == 0) atomicAdd(out + o, s);
}

// ---------------------------------------------------------------- thin first layers -------------------------------
// A stride-1 convolution over 1-3 input channels (the 1-channel first layer of BASELINE config 2: C k k = 9) is 1/7 of one 64-wide
// k-block of the tensor-core path, which then spends its time on padding (48 us for 36 MFLOP). Direct fp32, window shape known at
// compile time (KS x KS taps, up to CM channels; no runtime index arithmetic in the inner loops):
//   forward          one thread per output pixel keeps its window in registers and walks the output channels (weights + bias in
//                    shared memory, stores coalesced along the pixels of a channel plane)
//   backward-weight  one CTA per image stages gy[n] and the zero-haloed x[n] in shared memory; thread (o, row group) keeps the
//                    whole window's partial sums (and the bias sum) in registers over its pixels and over the CTA's images, then
//                    shared-memory atomics per CTA, one global atomicAdd per CTA and weight
constexpr int kThinMaxO = 64;

template <int KS, int CM>
__global__ void __launch_bounds__(256) k_conv_thin_fwd(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ y, ConvGeom g) {
  constexpr int WIN = CM * KS * KS;
  __shared__ float ws[kThinMaxO * WIN + kThinMaxO];
  const int C = (int)g.C, O = (int)g.O, win = C * KS * KS;
  for (int i = threadIdx.x; i < O * WIN; i += blockDim.x) {
    const int o = i / WIN, r = i - o * WIN;
    ws[i] = r < win ? w[o * win + r] : 0.f;
  }
  for (int i = threadIdx.x; i < O; i += blockDim.x) ws[kThinMaxO * WIN + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int64_t hw = g.oh * g.ow, total = g.N * hw;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = m / hw;
    const int     pix = (int)(m - n * hw), oy = pix / (int)g.ow, ox = pix - oy * (int)g.ow;
    const int     y0 = oy - g.pad, x0 = ox - g.pad;
    float         v[WIN];
#pragma unroll
    for (int c = 0; c < CM; ++c)
#pragma unroll
      for (int ky = 0; ky < KS; ++ky)
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const int yy = y0 + ky, xx = x0 + kx;
          v[(c * KS + ky) * KS + kx] =
              (c < C && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? __ldg(x + ((n * C + c) * g.H + yy) * g.W + xx) : 0.f;
        }
    float* yo = y + n * O * hw + pix;
    for (int o = 0; o < O; ++o) {
      const float* wo = ws + o * WIN;
      float        acc = ws[kThinMaxO * WIN + o];
#pragma unroll
      for (int i = 0; i < WIN; ++i) acc = fmaf(v[i], wo[i], acc);
      yo[(int64_t)o * hw] = acc;
    }
  }
}

// dynamic shared memory: gy[n] as [O][oh*ow + 1] | x[n] with a zero halo as [C][H + 2 pad][W + 2 pad] | dw partials [O][WIN + 1]
template <int KS, int CM>
__global__ void __launch_bounds__(256) k_conv_thin_bwd_weight(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ dw,
                                                              float* __restrict__ dbias, ConvGeom g) {
  constexpr int WIN = CM * KS * KS;
  extern __shared__ float sm[];
  const int C = (int)g.C, O = (int)g.O, ow = (int)g.ow, oh = (int)g.oh, hw = oh * ow, gst = hw + 1;
  const int Hp = (int)g.H + 2 * g.pad, Wp = (int)g.W + 2 * g.pad, win = C * KS * KS;
  float *gs = sm, *xs = gs + O * gst, *part = xs + C * Hp * Wp;
  const int G = 256 / O;  // row groups per output channel
  const int o = threadIdx.x % O, grp = threadIdx.x / O;
  const bool worker = grp < G;
  float acc[WIN], accb = 0.f;
#pragma unroll
  for (int i = 0; i < WIN; ++i) acc[i] = 0.f;
  for (int i = threadIdx.x; i < C * Hp * Wp; i += blockDim.x) xs[i] = 0.f;  // the halo stays zero for every image
  for (int i = threadIdx.x; i < O * (WIN + 1); i += blockDim.x) part[i] = 0.f;
  for (int64_t n = blockIdx.x; n < g.N; n += gridDim.x) {
    __syncthreads();
    if ((hw & 3) == 0 && ((uintptr_t)gy & 15) == 0) {  // 128-bit loads, several in flight per thread (the copy is latency-bound)
      const float4* src = reinterpret_cast<const float4*>(gy + n * O * hw);
      const int     n4 = O * hw / 4, hw4 = hw / 4;
#pragma unroll 4
      for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 v = __ldg(src + i);
        const int    o2 = i / hw4;
        float*       d = gs + o2 * gst + (i - o2 * hw4) * 4;
        d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
      }
    } else {
      for (int i = threadIdx.x; i < O * hw; i += blockDim.x) gs[(i / hw) * gst + i % hw] = gy[n * O * hw + i];
    }
    for (int i = threadIdx.x; i < C * (int)(g.H * g.W); i += blockDim.x) {
      const int c = i / (int)(g.H * g.W), r = i - c * (int)(g.H * g.W), yy = r / (int)g.W, xx = r - yy * (int)g.W;
      xs[(c * Hp + yy + g.pad) * Wp + xx + g.pad] = x[n * C * g.H * g.W + i];
    }
    __syncthreads();
    if (worker) {
      for (int oy = grp; oy < oh; oy += G) {
        const float* grow = gs + o * gst + oy * ow;
        for (int ox = 0; ox < ow; ++ox) {
          const float gv = grow[ox];
          accb += gv;
#pragma unroll
          for (int c = 0; c < CM; ++c)
            if (c < C) {
#pragma unroll
              for (int ky = 0; ky < KS; ++ky)
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) acc[(c * KS + ky) * KS + kx] = fmaf(gv, xs[(c * Hp + oy + ky) * Wp + ox + kx], acc[(c * KS + ky) * KS + kx]);
            }
        }
      }
    }
  }
  if (worker) {
#pragma unroll
    for (int i = 0; i < WIN; ++i)
      if (i < win) atomicAdd(part + o * (WIN + 1) + i, acc[i]);
    atomicAdd(part + o * (WIN + 1) + WIN, accb);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < O * win; i += blockDim.x) atomicAdd(dw + i, part[(i / win) * (WIN + 1) + i % win]);
  if (dbias)
    for (int i = threadIdx.x; i < O; i += blockDim.x) atomicAdd(dbias + i, part[i * (WIN + 1) + WIN]);
}

// window shapes the direct kernels are instantiated for: 3x3 over <= 3 channels, 5x5 over 1 channel
static int conv_thin_kind(const ConvGeom& g) {
  static const bool off = getenv("PDN_CONV_THIN") && getenv("PDN_CONV_THIN")[0] == '0';
  if (off || g.stride != 1 || g.O > kThinMaxO || g.O < 1) return 0;
  if (g.k == 3 && g.C == 1) return 31;
  if (g.k == 3 && g.C <= 3) return 33;
  if (g.k == 5 && g.C == 1) return 51;
  return 0;
}
static size_t conv_thin_bw_smem(const ConvGeom& g, int WIN) {
  return (size_t)(g.O * (g.oh * g.ow + 1) + g.C * (g.H + 2 * g.pad) * (g.W + 2 * g.pad) + g.O * (WIN + 1)) * sizeof(float);
}

// ---------------------------------------------------------------- pooling ----------------------------------------
struct PoolGeom {
  int64_t N, C, H, W, oh, ow;
  int     k, stride, pad, mode;
};

__global__ void __launch_bounds__(256) k_pool_fwd(const float* __restrict__ x, float* __restrict__ y, PoolGeom g) {
  const unsigned total = (unsigned)(g.N * g.C * g.oh * g.ow);  // 32-bit index arithmetic (checked on the host)
  const unsigned ow = (unsigned)g.ow, oh = (unsigned)g.oh;
  const int      H = (int)g.H, W = (int)g.W, k = g.k;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned r = i / ow, ox = i - r * ow, nc = r / oh, oy = r - nc * oh;
    const float*   xp = x + (size_t)nc * H * W;
    float          acc = g.mode == 0 ? -INFINITY : 0.f;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        const int iy = (int)oy * g.stride + ky - g.pad, ix = (int)ox * g.stride + kx - g.pad;
        // zero padding takes part in the max / mean exactly like the reference's xp.pad (functional.py:235-251)
        const float v = (iy < 0 || iy >= H || ix < 0 || ix >= W) ? 0.f : __ldg(xp + iy * W + ix);
        acc = g.mode == 0 ? fmaxf(acc, v) : acc + v;
      }
    y[i] = g.mode == 0 ? acc : acc / (float)(k * k);
  }
}

// one thread per INPUT element gathers from every window that contains it (no atomics): max mode gives the window's full
// gradient to every element equal to the window max (tensor.py:741-747). Index arithmetic is 32-bit (the host checks the tensor has
// fewer than 2^31 elements): the first version did its div / mod in 64 bits inside the window loops and took 76 us for the 4 M-element
// tensor of LeNet's first pooling layer - 19 % of the recorded training step.
__global__ void __launch_bounds__(256) k_pool_bwd(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                                                  float* __restrict__ dx, PoolGeom g) {
  const unsigned total = (unsigned)(g.N * g.C * g.H * g.W);
  const unsigned W = (unsigned)g.W, H = (unsigned)g.H, ow = (unsigned)g.ow, oh = (unsigned)g.oh;
  const int      k = g.k, stride = g.stride, pad = g.pad;
  const float    kk = (float)(k * k);
  const bool     tiled = (stride == k) && pad == 0;  // non-overlapping windows: an element belongs to at most one
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned r = i / W, ix = i - r * W, nc = r / H, iy = r - nc * H;
    const float    xv = x[i];
    const float *  yp = y + (size_t)nc * oh * ow, *gp = gy + (size_t)nc * oh * ow;
    float          acc = 0.f;
    if (tiled) {
      const unsigned oy = iy / (unsigned)k, ox = ix / (unsigned)k;
      if (oy < oh && ox < ow) {
        const float gg = __ldg(gp + oy * ow + ox);
        acc = g.mode == 0 ? ((__ldg(yp + oy * ow + ox) == xv) ? gg : 0.f) : gg / kk;
      }
    } else {
      for (int ky = 0; ky < k; ++ky) {
        const int ty = (int)iy + pad - ky;
        if (ty < 0 || ty % stride) continue;
        const int oy = ty / stride;
        if (oy >= (int)oh) continue;
        for (int kx = 0; kx < k; ++kx) {
          const int tx = (int)ix + pad - kx;
          if (tx < 0 || tx % stride) continue;
          const int ox = tx / stride;
          if (ox >= (int)ow) continue;
          const float gg = __ldg(gp + oy * ow + ox);
          if (g.mode == 0) acc += (__ldg(yp + oy * ow + ox) == xv) ? gg : 0.f;
          else acc += gg / kk;
        }
      }
    }
    dx[i] = acc;
  }
}

}  // namespace pdn

using namespace pdn;

extern "C" {

int pdn_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int k,
                   int stride, int pad, int64_t x_version) {
  PDN_TRY(ensure_init());
  ConvGeom g;
  PDN_TRY(make_geom(g, N, C, H, W, O, k, stride, pad));
  const int64_t M = N * g.oh * g.ow, K = C * k * k;
  if (M == 0 || O == 0) return 0;
  Scratch       bufA, bufB;
  PackedOperand A, B;
  const int64_t one[3] = {1, 1, 1}, zero[3] = {0, 0, 0};
  if (const int kind = conv_thin_kind(g)) {
    if (kind == 31) k_conv_thin_fwd<3, 1><<<grid_for(M, 256), 256, 0, stream()>>>(x, w, bias, y, g);
    else if (kind == 33) k_conv_thin_fwd<3, 3><<<grid_for(M, 256), 256, 0, stream()>>>(x, w, bias, y, g);
    else k_conv_thin_fwd<5, 1><<<grid_for(M, 256), 256, 0, stream()>>>(x, w, bias, y, g);
    PDN_LAUNCHED("conv_thin_fwd");
    return 0;
  }
  if (conv_tma_ok(C, stride, N, C, H, W)) {
    // weights per tap as K-major planes [tap][O][C]: w[o, c, ky, kx] -> rows o (stride C k k), contraction c (stride k k), batch tap
    const int64_t kk = (int64_t)k * k, tnb[3] = {1, 1, kk}, tbs[3] = {0, 0, 1};
    PDN_TRY(pack_operand_ex(w, O, C, C * kk, kk, 0, 0, tnb, tbs, &bufB, &B));
    return conv_tma_forward(x, N, C, H, W, B, bias, y, O, g.oh, g.ow, k, pad, +1, x_version);
  }
  PDN_TRY(pack_operand_ex(w, O, K, K, 1, 0, 0, one, zero, &bufB, &B));
  TcArgs t;
  tc_defaults(t);
  t.C = y; t.bias = bias; t.M = M; t.N = O; t.K = K; t.ldc = O;
  t.nchw_hw = g.oh * g.ow;
  t.c_clear_bytes = (size_t)M * O * sizeof(float);
  if (!getenv("PDN_CONV_EXPLICIT")) return gemm_tc_conv(x, g, 1, M, (int)K, B, t);  // implicit GEMM: no column matrix in HBM
  PDN_TRY((conv_pack<0, true>(x, g, M, K, &bufA, &A)));
  return gemm_tc_packed(A, B, t, 1);
}

int pdn_conv2d_bwd_data(const float* gy, const float* w, float* dx, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int k, int stride,
                        int pad, int64_t gy_version) {
  PDN_TRY(ensure_init());
  ConvGeom g;
  PDN_TRY(make_geom(g, N, C, H, W, O, k, stride, pad));
  const int64_t M = N * H * W, K = O * k * k;
  if (M == 0 || C == 0) return 0;
  Scratch       bufA, bufB;
  PackedOperand A, B;
  if (conv_tma_ok(O, stride, N, O, g.oh, g.ow)) {
    // dx[n, c, y, x] = Σ_{tap, o} gy[n, o, y - ky + pad, x - kx + pad] · w[o, c, ky, kx]: planes [tap][C rows][O contraction]
    const int64_t kk = (int64_t)k * k, tnb[3] = {1, 1, kk}, tbs[3] = {0, 0, 1};
    PDN_TRY(pack_operand_ex(w, C, O, kk, C * kk, 0, 0, tnb, tbs, &bufB, &B));
    return conv_tma_forward(gy, N, O, g.oh, g.ow, B, nullptr, dx, C, H, W, k, pad, -1, gy_version);
  }
  // B rows = input channel c; k index (o, ky, kx) -> W[o, c, ky, kx]
  const int64_t one[3] = {1, 1, 1}, zero[3] = {0, 0, 0};
  PDN_TRY(pack_operand_ex(w, C, K, (int64_t)k * k, 1, (int64_t)k * k, C * (int64_t)k * k, one, zero, &bufB, &B));
  TcArgs t;
  tc_defaults(t);
  t.C = dx; t.M = M; t.N = C; t.K = K; t.ldc = C;
  t.nchw_hw = H * W;
  t.c_clear_bytes = (size_t)M * C * sizeof(float);
  if (!getenv("PDN_CONV_EXPLICIT")) return gemm_tc_conv(gy, g, 2, M, (int)K, B, t);
  PDN_TRY((conv_pack<1, true>(gy, g, M, K, &bufA, &A)));
  return gemm_tc_packed(A, B, t, 1);
}

int pdn_conv2d_bwd_weight(const float* x, const float* gy, float* dw, float* dbias, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int k,
                          int stride, int pad, int64_t x_version, int64_t gy_version) {
  PDN_TRY(ensure_init());
  ConvGeom g;
  PDN_TRY(make_geom(g, N, C, H, W, O, k, stride, pad));
  const int64_t hw = g.oh * g.ow, M = N * hw, K = C * k * k;
  if (const int kind = (dw && M > 0) ? conv_thin_kind(g) : 0) {
    const int    WIN = kind == 31 ? 9 : (kind == 33 ? 27 : 25), slot = kind == 31 ? 0 : (kind == 33 ? 1 : 2);
    const size_t smem = conv_thin_bw_smem(g, WIN);
    if (smem <= 200 * 1024) {
      auto fn = kind == 31 ? k_conv_thin_bwd_weight<3, 1> : (kind == 33 ? k_conv_thin_bwd_weight<3, 3> : k_conv_thin_bwd_weight<5, 1>);
      static size_t smem_set[3] = {0, 0, 0};
      if (smem > smem_set[slot]) {
        PDN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[slot] = smem;
      }
      PDN_CUDA(cudaMemsetAsync(dw, 0, (size_t)O * K * sizeof(float), stream()));
      if (dbias) PDN_CUDA(cudaMemsetAsync(dbias, 0, (size_t)O * sizeof(float), stream()));
      const int64_t ctas = N < 2 * sm_count() ? N : 2 * sm_count();
      fn<<<(unsigned)ctas, 256, smem, stream()>>>(x, gy, dw, dbias, g);
      PDN_LAUNCHED("conv_thin_bwd_weight");
      return 0;
    }
  }
  if (dbias && O > 0) {
    PDN_CUDA(cudaMemsetAsync(dbias, 0, (size_t)O * sizeof(float), stream()));
    int64_t chunks = (sm_count() * 8 + O - 1) / O;
    if (chunks > N) chunks = N;
    if (chunks < 1) chunks = 1;
    k_channel_sum<<<dim3((unsigned)O, (unsigned)chunks), 256, 0, stream()>>>(gy, dbias, N, O, hw);
    PDN_LAUNCHED("channel_sum");
  }
  if (!dw || O == 0 || K == 0) return 0;
  if (M == 0) {
    PDN_CUDA(cudaMemsetAsync(dw, 0, (size_t)O * K * sizeof(float), stream()));
    return 0;
  }
  if (conv_tma_ok(16, stride, N, C, H, W) && C >= 8 && O >= 8 && N * O * g.oh < 0x7fffffff)
    return conv_tma_bwd_weight(x, gy, dw, N, C, H, W, O, g.oh, g.ow, k, pad, x_version, gy_version);
  Scratch       bufA, bufB;
  PackedOperand A, B;
  // A rows = o, contraction index m = (n, pix): g[n, o, pix]
  const int64_t one[3] = {1, 1, 1}, zero[3] = {0, 0, 0};
  PDN_TRY(pack_operand_ex(gy, O, M, hw, 1, hw, O * hw, one, zero, &bufA, &A));
  PDN_TRY((conv_pack<0, false>(x, g, M, K, &bufB, &B)));
  TcArgs t;
  tc_defaults(t);
  t.C = dw; t.M = O; t.N = K; t.K = M; t.ldc = K;
  return gemm_tc_packed(A, B, t, 0);
}

int pdn_pool2d_fwd(const float* x, float* y, int64_t N, int64_t C, int64_t H, int64_t W, int k, int stride, int pad, int mode) {
  PDN_TRY(ensure_init());
  PDN_CHECK(k >= 1 && stride >= 1 && pad >= 0 && (mode == 0 || mode == 1), "pool: bad arguments");
  PoolGeom g{N, C, H, W, (H + 2 * pad - k) / stride + 1, (W + 2 * pad - k) / stride + 1, k, stride, pad, mode};
  const int64_t total = N * C * g.oh * g.ow;
  if (total == 0) return 0;
  PDN_CHECK(N * C * H * W < (int64_t)1 << 31, "pool2d_fwd: tensors of 2^31 elements or more are not supported");
  k_pool_fwd<<<grid_for(total, 256), 256, 0, stream()>>>(x, y, g);
  PDN_LAUNCHED("pool_fwd");
  return 0;
}

int pdn_pool2d_bwd(const float* x, const float* y, const float* gy, float* dx, int64_t N, int64_t C, int64_t H, int64_t W, int k, int stride,
                   int pad, int mode) {
  PDN_TRY(ensure_init());
  PDN_CHECK(k >= 1 && stride >= 1 && pad >= 0 && (mode == 0 || mode == 1), "pool: bad arguments");
  PoolGeom g{N, C, H, W, (H + 2 * pad - k) / stride + 1, (W + 2 * pad - k) / stride + 1, k, stride, pad, mode};
  const int64_t total = N * C * H * W;
  if (total == 0) return 0;
  PDN_CHECK(total < (int64_t)1 << 31, "pool2d_bwd: tensors of 2^31 elements or more are not supported");
  k_pool_bwd<<<grid_for(total, 256), 256, 0, stream()>>>(x, y, gy, dx, g);
  PDN_LAUNCHED("pool_bwd");
  return 0;
}

}  // extern "C"
