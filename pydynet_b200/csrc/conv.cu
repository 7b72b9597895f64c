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
