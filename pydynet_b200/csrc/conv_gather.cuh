// conv_gather.cuh — convolution index arithmetic shared by the operand-pack kernels (conv.cu) and the implicit-GEMM
// producer warps of the tcgen05 GEMM (gemm_tc.cu): which source element feeds row m / column kk of the im2col matrix.
#pragma once
#include "common.cuh"

namespace pdn {

struct ConvGeom {
  int64_t N, C, H, W, O, oh, ow;
  int     k, stride, pad;
};

// MODE 0: im2col of x — row m = (n, oy, ox), column kk = (c, ky, kx)
// MODE 1: transposed-conv gather of g — row m = (n, y, x) of the INPUT grid, column kk = (o, ky, kx)
// The row index is decomposed once per thread (it is fixed across the tile's columns); columns cost three small
// 32-bit divisions each.
struct RowPos { int64_t base; int y, x; bool ok; };  // base = element offset of (n, channel 0, 0, 0) in the source

template <int MODE>
__device__ __forceinline__ RowPos conv_row(const ConvGeom& g, int64_t m, int64_t Mtot) {
  RowPos r;
  r.ok = m < Mtot;
  if (!r.ok) { r.base = 0; r.y = r.x = 0; return r; }
  if (MODE == 0) {
    const int64_t hw = g.oh * g.ow, n = m / hw;
    const int pix = (int)(m - n * hw);
    r.y = (pix / (int)g.ow) * g.stride - g.pad;   // top-left input coordinate of the window
    r.x = (pix % (int)g.ow) * g.stride - g.pad;
    r.base = n * g.C * g.H * g.W;
  } else {
    const int64_t hw = g.H * g.W, n = m / hw;
    const int pix = (int)(m - n * hw);
    r.y = pix / (int)g.W + g.pad;
    r.x = pix % (int)g.W + g.pad;
    r.base = n * g.O * g.oh * g.ow;
  }
  return r;
}

template <int MODE>
__device__ __forceinline__ float conv_fetch(const float* __restrict__ src, const ConvGeom& g, const RowPos& r, int kk, int Ktot) {
  if (!r.ok || kk >= Ktot) return 0.f;
  const int kx = kk % g.k, t = kk / g.k, ky = t % g.k, ch = t / g.k;
  if (MODE == 0) {
    const int iy = r.y + ky, ix = r.x + kx;
    if (iy < 0 || iy >= (int)g.H || ix < 0 || ix >= (int)g.W) return 0.f;
    return __ldg(src + r.base + ((int64_t)ch * g.H + iy) * g.W + ix);
  } else {
    const int ty = r.y - ky, tx = r.x - kx;
    if (ty < 0 || tx < 0) return 0.f;
    int oy = ty, ox = tx;
    if (g.stride != 1) {
      if (ty % g.stride || tx % g.stride) return 0.f;
      oy = ty / g.stride; ox = tx / g.stride;
    }
    if (oy >= (int)g.oh || ox >= (int)g.ow) return 0.f;
    return __ldg(src + r.base + ((int64_t)ch * g.oh + oy) * g.ow + ox);
  }
}


}  // namespace pdn
