// gemm.cu — pdn_gemm dispatcher + the fp32-FFMA tiled kernel and the skinny (M<=16) kernel.
// The tensor-core path (tcgen05/TMEM/TMA, BF16x3 operand split) lives in gemm_tc.cu.
// Replaces `x.data @ y.data` and the swapaxes-view grads of the reference's matmul operator
// (reference pydynet/core/tensor.py:657-676).
#include "common.cuh"
#include "gemm_args.h"
#include <stdlib.h>
#include <string.h>

namespace pdn {



static int g_last_path = 0;

__device__ __forceinline__ void batch_offsets(const GemmArgs& g, int64_t z, int64_t& oa, int64_t& ob, int64_t& oc) {
  int64_t i2 = z % g.nb[2]; z /= g.nb[2];
  int64_t i1 = z % g.nb[1]; z /= g.nb[1];
  int64_t i0 = z;
  oa = i0 * g.a_bs[0] + i1 * g.a_bs[1] + i2 * g.a_bs[2];
  ob = i0 * g.b_bs[0] + i1 * g.b_bs[1] + i2 * g.b_bs[2];
  oc = i0 * g.c_bs[0] + i1 * g.c_bs[1] + i2 * g.c_bs[2];
}

// ---- tiled FFMA kernel: 64x64 tile, BK=16, 256 threads, 4x4 micro-tile, any strides, any dtype ---
template <typename T>
__global__ void __launch_bounds__(256) k_gemm_tiled(GemmArgs g) {
  using A = typename Acc<T>::type;
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ A As[BK][BM + 4];
  __shared__ A Bs[BK][BN + 4];
  int64_t oa, ob, oc;
  batch_offsets(g, blockIdx.z, oa, ob, oc);
  const T* Ap = (const T*)g.A + oa;
  const T* Bp = (const T*)g.B + ob;
  T*       Cp = (T*)g.C + oc;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  // A tile loader mapping: unit stride along k -> k fastest, else m fastest
  const bool a_kfast = (g.a_cs == 1);
  const bool b_nfast = (g.b_cs == 1);
  A acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (A)0;

  for (int64_t k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {  // 64*16 = 1024 elements / 256 threads
      int l = tid + e * 256;
      int mm, kk;
      if (a_kfast) { kk = l % BK; mm = l / BK; } else { mm = l % BM; kk = l / BM; }
      int64_t gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < g.M && gk < g.K) ? ld<T>(Ap + gm * g.a_rs + gk * g.a_cs) : (A)0;
      int nn, kb;
      if (b_nfast) { nn = l % BN; kb = l / BN; } else { kb = l % BK; nn = l / BK; }
      int64_t gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < g.N && gkb < g.K) ? ld<T>(Bp + gkb * g.b_rs + gn * g.b_cs) : (A)0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      A a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      A v = acc[i][j];
      if (g.bias) v += ld<T>((const T*)g.bias + gn);
      T* c = Cp + gm * g.ldc + gn;
      if (g.accumulate) v += ld<T>(c);
      st<T>(c, v);
    }
  }
}

// ---- skinny fp32 kernel: M <= 16, B row-major with unit column stride (F.linear weights (in,out)) --
// block = 32 column-quads (128 columns) x 8 k-slices; HBM/L2-bound GEMV-like path of Llama decode.
template <int MT>
__global__ void __launch_bounds__(256) k_gemm_skinny_f32(GemmArgs g) {
  constexpr int KC = 256;  // k-chunk staged in smem
  __shared__ float xs[MT][KC];
  __shared__ float red[8][132];
  int64_t oa, ob, oc;
  batch_offsets(g, blockIdx.z, oa, ob, oc);
  const float* Ap = (const float*)g.A + oa;
  const float* Bp = (const float*)g.B + ob;
  float*       Cp = (float*)g.C + oc;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * 128 + lane * 4;
  const bool vec_ok = (n + 3 < g.N) && ((g.b_rs & 3) == 0) && ((((uintptr_t)Bp) & 15) == 0);
  float acc[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
  for (int64_t k0 = 0; k0 < g.K; k0 += KC) {
    int kc = (int)((g.K - k0) < KC ? (g.K - k0) : KC);
    __syncthreads();
    for (int l = threadIdx.x; l < MT * KC; l += 256) {
      int m = l / KC, kk = l % KC;
      xs[m][kk] = (m < g.M && kk < kc) ? Ap[m * g.a_rs + (k0 + kk) * g.a_cs] : 0.f;
    }
    __syncthreads();
    for (int kk = slice; kk < kc; kk += 8) {
      float4 w;
      const float* wp = Bp + (k0 + kk) * g.b_rs + n;
      if (vec_ok) w = __ldg((const float4*)wp);
      else {
        w.x = (n + 0 < g.N) ? wp[0] : 0.f; w.y = (n + 1 < g.N) ? wp[1] : 0.f;
        w.z = (n + 2 < g.N) ? wp[2] : 0.f; w.w = (n + 3 < g.N) ? wp[3] : 0.f;
      }
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float x = xs[m][kk];
        acc[m][0] += x * w.x; acc[m][1] += x * w.y; acc[m][2] += x * w.z; acc[m][3] += x * w.w;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[slice][lane * 4 + j] = acc[m][j];
    __syncthreads();
    if (threadIdx.x < 128 && m < g.M) {
      int     c = threadIdx.x;
      int64_t gn = (int64_t)blockIdx.x * 128 + c;
      if (gn < g.N) {
        float v = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) v += red[s][c];
        if (g.bias) v += ((const float*)g.bias)[gn];
        float* cp = Cp + m * g.ldc + gn;
        if (g.accumulate) v += *cp;
        *cp = v;
      }
    }
  }
}

// ---- small products the tensor-core path does not take (N or K below 32: a classifier head with 10 classes and its two gradient
// products, LeNet fc2): the 64x64-tile kernel runs them on 4-8 CTAs with an un-pipelined K loop (38 us each). Here one WARP owns an
// output element when K >= 32 (lanes stride over k, xor-shuffle tree) and one THREAD when K is short; any strides, fp32, fixed
// summation order.
template <bool WARP>
__global__ void __launch_bounds__(256) k_gemm_small_f32(GemmArgs g) {
  int64_t oa, ob, oc;
  batch_offsets(g, blockIdx.z, oa, ob, oc);
  const float* Ap = (const float*)g.A + oa;
  const float* Bp = (const float*)g.B + ob;
  float*       Cp = (float*)g.C + oc;
  const int    M = (int)g.M, N = (int)g.N, K = (int)g.K;
  const int    lane = threadIdx.x & 31;
  const int    per_block = WARP ? 8 : 256;
  for (int o = blockIdx.x * per_block + (WARP ? (int)(threadIdx.x >> 5) : (int)threadIdx.x); o < M * N; o += gridDim.x * per_block) {
    const int    m = o / N, n = o - m * N;
    const float *ap = Ap + (int64_t)m * g.a_rs, *bp = Bp + (int64_t)n * g.b_cs;
    float        acc0 = 0.f, acc1 = 0.f;
    if (WARP) {
      int k = lane;
      for (; k + 32 < K; k += 64) {
        acc0 = fmaf(__ldg(ap + (int64_t)k * g.a_cs), __ldg(bp + (int64_t)k * g.b_rs), acc0);
        acc1 = fmaf(__ldg(ap + (int64_t)(k + 32) * g.a_cs), __ldg(bp + (int64_t)(k + 32) * g.b_rs), acc1);
      }
      if (k < K) acc0 = fmaf(__ldg(ap + (int64_t)k * g.a_cs), __ldg(bp + (int64_t)k * g.b_rs), acc0);
      acc0 = warp_sum(acc0 + acc1);
      if (lane != 0) continue;
    } else {
      for (int k = 0; k < K; ++k) acc0 = fmaf(__ldg(ap + (int64_t)k * g.a_cs), __ldg(bp + (int64_t)k * g.b_rs), acc0);
    }
    if (g.bias) acc0 += ((const float*)g.bias)[n];
    float* cp = Cp + (int64_t)m * g.ldc + n;
    if (g.accumulate) acc0 += *cp;
    *cp = acc0;
  }
}

}  // namespace pdn

using namespace pdn;

extern "C" {

int pdn_gemm_last_path(void) { return g_last_path; }

int pdn_gemm(int dtype, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t a_rs, int64_t a_cs,
             int64_t b_rs, int64_t b_cs, int64_t ldc, const int64_t* nb, const int64_t* a_bs, const int64_t* b_bs,
             const int64_t* c_bs, const void* bias, int accumulate, int prec) {
  return pdn_gemm_cached(dtype, A, B, C, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, nb, a_bs, b_bs, c_bs, bias, accumulate, prec, -1, -1);
}

int pdn_plane_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* bytes, uint64_t* entries) {
  plane_cache_stats(hits, misses, bytes, entries);
  return 0;
}

int pdn_gemm_cached(int dtype, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t a_rs, int64_t a_cs,
                    int64_t b_rs, int64_t b_cs, int64_t ldc, const int64_t* nb, const int64_t* a_bs, const int64_t* b_bs,
                    const int64_t* c_bs, const void* bias, int accumulate, int prec, int64_t a_version, int64_t b_version) {
  PDN_TRY(ensure_init());
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.bias = bias;
  g.M = M; g.N = N; g.K = K;
  g.a_rs = a_rs; g.a_cs = a_cs; g.b_rs = b_rs; g.b_cs = b_cs; g.ldc = ldc;
  int64_t nbatch = 1;
  for (int i = 0; i < 3; ++i) {
    g.nb[i] = nb ? nb[i] : 1;
    g.a_bs[i] = a_bs ? a_bs[i] : 0;
    g.b_bs[i] = b_bs ? b_bs[i] : 0;
    g.c_bs[i] = c_bs ? c_bs[i] : 0;
    nbatch *= g.nb[i];
  }
  g.accumulate = accumulate;
  if (M == 0 || N == 0 || nbatch == 0) return 0;
  PDN_CHECK(nbatch <= 65535 * 64, "too many GEMM batches");
  if (K == 0 && !accumulate) {
    // x @ y with an empty contraction is zeros (+ bias)
    K = 0;
  }
  static int env_mode = -1;
  if (env_mode < 0) {
    const char* e = getenv("PDN_GEMM");
    env_mode = (e && !strcmp(e, "simt")) ? 1 : ((e && !strcmp(e, "tc")) ? 2 : 0);
  }
  if (prec == 0 && env_mode == 1) prec = 1;
  if (dtype == PDN_F32 && prec != 1 && gemm_tc_eligible(g)) {
    g_last_path = 1;
    return gemm_tc_launch(g, a_version, b_version);
  }
  PDN_CHECK(prec != 2, "pdn_gemm: tcgen05 path forced but shape/strides are not eligible");
  if (dtype == PDN_F32 && M <= 16 && b_cs == 1 && K >= 32 && nbatch <= 65535) {
    dim3 grd((unsigned)((N + 127) / 128), 1, (unsigned)nbatch);
    if (M <= 1) k_gemm_skinny_f32<1><<<grd, 256, 0, stream()>>>(g);
    else if (M <= 4) k_gemm_skinny_f32<4><<<grd, 256, 0, stream()>>>(g);
    else if (M <= 8) k_gemm_skinny_f32<8><<<grd, 256, 0, stream()>>>(g);
    else k_gemm_skinny_f32<16><<<grd, 256, 0, stream()>>>(g);
    PDN_LAUNCHED("gemm_skinny_f32");
    g_last_path = 2;
    return 0;
  }
  PDN_CHECK(nbatch <= 65535, "too many GEMM batches for the tiled kernel (%lld)", (long long)nbatch);
  if (dtype == PDN_F32 && (N < 32 || K < 32) && M * N <= (1 << 22) && (double)M * N * K <= (double)(1 << 26)) {
    const bool    warp = K >= 32;
    const int64_t blocks = warp ? (M * N + 7) / 8 : (M * N + 255) / 256;
    dim3          grd((unsigned)(blocks < 8192 ? blocks : 8192), 1, (unsigned)nbatch);
    if (warp) k_gemm_small_f32<true><<<grd, 256, 0, stream()>>>(g);
    else k_gemm_small_f32<false><<<grd, 256, 0, stream()>>>(g);
    PDN_LAUNCHED("gemm_small_f32");
    g_last_path = 3;
    return 0;
  }
  dim3 grd((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64), (unsigned)nbatch);
  PDN_CHECK((M + 63) / 64 <= 65535, "M too large for the tiled kernel grid");
  switch (dtype) {
    case PDN_F32: k_gemm_tiled<float><<<grd, 256, 0, stream()>>>(g); break;
    case PDN_F64: k_gemm_tiled<double><<<grd, 256, 0, stream()>>>(g); break;
    case PDN_F16: k_gemm_tiled<__half><<<grd, 256, 0, stream()>>>(g); break;
    default: set_error("gemm: unsupported dtype %d", dtype); return PDN_ERR_UNSUPPORTED;
  }
  PDN_LAUNCHED("gemm_tiled");
  g_last_path = 0;
  return 0;
}

}  // extern "C"
