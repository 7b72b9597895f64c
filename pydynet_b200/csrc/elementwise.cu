// elementwise.cu — broadcast-aware, stride-aware elementwise kernels with a 128-bit vectorised dense path.
// One HBM pass per operator node; replaces the reference's per-node NumPy/CuPy expressions
// (reference pydynet/core/tensor.py:535-641, 679-692, 776-832, 996-1019).
#include "common.cuh"

namespace pdn {

// ------------------------------------------------------------------ functors ------------------
template <typename A> __device__ __forceinline__ A pow_(A x, A y);
template <> __device__ __forceinline__ float pow_<float>(float x, float y) {
  if (y == 2.0f) return x * x;
  if (y == 0.5f) return sqrtf(x);
  return powf(x, y);
}
template <> __device__ __forceinline__ double pow_<double>(double x, double y) {
  if (y == 2.0) return x * x;
  if (y == 0.5) return sqrt(x);
  return pow(x, y);
}
template <> __device__ __forceinline__ long long pow_<long long>(long long x, long long y) {
  long long r = 1;
  for (long long i = 0; i < y; ++i) r *= x;
  return r;
}
template <> __device__ __forceinline__ int pow_<int>(int x, int y) {
  int r = 1;
  for (int i = 0; i < y; ++i) r *= x;
  return r;
}

template <typename A> __device__ __forceinline__ A div_(A x, A y) { return x / y; }
template <> __device__ __forceinline__ long long div_<long long>(long long x, long long y) { return y ? x / y : 0; }
template <> __device__ __forceinline__ int div_<int>(int x, int y) { return y ? x / y : 0; }

// NumPy maximum/minimum propagate NaN
template <typename A> __device__ __forceinline__ A max_(A x, A y) { return (x != x) ? x : ((y != y) ? y : (x > y ? x : y)); }
template <typename A> __device__ __forceinline__ A min_(A x, A y) { return (x != x) ? x : ((y != y) ? y : (x < y ? x : y)); }

template <typename A>
__device__ __forceinline__ A binary_arith(int op, A x, A y) {
  switch (op) {
    case PDN_ADD: return x + y;
    case PDN_SUB: return x - y;
    case PDN_MUL: return x * y;
    case PDN_DIV: return div_<A>(x, y);
    case PDN_POW: return pow_<A>(x, y);
    case PDN_MAXIMUM: return max_<A>(x, y);
    case PDN_MINIMUM: return min_<A>(x, y);
  }
  return x;
}
template <typename A>
__device__ __forceinline__ int binary_cmp(int op, A x, A y) {
  switch (op) {
    case PDN_EQ: return x == y;
    case PDN_NE: return x != y;
    case PDN_LT: return x < y;
    case PDN_LE: return x <= y;
    case PDN_GT: return x > y;
    case PDN_GE: return x >= y;
  }
  return 0;
}

// exp/log helpers per compute type
__device__ __forceinline__ float  exp_(float x) { return expf(x); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float  log_(float x) { return logf(x); }
__device__ __forceinline__ double log_(double x) { return log(x); }
__device__ __forceinline__ float  sqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }

// piecewise overflow-safe forms, reference tensor.py:999-1002 and :1012-1015
template <typename A> __device__ __forceinline__ A sigmoid_(A x) {
  return x > (A)0 ? (A)1 / ((A)1 + exp_(-x)) : (A)1 - (A)1 / ((A)1 + exp_(x));
}
template <typename A> __device__ __forceinline__ A tanh_(A x) {
  return x > (A)0 ? (A)2 / ((A)1 + exp_((A)-2 * x)) - (A)1 : (A)1 - (A)2 / ((A)1 + exp_((A)2 * x));
}

template <typename A>
__device__ __forceinline__ A unary_float(int op, A x) {
  switch (op) {
    case PDN_NEG: return -x;
    case PDN_EXP: return exp_(x);
    case PDN_LOG: return log_(x);
    case PDN_ABS: return x < (A)0 ? -x : x;
    case PDN_SIGN: return x > (A)0 ? (A)1 : (x < (A)0 ? (A)-1 : x);  // NaN stays NaN, 0 stays 0
    case PDN_SIGMOID: return sigmoid_<A>(x);
    case PDN_TANH: return tanh_<A>(x);
    case PDN_SQRT: return sqrt_(x);
    case PDN_SQUARE: return x * x;
    case PDN_RECIP: return (A)1 / x;
    case PDN_SILU: return x / ((A)1 + exp_(-x));  // functional.py:39-40
    case PDN_RELU: return max_<A>((A)0, x);       // functional.py:31-32 maximum(0., x)
  }
  return x;
}
template <typename A>
__device__ __forceinline__ A unary_int(int op, A x) {
  switch (op) {
    case PDN_NEG: return -x;
    case PDN_ABS: return x < 0 ? -x : x;
    case PDN_SIGN: return x > 0 ? 1 : (x < 0 ? -1 : 0);
    case PDN_SQUARE: return x * x;
    case PDN_RELU: return x > 0 ? x : 0;
  }
  return x;
}

template <typename A>
__device__ __forceinline__ A ternary_float(int op, A a, A b, A c) {
  switch (op) {
    case PDN_T_EQ_MUL: return (a == b) ? c : (A)0 * c;
    case PDN_T_DIV_GRAD_Y: return -a * b / c;
    case PDN_T_POW_GRAD_X: return a * b / c;
    case PDN_T_SIGMOID_GRAD: return a * ((A)1 - a) * b;
    case PDN_T_TANH_GRAD: return ((A)1 - a * a) * b;
    case PDN_T_FMA: return a * b + c;
    case PDN_T_SILU_GRAD: {
      A s = (A)1 / ((A)1 + exp_(-a));
      return (s + a * s * ((A)1 - s)) * b;
    }
    case PDN_T_WHERE: return a != (A)0 ? b : c;
  }
  return a;
}

// ------------------------------------------------------------------ kernels -------------------
template <typename T, typename TO, bool CMP>
__global__ void __launch_bounds__(256) k_binary_strided(int op, const T* a, const T* b, TO* out, StridedDesc d) {
  using A = typename Acc<T>::type;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t off[3];
    decompose<3>(d, i, off);
    A x = ld<T>(a + off[0]), y = ld<T>(b + off[1]);
    if (CMP) st<TO>(out + off[2], (typename Acc<TO>::type)binary_cmp<A>(op, x, y));
    else st<TO>(out + off[2], (typename Acc<TO>::type)binary_arith<A>(op, x, y));
  }
}

// dense fp32 path: 4 floats per thread per trip, 128-bit loads/stores
__global__ void __launch_bounds__(256) k_binary_dense_f32(int op, const float4* __restrict__ a, const float4* __restrict__ b,
                                                         float4* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = a[i], y = b[i], r;
    r.x = binary_arith<float>(op, x.x, y.x);
    r.y = binary_arith<float>(op, x.y, y.y);
    r.z = binary_arith<float>(op, x.z, y.z);
    r.w = binary_arith<float>(op, x.w, y.w);
    out[i] = r;
  }
}
// dense fp32 operands where any input may instead be ONE broadcast element (a 0-d / size-1 device array — what the reference's
// scalar wrapping produces, tensor.py:488-493: relu = maximum(Tensor(0.), x)); avoids the generic strided walk
__global__ void __launch_bounds__(256) k_binary_mixed_f32(int op, const float* __restrict__ a, const float* __restrict__ b, float4* __restrict__ out,
                                                         int64_t n4, int a_one, int b_one) {
  const float a0 = a_one ? __ldg(a) : 0.f, b0 = b_one ? __ldg(b) : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = a_one ? make_float4(a0, a0, a0, a0) : __ldg(reinterpret_cast<const float4*>(a) + i);
    const float4 y = b_one ? make_float4(b0, b0, b0, b0) : __ldg(reinterpret_cast<const float4*>(b) + i);
    float4 r;
    r.x = binary_arith<float>(op, x.x, y.x);
    r.y = binary_arith<float>(op, x.y, y.y);
    r.z = binary_arith<float>(op, x.z, y.z);
    r.w = binary_arith<float>(op, x.w, y.w);
    out[i] = r;
  }
}
__global__ void __launch_bounds__(256) k_ternary_mixed_f32(int op, const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                                          float4* __restrict__ out, int64_t n4, int a_one, int b_one, int c_one) {
  const float a0 = a_one ? __ldg(a) : 0.f, b0 = b_one ? __ldg(b) : 0.f, c0 = c_one ? __ldg(c) : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = a_one ? make_float4(a0, a0, a0, a0) : __ldg(reinterpret_cast<const float4*>(a) + i);
    const float4 y = b_one ? make_float4(b0, b0, b0, b0) : __ldg(reinterpret_cast<const float4*>(b) + i);
    const float4 z = c_one ? make_float4(c0, c0, c0, c0) : __ldg(reinterpret_cast<const float4*>(c) + i);
    float4 r;
    r.x = ternary_float<float>(op, x.x, y.x, z.x);
    r.y = ternary_float<float>(op, x.y, y.y, z.y);
    r.z = ternary_float<float>(op, x.z, y.z, z.z);
    r.w = ternary_float<float>(op, x.w, y.w, z.w);
    out[i] = r;
  }
}
// dense rows + broadcast vector over the last dim (bias add, scale mul): a[r, c] op b[c]
__global__ void __launch_bounds__(256) k_binary_rowvec_f32(int op, const float4* __restrict__ a, const float4* __restrict__ b,
                                                          float4* __restrict__ out, int64_t n4, int cols4, int b_left) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = a[i], y = __ldg(&b[i % cols4]), r;
    if (b_left) { float4 t = x; x = y; y = t; }
    r.x = binary_arith<float>(op, x.x, y.x);
    r.y = binary_arith<float>(op, x.y, y.y);
    r.z = binary_arith<float>(op, x.z, y.z);
    r.w = binary_arith<float>(op, x.w, y.w);
    out[i] = r;
  }
}

// dense rows + one broadcast value per row: a[r, c] op b[r] (padding masks, per-row scales: F.embedding's pad mask functional.py:17-19)
__global__ void __launch_bounds__(256) k_binary_colvec_f32(int op, const float4* __restrict__ a, const float* __restrict__ b,
                                                          float4* __restrict__ out, int64_t n4, int cols4, int64_t b_stride, int b_left) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float  s = __ldg(b + (i / cols4) * b_stride);
    float4 x = a[i], y = make_float4(s, s, s, s), r;
    if (b_left) { float4 t = x; x = y; y = t; }
    r.x = binary_arith<float>(op, x.x, y.x);
    r.y = binary_arith<float>(op, x.y, y.y);
    r.z = binary_arith<float>(op, x.z, y.z);
    r.w = binary_arith<float>(op, x.w, y.w);
    out[i] = r;
  }
}

template <typename T, typename TO, bool CMP>
__global__ void __launch_bounds__(256) k_binary_scalar(int op, const T* a, double scalar, int reverse, TO* out, StridedDesc d) {
  using A = typename Acc<T>::type;
  A s = (A)scalar;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t off[2];
    decompose<2>(d, i, off);
    A x = ld<T>(a + off[0]);
    A l = reverse ? s : x, r = reverse ? x : s;
    if (CMP) st<TO>(out + off[1], (typename Acc<TO>::type)binary_cmp<A>(op, l, r));
    else st<TO>(out + off[1], (typename Acc<TO>::type)binary_arith<A>(op, l, r));
  }
}
__global__ void __launch_bounds__(256) k_binary_scalar_dense_f32(int op, const float4* __restrict__ a, float s, int reverse,
                                                                float4* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = a[i], r;
    if (reverse) {
      r.x = binary_arith<float>(op, s, x.x); r.y = binary_arith<float>(op, s, x.y);
      r.z = binary_arith<float>(op, s, x.z); r.w = binary_arith<float>(op, s, x.w);
    } else {
      r.x = binary_arith<float>(op, x.x, s); r.y = binary_arith<float>(op, x.y, s);
      r.z = binary_arith<float>(op, x.z, s); r.w = binary_arith<float>(op, x.w, s);
    }
    out[i] = r;
  }
}

template <typename T, bool FLOATING>
__global__ void __launch_bounds__(256) k_unary_strided(int op, const T* a, T* out, StridedDesc d) {
  using A = typename Acc<T>::type;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t off[2];
    decompose<2>(d, i, off);
    A x = ld<T>(a + off[0]);
    if constexpr (FLOATING) st<T>(out + off[1], unary_float<A>(op, x));
    else st<T>(out + off[1], unary_int<A>(op, x));
  }
}
__global__ void __launch_bounds__(256) k_unary_dense_f32(int op, const float4* __restrict__ a, float4* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = a[i], r;
    r.x = unary_float<float>(op, x.x); r.y = unary_float<float>(op, x.y);
    r.z = unary_float<float>(op, x.z); r.w = unary_float<float>(op, x.w);
    out[i] = r;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_ternary_strided(int op, const T* a, const T* b, const T* c, T* out, StridedDesc d) {
  using A = typename Acc<T>::type;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t off[4];
    decompose<4>(d, i, off);
    st<T>(out + off[3], ternary_float<A>(op, ld<T>(a + off[0]), ld<T>(b + off[1]), ld<T>(c + off[2])));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_fill(T* out, StridedDesc d, double v) {
  using A = typename Acc<T>::type;
  A val = (A)v;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t off[1];
    decompose<1>(d, i, off);
    st<T>(out + off[0], val);
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) k_copy(const TS* src, TD* dst, StridedDesc d) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t off[2];
    decompose<2>(d, i, off);
    st<TD>(dst + off[1], (typename Acc<TD>::type)ld<TS>(src + off[0]));
  }
}
// bool destination: nonzero -> 1
template <typename TS>
__global__ void __launch_bounds__(256) k_copy_to_bool(const TS* src, bool* dst, StridedDesc d) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t off[2];
    decompose<2>(d, i, off);
    dst[off[1]] = ld<TS>(src + off[0]) != (typename Acc<TS>::type)0;
  }
}
// 2-D transposing copy through shared memory (both sides coalesced): dst[c, r] = src[r, c]
__global__ void __launch_bounds__(256) k_copy_f32_dense(const float4* __restrict__ s, float4* __restrict__ d, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) d[i] = s[i];
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace pdn

using namespace pdn;

#define DISPATCH_ALL(dt, ...)                                                       \
  switch (dt) {                                                                     \
    case PDN_F32: { using T = float; constexpr bool FL = true; (void)FL; __VA_ARGS__; break; }        \
    case PDN_F64: { using T = double; constexpr bool FL = true; (void)FL; __VA_ARGS__; break; }       \
    case PDN_F16: { using T = __half; constexpr bool FL = true; (void)FL; __VA_ARGS__; break; }       \
    case PDN_I64: { using T = long long; constexpr bool FL = false; (void)FL; __VA_ARGS__; break; }   \
    case PDN_I32: { using T = int; constexpr bool FL = false; (void)FL; __VA_ARGS__; break; }         \
    case PDN_BOOL: { using T = bool; constexpr bool FL = false; (void)FL; __VA_ARGS__; break; }       \
    default: pdn::set_error("unsupported dtype %d", dt); return PDN_ERR_UNSUPPORTED;                 \
  }

extern "C" {

int pdn_fill(void* out, int dtype, int ndim, const int64_t* shape, const int64_t* so, double value) {
  PDN_TRY(ensure_init());
  StridedDesc    d;
  const int64_t* st[1] = {so};
  PDN_TRY(make_desc(ndim, shape, 1, st, &d));
  if (d.n == 0) return 0;
  if (desc_dense(d, 0) && value == 0.0) {  // xp.zeros
    PDN_CUDA(cudaMemsetAsync(out, 0, (size_t)d.n * dtype_size(dtype), stream()));
    return 0;
  }
  DISPATCH_ALL(dtype, (k_fill<T><<<grid_for(d.n, 256, 4), 256, 0, stream()>>>((T*)out, d, value)));
  PDN_LAUNCHED("fill");
  return 0;
}

}  // extern "C"

template <typename TS>
static int copy_from(const void* src, void* dst, int ddtype, const StridedDesc& d) {
  int g = grid_for(d.n, 256, 4);
  switch (ddtype) {
    case PDN_F32: k_copy<TS, float><<<g, 256, 0, stream()>>>((const TS*)src, (float*)dst, d); break;
    case PDN_F64: k_copy<TS, double><<<g, 256, 0, stream()>>>((const TS*)src, (double*)dst, d); break;
    case PDN_F16: k_copy<TS, __half><<<g, 256, 0, stream()>>>((const TS*)src, (__half*)dst, d); break;
    case PDN_I64: k_copy<TS, long long><<<g, 256, 0, stream()>>>((const TS*)src, (long long*)dst, d); break;
    case PDN_I32: k_copy<TS, int><<<g, 256, 0, stream()>>>((const TS*)src, (int*)dst, d); break;
    case PDN_BOOL: k_copy_to_bool<TS><<<g, 256, 0, stream()>>>((const TS*)src, (bool*)dst, d); break;
    default: set_error("unsupported destination dtype %d", ddtype); return PDN_ERR_UNSUPPORTED;
  }
  PDN_LAUNCHED("copy");
  return 0;
}

extern "C" {

// dst[i] = *src : a broadcast scalar materialised (the upstream gradient of `.sum()`: ones expanded to the operand's shape)
__global__ void __launch_bounds__(256) k_bcast_scalar_f32(const float* __restrict__ s, float4* __restrict__ d, int64_t n4) {
  const float  v = __ldg(s);
  const float4 v4 = make_float4(v, v, v, v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) d[i] = v4;
}

int pdn_copy(const void* src, int sdtype, void* dst, int ddtype, int ndim, const int64_t* shape, const int64_t* ss,
             const int64_t* ds) {
  PDN_TRY(ensure_init());
  StridedDesc    d;
  const int64_t* st[2] = {ss, ds};
  PDN_TRY(make_desc(ndim, shape, 2, st, &d));
  if (d.n == 0) return 0;
  if (sdtype == ddtype && desc_dense(d, 0) && desc_dense(d, 1)) {
    size_t bytes = (size_t)d.n * dtype_size(sdtype);
    if (aligned16(src) && aligned16(dst) && bytes % 16 == 0) {
      k_copy_f32_dense<<<grid_for(bytes / 16, 256, 4), 256, 0, stream()>>>((const float4*)src, (float4*)dst, bytes / 16);
      PDN_LAUNCHED("copy_dense");
    } else {
      PDN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream()));
    }
    return 0;
  }
  if (sdtype == PDN_F32 && ddtype == PDN_F32 && d.ndim == 1 && d.s[0][0] == 0 && desc_dense(d, 1) && aligned16(dst) && d.n % 4 == 0) {
    k_bcast_scalar_f32<<<grid_for(d.n / 4, 256, 4), 256, 0, stream()>>>((const float*)src, (float4*)dst, d.n / 4);
    PDN_LAUNCHED("bcast_scalar");
    return 0;
  }
  DISPATCH_ALL(sdtype, return copy_from<T>(src, dst, ddtype, d));
  return 0;
}

int pdn_ew_binary(int op, int dtype, const void* a, const void* b, void* out, int ndim, const int64_t* shape,
                  const int64_t* sa, const int64_t* sb, const int64_t* so) {
  PDN_TRY(ensure_init());
  StridedDesc    d;
  const int64_t* st[3] = {sa, sb, so};
  PDN_TRY(make_desc(ndim, shape, 3, st, &d));
  if (d.n == 0) return 0;
  bool cmp = op >= PDN_EQ;
  if (!cmp && dtype == PDN_F32 && aligned16(a) && aligned16(b) && aligned16(out)) {
    if (desc_dense(d, 0) && desc_dense(d, 1) && desc_dense(d, 2) && d.n % 4 == 0) {
      k_binary_dense_f32<<<grid_for(d.n / 4, 256, 2), 256, 0, stream()>>>(op, (const float4*)a, (const float4*)b, (float4*)out, d.n / 4);
      PDN_LAUNCHED("binary_dense_f32");
      return 0;
    }
    {  // every input dense or a single broadcast element, output dense
      auto one = [&](int o) { return d.ndim == 0 || (d.ndim == 1 && d.s[o][0] == 0); };
      if (desc_dense(d, 2) && d.n % 4 == 0 && d.ndim <= 1 && (desc_dense(d, 0) || one(0)) && (desc_dense(d, 1) || one(1))) {
        k_binary_mixed_f32<<<grid_for(d.n / 4, 256, 2), 256, 0, stream()>>>(op, (const float*)a, (const float*)b, (float4*)out, d.n / 4,
                                                                           one(0) && !desc_dense(d, 0), one(1) && !desc_dense(d, 1));
        PDN_LAUNCHED("binary_mixed_f32");
        return 0;
      }
    }
    // [rows, cols] op [cols] with dense rows (bias / scale vectors)
    if (d.ndim == 2 && d.s[2][1] == 1 && d.s[2][0] == d.shape[1] && d.shape[1] % 4 == 0) {
      bool a_full = d.s[0][1] == 1 && d.s[0][0] == d.shape[1], b_full = d.s[1][1] == 1 && d.s[1][0] == d.shape[1];
      bool a_vec = d.s[0][1] == 1 && d.s[0][0] == 0, b_vec = d.s[1][1] == 1 && d.s[1][0] == 0;
      if ((a_full && b_vec) || (a_vec && b_full)) {
        const void* full = a_full ? a : b;
        const void* vec = a_full ? b : a;
        k_binary_rowvec_f32<<<grid_for(d.n / 4, 256, 2), 256, 0, stream()>>>(op, (const float4*)full, (const float4*)vec, (float4*)out,
                                                                            d.n / 4, (int)(d.shape[1] / 4), a_full ? 0 : 1);
        PDN_LAUNCHED("binary_rowvec_f32");
        return 0;
      }
      // [rows, cols] op [rows, 1]
      bool a_col = d.s[0][1] == 0, b_col = d.s[1][1] == 0;
      if ((a_full && b_col) || (a_col && b_full)) {
        const void* full = a_full ? a : b;
        const void* vec = a_full ? b : a;
        const int64_t vs = a_full ? d.s[1][0] : d.s[0][0];
        k_binary_colvec_f32<<<grid_for(d.n / 4, 256, 2), 256, 0, stream()>>>(op, (const float4*)full, (const float*)vec, (float4*)out, d.n / 4,
                                                                            (int)(d.shape[1] / 4), vs, a_full ? 0 : 1);
        PDN_LAUNCHED("binary_colvec_f32");
        return 0;
      }
    }
  }
  int g = grid_for(d.n, 256, 2);
  if (cmp) {
    DISPATCH_ALL(dtype, (k_binary_strided<T, bool, true><<<g, 256, 0, stream()>>>(op, (const T*)a, (const T*)b, (bool*)out, d)));
  } else {
    PDN_CHECK(dtype != PDN_BOOL, "arithmetic on bool arrays must be promoted by the caller");
    DISPATCH_ALL(dtype, (k_binary_strided<T, T, false><<<g, 256, 0, stream()>>>(op, (const T*)a, (const T*)b, (T*)out, d)));
  }
  PDN_LAUNCHED("binary_strided");
  return 0;
}

int pdn_ew_binary_scalar(int op, int dtype, const void* a, double scalar, int reverse, void* out, int ndim,
                         const int64_t* shape, const int64_t* sa, const int64_t* so) {
  PDN_TRY(ensure_init());
  StridedDesc    d;
  const int64_t* st[2] = {sa, so};
  PDN_TRY(make_desc(ndim, shape, 2, st, &d));
  if (d.n == 0) return 0;
  bool cmp = op >= PDN_EQ;
  if (!cmp && dtype == PDN_F32 && desc_dense(d, 0) && desc_dense(d, 1) && d.n % 4 == 0 && aligned16(a) && aligned16(out)) {
    k_binary_scalar_dense_f32<<<grid_for(d.n / 4, 256, 2), 256, 0, stream()>>>(op, (const float4*)a, (float)scalar, reverse, (float4*)out, d.n / 4);
    PDN_LAUNCHED("binary_scalar_dense_f32");
    return 0;
  }
  int g = grid_for(d.n, 256, 2);
  if (cmp) {
    DISPATCH_ALL(dtype, (k_binary_scalar<T, bool, true><<<g, 256, 0, stream()>>>(op, (const T*)a, scalar, reverse, (bool*)out, d)));
  } else {
    PDN_CHECK(dtype != PDN_BOOL, "arithmetic on bool arrays must be promoted by the caller");
    DISPATCH_ALL(dtype, (k_binary_scalar<T, T, false><<<g, 256, 0, stream()>>>(op, (const T*)a, scalar, reverse, (T*)out, d)));
  }
  PDN_LAUNCHED("binary_scalar");
  return 0;
}

int pdn_ew_unary(int op, int dtype, const void* a, void* out, int ndim, const int64_t* shape, const int64_t* sa,
                 const int64_t* so) {
  PDN_TRY(ensure_init());
  StridedDesc    d;
  const int64_t* st[2] = {sa, so};
  PDN_TRY(make_desc(ndim, shape, 2, st, &d));
  if (d.n == 0) return 0;
  if (dtype == PDN_F32 && desc_dense(d, 0) && desc_dense(d, 1) && d.n % 4 == 0 && aligned16(a) && aligned16(out)) {
    k_unary_dense_f32<<<grid_for(d.n / 4, 256, 2), 256, 0, stream()>>>(op, (const float4*)a, (float4*)out, d.n / 4);
    PDN_LAUNCHED("unary_dense_f32");
    return 0;
  }
  PDN_CHECK(dtype != PDN_BOOL, "unary math on bool arrays is not defined");
  int g = grid_for(d.n, 256, 2);
  DISPATCH_ALL(dtype, (k_unary_strided<T, FL><<<g, 256, 0, stream()>>>(op, (const T*)a, (T*)out, d)));
  PDN_LAUNCHED("unary_strided");
  return 0;
}

int pdn_ew_ternary(int op, int dtype, const void* a, const void* b, const void* c, void* out, int ndim,
                   const int64_t* shape, const int64_t* sa, const int64_t* sb, const int64_t* sc, const int64_t* so) {
  PDN_TRY(ensure_init());
  StridedDesc    d;
  const int64_t* st[4] = {sa, sb, sc, so};
  PDN_TRY(make_desc(ndim, shape, 4, st, &d));
  if (d.n == 0) return 0;
  if (dtype == PDN_F32 && d.ndim <= 1 && d.n % 4 == 0 && desc_dense(d, 3) && aligned16(a) && aligned16(b) && aligned16(c) && aligned16(out)) {
    auto one = [&](int o) { return d.ndim == 0 || (d.ndim == 1 && d.s[o][0] == 0); };
    if ((desc_dense(d, 0) || one(0)) && (desc_dense(d, 1) || one(1)) && (desc_dense(d, 2) || one(2))) {
      k_ternary_mixed_f32<<<grid_for(d.n / 4, 256, 2), 256, 0, stream()>>>(op, (const float*)a, (const float*)b, (const float*)c, (float4*)out,
                                                                          d.n / 4, one(0) && !desc_dense(d, 0), one(1) && !desc_dense(d, 1),
                                                                          one(2) && !desc_dense(d, 2));
      PDN_LAUNCHED("ternary_mixed_f32");
      return 0;
    }
  }
  int g = grid_for(d.n, 256, 2);
  switch (dtype) {
    case PDN_F32: k_ternary_strided<float><<<g, 256, 0, stream()>>>(op, (const float*)a, (const float*)b, (const float*)c, (float*)out, d); break;
    case PDN_F64: k_ternary_strided<double><<<g, 256, 0, stream()>>>(op, (const double*)a, (const double*)b, (const double*)c, (double*)out, d); break;
    case PDN_F16: k_ternary_strided<__half><<<g, 256, 0, stream()>>>(op, (const __half*)a, (const __half*)b, (const __half*)c, (__half*)out, d); break;
    default: set_error("ternary ops need a floating dtype"); return PDN_ERR_UNSUPPORTED;
  }
  PDN_LAUNCHED("ternary_strided");
  return 0;
}

}  // extern "C"
