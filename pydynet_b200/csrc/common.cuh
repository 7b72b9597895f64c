// pydynet_b200 — shared declarations for the sm_100a backend library (libpdn_b200.so).
//
// Everything in csrc/ is written from scratch for B200; the reference (WeltXing/PyDyNet) ships no
// native code at all — its GPU path is "xp = cupy" (pydynet/cuda.py:90-91).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pdn_b200.h"

#define PDN_MAXD 8

namespace pdn {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define PDN_CUDA(expr)                                                           \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return pdn::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define PDN_CHECK(cond, ...)       \
  do {                             \
    if (!(cond)) {                 \
      pdn::set_error(__VA_ARGS__); \
      return PDN_ERR_INVALID;      \
    }                              \
  } while (0)

#define PDN_TRY(expr)      \
  do {                     \
    int _r = (expr);       \
    if (_r) return _r;     \
  } while (0)

// Checks the launch that just happened; every kernel launch in the library goes through this so
// that pdn_kernel_launch_count() is an honest count of OUR kernels (bench.py "gpu_launches").
int after_launch(const char* name);
#define PDN_LAUNCHED(name)            \
  do {                                \
    int _r = pdn::after_launch(name); \
    if (_r) return _r;                \
  } while (0)

cudaStream_t stream();       // compute stream of the current device
cudaStream_t comm_stream();  // side stream used by the NCCL all-reduce
int          sm_count();
int          ensure_init();

// scratch allocations served by the caching allocator (runtime.cu)
int  dev_alloc(void** p, size_t bytes);
void dev_free(void* p);
size_t dev_block_size(void* p);  // size of the live allocation starting at p (0 if unknown)
// operand-plane cache of the tcgen05 GEMM (gemm_tc.cu): entries derived from memory inside [p, p + n) are dropped when it is freed
void plane_cache_drop_range(const void* p, size_t n);
bool is_capturing();

struct Scratch {  // RAII scratch buffer
  void* p = nullptr;
  int   alloc(size_t bytes) { return dev_alloc(&p, bytes); }
  ~Scratch() {
    if (p) dev_free(p);
  }
};

// ---- dtype helpers --------------------------------------------------------------------------
__host__ __device__ inline int dtype_size(int dt) {
  switch (dt) {
    case PDN_F32: return 4;
    case PDN_F64: return 8;
    case PDN_F16: return 2;
    case PDN_I64: return 8;
    case PDN_I32: return 4;
    case PDN_BOOL: return 1;
    case PDN_BF16: return 2;
    case PDN_U8: return 1;
    default: return 0;
  }
}

// storage type -> compute type
template <typename T> struct Acc { using type = T; };
template <> struct Acc<__half> { using type = float; };
template <> struct Acc<__nv_bfloat16> { using type = float; };
template <> struct Acc<bool> { using type = int; };
template <> struct Acc<unsigned char> { using type = int; };

template <typename T> __device__ __forceinline__ typename Acc<T>::type ld(const T* p) { return (typename Acc<T>::type)(*p); }
template <> __device__ __forceinline__ float ld<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st(T* p, typename Acc<T>::type v) { *p = (T)v; }
template <> __device__ __forceinline__ void st<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void st<bool>(bool* p, int v) { *p = (v != 0); }

// ---- strided indexing -----------------------------------------------------------------------
// Up to four operands walk the same logical shape with their own element strides.
struct StridedDesc {
  int     ndim;
  int64_t n;  // total elements
  int64_t shape[PDN_MAXD];
  int64_t s[4][PDN_MAXD];
};

// Collapses adjacent dimensions that are jointly contiguous for all operands and drops size-1
// dims; returns 0 or an error. nops <= 4; strides[i] may not be nullptr for i < nops.
int make_desc(int ndim, const int64_t* shape, int nops, const int64_t* const* strides, StridedDesc* out);
// true if operand `op` of the collapsed descriptor is one dense run (ndim<=1 and stride 1)
inline bool desc_dense(const StridedDesc& d, int op) {
  return d.ndim == 0 || (d.ndim == 1 && d.s[op][0] == 1);
}

template <int NOPS>
__device__ __forceinline__ void decompose(const StridedDesc& d, int64_t i, int64_t* off) {
#pragma unroll
  for (int o = 0; o < NOPS; ++o) off[o] = 0;
  if (d.n <= 0x7fffffffLL) {
    uint32_t r = (uint32_t)i;
#pragma unroll
    for (int k = PDN_MAXD - 1; k >= 0; --k) {
      if (k < d.ndim) {
        uint32_t sh = (uint32_t)d.shape[k];
        uint32_t q = r / sh;
        uint32_t m = r - q * sh;
#pragma unroll
        for (int o = 0; o < NOPS; ++o) off[o] += (int64_t)m * d.s[o][k];
        r = q;
      }
    }
  } else {
#pragma unroll
    for (int k = PDN_MAXD - 1; k >= 0; --k) {
      if (k < d.ndim) {
        int64_t q = i / d.shape[k];
        int64_t m = i - q * d.shape[k];
#pragma unroll
        for (int o = 0; o < NOPS; ++o) off[o] += m * d.s[o][k];
        i = q;
      }
    }
  }
}

inline int grid_for(int64_t n, int block, int per_thread = 1) {
  int64_t b = (n + (int64_t)block * per_thread - 1) / ((int64_t)block * per_thread);
  int64_t cap = (int64_t)sm_count() * 16;  // grid-stride loops take the rest
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- warp / block reductions -----------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in every thread.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  T r = (lane < nw) ? smem32[lane] : (T)0;
  r = warp_sum(r);
  return r;
}
template <typename T>
__device__ __forceinline__ T block_max(T v, T* smem32, T lowest) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  T r = (lane < nw) ? smem32[lane] : lowest;
  r = warp_max(r);
  return r;
}

}  // namespace pdn
