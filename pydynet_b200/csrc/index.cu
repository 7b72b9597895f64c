// index.cu — integer-array indexing: gather (x[key]) and scatter (full[key] = g).
// Replaces NumPy fancy indexing behind the reference's _get_slice operator (reference
// pydynet/core/tensor.py:934-940), F.embedding (functional.py:14-20) and the CE-loss fancy index
// (functional.py:372).  Scatter reproduces NumPy ASSIGNMENT semantics: with duplicate indices the last
// occurrence wins — made deterministic here with an atomicMax "winner" pass.
#include "common.cuh"

namespace pdn {

struct IdxDesc {
  int          K;
  const long long* idx[4];
  int64_t      dim[4], stride[4];
  int64_t      J;
  int          n_outer, n_inner;
  int64_t      os[4], ost[4], is[4], ist[4];
  int64_t      outer_n, inner_n;
};

__device__ __forceinline__ int64_t idx_off(const IdxDesc& d, int64_t j, bool* ok) {
  int64_t off = 0;
  *ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < d.K) {
      long long v = d.idx[k][j];
      if (v < 0) v += d.dim[k];
      if (v < 0 || v >= d.dim[k]) *ok = false;
      off += v * d.stride[k];
    }
  }
  return off;
}
__device__ __forceinline__ int64_t sub_off(int n, const int64_t* shape, const int64_t* stride, int64_t i) {
  int64_t off = 0;
#pragma unroll
  for (int k = 3; k >= 0; --k) {
    if (k < n) {
      int64_t q = i / shape[k];
      off += (i - q * shape[k]) * stride[k];
      i = q;
    }
  }
  return off;
}
// linear id of the index tuple (for the winner table)
__device__ __forceinline__ int64_t idx_slot(const IdxDesc& d, int64_t j) {
  int64_t slot = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < d.K) {
      long long v = d.idx[k][j];
      if (v < 0) v += d.dim[k];
      if (v < 0 || v >= d.dim[k]) return -1;
      slot = slot * d.dim[k] + v;
    }
  }
  return slot;
}

template <typename T>
__global__ void __launch_bounds__(256) k_gather(const T* src, T* out, IdxDesc d) {
  int64_t total = d.outer_n * d.J * d.inner_n;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % d.inner_n, rest = t / d.inner_n;
    int64_t j = rest % d.J, o = rest / d.J;
    bool    ok;
    int64_t off = idx_off(d, j, &ok) + sub_off(d.n_outer, d.os, d.ost, o) + sub_off(d.n_inner, d.is, d.ist, i);
    if (ok) out[t] = src[off];
  }
}

// Row form of both kernels (F.embedding, functional.py:14-20, and its gradient): one index array over the leading axis of a
// matrix whose rows are contiguous 16-byte multiples — a warp copies whole rows with 128-bit accesses, no per-element index math.
// (The generic kernels above spent ~15 64-bit divisions per 4-byte element: 1.2 TB/s on the encoder's embedding lookup.)
__global__ void __launch_bounds__(256) k_gather_rows16(const uint4* __restrict__ src, uint4* __restrict__ out, const long long* __restrict__ idx,
                                                       int64_t J, int64_t dim, int64_t src_row16, int row16) {
  const int64_t total = J * row16;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / row16;
    const int     i = (int)(t - j * row16);
    long long v = __ldg(idx + j);
    if (v < 0) v += dim;
    if (v >= 0 && v < dim) out[t] = __ldg(src + v * src_row16 + i);
  }
}
__global__ void __launch_bounds__(256) k_scatter_rows16(uint4* __restrict__ dst, const uint4* __restrict__ values, const long long* __restrict__ idx,
                                                        const long long* __restrict__ winner, int64_t J, int64_t dim, int64_t dst_row16, int row16) {
  const int64_t total = J * row16;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / row16;
    const int     i = (int)(t - j * row16);
    long long v = __ldg(idx + j);
    if (v < 0) v += dim;
    if (v >= 0 && v < dim && __ldg(winner + v) == j) dst[v * dst_row16 + i] = __ldg(values + t);  // last occurrence wins
  }
}

__global__ void __launch_bounds__(256) k_winner(long long* winner, IdxDesc d) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < d.J; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t slot = idx_slot(d, j);
    if (slot >= 0) atomicMax((long long*)&winner[slot], (long long)j);
  }
}

template <typename T, bool ACC>
__global__ void __launch_bounds__(256) k_scatter(T* dst, const T* values, const long long* winner, IdxDesc d) {
  int64_t total = d.outer_n * d.J * d.inner_n;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % d.inner_n, rest = t / d.inner_n;
    int64_t j = rest % d.J, o = rest / d.J;
    bool    ok;
    int64_t off = idx_off(d, j, &ok) + sub_off(d.n_outer, d.os, d.ost, o) + sub_off(d.n_inner, d.is, d.ist, i);
    if (!ok) continue;
    if constexpr (ACC) {
      atomicAdd(dst + off, values[t]);
    } else {
      if (winner[idx_slot(d, j)] == j) dst[off] = values[t];
    }
  }
}

static int fill_desc(IdxDesc* d, int K, const void* const* idx, const int64_t* idx_dim, const int64_t* idx_stride, int64_t J,
                     int n_outer, const int64_t* outer_shape, const int64_t* outer_stride, int n_inner,
                     const int64_t* inner_shape, const int64_t* inner_stride) {
  PDN_CHECK(K >= 1 && K <= 4, "index: 1..4 index arrays supported, got %d", K);
  PDN_CHECK(n_outer >= 0 && n_outer <= 4 && n_inner >= 0 && n_inner <= 4, "index: at most 4 outer and 4 inner dims");
  d->K = K;
  d->J = J;
  for (int k = 0; k < 4; ++k) {
    d->idx[k] = k < K ? (const long long*)idx[k] : nullptr;
    d->dim[k] = k < K ? idx_dim[k] : 1;
    d->stride[k] = k < K ? idx_stride[k] : 0;
  }
  d->n_outer = n_outer;
  d->n_inner = n_inner;
  d->outer_n = d->inner_n = 1;
  for (int k = 0; k < 4; ++k) {
    d->os[k] = k < n_outer ? outer_shape[k] : 1;
    d->ost[k] = k < n_outer ? outer_stride[k] : 0;
    d->is[k] = k < n_inner ? inner_shape[k] : 1;
    d->ist[k] = k < n_inner ? inner_stride[k] : 0;
    d->outer_n *= d->os[k];
    d->inner_n *= d->is[k];
  }
  return 0;
}

}  // namespace pdn

using namespace pdn;

extern "C" {

int pdn_index_gather(const void* src, int dtype, void* out, int K, const void* const* idx, const int64_t* idx_dim,
                     const int64_t* idx_stride, int64_t J, int n_outer, const int64_t* outer_shape,
                     const int64_t* outer_stride, int n_inner, const int64_t* inner_shape, const int64_t* inner_stride) {
  PDN_TRY(ensure_init());
  IdxDesc d;
  PDN_TRY(fill_desc(&d, K, idx, idx_dim, idx_stride, J, n_outer, outer_shape, outer_stride, n_inner, inner_shape, inner_stride));
  int64_t total = d.outer_n * d.J * d.inner_n;
  if (total == 0) return 0;
  {
    const int64_t row_bytes = d.inner_n * dtype_size(dtype);
    if (K == 1 && n_outer == 0 && n_inner == 1 && d.ist[0] == 1 && row_bytes % 16 == 0 && row_bytes / 16 < 0x7fffffff &&
        (d.stride[0] * dtype_size(dtype)) % 16 == 0 && ((((uintptr_t)src) | ((uintptr_t)out)) & 15) == 0) {
      const int row16 = (int)(row_bytes / 16);
      k_gather_rows16<<<grid_for(d.J * row16, 256, 2), 256, 0, stream()>>>((const uint4*)src, (uint4*)out, d.idx[0], d.J, d.dim[0],
                                                                            d.stride[0] * dtype_size(dtype) / 16, row16);
      PDN_LAUNCHED("index_gather_rows");
      return 0;
    }
  }
  int g = grid_for(total, 256, 2);
  switch (dtype_size(dtype)) {
    case 1: k_gather<unsigned char><<<g, 256, 0, stream()>>>((const unsigned char*)src, (unsigned char*)out, d); break;
    case 2: k_gather<unsigned short><<<g, 256, 0, stream()>>>((const unsigned short*)src, (unsigned short*)out, d); break;
    case 4: k_gather<unsigned int><<<g, 256, 0, stream()>>>((const unsigned int*)src, (unsigned int*)out, d); break;
    case 8: k_gather<unsigned long long><<<g, 256, 0, stream()>>>((const unsigned long long*)src, (unsigned long long*)out, d); break;
    default: set_error("gather: bad dtype %d", dtype); return PDN_ERR_UNSUPPORTED;
  }
  PDN_LAUNCHED("index_gather");
  return 0;
}

int pdn_index_scatter(void* dst, int dtype, const void* values, int K, const void* const* idx, const int64_t* idx_dim,
                      const int64_t* idx_stride, int64_t J, int n_outer, const int64_t* outer_shape,
                      const int64_t* outer_stride, int n_inner, const int64_t* inner_shape, const int64_t* inner_stride,
                      int accumulate) {
  PDN_TRY(ensure_init());
  IdxDesc d;
  PDN_TRY(fill_desc(&d, K, idx, idx_dim, idx_stride, J, n_outer, outer_shape, outer_stride, n_inner, inner_shape, inner_stride));
  int64_t total = d.outer_n * d.J * d.inner_n;
  if (total == 0) return 0;
  int g = grid_for(total, 256, 2);
  if (accumulate) {
    switch (dtype) {
      case PDN_F32: k_scatter<float, true><<<g, 256, 0, stream()>>>((float*)dst, (const float*)values, nullptr, d); break;
      case PDN_F64: k_scatter<double, true><<<g, 256, 0, stream()>>>((double*)dst, (const double*)values, nullptr, d); break;
      default: set_error("scatter-add: only f32/f64"); return PDN_ERR_UNSUPPORTED;
    }
    PDN_LAUNCHED("index_scatter_add");
    return 0;
  }
  int64_t slots = 1;
  for (int k = 0; k < K; ++k) slots *= idx_dim[k];
  Scratch win;
  PDN_TRY(win.alloc(sizeof(long long) * (size_t)slots));
  PDN_CUDA(cudaMemsetAsync(win.p, 0xff, sizeof(long long) * (size_t)slots, stream()));  // -1
  k_winner<<<grid_for(J, 256, 1), 256, 0, stream()>>>((long long*)win.p, d);
  PDN_LAUNCHED("index_winner");
  {
    const int64_t row_bytes = d.inner_n * dtype_size(dtype);
    if (K == 1 && n_outer == 0 && n_inner == 1 && d.ist[0] == 1 && row_bytes % 16 == 0 && row_bytes / 16 < 0x7fffffff &&
        (d.stride[0] * dtype_size(dtype)) % 16 == 0 && ((((uintptr_t)dst) | ((uintptr_t)values)) & 15) == 0) {
      const int row16 = (int)(row_bytes / 16);
      k_scatter_rows16<<<grid_for(d.J * row16, 256, 2), 256, 0, stream()>>>((uint4*)dst, (const uint4*)values, d.idx[0], (const long long*)win.p, d.J,
                                                                             d.dim[0], d.stride[0] * dtype_size(dtype) / 16, row16);
      PDN_LAUNCHED("index_scatter_rows");
      return 0;
    }
  }
  switch (dtype_size(dtype)) {
    case 1: k_scatter<unsigned char, false><<<g, 256, 0, stream()>>>((unsigned char*)dst, (const unsigned char*)values, (const long long*)win.p, d); break;
    case 2: k_scatter<unsigned short, false><<<g, 256, 0, stream()>>>((unsigned short*)dst, (const unsigned short*)values, (const long long*)win.p, d); break;
    case 4: k_scatter<unsigned int, false><<<g, 256, 0, stream()>>>((unsigned int*)dst, (const unsigned int*)values, (const long long*)win.p, d); break;
    case 8: k_scatter<unsigned long long, false><<<g, 256, 0, stream()>>>((unsigned long long*)dst, (const unsigned long long*)values, (const long long*)win.p, d); break;
    default: set_error("scatter: bad dtype %d", dtype); return PDN_ERR_UNSUPPORTED;
  }
  PDN_LAUNCHED("index_scatter");
  return 0;
}

}  // extern "C"
