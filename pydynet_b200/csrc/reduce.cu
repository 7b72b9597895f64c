// reduce.cu — sum/mean/max/min/argmax/argmin over arbitrary axis sets of a strided array.
// Replaces xp.sum/mean/max/min/argmax/argmin behind the reference's _ReduceOperator
// (reference pydynet/core/tensor.py:695-773) and the engine's un-broadcast sums (tensor.py:360-370).
#include "common.cuh"
#include <float.h>

namespace pdn {

struct RedDesc {
  int     nk, nr;             // kept / reduced dim counts (after collapsing)
  int64_t n_out, n_red;
  int64_t ks[PDN_MAXD], kst[PDN_MAXD];  // kept shape / input strides
  int64_t rs[PDN_MAXD], rst[PDN_MAXD];  // reduced shape / input strides
};

__device__ __forceinline__ int64_t off_of(int nd, const int64_t* shape, const int64_t* stride, int64_t i) {
  int64_t off = 0;
#pragma unroll
  for (int k = PDN_MAXD - 1; k >= 0; --k) {
    if (k < nd) {
      int64_t q = i / shape[k];
      off += (i - q * shape[k]) * stride[k];
      i = q;
    }
  }
  return off;
}

template <typename A> struct Lim;
template <> struct Lim<float> { static __device__ float lo() { return -INFINITY; } static __device__ float hi() { return INFINITY; } };
template <> struct Lim<double> { static __device__ double lo() { return -INFINITY; } static __device__ double hi() { return INFINITY; } };
template <> struct Lim<long long> { static __device__ long long lo() { return LLONG_MIN; } static __device__ long long hi() { return LLONG_MAX; } };
template <> struct Lim<int> { static __device__ int lo() { return INT_MIN; } static __device__ int hi() { return INT_MAX; } };

template <typename A, int OP>
struct Red {
  static __device__ __forceinline__ A init() {
    if (OP == PDN_MAX) return Lim<A>::lo();
    if (OP == PDN_MIN) return Lim<A>::hi();
    return (A)0;
  }
  static __device__ __forceinline__ A comb(A a, A b) {
    if (OP == PDN_MAX) return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
    if (OP == PDN_MIN) return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
    return a + b;
  }
};

// ---- rows strategy: a group of threads per output, strided walk over the reduced index space ---
// split > 1: output o of split s covers r in [s*chunk, min((s+1)*chunk, n_red)) and lands in out[s*n_out + o]
template <typename T, typename TO, int OP, int GROUP>  // GROUP = 32 (warp per output) or 256 (block per output)
__global__ void __launch_bounds__(256) k_reduce_rows(const T* x, TO* out, RedDesc d, int split, int64_t chunk, double scale) {
  using A = typename Acc<T>::type;
  __shared__ A sm[32];
  const int     groups_per_block = 256 / GROUP;
  const int     gid = threadIdx.x / GROUP, lane = threadIdx.x % GROUP;
  const int64_t total = d.n_out * split;
  for (int64_t w = (int64_t)blockIdx.x * groups_per_block + gid; w < total + (GROUP == 256 ? 0 : 0); w += (int64_t)gridDim.x * groups_per_block) {
    int64_t o = w % d.n_out, s = w / d.n_out;
    int64_t base = off_of(d.nk, d.ks, d.kst, o);
    int64_t r0 = s * chunk, r1 = r0 + chunk < d.n_red ? r0 + chunk : d.n_red;
    A       acc = Red<A, OP>::init();
    if (d.nr == 1) {
      const T* p = x + base;
      int64_t  st = d.rst[0];
      for (int64_t r = r0 + lane; r < r1; r += GROUP) acc = Red<A, OP>::comb(acc, ld<T>(p + r * st));
    } else {
      for (int64_t r = r0 + lane; r < r1; r += GROUP) acc = Red<A, OP>::comb(acc, ld<T>(x + base + off_of(d.nr, d.rs, d.rst, r)));
    }
    // combine within the group
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc = Red<A, OP>::comb(acc, __shfl_xor_sync(0xffffffffu, acc, m));
    if (GROUP == 256) {
      __syncthreads();
      if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
      __syncthreads();
      acc = (threadIdx.x < 8) ? sm[threadIdx.x] : Red<A, OP>::init();
      if (threadIdx.x < 32) {
#pragma unroll
        for (int m = 4; m > 0; m >>= 1) acc = Red<A, OP>::comb(acc, __shfl_xor_sync(0xffffffffu, acc, m));
      }
    }
    if (lane == 0) st<TO>(out + w, (typename Acc<TO>::type)(acc * (A)scale));
  }
}

// ---- cols strategy: thread per output (consecutive outputs are adjacent in memory), 8 slices of the
// reduced range per block combined through shared memory; blockIdx.y = split
template <typename T, typename TO, int OP>
__global__ void __launch_bounds__(256) k_reduce_cols(const T* x, TO* out, RedDesc d, int split, int64_t chunk, double scale) {
  using A = typename Acc<T>::type;
  __shared__ A sm[8][33];
  int64_t o = (int64_t)blockIdx.x * 32 + threadIdx.x;
  int     s = blockIdx.y;
  int64_t r0 = s * chunk, r1 = r0 + chunk < d.n_red ? r0 + chunk : d.n_red;
  A       acc = Red<A, OP>::init();
  if (o < d.n_out) {
    int64_t base = off_of(d.nk, d.ks, d.kst, o);
    if (d.nr == 1) {
      int64_t st = d.rst[0];
      for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc = Red<A, OP>::comb(acc, ld<T>(x + base + r * st));
    } else {
      for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc = Red<A, OP>::comb(acc, ld<T>(x + base + off_of(d.nr, d.rs, d.rst, r)));
    }
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && o < d.n_out) {
#pragma unroll
    for (int j = 1; j < 8; ++j) acc = Red<A, OP>::comb(acc, sm[j][threadIdx.x]);
    st<TO>(out + (int64_t)s * d.n_out + o, (typename Acc<TO>::type)(acc * (A)scale));
  }
}

// ---- arg reductions: block per output, (value, first index) pairs --------------------------------
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256) k_argreduce(const T* x, long long* out, RedDesc d) {
  using A = typename Acc<T>::type;
  __shared__ A         sv[256];
  __shared__ long long si[256];
  for (int64_t o = blockIdx.x; o < d.n_out; o += gridDim.x) {
    int64_t   base = off_of(d.nk, d.ks, d.kst, o);
    A         best = 0;
    long long bi = -1;
    const bool    flat = d.nr == 1;  // one (collapsed) reduced axis: plain strided walk, no index decomposition
    const int64_t rstride = d.rst[0];
    for (int64_t r = threadIdx.x; r < d.n_red; r += 256) {
      A    v = ld<T>(x + base + (flat ? r * rstride : off_of(d.nr, d.rs, d.rst, r)));
      bool better;
      if (bi < 0) better = true;
      else if (best != best) better = false;  // first NaN wins (NumPy)
      else if (v != v) better = true;
      else better = IS_MAX ? (v > best) : (v < best);
      if (better) { best = v; bi = r; }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int m = 128; m > 0; m >>= 1) {
      if (threadIdx.x < m) {
        A         v2 = sv[threadIdx.x + m], v1 = sv[threadIdx.x];
        long long i2 = si[threadIdx.x + m], i1 = si[threadIdx.x];
        bool      take2;
        if (i2 < 0) take2 = false;
        else if (i1 < 0) take2 = true;
        else {
          bool n1 = v1 != v1, n2 = v2 != v2;
          if (n1 || n2) take2 = n2 && (!n1 || i2 < i1);
          else if (v1 == v2) take2 = i2 < i1;
          else take2 = IS_MAX ? (v2 > v1) : (v2 < v1);
        }
        if (take2) { sv[threadIdx.x] = v2; si[threadIdx.x] = i2; }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) out[o] = si[0];
    __syncthreads();
  }
}

static int collapse(int n, int64_t* shape, int64_t* stride) {
  // drop size-1, merge adjacent dims that are jointly contiguous
  int w = 0;
  for (int k = 0; k < n; ++k) {
    if (shape[k] == 1) continue;
    if (w > 0 && stride[w - 1] == shape[k] * stride[k]) {
      shape[w - 1] *= shape[k];
      stride[w - 1] = stride[k];
    } else {
      shape[w] = shape[k];
      stride[w] = stride[k];
      ++w;
    }
  }
  return w;
}

template <typename T, int OP>
static int launch_reduce(const T* x, T* out, const RedDesc& d, double scale) {
  // choose strategy: rows when the reduced index space contains the unit-stride run, cols otherwise
  int64_t min_r = INT64_MAX, min_k = INT64_MAX;
  for (int i = 0; i < d.nr; ++i) { int64_t s = d.rst[i] < 0 ? -d.rst[i] : d.rst[i]; if (s < min_r) min_r = s; }
  for (int i = 0; i < d.nk; ++i) { int64_t s = d.kst[i] < 0 ? -d.kst[i] : d.kst[i]; if (s < min_k) min_k = s; }
  bool    rows = d.nk == 0 || min_r <= min_k;
  int     sms = sm_count();
  if (rows) {
    // full-ish reductions: split the reduced range over many blocks, then reduce the partials
    int     split = 1;
    int64_t per_block_work = d.n_red;
    if (d.n_out < sms * 2 && per_block_work > 16384) {
      split = (int)((sms * 4 + d.n_out - 1) / d.n_out);
      int64_t max_split = (d.n_red + 4095) / 4096;
      if (split > max_split) split = (int)max_split;
      if (split < 1) split = 1;
    }
    int64_t chunk = (d.n_red + split - 1) / split;
    if (split == 1) {
      if (d.n_red <= 1024) {
        int64_t blocks = (d.n_out + 7) / 8;
        int     g = (int)(blocks < (int64_t)sms * 32 ? blocks : (int64_t)sms * 32);
        k_reduce_rows<T, T, OP, 32><<<g, 256, 0, stream()>>>(x, out, d, 1, chunk, scale);
      } else {
        int g = (int)(d.n_out < (int64_t)sms * 32 ? d.n_out : (int64_t)sms * 32);
        k_reduce_rows<T, T, OP, 256><<<g, 256, 0, stream()>>>(x, out, d, 1, chunk, scale);
      }
      PDN_LAUNCHED("reduce_rows");
      return 0;
    }
    using A = typename Acc<T>::type;
    Scratch part;
    PDN_TRY(part.alloc(sizeof(A) * d.n_out * split));
    int64_t total = d.n_out * split;
    k_reduce_rows<T, A, OP, 256><<<(int)total, 256, 0, stream()>>>(x, (A*)part.p, d, split, chunk, 1.0);
    PDN_LAUNCHED("reduce_rows_split");
    RedDesc d2{};
    d2.nk = 1; d2.nr = 1; d2.n_out = d.n_out; d2.n_red = split;
    d2.ks[0] = d.n_out; d2.kst[0] = 1; d2.rs[0] = split; d2.rst[0] = d.n_out;
    dim3 blk(32, 8), grd((unsigned)((d.n_out + 31) / 32), 1);
    k_reduce_cols<A, T, OP><<<grd, blk, 0, stream()>>>((const A*)part.p, out, d2, 1, split, scale);
    PDN_LAUNCHED("reduce_final");
    return 0;
  }
  // cols
  int64_t xblocks = (d.n_out + 31) / 32;
  int     split = 1;
  if (xblocks < sms * 2 && d.n_red > 512) {
    split = (int)((sms * 4 + xblocks - 1) / xblocks);
    int64_t max_split = (d.n_red + 255) / 256;
    if (split > max_split) split = (int)max_split;
    if (split < 1) split = 1;
  }
  int64_t chunk = (d.n_red + split - 1) / split;
  dim3    blk(32, 8);
  if (split == 1) {
    dim3 grd((unsigned)xblocks, 1);
    k_reduce_cols<T, T, OP><<<grd, blk, 0, stream()>>>(x, out, d, 1, chunk, scale);
    PDN_LAUNCHED("reduce_cols");
    return 0;
  }
  using A = typename Acc<T>::type;
  Scratch part;
  PDN_TRY(part.alloc(sizeof(A) * d.n_out * split));
  dim3 grd((unsigned)xblocks, (unsigned)split);
  k_reduce_cols<T, A, OP><<<grd, blk, 0, stream()>>>(x, (A*)part.p, d, split, chunk, 1.0);
  PDN_LAUNCHED("reduce_cols_split");
  RedDesc d2{};
  d2.nk = 1; d2.nr = 1; d2.n_out = d.n_out; d2.n_red = split;
  d2.ks[0] = d.n_out; d2.kst[0] = 1; d2.rs[0] = split; d2.rst[0] = d.n_out;
  dim3 grd2((unsigned)xblocks, 1);
  k_reduce_cols<A, T, OP><<<grd2, blk, 0, stream()>>>((const A*)part.p, out, d2, 1, split, scale);
  PDN_LAUNCHED("reduce_final");
  return 0;
}

template <typename T>
static int reduce_typed(int op, const void* x, void* out, const RedDesc& d) {
  switch (op) {
    case PDN_SUM: return launch_reduce<T, PDN_SUM>((const T*)x, (T*)out, d, 1.0);
    case PDN_MEAN: return launch_reduce<T, PDN_SUM>((const T*)x, (T*)out, d, 1.0 / (double)d.n_red);
    case PDN_MAX: return launch_reduce<T, PDN_MAX>((const T*)x, (T*)out, d, 1.0);
    case PDN_MIN: return launch_reduce<T, PDN_MIN>((const T*)x, (T*)out, d, 1.0);
    case PDN_ARGMAX: {
      int g = (int)(d.n_out < (int64_t)sm_count() * 16 ? d.n_out : (int64_t)sm_count() * 16);
      k_argreduce<T, true><<<g, 256, 0, stream()>>>((const T*)x, (long long*)out, d);
      PDN_LAUNCHED("argmax");
      return 0;
    }
    case PDN_ARGMIN: {
      int g = (int)(d.n_out < (int64_t)sm_count() * 16 ? d.n_out : (int64_t)sm_count() * 16);
      k_argreduce<T, false><<<g, 256, 0, stream()>>>((const T*)x, (long long*)out, d);
      PDN_LAUNCHED("argmin");
      return 0;
    }
  }
  set_error("unknown reduce op %d", op);
  return PDN_ERR_INVALID;
}

}  // namespace pdn

using namespace pdn;

extern "C" int pdn_reduce(int op, int dtype, const void* x, void* out, int ndim, const int64_t* shape, const int64_t* sx,
                          uint32_t reduce_mask) {
  PDN_TRY(ensure_init());
  PDN_CHECK(ndim >= 0 && ndim <= PDN_MAXD, "ndim %d out of range", ndim);
  RedDesc d{};
  d.n_out = 1;
  d.n_red = 1;
  for (int i = 0; i < ndim; ++i) {
    if (reduce_mask & (1u << i)) {
      d.rs[d.nr] = shape[i]; d.rst[d.nr] = sx[i]; d.nr++; d.n_red *= shape[i];
    } else {
      d.ks[d.nk] = shape[i]; d.kst[d.nk] = sx[i]; d.nk++; d.n_out *= shape[i];
    }
  }
  if (d.n_out == 0) return 0;
  PDN_CHECK(d.n_red > 0 || op == PDN_SUM, "zero-size reduction has no identity for op %d", op);
  d.nk = collapse(d.nk, d.ks, d.kst);
  d.nr = collapse(d.nr, d.rs, d.rst);
  if (d.nr == 0) { d.nr = 1; d.rs[0] = d.n_red; d.rst[0] = 0; }  // all reduced dims had size 1 (or none)
  switch (dtype) {
    case PDN_F32: return reduce_typed<float>(op, x, out, d);
    case PDN_F64: return reduce_typed<double>(op, x, out, d);
    case PDN_F16: return reduce_typed<__half>(op, x, out, d);
    case PDN_I64: return reduce_typed<long long>(op, x, out, d);
    case PDN_I32: return reduce_typed<int>(op, x, out, d);
  }
  set_error("reduce: unsupported dtype %d", dtype);
  return PDN_ERR_UNSUPPORTED;
}
