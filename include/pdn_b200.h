/* pdn_b200.h — C ABI of libpdn_b200.so, the sm_100a dense-tensor backend behind PyDyNet's Tensor surface.
 *
 * The reference (WeltXing/PyDyNet) has NO native interface: its GPU path is "xp = cupy"
 * (reference pydynet/cuda.py:90-91) and every operator is a forward_/grad_fn pair of array-library
 * calls (reference pydynet/core/tensor.py:416-533).  The entry points below are what a ctypes/cffi
 * binding for that seam binds instead of CuPy; each one cites the reference call site it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a PDN_ERR_* code otherwise; the message is
 *     pdn_last_error().  Nothing throws, nothing aborts.
 *   - pointers named d_* / x / out are DEVICE pointers unless the comment says host.
 *   - shapes and strides are int64 arrays of `ndim` entries; strides are in ELEMENTS, may be 0
 *     (broadcast) or negative.
 *   - launches are asynchronous on the library's per-device compute stream; host reads go through
 *     pdn_memcpy_d2h (which synchronises), mirroring the reference's only sync points
 *     .item()/.numpy() (tensor.py:385-393).
 */
#ifndef PDN_B200_H
#define PDN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PDN_OK = 0, PDN_ERR_CUDA = 1, PDN_ERR_INVALID = 2, PDN_ERR_OOM = 3, PDN_ERR_NCCL = 4, PDN_ERR_UNSUPPORTED = 5 };

/* numeric types (numpy dtypes the reference tests pin: fp16/32/64 + int64 indices + bool masks) */
enum { PDN_F32 = 0, PDN_F64 = 1, PDN_F16 = 2, PDN_I64 = 3, PDN_I32 = 4, PDN_BOOL = 5, PDN_BF16 = 6, PDN_U8 = 7 };

/* binary elementwise ops — reference tensor.py:535-641 (add sub mul div pow), :808-823 (maximum minimum),
 * comparisons tensor.py:288-324 */
enum { PDN_ADD = 0, PDN_SUB, PDN_MUL, PDN_DIV, PDN_POW, PDN_MAXIMUM, PDN_MINIMUM,
       PDN_EQ = 16, PDN_NE, PDN_LT, PDN_LE, PDN_GT, PDN_GE };

/* unary elementwise ops — reference tensor.py:679-692 (abs), :776-805 (exp log), :826-832 (sign),
 * :996-1019 (sigmoid tanh, piecewise overflow-safe forms), function.py:4-11 (sqrt square) */
enum { PDN_NEG = 0, PDN_EXP, PDN_LOG, PDN_ABS, PDN_SIGN, PDN_SIGMOID, PDN_TANH, PDN_SQRT, PDN_SQUARE, PDN_RECIP,
       PDN_SILU, PDN_RELU };

/* fused three-operand ops used by grad_fn bodies (one pass instead of the reference's 2-3 array expressions) */
enum {
  PDN_T_EQ_MUL = 0,      /* (a == b) * c            max/min/maximum/minimum grad, tensor.py:741-747,812-823 */
  PDN_T_DIV_GRAD_Y,      /* -a * b / c              div grad wrt y: -out * grad / y, tensor.py:611-615     */
  PDN_T_POW_GRAD_X,      /* a * b / c  (* handled by caller) */
  PDN_T_SIGMOID_GRAD,    /* a * (1 - a) * b         tensor.py:1004-1005 (c unused) */
  PDN_T_TANH_GRAD,       /* (1 - a*a) * b           tensor.py:1017-1018 (c unused) */
  PDN_T_FMA,             /* a * b + c */
  PDN_T_SILU_GRAD,       /* d/dx [x/(1+exp(-x))] (a=x) * b */
  PDN_T_WHERE            /* a ? b : c  (a is same dtype, nonzero = true) */
};

/* reductions — reference tensor.py:695-773 */
enum { PDN_SUM = 0, PDN_MEAN, PDN_MAX, PDN_MIN, PDN_ARGMAX, PDN_ARGMIN };

/* ---------------------------------------------------------------- runtime ------------------- */
/* replaces cupy.cuda.runtime.getDeviceCount/getDevice/setDevice used by reference cuda.py:16-32 */
int pdn_device_count(int* n);
int pdn_init(int device);
int pdn_set_device(int device);
int pdn_get_device(int* device);
int pdn_device_name(char* buf, int buflen);
int pdn_sm_count(int* n);
const char* pdn_last_error(void);

/* stream-ordered caching allocator (replaces cupy's memory pool behind xp.zeros/xp.array, tensor.py:80,90) */
int pdn_malloc(void** p, size_t bytes);
int pdn_free(void* p);
int pdn_malloc_host(void** p, size_t bytes); /* pinned host memory */
int pdn_free_host(void* p);
int pdn_mem_stats(uint64_t* bytes_in_use, uint64_t* bytes_cached, uint64_t* n_cuda_malloc);
int pdn_empty_cache(void);

/* Tensor.to / .numpy / .item (tensor.py:385-407): H2D, D2H (synchronising), D2D */
int pdn_memcpy_h2d(void* dst, const void* host_src, size_t bytes);
int pdn_memcpy_d2h(void* host_dst, const void* src, size_t bytes);
int pdn_memcpy_d2d(void* dst, const void* src, size_t bytes);
int pdn_memcpy_h2d_async(void* dst, const void* pinned_src, size_t bytes);
/* Input pipeline (pydynet/data.py:73-123 DataLoader, examples/pydynet/mnist.py:161-162: every batch is a synchronous
 * Tensor(numpy) upload in the reference): the batch is copied from PINNED memory on a dedicated copy stream while the compute stream
 * works on the previous one; `done` is an event from pdn_event_create. pdn_stream_wait_event(done) orders the compute stream after
 * the copy; pdn_event_synchronize(done) tells the host the pinned staging buffer may be overwritten. */
int pdn_prefetch_h2d(void* dst, const void* pinned_src, size_t bytes, void* done);
int pdn_stream_wait_event(void* ev);
int pdn_event_synchronize(void* ev);
int pdn_memcpy_d2h_async(void* pinned_dst, const void* src, size_t bytes);
int pdn_memset(void* dst, int byte, size_t bytes);
int pdn_sync(void);

/* NVTX ranges around the phases of a step (forward / backward / optimizer / all-reduce; pydynet_b200.cuda.nvtx_range, active with
 * PDN_NVTX=1, which also marks every kernel launch with its entry-point name) for `ncu --nvtx --nvtx-include` and nsys timelines.
 * Nothing in the reference corresponds to this (it has no tracing). */
int pdn_nvtx_push(const char* name);
int pdn_nvtx_pop(void);

/* Recorded fork regions: inside pdn_graph_begin .. pdn_graph_end, fork the compute stream into n <= 4 branches, direct the
 * following launches and allocations of the calling thread to branch i, add an ordering edge between two branches (mark / wait), join. The
 * branches of a replayed graph run concurrently (independent batch slices of the decode step: the launch-latency-bound chain of one
 * slice fills the gaps of the others). No-ops outside a recording. Nothing in the reference corresponds to this (one NumPy thread). */
int pdn_branch_begin(int n);
int pdn_branch_select(int i);
int pdn_branch_mark(int* token);  /* edge source: the current tail of the selected branch */
int pdn_branch_wait(int token);   /* the selected branch's next work waits for that point */
int pdn_branch_end(void);

/* launch accounting (bench.py "gpu_launches") and device-side timing on the library stream */
uint64_t pdn_kernel_launch_count(void);
void pdn_reset_launch_count(void);
/* test aid: count the launches whose entry-point name equals `name` from now on (nullptr = stop watching); the count so far */
void pdn_watch_launches(const char* name);
uint64_t pdn_watched_launch_count(void);
int pdn_event_create(void** ev);
int pdn_event_destroy(void* ev);
int pdn_event_record(void* ev); /* while a graph is being recorded: an event-record NODE (cudaEventRecordExternal), re-stamped by every replay */
int pdn_event_elapsed_ms(void* start, void* stop, float* ms); /* synchronises on stop */

/* CUDA-graph capture of a launch sequence on the library stream (decode loop replay) */
int pdn_graph_begin(void);
int pdn_graph_status(int* status); /* 0 none, 1 recording, 2 recording invalidated (an operation that cannot be recorded was issued) */
int pdn_graph_end(void** graph_exec);
int pdn_graph_launch(void* graph_exec);
int pdn_graph_destroy(void* graph_exec);

/* ---------------------------------------------------------------- elementwise --------------- */
/* strided fill: xp.zeros/ones, data[...] = val (tensor.py:90,355,383; init.py:42-44) */
int pdn_fill(void* out, int dtype, int ndim, const int64_t* shape, const int64_t* so, double value);
/* strided copy with cast: astype (tensor.py:165-174), .copy(), __setitem__ of arrays (tensor.py:278-279),
 * xp.concatenate pieces (tensor.py:982-985), xp.broadcast_to materialisation */
int pdn_copy(const void* src, int sdtype, void* dst, int ddtype, int ndim, const int64_t* shape,
             const int64_t* ss, const int64_t* ds);
/* out = a (op) b with NumPy broadcasting expressed through 0 strides; inputs share `dtype`,
 * out has `dtype` (arithmetic) or PDN_BOOL (comparisons). out may alias a (in-place +=, tensor.py:281-294,371) */
int pdn_ew_binary(int op, int dtype, const void* a, const void* b, void* out, int ndim, const int64_t* shape,
                  const int64_t* sa, const int64_t* sb, const int64_t* so);
/* out = a (op) scalar, or scalar (op) a when reverse != 0 — the reference wraps Python scalars as
 * 0-d Tensors of the other operand's dtype (tensor.py:488-493) */
int pdn_ew_binary_scalar(int op, int dtype, const void* a, double scalar, int reverse, void* out, int ndim,
                         const int64_t* shape, const int64_t* sa, const int64_t* so);
int pdn_ew_unary(int op, int dtype, const void* a, void* out, int ndim, const int64_t* shape, const int64_t* sa,
                 const int64_t* so);
int pdn_ew_ternary(int op, int dtype, const void* a, const void* b, const void* c, void* out, int ndim,
                   const int64_t* shape, const int64_t* sa, const int64_t* sb, const int64_t* sc,
                   const int64_t* so);

/* ---------------------------------------------------------------- reductions ---------------- */
/* x.sum/mean/max/min/argmax/argmin(axis, keepdims) (tensor.py:695-773). reduce_mask bit i set = dim i
 * reduced. out is C-contiguous over the kept dims (dtype = input dtype, or PDN_I64 for arg ops —
 * arg ops take exactly one reduced dim, or all dims for axis=None, NumPy first-occurrence rule). */
int pdn_reduce(int op, int dtype, const void* x, void* out, int ndim, const int64_t* shape, const int64_t* sx,
               uint32_t reduce_mask);

/* ---------------------------------------------------------------- indexing ------------------ */
/* x[key] with integer-array keys (tensor.py:934-935; F.embedding functional.py:14-20; CE fancy index
 * functional.py:372). src dims are split [outer | K indexed dims | inner]; idx[k] are device int64
 * arrays of J entries (negative values wrap); out is C-contiguous [outer, J, inner]. */
int pdn_index_gather(const void* src, int dtype, void* out, int K, const void* const* idx, const int64_t* idx_dim,
                     const int64_t* idx_stride, int64_t J, int n_outer, const int64_t* outer_shape,
                     const int64_t* outer_stride, int n_inner, const int64_t* inner_shape,
                     const int64_t* inner_stride);
/* full[key] = values (tensor.py:937-940, __setitem__ tensor.py:278-279): NumPy assignment semantics,
 * i.e. for duplicate indices the LAST occurrence wins (deterministic here: highest j wins).
 * values is C-contiguous [outer, J, inner]. accumulate != 0 gives np.add.at semantics instead. */
int pdn_index_scatter(void* dst, int dtype, const void* values, int K, const void* const* idx,
                      const int64_t* idx_dim, const int64_t* idx_stride, int64_t J, int n_outer,
                      const int64_t* outer_shape, const int64_t* outer_stride, int n_inner,
                      const int64_t* inner_shape, const int64_t* inner_stride, int accumulate);

/* ---------------------------------------------------------------- GEMM ---------------------- */
/* x.data @ y.data (tensor.py:657-659) and its grads g @ Bᵀ, Aᵀ @ g on swapaxes views (tensor.py:661-676).
 * C[b0,b1,b2] (M×N, row-major, ldc) = A[..] (M×K, element strides a_rs/a_cs) @ B[..] (K×N, strides b_rs/b_cs)
 * with up to three batch dims (shape nb[3], per-operand batch strides, 0 = broadcast).
 * prec: 0 = default (tcgen05 BF16x3 split for F32 when the shape qualifies, else fp32 FFMA tiles),
 *       1 = force FFMA path, 2 = force tcgen05 path (error if not eligible).
 * If accumulate != 0, C += A@B.  bias (may be NULL) is a length-N vector added to every row (F.linear,
 * functional.py:7-11). */
int pdn_gemm(int dtype, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t a_rs,
             int64_t a_cs, int64_t b_rs, int64_t b_cs, int64_t ldc, const int64_t* nb, const int64_t* a_bs,
             const int64_t* b_bs, const int64_t* c_bs, const void* bias, int accumulate, int prec);
/* Same product with operand-plane caching for the tcgen05 path. Training multiplies the same buffers several times per step (x in
 * the Q/K/V projections and again in dW = x^T @ g, tensor.py:672-676; g in both gradient products; W in forward and dX), and every
 * use reads the same bf16 hi/lo planes because operands are packed in their own orientation. a_version / b_version are the
 * caller's write counters of the buffers holding A / B (pydynet_b200: DeviceBuffer.version): >= 0 lets the library keep the planes
 * until the counter changes or the memory is freed (pdn_free); -1 = transient operand (pdn_gemm). Capacity: PDN_PLANE_CACHE_MB
 * (default 16384), PDN_PLANE_CACHE=0 disables. */
int pdn_gemm_cached(int dtype, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t a_rs,
                    int64_t a_cs, int64_t b_rs, int64_t b_cs, int64_t ldc, const int64_t* nb, const int64_t* a_bs,
                    const int64_t* b_bs, const int64_t* c_bs, const void* bias, int accumulate, int prec, int64_t a_version,
                    int64_t b_version);
int pdn_plane_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* bytes, uint64_t* entries);
/* Inference: pack a constant fp32 weight matrix B (K x N, element strides b_rs/b_cs) once into the tcgen05 operand format
 * and reuse it: C[M,N] (+)= A[M,K] @ B (+ bias). The handle owns device memory until pdn_gemm_prepack_free. The caller
 * must re-pack when the weight values change (pydynet_b200 tracks a per-buffer version for that). */
int pdn_gemm_prepack(const float* B, int64_t K, int64_t N, int64_t b_rs, int64_t b_cs, void** handle);
int pdn_gemm_prepacked(const float* A, void* handle, float* C, int64_t M, int64_t a_rs, int64_t a_cs, int64_t ldc,
                       const float* bias, int accumulate);
/* same with the A operand already emitted as bf16 hi/lo planes [2][M][Kp] by a producer kernel (pdn_rmsnorm_planes,
 * pdn_swiglu_rows_planes, pdn_attention_fwd* with out_planes): no fp32 round trip, no pack launch. */
int pdn_gemm_prepacked_planes(const void* A_planes, int64_t M, int64_t Kp, void* handle, float* C, int64_t ldc,
                              const float* bias, int accumulate);
/* greedy decoding: out_idx[m] = argmax_n (A @ B + bias)[m, n] (first occurrence), computed in the GEMM epilogue + a tiny
 * second stage; the [M, N] logits are never written (reference llm/llama/model.py:268: logits[:, -1, :].argmax(-1)) */
int pdn_gemm_prepacked_planes_argmax(const void* A_planes, int64_t M, int64_t Kp, void* handle, const float* bias,
                                     int64_t* out_idx);
int pdn_gemm_prepack_free(void* handle);
/* which path the last pdn_gemm call took: 0 = FFMA tiles, 1 = tcgen05, 2 = skinny (M<=16) */
int pdn_gemm_last_path(void);

/* ---------------------------------------------------------------- fused rows ---------------- */
/* F.softmax / F.log_softmax over the last axis of a C-contiguous [rows, n] view (functional.py:43-58):
 * max under no_grad, sub, exp, sum, div in one pass. mask (nullable) is an additive [mask_rows, n] term
 * broadcast over rows as row % mask_rows (attention's causal mask, llm/llama/model.py:113-117). */
int pdn_softmax_fwd(int dtype, const void* x, void* y, int64_t rows, int64_t n, int log_mode);
/* dx = y * (g - sum(g*y))  (log_mode: dx = g - exp(y) * sum(g)) — the composite's grad, functional.py:43-58 */
int pdn_softmax_bwd(int dtype, const void* y, const void* g, void* dx, int64_t rows, int64_t n, int log_mode);

/* RMSNorm over the last axis (norm.py:245-248): y = x / sqrt(mean(x^2) + eps) * w ; rstd[rows] saved */
int pdn_rmsnorm_fwd(const float* x, const float* w, float* y, float* rstd, int64_t rows, int64_t n, float eps);
/* dx (nullable) and dw[n] (nullable; zeroed here, accumulated with atomics) of the composite */
/* inference: y emitted directly as GEMM operand planes [2][rows][Kp] (pad columns zero) */
int pdn_rmsnorm_planes(const float* x, const float* w, void* planes, int64_t rows, int64_t n, int64_t Kp, float eps);
int pdn_rmsnorm_bwd(const float* x, const float* w, const float* rstd, const float* g, float* dx, float* dw, int64_t rows,
                    int64_t n);

/* Batch-statistic normalisation shared by BatchNorm1d/2d and the reference's "LayerNorm" (norm.py:58-73,
 * 132-147,203-218): x viewed as [outer, C, inner], statistics per channel c over outer*inner elements
 * (biased variance), y = (x-mean)/sqrt(var+eps)*scale+shift. mean/var out are [C]. */
int pdn_bnorm_stats(const float* x, float* mean, float* var, int64_t outer, int64_t C, int64_t inner);
/* running statistics in place: r <- (1 - momentum) r + momentum stat (norm.py:66-69, 140-143, 211-214), both vectors in one launch */
int pdn_bnorm_running(float* running_mean, float* running_var, const float* mean, const float* var, float momentum, int64_t C);
int pdn_bnorm_apply(const float* x, const float* mean, const float* var, const float* scale, const float* shift,
                    float* y, int64_t outer, int64_t C, int64_t inner, float eps);
/* backward of the composite: given g, x, mean, var → dx, dscale[C], dshift[C] */
int pdn_bnorm_bwd(const float* x, const float* mean, const float* var, const float* scale, const float* g,
                  float* dx, float* dscale, float* dshift, int64_t outer, int64_t C, int64_t inner, float eps);

/* Data-parallel variants (equal shards): local partial sums pre-scaled by 1/m_global, to be summed across ranks with
 * pdn_allreduce_sum_f32_inline so that every rank normalises with the GLOBAL batch statistics — what the single-process
 * reference computes on the whole batch. which = 0: mean partial; 1: centred-square partial. */
int pdn_bnorm_partial(const float* x, const float* mean, float* out, int64_t outer, int64_t C, int64_t inner, int which,
                      float inv_m_global);
int pdn_bnorm_bwd_reduce(const float* x, const float* mean, const float* var, const float* g, float* mg, float* mgx,
                         int64_t outer, int64_t C, int64_t inner, float eps, float inv_m_global);
int pdn_bnorm_bwd_dx(const float* x, const float* mean, const float* var, const float* scale, const float* g,
                     const float* mg, const float* mgx, float* dx, int64_t outer, int64_t C, int64_t inner, float eps);

/* ---------------------------------------------------------------- conv / pool --------------- */
/* F.conv2d (functional.py:254-281): pad → im2col → GEMM → NCHW, fused; x [N,C,H,W], w [O,C,k,k],
 * bias nullable [O] (Conv2d.forward conv.py:99-103), y [N,O,oh,ow] all C-contiguous fp32.
 * Stride-1 convolutions with >= 16 contraction channels run as TMA-tiled implicit GEMMs (csrc/conv_tma.cu: shifted 5-D TMA boxes of
 * the activation's bf16 hi/lo NCHW planes, zero padding by out-of-bounds fill); others through the gather producer of
 * csrc/gemm_tc.cu. *_version: write counters of the activation buffers (>= 0: their operand planes stay in the plane cache so
 * backward-weight re-uses the forward's pack of x and backward-data's pack of g), -1 = transient. */
int pdn_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int64_t N, int64_t C, int64_t H,
                   int64_t W, int64_t O, int k, int stride, int pad, int64_t x_version);
/* dx = col2im(g_col @ Wmat) — replaces xp.add.at (functional.py:224-232) */
int pdn_conv2d_bwd_data(const float* g, const float* w, float* dx, int64_t N, int64_t C, int64_t H, int64_t W,
                        int64_t O, int k, int stride, int pad, int64_t g_version);
/* dw = colᵀ @ g ; dbias = sum over N,oh,ow (nullable) */
int pdn_conv2d_bwd_weight(const float* x, const float* g, float* dw, float* dbias, int64_t N, int64_t C, int64_t H,
                          int64_t W, int64_t O, int k, int stride, int pad, int64_t x_version, int64_t g_version);
/* F.max_pool2d / avg_pool2d (functional.py:284-339); mode 0 = max, 1 = avg. Padding is zero padding that
 * takes part in max/mean like the reference's xp.pad. bwd for max: EVERY element equal to the window max
 * receives the window's gradient (tensor.py:741-747). */
int pdn_pool2d_fwd(const float* x, float* y, int64_t N, int64_t C, int64_t H, int64_t W, int k, int stride, int pad,
                   int mode);
int pdn_pool2d_bwd(const float* x, const float* y, const float* g, float* dx, int64_t N, int64_t C, int64_t H,
                   int64_t W, int k, int stride, int pad, int mode);

/* ---------------------------------------------------------------- attention ----------------- */
/* softmax(q kᵀ * scale + mask) v, per (batch, head) — llm/llama/model.py:112-121,
 * examples/pydynet/transformer.py:93-104. q [B,H,Lq,D], k/v [B,H,Lk,D] given by element strides
 * (batch, head, row; the D axis is unit-stride) so head-split views of [B,L,H*D] projections and KV-cache
 * views are consumed in place. mask: nullable additive term (−inf allowed), unit stride along keys, element strides
 * mask_str = {batch, query-row} (0 = broadcast) — covers the causal [Lq,Lk] mask of Llama and the [B,1,1,Lk] padding
 * mask of the encoder. out [B,Lq,H,D] contiguous (= the transpose(0,2,1,3).reshape(B,L,-1) the models apply next).
 * lse [B,H,Lq] saved for backward. D <= 128. */
int pdn_attention_fwd(const float* q, const float* k, const float* v, const float* mask, float* out, float* lse,
                      int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const int64_t* q_str,
                      const int64_t* k_str, const int64_t* v_str, const int64_t* mask_str, float scale,
                      void* out_planes, int64_t planes_kp);
/* (out_planes != NULL: instead of fp32 `out`, write the [B*Lq, H*D] result as GEMM operand planes [2][B*Lq][planes_kp]) */
/* dq [B,Lq,H,D], dk/dv [B,Lk,H,D] contiguous (each nullable; dk/dv are zeroed here and accumulated atomically) */
int pdn_attention_bwd(const float* q, const float* k, const float* v, const float* mask, const float* out,
                      const float* lse, const float* g_out, float* dq, float* dk, float* dv, int64_t B, int64_t H,
                      int64_t Lq, int64_t Lk, int64_t D, const int64_t* q_str, const int64_t* k_str,
                      const int64_t* v_str, const int64_t* mask_str, float scale);

/* Tensor-core form of the same operator (D <= 64): scores live in TMEM only, P / dS are written back to TMEM as bf16 hi/lo
 * planes and read by the accumulate MMA from there (TS-mode tcgen05.mma), V / K / Q / dO tiles are consumed as MN-major B
 * operands (no transposed copies), fp32 parity via the BF16x3 split; same arguments and layouts as pdn_attention_fwd / _bwd
 * (csrc/attention_tc.cu). Chosen by the Python layer for training-sized problems. versions (nullable) = write counters of the
 * buffers holding q, k, v: >= 0 keeps their operand planes in the plane cache so that backward re-uses the forward's packs. */
int pdn_attention_tc_fwd(const float* q, const float* k, const float* v, const float* mask, float* out, float* lse,
                         int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t D, const int64_t* q_str,
                         const int64_t* k_str, const int64_t* v_str, const int64_t* mask_str, float scale,
                         const int64_t* versions);
int pdn_attention_tc_bwd(const float* q, const float* k, const float* v, const float* mask, const float* out,
                         const float* lse, const float* g_out, float* dq, float* dk, float* dv, int64_t B, int64_t H,
                         int64_t Lq, int64_t Lk, int64_t D, const int64_t* q_str, const int64_t* k_str,
                         const int64_t* v_str, const int64_t* mask_str, float scale, const int64_t* versions);

/* ---------------------------------------------------------------- recurrent ----------------- */
/* GRU sequence (rnn.py:529-544 cell, :702-708 loop). xp1 [T,B,2H] = x@Wx1+b1 and xp2 [T,B,H] = x@Wx2+b2 are
 * hoisted input projections (one GEMM over all T); h0 [B,H]; Wh1 [H,2H], Wh2 [H,H]. Writes hs [T,B,H] and the
 * gate activations zr [T,B,2H], n [T,B,H], rhWh2-free recompute inputs needed by backward. */
int pdn_gru_seq_fwd(const float* xp1, const float* xp2, const float* h0, const float* Wh1, const float* Wh2,
                    float* hs, float* zr, float* nn, int64_t T, int64_t B, int64_t H);
/* backward through time: g_hs [T,B,H] (grad of every output step; may be all-zero but not NULL), returns
 * dxp1 [T,B,2H], dxp2 [T,B,H], dh0 [B,H], dWh1 [H,2H], dWh2 [H,H] (overwritten; nullable). */
int pdn_gru_seq_bwd(const float* g_hs, const float* h0, const float* hs, const float* zr, const float* nn,
                    const float* Wh1, const float* Wh2, float* dxp1, float* dxp2, float* dh0, float* dWh1,
                    float* dWh2, int64_t T, int64_t B, int64_t H);
/* Plain RNN sequence (rnn.py:38-49 cell, :196-214 loop): xp [T,B,H] = x@Wx+b hoisted; hs[t] = act(xp[t] + hs[t-1]@Wh),
 * act = tanh (relu = 0) or relu (relu = 1). Backward: g_hs [T,B,H] (nullable = zeros) -> dxp [T,B,H], dh0 [B,H], dWh [H,H]
 * (nullable). H a multiple of 64 (<= 512): one persistent cooperative launch per direction (rnn_persist.cu). */
int pdn_rnn_seq_fwd(const float* xp, const float* h0, const float* Wh, float* hs, int64_t T, int64_t B, int64_t H, int relu);
int pdn_rnn_seq_bwd(const float* g_hs, const float* h0, const float* hs, const float* Wh, float* dxp, float* dh0, float* dWh,
                    int64_t T, int64_t B, int64_t H, int relu);
/* LSTM sequence (rnn.py:268-288): xp [T,B,4H] = x@Wx+b hoisted, gate order f,i,o,g. */
int pdn_lstm_seq_fwd(const float* xp, const float* h0, const float* c0, const float* Wh, float* hs, float* cs,
                     float* gates, int64_t T, int64_t B, int64_t H);
int pdn_lstm_seq_bwd(const float* g_hs, const float* g_cT, const float* h0, const float* c0, const float* hs,
                     const float* cs, const float* gates, const float* Wh, float* dxp, float* dh0, float* dc0,
                     float* dWh, int64_t T, int64_t B, int64_t H);

/* ---------------------------------------------------------------- loss / optimiser ---------- */
/* F.cross_entropy_loss with int targets, reduction mean|sum (functional.py:364-381): loss[1];
 * saves per-row log-sum-exp. dlogits = (softmax - onehot) * gscale. */
int pdn_ce_loss_fwd(const float* logits, const int64_t* target, float* loss, float* lse, int64_t N, int64_t C,
                    int mean);
int pdn_ce_loss_bwd(const float* logits, const int64_t* target, const float* lse, const float* gloss, float* dlogits,
                    int64_t N, int64_t C, int mean);
/* Adam.step (optimizer.py:185-196) over one flat segment: g = grad*grad_scale + wd*p; m,v update;
 * p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps). t is the reference's shared step counter (starts at 1). */
int pdn_adam_step(float* p, const float* grad, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                  float wd, int t, float grad_scale);
/* multi-tensor form: arrays of n_tensors device pointers/sizes (host arrays), one launch. */
int pdn_adam_multi(int n_tensors, float* const* p, const float* const* grad, float* const* m, float* const* v,
                   const int64_t* sizes, float lr, float b1, float b2, float eps, float wd, int t, float grad_scale);
/* The same update with its host arithmetic moved to the device, for steps recorded into a CUDA graph (pydynet_b200.cuda.graphed_step):
 * t_dev (int, starts at 1) and lr_dev (float) live in device memory; a one-thread kernel writes step_dev = lr * sqrt(1 - b2^t) /
 * (1 - b1^t) and advances t before the update kernel reads it, so every replay performs the next optimizer step. */
int pdn_adam_step_dev(float* p, const float* grad, float* m, float* v, int64_t n, float b1, float b2, float eps, float wd, float grad_scale,
                      int* t_dev, const float* lr_dev, float* step_dev);

/* ---------------------------------------------------------------- Llama decode fast path ---- */
/* interleaved-pair RoPE (llm/llama/model.py:23-44) applied in place to q and k rows [rows, H, D] using
 * cos/sin [max_seq, D/2] at positions pos0 + (row % L) ; then k,v rows appended to the KV cache
 * [Bmax, S, H, D] at [b, pos0 + l] (model.py:105-107). */
int pdn_rope_kv_append(float* q, float* k, const float* v, const float* cosT, const float* sinT, float* cache_k,
                       float* cache_v, int64_t B, int64_t L, int64_t H, int64_t D, int64_t S, int64_t pos0, int64_t ld);
/* ld = elements between consecutive q/k/v rows: H*D for separate projections (0 selects it), 3*H*D when q, k, v are the
 * three column blocks of one fused QKV projection output. */
/* Device-scalar variants for CUDA-graph replay of one decode step: the sequence position is read from device memory
 * (pos_dev) instead of being baked into the launch; attention runs over Lk = *pos_dev + lk_add cached keys, no mask. */
int pdn_rope_kv_append_dev(float* q, float* k, const float* v, const float* cosT, const float* sinT, float* cache_k,
                           float* cache_v, int64_t B, int64_t L, int64_t H, int64_t D, int64_t S, const int64_t* pos_dev,
                           int64_t ld);
int pdn_attention_fwd_dev(const float* q, const float* k, const float* v, float* out, int64_t B, int64_t H, int64_t Lq,
                          int64_t D, const int64_t* q_str, const int64_t* k_str, const int64_t* v_str, float scale,
                          const int64_t* pos_dev, int64_t lk_add, void* out_planes, int64_t planes_kp);
/* out = silu(gate) * up  (FeedForward.forward model.py:56-58), gate/up are the two halves [rows, F] */
int pdn_swiglu(const float* gate, const float* up, float* out, int64_t n);
/* same on the [rows, 2F] output of a fused gate|up projection: out[r, j] = silu(gu[r, j]) * gu[r, F + j] */
int pdn_swiglu_rows(const float* gu, float* out, int64_t rows, int64_t F);
int pdn_swiglu_rows_planes(const float* gu, void* planes, int64_t rows, int64_t F, int64_t Kp);
int pdn_swiglu_bwd(const float* gate, const float* up, const float* g, float* dgate, float* dup, int64_t n);

/* Whole decode step for small batches (rows < 32) as ONE cooperative persistent kernel (csrc/decode_mega.cu): replaces the
 * ~490 eager array expressions per token of the reference's own B = 1 loop — llm/llama/model.py:254-256 (forward), 192-207
 * (_forward_hidden), 142-150 (block), 95-121 (attention with KV cache), 23-44 (RoPE), 56-58 (SwiGLU), norm.py:245-248, and the
 * greedy argmax of model.py:268. Weights are passed as re-laid-out copies (made by the caller, rebuilt when a weight changes):
 * layer_ptrs[l*8 + {0..7}] = {[Wq|Wk|Wv]ᵀ [3*dim][dim], Wo by head [H][dim][hd] (element [h][n][d] = Wo[h*hd + d][n]), gate/up
 * rows interleaved [2*FF][dim], W_downᵀ [dim][FF], input_norm weight, post_attn_norm weight, cache_k, cache_v ([Bmax][S][H][hd]
 * fp32, contiguous)}; layer_eps[l*2 + {0,1}] the two RMSNorm eps; wlm_t = lm_headᵀ [V][dim]. B <= 8, head size 32 / 48 / 64,
 * dim and FF multiples of 4 and <= 1024, S <= 2048; an unsupported model is an error (PDN_ERR_INVALID), never a silent fallback.
 * The handle owns its scratch (residual rows, attention partials, grid barrier) until pdn_decoder_destroy. */
int pdn_decoder_create(void** handle, int n_layers, int B, int dim, int H, int FF, int V, int S, const void* const* layer_ptrs,
                       const float* layer_eps, const float* emb, const float* cosT, const float* sinT, const float* norm_w, float eps_f,
                       const float* wlm_t, const float* lm_bias);
/* One token for each of the B sequences: ids[b * ids_stride] at position pos -> KV cache updated at [b, pos], logits [B][V]
 * and ids_out[b] = argmax (first occurrence). logits == ids_out == NULL: cache update only (prompt positions before the last). */
int pdn_decoder_step(void* handle, const int64_t* ids, int64_t ids_stride, int64_t pos, float* logits, int64_t* ids_out);
int pdn_decoder_destroy(void* handle);

/* ---------------------------------------------------------------- data-parallel comm -------- */
/* Not in the reference (single process); defined by BASELINE north_star: NCCL all-reduce of the flat
 * parameter-gradient bucket over NVLink. */
int pdn_nccl_unique_id(char* id128);
int pdn_nccl_init(int rank, int world, const char* id128);
int pdn_nccl_world(int* rank, int* world);
int pdn_allreduce_sum_f32(float* buf, int64_t n);      /* on the comm stream, ordered after compute stream */
int pdn_allreduce_wait(void);                          /* compute stream waits for the comm stream */
int pdn_allreduce_sum_f32_inline(float* buf, int64_t n); /* small reductions on the compute stream itself */
int pdn_nccl_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* PDN_B200_H */
