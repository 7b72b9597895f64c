// gemm_tc.cu — fp32-parity GEMM on the Blackwell tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA into 128B-swizzled shared memory, mbarrier producer/consumer pipeline).
//
// fp32 parity (<=1e-4 normwise vs the reference NumPy sgemm) comes from a 2-term BF16 split
// of each operand: x = hi + lo, C = Ahi·Bhi + Ahi·Blo + Alo·Bhi accumulated in fp32 TMEM (3 BF16 MMAs per
// product, ~16 effective mantissa bits, 4.5e-6 measured normwise error in emulation).
//
// Pieces:
//   k_pack_split(_t)  fp32 operand with arbitrary (row, col, batch) element strides -> bf16 planes [batch][hi|lo][rows][Cp] in the
//                     operand's OWN orientation (strided / broadcast batch views of `@`, tensor.py:657-676, are absorbed here)
//   planes_cached     operand-plane cache: one pack per source buffer and write-version, shared by every product that reads it
//   k_gemm_tc         persistent CTAs walking 128xBN output tiles; warp 0 = TMA producer, warp 1 = MMA issuer (elected lane,
//                     warp-uniform operands), warp 2 = TMEM allocator, warps 4-7 = epilogue over two TMEM accumulators. Each operand
//                     is read K-major (contraction axis contiguous) or MN-major (rows contiguous: W [K][N] in x @ W, x [M][K] in
//                     x^T @ g) — a descriptor flag, never a transposed copy. GATHER = 1/2: implicit-GEMM producer warps for the
//                     convolutions conv_tma.cu does not take (stride > 1, < 16 channels).
#include "common.cuh"
#include "gemm_args.h"
#include "gemm_tc.h"
#include "tc_ptx.cuh"
#include "conv_gather.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

namespace pdn {


// ------------------------------------------------------------------ operand packing -----------
struct PackArgs {
  const float* src;
  __nv_bfloat16* dst;      // [pbatch][2][R][Kp]
  int64_t R, K, Kp;
  int64_t r_stride, k_stride;  // element strides of the source along rows / k
  int64_t k_inner, k_outer_stride;  // k = ko * k_inner + ki -> offset ko * k_outer_stride + ki * k_stride (k_inner = K: flat)
  int64_t nb[3], bs[3];        // source batch shape / strides (only dims with stride != 0 are walked)
};

__global__ void __launch_bounds__(256) k_pack_split(PackArgs p) {
  // 32 (rows) x 64 (k) tile through shared memory: global reads follow the source's unit-stride axis, plane writes are
  // bf16x2 along k (one 128-byte segment per warp and row)
  __shared__ float tile[32][65];
  // blockIdx.z enumerates the operand's own distinct batches
  int64_t z = blockIdx.z, off = 0;
  {
    int64_t rem = z;
    for (int d = 2; d >= 0; --d) {
      int64_t n = (p.bs[d] != 0 && p.nb[d] > 1) ? p.nb[d] : 1;
      off += (rem % n) * p.bs[d];
      rem /= n;
    }
  }
  const float* src = p.src + off;
  __nv_bfloat16* hi = p.dst + (size_t)z * 2 * p.R * p.Kp;
  __nv_bfloat16* lo = hi + (size_t)p.R * p.Kp;
  const int64_t r0 = (int64_t)blockIdx.y * 32, k0 = (int64_t)blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const bool k_fast = (p.k_stride == 1) || (p.r_stride != 1);
  const bool flat_k = p.k_inner >= p.K;
  auto koff = [&](int64_t k) { return flat_k ? k * p.k_stride : (k / p.k_inner) * p.k_outer_stride + (k % p.k_inner) * p.k_stride; };
  if (p.k_stride == 1 && flat_k && (p.r_stride & 3) == 0 && ((((uintptr_t)src) & 15) == 0) && k0 + 64 <= p.K) {
    // contiguous k: the 32 x 64 tile is 32 rows of 16 float4 — half a warp per row, 128-bit loads
    const int c4 = threadIdx.x & 15, rr = threadIdx.x >> 4;  // 16 x 16
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t r = r0 + rr + i * 16;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < p.R) v = __ldg(reinterpret_cast<const float4*>(src + r * p.r_stride + k0) + c4);
      float* t = &tile[rr + i * 16][c4 * 4];
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
  } else if (k_fast) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + ty + i * 8;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t k = k0 + tx + h * 32;
        tile[ty + i * 8][tx + h * 32] = (r < p.R && k < p.K) ? src[r * p.r_stride + koff(k)] : 0.f;
      }
    }
  } else {  // rows are the unit-stride axis of the source: read along rows, transpose through smem
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = r0 + tx, k = k0 + ty + i * 8;
      tile[tx][ty + i * 8] = (r < p.R && k < p.K) ? src[r * p.r_stride + koff(k)] : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + i * 8, k = k0 + 2 * tx;
    if (r < p.R && k < p.Kp) {  // Kp is even
      const float x0 = tile[ty + i * 8][2 * tx], x1 = tile[ty + i * 8][2 * tx + 1];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
      __nv_bfloat162 H2, L2;
      H2.x = h0; H2.y = h1;
      L2.x = __float2bfloat16_rn(x0 - __bfloat162float(h0));
      L2.y = __float2bfloat16_rn(x1 - __bfloat162float(h1));
      *reinterpret_cast<__nv_bfloat162*>(hi + r * p.Kp + k) = H2;
      *reinterpret_cast<__nv_bfloat162*>(lo + r * p.Kp + k) = L2;
    }
  }
}

// Streaming form for the common training operand — one batch, k contiguous, K a multiple of 4 (= Kp), 16-byte aligned rows: a thread
// converts one float4 into 4 hi + 4 lo bf16 (two packed cvt.rn.bf16x2 each) and stores 8 bytes per plane, a warp 256 contiguous
// bytes per plane; no shared-memory tile, no barrier (the tiled kernel above ran at 80 % of HBM peak on [65536, 512]).
__global__ void __launch_bounds__(256) k_pack_split_rows4(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                          int64_t R, int64_t K4, int64_t r_stride, int64_t Kp) {
  const int64_t n = R * K4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / K4, c = i - r * K4;
    const float4  v = __ldcs(reinterpret_cast<const float4*>(src + r * r_stride) + c);
    uint32_t h01, h23, l01, l23;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h01) : "f"(v.y), "f"(v.x));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h23) : "f"(v.w), "f"(v.z));
    const float r0 = v.x - __uint_as_float(h01 << 16), r1 = v.y - __uint_as_float(h01 & 0xffff0000u);
    const float r2 = v.z - __uint_as_float(h23 << 16), r3 = v.w - __uint_as_float(h23 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l01) : "f"(r1), "f"(r0));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l23) : "f"(r3), "f"(r2));
    *reinterpret_cast<uint2*>(hi + r * Kp + 4 * c) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(lo + r * Kp + 4 * c) = make_uint2(l01, l23);
  }
}

// Transposing form for sources whose ROWS are the unit-stride axis (channels-last conv planes from NCHW activations): 64 (rows) x 64
// (k) tile, 128-bit loads along the rows, 16 bf16 (two 128-bit stores) per thread and plane along k. The generic kernel above moves
// 4 bytes per thread access in this orientation (2.3 TB/s).
__global__ void __launch_bounds__(256) k_pack_split_t(PackArgs p) {
  __shared__ float tile[64][65];  // [k][r]
  int64_t z = blockIdx.z, off = 0;
  {
    int64_t rem = z;
    for (int d = 2; d >= 0; --d) {
      int64_t n = (p.bs[d] != 0 && p.nb[d] > 1) ? p.nb[d] : 1;
      off += (rem % n) * p.bs[d];
      rem /= n;
    }
  }
  const float* src = p.src + off;
  __nv_bfloat16* hi = p.dst + (size_t)z * 2 * p.R * p.Kp;
  __nv_bfloat16* lo = hi + (size_t)p.R * p.Kp;
  const int64_t r0 = (int64_t)blockIdx.y * 64, k0 = (int64_t)blockIdx.x * 64;
  {
    const int r4 = threadIdx.x & 15, kk = threadIdx.x >> 4;
    const int64_t r = r0 + 4 * r4;
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      const int     kl = kk + 16 * ps;
      const int64_t k = k0 + kl;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < p.K) {
        const float* q = src + k * p.k_stride + r;
        if (r + 3 < p.R) v = __ldg(reinterpret_cast<const float4*>(q));
        else {
          if (r < p.R) v.x = __ldg(q);
          if (r + 1 < p.R) v.y = __ldg(q + 1);
          if (r + 2 < p.R) v.z = __ldg(q + 2);
        }
      }
      float* t = &tile[kl][4 * r4];
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
  }
  __syncthreads();
  const int rl = threadIdx.x >> 2, kq = threadIdx.x & 3;
  const int64_t r = r0 + rl;
  if (r >= p.R) return;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t k = k0 + kq * 16 + 8 * h;
    if (k >= p.Kp) break;  // Kp is a multiple of 8
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float x0 = tile[kq * 16 + 8 * h + 2 * u][rl], x1 = tile[kq * 16 + 8 * h + 2 * u + 1][rl];
      const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
      const float2         hf = __bfloat1622float2(hh);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
      hw[u] = *reinterpret_cast<const uint32_t*>(&hh);
      lw[u] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(hi + r * p.Kp + k) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo + r * p.Kp + k) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// ------------------------------------------------------------------ the MMA kernel -------------
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // 64 bf16 = 128 B = one swizzle span

template <int BN>
struct TcCfg {
  static constexpr int kStageBytes = 2 * (TC_BM * TC_BK * 2) + 2 * (BN * TC_BK * 2);  // Ahi, Alo, Bhi, Blo
  static constexpr int kStages = (BN == 64) ? 4 : ((BN == 128) ? 3 : 2);
  static constexpr int kEpiBytes = 4 * 32 * 32 * 4;  // per-epilogue-warp 32x32 fp32 transpose tile (XOR-swizzled float4 chunks)
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;  // two fp32 accumulators: the epilogue of tile i overlaps the MMAs of tile i+1
};

// Persistent kernel: each CTA walks output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; tile index decodes to
// (n_blk fastest, m_blk, batch/split) so CTAs running at the same time share A rows in L2.
//   warp 0     TMA producer   (smem ring: full/empty mbarriers, runs ahead across tile boundaries)
//   warp 1     MMA issuer     (one elected lane; accumulator ping-pong in TMEM: tmem_full/tmem_empty mbarriers)
//   warp 2     TMEM allocator
//   warps 4-7  epilogue       (tcgen05.ld -> +bias -> smem transpose -> coalesced stores | NCHW scatter | split-K atomics | argmax)
// GATHER = 0: A tiles arrive by TMA from packed planes. GATHER = 1 / 2: implicit-GEMM convolution — 8 extra producer warps
// gather the im2col (1) or transposed-conv (2) rows of the A tile straight from the fp32 NCHW tensor, split them into bf16
// hi/lo and write them in the swizzled layout the MMA expects; the column matrix never exists in HBM.
struct GatherArgs {
  const float* src;
  ConvGeom     geom;
  int64_t      Mtot;
  int          Ktot;
  const int2*  tab;  // per column kk of the im2col matrix: {channel * plane_size, (ky << 16) | kx}, built once per call
};

// kk -> (channel plane offset, window offsets): hoists three divisions out of the gather loop (the first ncu capture of the
// producer warps showed ~70 instructions per gathered element and instruction-fetch stalls, profiles/r1_ncu_summary.md)
__global__ void k_conv_table(int2* tab, int Kpad, int Ktot, int k, int plane) {
  const int kk = blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= Kpad) return;
  const int kx = kk % k, t = kk / k, ky = t % k, ch = t / k;
  tab[kk] = kk < Ktot ? make_int2(ch * plane, (ky << 16) | kx) : make_int2(0, 0x7fff7fff);  // out of range -> fails the bounds test
}

template <int BN, int GATHER>
__global__ void __launch_bounds__(GATHER ? 512 : 256, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcArgs g, int n_tiles_n, int n_tiles_m,
          long long total_tiles, GatherArgs ga) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t*  smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float*    epi_smem = (float*)(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = (uint64_t*)(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kEpiBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;  // [2] accumulator ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;            // [2] accumulator drained by the epilogue
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const bool tr = g.trace != nullptr && blockIdx.x == 0 && lane == 0;
#define TC_STAMP(i) do { if (tr) g.trace[i] = clock64(); } while (0)
  if (warp == 0) TC_STAMP(0);
  const int all_kb = (int)((g.K + TC_BK - 1) / TC_BK);
  const int kb_per = (all_kb + g.splits - 1) / g.splits;  // the host guarantees every split owns >= 1 k-block

  if (warp == 0 && lane == 0) {
    if (!GATHER) tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], GATHER ? 9 : 1);  // TMA producer (+ 8 gather warps)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) TC_STAMP(1);

  // tile -> coordinates
  // 32-bit arithmetic (the host guarantees total_tiles < 2^31): 64-bit divisions cost the producer ~2000 cycles before its first TMA
  auto decode = [&](uint32_t t, int& n_blk, int& m_blk, int& split, int64_t& i0, int64_t& i1, int64_t& i2) {
    const uint32_t tn = (uint32_t)n_tiles_n, tm = (uint32_t)n_tiles_m;
    uint32_t r = t / tn;
    n_blk = (int)(t - r * tn);
    t = r; r = t / tm;
    m_blk = (int)(t - r * tm);
    t = r;
    if (g.splits > 1) { r = t / (uint32_t)g.splits; split = (int)(t - r * (uint32_t)g.splits); t = r; } else split = 0;
    if (t == 0) { i0 = i1 = i2 = 0; return; }
    const uint32_t n2 = (uint32_t)g.nb[2], n1 = (uint32_t)g.nb[1];
    r = t / n2; i2 = t - r * n2; t = r;
    r = t / n1; i1 = t - r * n1;
    i0 = r;
  };

  if (warp == 0) {
    // ===== TMA producer (the whole warp walks the loop, the elected lane issues: operands stay in uniform registers) =====
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (uint32_t t = blockIdx.x; t < (uint32_t)total_tiles; t += gridDim.x) {
      int n_blk, m_blk, split; int64_t i0, i1, i2;
      decode(t, n_blk, m_blk, split, i0, i1, i2);
      const int a_batch = (int)(i0 * g.a_pbs[0] + i1 * g.a_pbs[1] + i2 * g.a_pbs[2]);
      const int b_batch = (int)(i0 * g.b_pbs[0] + i1 * g.b_pbs[1] + i2 * g.b_pbs[2]);
      const int kb_begin = split * kb_per;
      const int num_kb = min(all_kb, kb_begin + kb_per) - kb_begin;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* st = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[stage], GATHER ? 2 * (BN * TC_BK * 2) : Cfg::kStageBytes);
          const int k = (kb_begin + kb) * TC_BK;
          if (!GATHER) {
            if (!g.a_mn) {
              tma_load_4d(&mapA, &full_bar[stage], st, k, m_blk * TC_BM, 0, a_batch);
              tma_load_4d(&mapA, &full_bar[stage], st + TC_BM * TC_BK * 2, k, m_blk * TC_BM, 1, a_batch);
            } else {  // MN-major: [64 k rows][64 operand rows = 128 B] blocks, 8 KB each
#pragma unroll
              for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                for (int i = 0; i < TC_BM / 64; ++i)
                  tma_load_4d(&mapA, &full_bar[stage], st + pl * (TC_BM * TC_BK * 2) + i * 8192, m_blk * TC_BM + 64 * i, k, pl, a_batch);
            }
          }
          uint8_t* sb = st + 2 * TC_BM * TC_BK * 2;
          if (!g.b_mn) {
            tma_load_4d(&mapB, &full_bar[stage], sb, k, n_blk * BN, 0, b_batch);
            tma_load_4d(&mapB, &full_bar[stage], sb + BN * TC_BK * 2, k, n_blk * BN, 1, b_batch);
          } else {
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_4d(&mapB, &full_bar[stage], sb + pl * (BN * TC_BK * 2) + i * 8192, n_blk * BN + 64 * i, k, pl, b_batch);
          }
        }
        if (kb == 0 && t == blockIdx.x) TC_STAMP(2);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (the whole warp walks the loop, the elected lane issues) =====
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(TC_BM, BN) | (g.a_mn ? IDESC_A_MN_MAJOR : 0u) | (g.b_mn ? IDESC_B_MN_MAJOR : 0u);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (uint32_t t = blockIdx.x; t < (uint32_t)total_tiles; t += gridDim.x) {
      int n_blk, m_blk, split; int64_t i0, i1, i2;
      decode(t, n_blk, m_blk, split, i0, i1, i2);
      const int kb_begin = split * kb_per;
      const int num_kb = min(all_kb, kb_begin + kb_per) - kb_begin;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // the epilogue has drained this accumulator (passes on first use)
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (kb == 0 && t == blockIdx.x) TC_STAMP(3);
        if (leader) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes), sb = sa + 2 * TC_BM * TC_BK * 2;
          const uint64_t d_ahi = g.a_mn ? make_smem_desc_sw128_mn(sa, 8192) : make_smem_desc_sw128(sa);
          const uint64_t d_alo = g.a_mn ? make_smem_desc_sw128_mn(sa + TC_BM * TC_BK * 2, 8192) : make_smem_desc_sw128(sa + TC_BM * TC_BK * 2);
          const uint64_t d_bhi = g.b_mn ? make_smem_desc_sw128_mn(sb, 8192) : make_smem_desc_sw128(sb);
          const uint64_t d_blo = g.b_mn ? make_smem_desc_sw128_mn(sb + BN * TC_BK * 2, 8192) : make_smem_desc_sw128(sb + BN * TC_BK * 2);
          // one UMMA_K = 16 step: 32 B inside the swizzle span (K-major) or 16 rows of 128 B (MN-major), in 16-byte units
          const uint64_t a_step = g.a_mn ? 128 : 2, b_step = g.b_mn ? 128 : 2;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t aa = a_step * k, bb = b_step * k;
            // small cross terms first, then the leading term
            umma_bf16(tmem_d, d_alo + aa, d_bhi + bb, idesc, (kb | k) ? 1u : 0u);
            umma_bf16(tmem_d, d_ahi + aa, d_blo + bb, idesc, 1u);
            umma_bf16(tmem_d, d_ahi + aa, d_bhi + bb, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&tfull_bar[acc]);  // accumulator complete
      if (t == blockIdx.x) TC_STAMP(4);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue: TMEM -> registers -> (+bias) -> {swizzled smem transpose -> 128-bit coalesced stores | NCHW | argmax} =====
    // Everything per element is 32-bit and predicate-free on full 32-column chunks; the first timeline trace of this kernel
    // (PDN_TC_TRACE) showed the previous scalar epilogue at ~2100 cycles per 32-column chunk — more than the MMAs of a K = 288 tile.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    float*    stg = epi_smem + q * (32 * 32);  // 32 rows x 32 floats, float4 chunk c of row r at r*32 + ((c ^ (r & 7)) << 2)
    int       acc = 0;
    uint32_t  acc_phase = 0;
    const bool atomic = g.splits > 1;
    const int  N = (int)g.N;
    const bool bias_vec = (((uintptr_t)g.bias) & 15) == 0;
    for (uint32_t t = blockIdx.x; t < (uint32_t)total_tiles; t += gridDim.x) {
      int n_blk, m_blk, split; int64_t i0, i1, i2;
      decode(t, n_blk, m_blk, split, i0, i1, i2);
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
      const int64_t row0 = (int64_t)m_blk * TC_BM + q * 32;
      const int     m_rem = (int)((g.M - row0) < 32 ? (g.M - row0) : 32);  // valid rows of this warp's quarter (may be <= 0)
      float*        cbase = g.C + i0 * g.c_bs[0] + i1 * g.c_bs[1] + i2 * g.c_bs[2];
      const bool    c_vec = ((g.ldc & 3) == 0) && ((((uintptr_t)cbase) & 15) == 0);
      const bool    add_bias = g.bias != nullptr && split == 0;
      float*        nchw_p0 = nullptr;
      if (g.nchw_hw > 0 && lane < m_rem) {
        const int64_t row = row0 + lane, img = row / g.nchw_hw;
        nchw_p0 = cbase + img * g.N * g.nchw_hw + (row - img * g.nchw_hw);
      }
      float best_v = -INFINITY;
      int   best_i = n_blk * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (warp == 4 && t == blockIdx.x) TC_STAMP(5);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int col0 = n_blk * BN + c0;
        if (col0 >= N) break;  // uniform across the CTA
        const int ncol = (N - col0) < 32 ? (N - col0) : 32;
        float v[32];
        const bool tr0 = warp == 4 && t == blockIdx.x && c0 == 0;
        if (tr0) TC_STAMP(9);
        tmem_ld_32x32(tmem_d + (uint32_t)c0, v);
        tmem_ld_wait();
        if (tr0) TC_STAMP(10);
        if (c0 + 32 >= BN || col0 + 32 >= N) {
          // last chunk of this tile is in registers: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        if (add_bias) {
          if (ncol == 32 && bias_vec) {  // col0 is a multiple of 32: 8 broadcast 128-bit loads
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0) + c);
              v[4 * c] += b4.x; v[4 * c + 1] += b4.y; v[4 * c + 2] += b4.z; v[4 * c + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) v[j] += __ldg(g.bias + col0 + j);
          }
        }
        if (g.amax_val) {
          // running (max, first column) of this row over the tile's columns; nothing is stored to C
          if (ncol == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (v[j] > best_v) { best_v = v[j]; best_i = col0 + j; }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol && v[j] > best_v) { best_v = v[j]; best_i = col0 + j; }
          }
        } else if (g.nchw_hw > 0) {
          // NCHW scatter: for a fixed column (channel) consecutive lanes hold consecutive pixels -> coalesced
          if (nchw_p0) {
            float* p = nchw_p0 + (int64_t)col0 * g.nchw_hw;
            if (atomic) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncol) atomicAdd(p + (int64_t)j * g.nchw_hw, v[j]);
            } else if (g.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncol) p[(int64_t)j * g.nchw_hw] += v[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncol) p[(int64_t)j * g.nchw_hw] = v[j];
            }
          }
        } else {
          // row-major C: a thread owns a ROW of the accumulator; transpose the 32x32 chunk through XOR-swizzled shared memory
          // (8 conflict-free 128-bit stores) so that every global access is a 128-bit piece of a fully covered 128-byte row segment
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          __syncwarp();
          if (tr0) TC_STAMP(11);
          if (ncol == 32 && c_vec) {
            const int c = lane & 7, rs = lane >> 3;  // lane -> (float4 column chunk, row within a group of 4 rows)
            float*    p = cbase + (row0 + rs) * g.ldc + col0 + 4 * c;
            const int64_t step = 4 * g.ldc;
            float4    x[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + rs;
              x[it] = *reinterpret_cast<const float4*>(stg + r * 32 + ((c ^ (r & 7)) << 2));
            }
            if (atomic) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (it * 4 + rs < m_rem) atomicAdd(reinterpret_cast<float4*>(p + it * step), x[it]);
            } else {
              if (g.accumulate) {  // all loads of the old C values are issued before the first store
                float4 o[8];
#pragma unroll
                for (int it = 0; it < 8; ++it)
                  o[it] = (it * 4 + rs < m_rem) ? __ldcg(reinterpret_cast<const float4*>(p + it * step)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int it = 0; it < 8; ++it) { x[it].x += o[it].x; x[it].y += o[it].y; x[it].z += o[it].z; x[it].w += o[it].w; }
              }
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (it * 4 + rs < m_rem) *reinterpret_cast<float4*>(p + it * step) = x[it];
            }
          } else if (lane < ncol) {  // edge chunk or unaligned C: scalar, lane = column
            float* p = cbase + row0 * g.ldc + col0 + lane;
            const int sw = lane >> 2, lo = lane & 3;
#pragma unroll 4
            for (int r = 0; r < m_rem; ++r) {
              const float x = stg[r * 32 + ((sw ^ (r & 7)) << 2) + lo];
              if (atomic) atomicAdd(p + (int64_t)r * g.ldc, x);
              else if (g.accumulate) p[(int64_t)r * g.ldc] += x;
              else p[(int64_t)r * g.ldc] = x;
            }
          }
          __syncwarp();
        }
        if (tr0) TC_STAMP(12);
      }
      if (g.amax_val && lane < m_rem) {
        g.amax_val[(row0 + lane) * n_tiles_n + n_blk] = best_v;
        g.amax_idx[(row0 + lane) * n_tiles_n + n_blk] = (long long)best_i;
      }
      if (warp == 4 && t == blockIdx.x) TC_STAMP(6);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  if (GATHER && warp >= 8) {
    // ===== implicit-GEMM A producer: thread = (tile row r, 32-column half of the 64-wide k-block) =====
    const int gw = warp - 8, r = (gw & 3) * 32 + lane, half = gw >> 2;
    int stage = 0;
    uint32_t phase = 0;
    for (uint32_t t = blockIdx.x; t < (uint32_t)total_tiles; t += gridDim.x) {
      int n_blk, m_blk, split; int64_t i0, i1, i2;
      decode(t, n_blk, m_blk, split, i0, i1, i2);
      const int kb_begin = split * kb_per;
      const int num_kb = min(all_kb, kb_begin + kb_per) - kb_begin;
      const RowPos rp = (GATHER == 1) ? conv_row<0>(ga.geom, (int64_t)m_blk * TC_BM + r, ga.Mtot) : conv_row<1>(ga.geom, (int64_t)m_blk * TC_BM + r, ga.Mtot);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int kk0 = (kb_begin + kb) * TC_BK + half * 32;
        // all 32 loads are issued before the first conversion (ILP hides the latency); per element: one broadcast table read,
        // two unsigned bounds tests, one 32-bit offset
        float v[32];
        const float* base = ga.src + rp.base;
        const int2*  tb = ga.tab + kk0;
        const int    Hs = (GATHER == 1) ? (int)ga.geom.H : (int)ga.geom.oh, Ws = (GATHER == 1) ? (int)ga.geom.W : (int)ga.geom.ow;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int2 te = __ldg(tb + e);
          const int  ky = te.y >> 16, kx = te.y & 0xffff;
          int yy, xx;
          bool ok = rp.ok;
          if (GATHER == 1) {
            yy = rp.y + ky; xx = rp.x + kx;
          } else {
            yy = rp.y - ky; xx = rp.x - kx;
            if (ga.geom.stride != 1) {
              ok = ok && yy >= 0 && xx >= 0 && (yy % ga.geom.stride == 0) && (xx % ga.geom.stride == 0);
              yy /= ga.geom.stride; xx /= ga.geom.stride;
            }
          }
          ok = ok && (unsigned)yy < (unsigned)Hs && (unsigned)xx < (unsigned)Ws;
          v[e] = ok ? __ldg(base + (te.x + yy * Ws + xx)) : 0.f;
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* ah = smem + stage * Cfg::kStageBytes;
        uint8_t* al = ah + TC_BM * TC_BK * 2;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float v0 = v[c8 * 8 + 2 * u], v1 = v[c8 * 8 + 2 * u + 1];
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
            hw[u] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lw[u] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          const int      c = half * 4 + c8;
          const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));  // 128-byte swizzle, K-major
          *reinterpret_cast<uint4*>(ah + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(al + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) TC_STAMP(7);
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
  if (warp == 2) TC_STAMP(8);
#undef TC_STAMP
}

// second stage of the fused GEMM + argmax: one warp per row over the per-tile partials (tiles are in column order, so the
// smallest tile index wins ties = NumPy's first-occurrence rule)
__global__ void __launch_bounds__(256) k_argmax_partials(const float* __restrict__ val, const long long* __restrict__ idx, long long* __restrict__ out,
                                                         int64_t M, int nt) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float     bv = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (int t = lane; t < nt; t += 32) {
    float v = val[row * nt + t];
    long long i = idx[row * nt + t];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float     ov = __shfl_xor_sync(0xffffffffu, bv, o);
    long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) out[row] = bi;
}

// ------------------------------------------------------------------ host side ------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int get_encode_fn() {
  if (g_encode) return 0;
  void*                           fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver (err %d, q %d)", (int)e, (int)qres);
    return PDN_ERR_CUDA;
  }
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  return 0;
}

// 4-D map over packed planes [batch][2][R][Kp] (bf16), box = 64 (k) x box_rows x 1 x 1, 128B swizzle
static int make_map(CUtensorMap* map, const void* base, int64_t R, int64_t K, int64_t Kp, int64_t nbatch, int box_rows) {
  cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)R, 2, (cuuint64_t)nbatch};
  cuuint64_t strides[3] = {(cuuint64_t)Kp * 2, (cuuint64_t)R * Kp * 2, (cuuint64_t)R * Kp * 4};
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult   r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (R=%lld K=%lld Kp=%lld batch=%lld)", (int)r, (long long)R, (long long)K,
              (long long)Kp, (long long)nbatch);
    return PDN_ERR_CUDA;
  }
  return 0;
}

int tc_make_map(CUtensorMap* map, const void* base, int64_t R, int64_t K, int64_t Kp, int64_t nbatch, int box_rows) {
  PDN_TRY(get_encode_fn());
  return make_map(map, base, R, K, Kp, nbatch, box_rows);
}

bool gemm_tc_eligible(const GemmArgs& g) {
  if (g.M < 64 || g.N < 32 || g.K < 32) return false;
  double work = (double)g.M * (double)g.N * (double)g.K;
  if (work < (double)(1 << 21)) return false;  // tiny products: launch-bound, FFMA tile kernel is fine
  int64_t nbatch = g.nb[0] * g.nb[1] * g.nb[2];
  if (nbatch > 65535) return false;
  if ((g.M + TC_BM - 1) / TC_BM > 65535) return false;
  if (g.M > 0x7fffffff || g.N > 0x7fffffff || g.K > 0x7fffffff) return false;
  return true;
}

// planes of `src` written to `dst` (or to a fresh allocation of `buf` when dst == nullptr)
static int pack_operand_to(const float* src, int64_t R, int64_t K, int64_t r_stride, int64_t k_stride, int64_t k_inner, int64_t k_outer_stride,
                           const int64_t* nb, const int64_t* bs, Scratch* buf, void* dst, PackedOperand* out) {
  int64_t Kp = (K + 7) & ~(int64_t)7;
  int64_t pb = 1;
  PackArgs p;
  for (int d = 2; d >= 0; --d) {
    bool walk = (bs[d] != 0 && nb[d] > 1);
    out->pbs[d] = walk ? pb : 0;
    if (walk) pb *= nb[d];
    p.nb[d] = nb[d];
    p.bs[d] = bs[d];
  }
  PDN_CHECK(pb <= 65535, "gemm_tc: too many operand batches");
  if (!dst) {
    PDN_TRY(buf->alloc((size_t)pb * 2 * R * Kp * sizeof(__nv_bfloat16)));
    dst = buf->p;
  }
  p.src = src;
  p.dst = (__nv_bfloat16*)dst;
  p.R = R; p.K = K; p.Kp = Kp;
  p.r_stride = r_stride; p.k_stride = k_stride;
  p.k_inner = k_inner > 0 ? k_inner : (K > 0 ? K : 1);
  p.k_outer_stride = k_outer_stride;
  bool bs_al = true;
  for (int d = 0; d < 3; ++d) bs_al = bs_al && (bs[d] & 3) == 0;
  static const bool rows4_off = getenv("PDN_PACK_TILED") != nullptr;
  if (!rows4_off && pb == 1 && k_stride == 1 && p.k_inner >= K && (K & 3) == 0 && Kp == K && (r_stride & 3) == 0 && ((((uintptr_t)src) & 15) == 0) &&
      R * (K / 4) >= 4096) {
    __nv_bfloat16* hi = p.dst;
    k_pack_split_rows4<<<grid_for(R * (K / 4), 256, 2), 256, 0, stream()>>>(src, hi, hi + (size_t)R * Kp, R, K / 4, r_stride, Kp);
    PDN_LAUNCHED("pack_split_rows4");
  } else if (r_stride == 1 && k_stride != 1 && p.k_inner >= K && (k_stride & 3) == 0 && bs_al && ((((uintptr_t)src) & 15) == 0) && R >= 16 &&
      (R + 63) / 64 <= 65535) {
    dim3 grd((unsigned)((Kp + 63) / 64), (unsigned)((R + 63) / 64), (unsigned)pb);
    k_pack_split_t<<<grd, 256, 0, stream()>>>(p);
    PDN_LAUNCHED("pack_split_t");
  } else {
    dim3 grd((unsigned)((Kp + 63) / 64), (unsigned)((R + 31) / 32), (unsigned)pb);
    PDN_CHECK(grd.y <= 65535, "gemm_tc: operand has too many rows for the pack grid");
    k_pack_split<<<grd, 256, 0, stream()>>>(p);
    PDN_LAUNCHED("pack_split");
  }
  out->planes = dst; out->R = R; out->K = K; out->Kp = Kp; out->nbatch = pb;
  return 0;
}

int pack_operand_ex(const float* src, int64_t R, int64_t K, int64_t r_stride, int64_t k_stride, int64_t k_inner, int64_t k_outer_stride,
                    const int64_t* nb, const int64_t* bs, Scratch* buf, PackedOperand* out) {
  return pack_operand_to(src, R, K, r_stride, k_stride, k_inner, k_outer_stride, nb, bs, buf, nullptr, out);
}

// ------------------------------------------------------------------ operand-plane cache --------
// Training re-uses GEMM operands: x feeds the Q/K/V projections and, in backward, dW = x^T g; g feeds dX = g W^T and dW; W feeds the
// forward and dX. With every operand packed in its own orientation (MN-major support above) all these uses read the SAME planes,
// so they are packed once and kept until the source buffer is written (the caller passes its write-version counter) or freed
// (allocator hook). First ncu launch list of the encoder step: k_pack_split = 33 % of GPU time, 48 launches per step.
struct PlaneEntry {
  const float* src; int64_t R, K, rs, ks; int64_t nb[3], bs[3];
  long long version; PackedOperand op; size_t bytes; uint64_t stamp;
};
// heap objects that are never destroyed: device arrays may still be freed (-> plane_cache_drop_range) while the process shuts down
static std::vector<PlaneEntry>& g_planes = *new std::vector<PlaneEntry>();
static std::mutex&              g_planes_mu = *new std::mutex();
static uint64_t                g_planes_clock = 0;
static size_t                  g_planes_bytes = 0;
static uint64_t                g_planes_hits = 0, g_planes_misses = 0;

static size_t plane_cache_cap() {
  static size_t cap = 0;
  if (!cap) {
    const char* e = getenv("PDN_PLANE_CACHE_MB");
    cap = (e ? (size_t)atoll(e) : (size_t)16384) << 20;
  }
  return cap;
}

void plane_cache_drop_range(const void* p, size_t n) {
  if (n == 0) n = 1;
  std::vector<void*> victims;
  {
    std::lock_guard<std::mutex> lk(g_planes_mu);
    if (g_planes.empty()) return;
    const char *lo = (const char*)p, *hi = lo + n;
    for (size_t i = 0; i < g_planes.size();) {
      const char* sp = (const char*)g_planes[i].src;
      if (sp >= lo && sp < hi) {
        victims.push_back(g_planes[i].op.planes);
        g_planes_bytes -= g_planes[i].bytes;
        g_planes[i] = g_planes.back();
        g_planes.pop_back();
      } else {
        ++i;
      }
    }
  }
  for (void* v : victims) dev_free(v);
}

// operand planes of `src` in its own orientation, from the cache when `version` >= 0 (the caller's write counter of the buffer)
int planes_cached(const float* src, int64_t rows, int64_t kc, int64_t r_stride, int64_t k_stride, const int64_t* nb, const int64_t* bs,
                  long long version, Scratch* buf, PackedOperand* out, bool force_kmajor) {
  const bool mn = !force_kmajor && (k_stride != 1 && r_stride == 1 && rows > 1 && kc > 1);
  const int64_t R = mn ? kc : rows, K = mn ? rows : kc, rs = mn ? k_stride : r_stride, ks = mn ? 1 : k_stride;
  static const bool off = getenv("PDN_PLANE_CACHE") != nullptr && atoi(getenv("PDN_PLANE_CACHE")) == 0;
  if (version < 0 || off || is_capturing()) {
    PDN_TRY(pack_operand_to(src, R, K, rs, ks, 0, 0, nb, bs, buf, nullptr, out));
    out->mn = mn ? 1 : 0;
    return 0;
  }
  void* reuse = nullptr;
  std::vector<void*> evict;
  {
    std::lock_guard<std::mutex> lk(g_planes_mu);
    for (auto& e : g_planes) {
      if (e.src != src || e.R != R || e.K != K || e.rs != rs || e.ks != ks) continue;
      bool same = true;
      for (int d = 0; d < 3; ++d) same = same && e.nb[d] == nb[d] && e.bs[d] == bs[d];
      if (!same) continue;
      e.stamp = ++g_planes_clock;
      if (e.version == version) {
        *out = e.op;
        out->mn = mn ? 1 : 0;  // the same planes serve a K-major use (rows x k) and an MN-major use (k x rows) of the buffer
        ++g_planes_hits;
        return 0;
      }
      e.version = version;  // the buffer was written since: refresh the planes in place
      reuse = e.op.planes;
      break;
    }
  }
  ++g_planes_misses;
  if (reuse) {
    PDN_TRY(pack_operand_to(src, R, K, rs, ks, 0, 0, nb, bs, nullptr, reuse, out));
    out->mn = mn ? 1 : 0;
    return 0;
  }
  void* mem = nullptr;
  int64_t pb = 1;
  for (int d = 0; d < 3; ++d)
    if (bs[d] != 0 && nb[d] > 1) pb *= nb[d];
  const size_t bytes = (size_t)pb * 2 * R * ((K + 7) & ~(int64_t)7) * sizeof(__nv_bfloat16);
  if (dev_alloc(&mem, bytes) != 0) {  // no room to keep it: transient planes
    PDN_TRY(pack_operand_to(src, R, K, rs, ks, 0, 0, nb, bs, buf, nullptr, out));
    out->mn = mn ? 1 : 0;
    return 0;
  }
  int r = pack_operand_to(src, R, K, rs, ks, 0, 0, nb, bs, nullptr, mem, out);
  if (r) { dev_free(mem); return r; }
  out->mn = mn ? 1 : 0;
  {
    std::lock_guard<std::mutex> lk(g_planes_mu);
    PlaneEntry e;
    e.src = src; e.R = R; e.K = K; e.rs = rs; e.ks = ks;
    for (int d = 0; d < 3; ++d) { e.nb[d] = nb[d]; e.bs[d] = bs[d]; }
    e.version = version; e.op = *out; e.bytes = bytes; e.stamp = ++g_planes_clock;
    g_planes.push_back(e);
    g_planes_bytes += bytes;
    while ((g_planes_bytes > plane_cache_cap() || g_planes.size() > 256) && g_planes.size() > 1) {  // evict least recently used (never the new one)
      size_t lru = 0;
      for (size_t i = 1; i + 1 < g_planes.size(); ++i)
        if (g_planes[i].stamp < g_planes[lru].stamp) lru = i;
      if (lru + 1 == g_planes.size()) break;
      evict.push_back(g_planes[lru].op.planes);
      g_planes_bytes -= g_planes[lru].bytes;
      g_planes[lru] = g_planes.back();
      g_planes.pop_back();
    }
  }
  // NOTE: planes evicted here may still be read by GEMMs already queued on the stream; the allocator is stream-ordered (single
  // compute stream), so handing the block to later work is safe.
  for (void* v : evict) dev_free(v);
  return 0;
}

int pack_operand_auto(const float* src, int64_t rows, int64_t kc, int64_t r_stride, int64_t k_stride, const int64_t* nb, const int64_t* bs,
                      Scratch* buf, PackedOperand* out) {
  if (k_stride != 1 && r_stride == 1 && rows > 1 && kc > 1) {
    PDN_TRY(pack_operand_ex(src, kc, rows, k_stride, 1, 0, 0, nb, bs, buf, out));
    out->mn = 1;
    return 0;
  }
  PDN_TRY(pack_operand_ex(src, rows, kc, r_stride, k_stride, 0, 0, nb, bs, buf, out));
  out->mn = 0;
  return 0;
}

template <int BN, int GATHER>
static int launch_tc(const CUtensorMap& mA, const CUtensorMap& mB, const TcArgs& t, const GatherArgs& ga) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    PDN_CUDA(cudaFuncSetAttribute(k_gemm_tc<BN, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int64_t n_tiles_n = (t.N + BN - 1) / BN, n_tiles_m = (t.M + TC_BM - 1) / TC_BM;
  const int64_t total = n_tiles_n * n_tiles_m * t.nb[0] * t.nb[1] * t.nb[2] * t.splits;
  PDN_CHECK(n_tiles_n <= 0x7fffffff && n_tiles_m <= 0x7fffffff && total < 0x7fffffff, "gemm_tc: too many tiles");
  const int64_t ctas = total < sm_count() ? total : sm_count();  // persistent: one CTA per SM walks the tile list
  static long long* trace_buf = nullptr;
  static const bool trace_on = getenv("PDN_TC_TRACE") != nullptr;
  TcArgs tt = t;
  tt.trace = nullptr;
  if (trace_on) {
    if (!trace_buf) PDN_CUDA(cudaMalloc(&trace_buf, 16 * sizeof(long long)));
    PDN_CUDA(cudaMemsetAsync(trace_buf, 0, 16 * sizeof(long long), stream()));
    tt.trace = trace_buf;
  }
  k_gemm_tc<BN, GATHER><<<(unsigned)ctas, GATHER ? 512 : 256, Cfg::kSmemBytes, stream()>>>(mA, mB, tt, (int)n_tiles_n, (int)n_tiles_m,
                                                                                             (long long)total, ga);
  PDN_LAUNCHED(GATHER ? "conv_gemm_tc" : "gemm_tc");
  if (trace_on) {
    long long h[16];
    PDN_CUDA(cudaStreamSynchronize(stream()));
    PDN_CUDA(cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[tc trace] BN=%d M=%lld N=%lld K=%lld ctas=%lld tiles=%lld splits=%d acc=%d | cycles from entry: setup %lld, tma0 %lld, full0 %lld, "
                    "mma_done_issue %lld, epi_start %lld, epi_end %lld, all_done %lld, dealloc %lld | chunk0: ld %lld, ld_done %lld, smem %lld, end %lld\n",
            BN, (long long)t.M, (long long)t.N, (long long)t.K, (long long)ctas, (long long)total, t.splits, t.accumulate, h[1] - h[0], h[2] - h[0],
            h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0], h[9] - h[0], h[10] - h[0], h[11] - h[0], h[12] - h[0]);
  }
  return 0;
}

template <int GATHER>
static int launch_bn(int BN, const CUtensorMap& mA, const CUtensorMap& mB, const TcArgs& t, const GatherArgs& ga) {
  if (BN == 256) return launch_tc<256, GATHER>(mA, mB, t, ga);
  if (BN == 128) return launch_tc<128, GATHER>(mA, mB, t, ga);
  return launch_tc<64, GATHER>(mA, mB, t, ga);
}

// C (+)= A·Bᵀ on pre-packed bf16 hi/lo planes. `t` carries C, bias, M/N/K, ldc, batches, accumulate, nchw_hw;
// splits <= 0 picks a split-K factor that fills the SMs when the tile grid is small and K is long.
int gemm_tc_packed(const PackedOperand& A, const PackedOperand& B, TcArgs t, int splits, int* n_tiles_out) {
  PDN_TRY(get_encode_fn());
  PDN_CHECK(t.K > 0 && t.M > 0 && t.N > 0, "gemm_tc: empty product (M=%lld N=%lld K=%lld)", (long long)t.M, (long long)t.N, (long long)t.K);
  const int64_t nbatch = t.nb[0] * t.nb[1] * t.nb[2];
  const int64_t m_tiles = (t.M + TC_BM - 1) / TC_BM;
  auto tiles_for = [&](int bn) { return m_tiles * ((t.N + bn - 1) / bn) * nbatch; };
  // widest N tile that still gives the grid enough CTAs; narrow tiles (and split-K below) for small problems
  const int sms = sm_count();
  int       BN = 64;
  if (getenv("PDN_TC_BN128") != nullptr) BN = 128;
  else if (t.N > 128 && tiles_for(256) >= (sms * 3) / 4) BN = 256;
  else if (t.N > 64 && tiles_for(128) >= sms / 2) BN = 128;
  else if (splits <= 0 && (t.K + TC_BK - 1) / TC_BK >= 16 && t.N > 64) BN = t.N > 128 ? 256 : 128;  // long K: wide tiles, split-K fills the SMs
  for (int i = 0; i < 3; ++i) { t.a_pbs[i] = A.pbs[i]; t.b_pbs[i] = B.pbs[i]; }
  const int64_t tiles = tiles_for(BN);
  const int     num_kb = (int)((t.K + TC_BK - 1) / TC_BK);
  if (splits <= 0) {
    splits = 1;
    if (tiles < sms / 2 && num_kb >= 8) {
      // tiles x splits work items run as ONE wave of persistent CTAs: rounding the split count UP put 8 tiles x 19 splits = 152
      // items on 148 SMs, i.e. a second wave for 4 items (dW = x^T g of the encoder, K = 65536: 0.53 of the ceiling); rounding down
      // keeps every item in the first wave (144 CTAs)
      int64_t want = sms / tiles;
      if (want < 1) want = 1;
      int64_t cap = num_kb / 4;
      splits = (int)(want < cap ? want : cap);
      if (splits < 1) splits = 1;
    }
  }
  if (splits > 1) {  // every split must own at least one k-block (the kernel's barriers assume it)
    const int kb_per = (num_kb + splits - 1) / splits;
    splits = (num_kb + kb_per - 1) / kb_per;
  }
  t.splits = splits;
  if (splits > 1 && !t.accumulate) {
    // atomics accumulate into C: clear the destination region first
    if (t.nchw_hw > 0 || nbatch > 1) {
      PDN_CHECK(t.c_clear_bytes > 0, "gemm_tc: split-K needs the destination extent");
      PDN_CUDA(cudaMemsetAsync(t.C, 0, t.c_clear_bytes, stream()));
    } else if (t.ldc == t.N) {
      PDN_CUDA(cudaMemsetAsync(t.C, 0, (size_t)t.M * t.N * sizeof(float), stream()));
    } else {
      PDN_CUDA(cudaMemset2DAsync(t.C, (size_t)t.ldc * sizeof(float), 0, (size_t)t.N * sizeof(float), (size_t)t.M, stream()));
    }
  }
  if (n_tiles_out) *n_tiles_out = (int)((t.N + BN - 1) / BN);
  CUtensorMap mA, mB;
  t.a_mn = A.mn; t.b_mn = B.mn;
  // MN-major operands are fetched as [64 contraction rows][64 operand rows] boxes of their own row-major planes
  PDN_TRY(make_map(&mA, A.planes, A.R, A.K, A.Kp, A.nbatch, A.mn ? TC_BK : TC_BM));
  PDN_TRY(make_map(&mB, B.planes, B.R, B.K, B.Kp, B.nbatch, B.mn ? TC_BK : BN));
  GatherArgs none{};
  return launch_bn<0>(BN, mA, mB, t, none);
}

// Implicit-GEMM convolution: C[(n,pix), o] = Σ_kk gather(src)[(n,pix), kk] · B[o, kk]; mode 1 = im2col of x (forward),
// mode 2 = transposed-conv gather of the output gradient (backward-data). B = pre-packed planes [rows = N][K].
int gemm_tc_conv(const float* src, const ConvGeom& geom, int mode, int64_t Mtot, int Ktot, const PackedOperand& B, TcArgs t) {
  PDN_TRY(get_encode_fn());
  PDN_CHECK(t.K > 0 && t.M > 0 && t.N > 0 && (mode == 1 || mode == 2), "conv_gemm_tc: bad arguments");
  const int64_t m_tiles = (t.M + TC_BM - 1) / TC_BM;
  const int sms = sm_count();
  int BN = 64;
  if (t.N > 128 && m_tiles * ((t.N + 255) / 256) >= (sms * 3) / 4) BN = 256;
  else if (t.N > 64 && m_tiles * ((t.N + 127) / 128) >= sms / 2) BN = 128;
  for (int i = 0; i < 3; ++i) { t.nb[i] = 1; t.a_pbs[i] = 0; t.b_pbs[i] = B.pbs[i]; }
  t.splits = 1;
  CUtensorMap mB;
  PDN_TRY(make_map(&mB, B.planes, B.R, B.K, B.Kp, B.nbatch, BN));
  GatherArgs ga;
  ga.src = src; ga.geom = geom; ga.Mtot = Mtot; ga.Ktot = Ktot;
  Scratch   stab;
  const int Kpad = ((Ktot + TC_BK - 1) / TC_BK) * TC_BK;
  PDN_TRY(stab.alloc((size_t)Kpad * sizeof(int2)));
  const int64_t plane = mode == 1 ? geom.H * geom.W : geom.oh * geom.ow;
  PDN_CHECK(plane * (mode == 1 ? geom.C : geom.O) < 0x7fffffff && geom.k < 0x7fff, "conv_gemm_tc: image too large for 32-bit offsets");
  k_conv_table<<<(Kpad + 255) / 256, 256, 0, stream()>>>((int2*)stab.p, Kpad, Ktot, geom.k, (int)plane);
  PDN_LAUNCHED("conv_table");
  ga.tab = (const int2*)stab.p;
  return mode == 1 ? launch_bn<1>(BN, mB, mB, t, ga) : launch_bn<2>(BN, mB, mB, t, ga);
}

int gemm_tc_launch(const GemmArgs& g, long long a_version, long long b_version) {
  Scratch       bufA, bufB;
  PackedOperand A, B;
  // Every operand is packed in its OWN memory orientation (no transposing pack): contraction axis unit-stride -> K-major planes
  // [rows][K]; operand rows unit-stride (W [K][N] as B of x @ W, x [M][K] as A of x^T @ g) -> MN-major planes [K][rows].
  // A: rows = M, k along a_cs.  B: rows = N, k along b_rs.
  PDN_TRY(planes_cached((const float*)g.A, g.M, g.K, g.a_rs, g.a_cs, g.nb, g.a_bs, a_version, &bufA, &A));
  PDN_TRY(planes_cached((const float*)g.B, g.N, g.K, g.b_cs, g.b_rs, g.nb, g.b_bs, b_version, &bufB, &B));
  TcArgs t;
  t.C = (float*)g.C;
  t.bias = (const float*)g.bias;
  t.M = g.M; t.N = g.N; t.K = g.K; t.ldc = g.ldc;
  int64_t nbatch = 1;
  for (int i = 0; i < 3; ++i) { t.nb[i] = g.nb[i]; t.c_bs[i] = g.c_bs[i]; nbatch *= g.nb[i]; }
  t.accumulate = g.accumulate;
  t.nchw_hw = 0;
  t.c_clear_bytes = 0;
  t.amax_val = nullptr; t.amax_idx = nullptr;
  // split-K only for single-batch products (batched ones already fill the grid or have strided C)
  return gemm_tc_packed(A, B, t, nbatch > 1 ? 1 : 0);
}

void plane_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* bytes, uint64_t* entries) {
  std::lock_guard<std::mutex> lk(g_planes_mu);
  *hits = g_planes_hits; *misses = g_planes_misses; *bytes = g_planes_bytes; *entries = g_planes.size();
}

}  // namespace pdn

// ------------------------------------------------------------------ pre-packed weights -----------
// Inference keeps a weight matrix's bf16 hi/lo planes resident (packed once) instead of re-packing 4 B/element on every
// GEMM call — for Llama decode this removes one pack launch and one full weight read+write per Linear per token.
namespace pdn {
struct Prepacked {
  Scratch       buf;
  PackedOperand op;
};
}  // namespace pdn

extern "C" {

int pdn_gemm_prepack(const float* B, int64_t K, int64_t N, int64_t b_rs, int64_t b_cs, void** handle) {
  PDN_TRY(pdn::ensure_init());
  PDN_CHECK(K > 0 && N > 0, "prepack: empty matrix");
  auto* h = new pdn::Prepacked();
  const int64_t one[3] = {1, 1, 1}, zero[3] = {0, 0, 0};
  int r = pdn::pack_operand_ex(B, N, K, b_cs, b_rs, 0, 0, one, zero, &h->buf, &h->op);
  if (r) { delete h; return r; }
  *handle = h;
  return 0;
}

int pdn_gemm_prepacked(const float* A, void* handle, float* C, int64_t M, int64_t a_rs, int64_t a_cs, int64_t ldc, const float* bias,
                       int accumulate) {
  PDN_TRY(pdn::ensure_init());
  auto* h = (pdn::Prepacked*)handle;
  PDN_CHECK(h != nullptr, "prepacked: null handle");
  if (M == 0) return 0;
  pdn::Scratch       bufA;
  pdn::PackedOperand Aop;
  const int64_t one[3] = {1, 1, 1}, zero[3] = {0, 0, 0};
  PDN_TRY(pdn::pack_operand_ex(A, M, h->op.K, a_rs, a_cs, 0, 0, one, zero, &bufA, &Aop));
  pdn::TcArgs t;
  t.C = C; t.bias = bias; t.M = M; t.N = h->op.R; t.K = h->op.K; t.ldc = ldc;
  for (int i = 0; i < 3; ++i) { t.nb[i] = 1; t.c_bs[i] = 0; }
  t.accumulate = accumulate; t.nchw_hw = 0; t.c_clear_bytes = 0;
  t.amax_val = nullptr; t.amax_idx = nullptr;
  return pdn::gemm_tc_packed(Aop, h->op, t, 0);
}

// A operand already in plane format [2][M][Kp] (emitted by k_rmsnorm_planes / k_swiglu_rows_planes / k_attention_fwd)
int pdn_gemm_prepacked_planes(const void* A_planes, int64_t M, int64_t Kp, void* handle, float* C, int64_t ldc, const float* bias,
                              int accumulate) {
  PDN_TRY(pdn::ensure_init());
  auto* h = (pdn::Prepacked*)handle;
  PDN_CHECK(h != nullptr && A_planes != nullptr, "prepacked_planes: null operand");
  PDN_CHECK(Kp >= h->op.K && (Kp & 7) == 0, "prepacked_planes: K padding %lld does not cover K=%lld", (long long)Kp, (long long)h->op.K);
  if (M == 0) return 0;
  pdn::PackedOperand Aop;
  Aop.planes = const_cast<void*>(A_planes); Aop.R = M; Aop.K = h->op.K; Aop.Kp = Kp; Aop.nbatch = 1;
  Aop.pbs[0] = Aop.pbs[1] = Aop.pbs[2] = 0;
  pdn::TcArgs t;
  t.C = C; t.bias = bias; t.M = M; t.N = h->op.R; t.K = h->op.K; t.ldc = ldc;
  for (int i = 0; i < 3; ++i) { t.nb[i] = 1; t.c_bs[i] = 0; }
  t.accumulate = accumulate; t.nchw_hw = 0; t.c_clear_bytes = 0;
  t.amax_val = nullptr; t.amax_idx = nullptr;
  return pdn::gemm_tc_packed(Aop, h->op, t, 0);
}

// out_idx[m] = argmax_n (A·B + bias)[m, n] without materialising the [M, N] product (greedy decoding over the vocabulary)
int pdn_gemm_prepacked_planes_argmax(const void* A_planes, int64_t M, int64_t Kp, void* handle, const float* bias, int64_t* out_idx) {
  PDN_TRY(pdn::ensure_init());
  auto* h = (pdn::Prepacked*)handle;
  PDN_CHECK(h != nullptr && A_planes != nullptr, "prepacked_argmax: null operand");
  PDN_CHECK(Kp >= h->op.K && (Kp & 7) == 0, "prepacked_argmax: bad K padding");
  if (M == 0) return 0;
  const int64_t N = h->op.R;
  const int64_t nt_max = (N + 63) / 64;  // upper bound on N tiles whatever BN the launcher picks
  pdn::Scratch sv, si;
  PDN_TRY(sv.alloc((size_t)M * nt_max * sizeof(float)));
  PDN_TRY(si.alloc((size_t)M * nt_max * sizeof(long long)));
  pdn::PackedOperand Aop;
  Aop.planes = const_cast<void*>(A_planes); Aop.R = M; Aop.K = h->op.K; Aop.Kp = Kp; Aop.nbatch = 1;
  Aop.pbs[0] = Aop.pbs[1] = Aop.pbs[2] = 0;
  pdn::TcArgs t;
  t.C = nullptr; t.bias = bias; t.M = M; t.N = N; t.K = h->op.K; t.ldc = N;
  for (int i = 0; i < 3; ++i) { t.nb[i] = 1; t.c_bs[i] = 0; }
  t.accumulate = 0; t.nchw_hw = 0; t.c_clear_bytes = 0;
  t.amax_val = (float*)sv.p; t.amax_idx = (long long*)si.p;
  int n_tiles = 0;
  PDN_TRY(pdn::gemm_tc_packed(Aop, h->op, t, 1, &n_tiles));
  pdn::k_argmax_partials<<<(unsigned)((M + 7) / 8), 256, 0, pdn::stream()>>>((const float*)sv.p, (const long long*)si.p, (long long*)out_idx, M, n_tiles);
  PDN_LAUNCHED("argmax_partials");
  return 0;
}

int pdn_gemm_prepack_free(void* handle) {
  delete (pdn::Prepacked*)handle;
  return 0;
}

}  // extern "C"

