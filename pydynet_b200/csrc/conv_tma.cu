// conv_tma.cu — stride-1 Conv2d forward / backward-data / backward-weight as implicit GEMMs whose operand tiles are fetched by
// TMA straight from channels-last (NHWC) bf16 hi/lo planes of the activation (reference: functional.py:194-281 pads, builds a 6-D
// strided window view, COPIES it into an (N·oh·ow, C·k·k) column matrix and multiplies; backward scatters with xp.add.at).
//
// A convolution is a sum over the k·k taps of shifted 1x1 convolutions. For one tap the operand "pixel patch x 64 channels" is a
// box {64 channels, 16 px, rows} of the channels-last activation, shifted by the tap offset — one 5-D TMA load, with the zero
// padding produced by TMA's out-of-bounds fill (negative pixel coordinates included; the shift must not be on the innermost
// dimension, whose start has to stay 16-byte aligned — probed in tools/probe/tma_probe.cu — hence channels-last planes). The box
// lands in shared memory as [pixel][64 channels = 128 B] with the 128-byte swizzle, which tcgen05.mma reads
//   * as a K-major A tile (pixels = M rows, channels = contraction) in forward / backward-data:
//         Y[pix, o] += X_tap[pix, c] · W_tap[o, c]
//   * as MN-major A and B tiles (pixels = contraction) in backward-weight:
//         dW_tap[o, c] += dY[pix, o]^T · X_tap[pix, c]
// so the column matrix never exists, nothing is gathered by threads, and the same planes of x (forward, backward-weight) and of
// dY (backward-data, backward-weight) serve all three passes. fp32 parity through the BF16x3 split like every other tensor-core
// kernel here. Warp roles as in gemm_tc.cu: warp 0 TMA producer, warp 1 MMA issuer (elected lane), warp 2 TMEM allocator,
// warps 4-7 epilogue; persistent CTAs, two TMEM accumulators.
#include "common.cuh"
#include "conv_gather.cuh"
#include "gemm_tc.h"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

namespace pdn {

enum { CT_F = 0, CT_W = 1, CT_W3 = 2 };  // W3: backward-weight with the 3 taps of one kernel row per tile sharing the dY tile

struct CtArgs {
  float*       out;    // F: NCHW [n_img][n_out][oh][ow];  W: tmp [taps][M = O][N = C] row-major (atomic accumulation, pre-zeroed)
  const float* bias;   // F only (nullable)
  int n_img, n_contr, n_out;  // F: contraction channels, output channels.  W: n_out = GEMM N extent (C), n_contr unused
  int m_rows;                 // W: GEMM M extent (O)
  int oh, ow;                 // pixel grid: F output grid, W contraction grid (= dY grid)
  int TY, TX;                 // F: 8x16 pixel tiles per image; W: 4x16 pixel patches per image
  int k, pad, sign;           // taps = k*k; tap (ky, kx) reads the activation at pixel + sign * (k? - pad)
  int cblks;                  // F: 64-channel blocks of the contraction
  int n_tiles_n, m_tiles;     // GEMM-N tiles (BN wide); W: 128-row tiles of O
  int splits, per_split;      // W: patches of the contraction per split
  unsigned total_tiles;
};

template <int BN, int MODE>
struct CtCfg {
  static constexpr int kTaps = MODE == CT_W3 ? 3 : 1;     // B tiles (and accumulators) per k-block
  static constexpr int kAcc = MODE == CT_W3 ? 1 : 2;      // accumulator sets (W3: one tile per CTA, no ping-pong needed)
  static constexpr int kABytes = 2 * 128 * 128;           // hi + lo, 128 rows (or 2 x 64-channel blocks) of 128 B
  static constexpr int kBBytes = 2 * BN * 128;
  static constexpr int kStageBytes = kABytes + kTaps * kBBytes;
  static constexpr int kStages = MODE == CT_W3 ? 2 : ((BN == 64) ? 4 : ((BN == 128) ? 3 : 2));
  static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
  static constexpr int kTmemCols = kTaps * kAcc * BN <= 128 ? 128 : (kTaps * kAcc * BN <= 256 ? 256 : 512);
};

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

#ifndef PDN_CONV_W3_WIDE
#define PDN_CONV_W3_WIDE 1
#endif
constexpr bool kWideW3 = PDN_CONV_W3_WIDE != 0;

template <int BN, int MODE>
__global__ void __launch_bounds__(256, 1)
k_conv_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, CtArgs g) {
  using Cfg = CtCfg<BN, MODE>;
  constexpr bool kW = MODE != CT_F;  // backward-weight family: MN-major operands, contraction over pixel patches
  constexpr int  T = Cfg::kTaps, NACC = Cfg::kAcc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t*  smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int taps = g.k * g.k;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }  // W3 uses set 0 only
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> coordinates.  F: (n_blk, tx, ty, image).  W: (n_blk, m_blk, split, tap)
  struct Tile { int n_blk, a, b, c; };
  auto decode = [&](uint32_t t) {
    Tile r;
    uint32_t q = t / (uint32_t)g.n_tiles_n;
    r.n_blk = (int)(t - q * (uint32_t)g.n_tiles_n);
    t = q;
    if (MODE == CT_F) {
      q = t / (uint32_t)g.TX; r.a = (int)(t - q * (uint32_t)g.TX); t = q;   // tx
      q = t / (uint32_t)g.TY; r.b = (int)(t - q * (uint32_t)g.TY);          // ty
      r.c = (int)q;                                                         // image
    } else {
      q = t / (uint32_t)g.m_tiles; r.a = (int)(t - q * (uint32_t)g.m_tiles); t = q;  // m_blk
      q = t / (uint32_t)g.splits; r.b = (int)(t - q * (uint32_t)g.splits);           // split
      r.c = (int)q;                                                                  // tap
    }
    return r;
  };
  auto kb_count = [&](const Tile& tl) {
    if (MODE == CT_F) return taps * g.cblks;  // (W / W3: pixel patches of this split)
    const int total = g.n_img * g.TY * g.TX, b0 = tl.b * g.per_split;
    const int e = b0 + g.per_split < total ? b0 + g.per_split : total;
    return e > b0 ? e - b0 : 0;
  };

  if (warp == 0) {
    // ===== TMA producer =====
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (uint32_t t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      const Tile tl = decode(t);
      const int  nkb = kb_count(tl);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (MODE == CT_F) {
            const int tap = kb / g.cblks, cb = kb - tap * g.cblks;
            const int ky = tap / g.k, kx = tap - ky * g.k;
            const int x0 = tl.a * 16 + g.sign * (kx - g.pad), y0 = tl.b * 8 + g.sign * (ky - g.pad);
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
              tma_load_5d(&mapA, &full_bar[stage], sa + pl * 16384, cb * 64, x0, y0, pl, tl.c);  // [8 rows][16 px][64 ch]
              tma_load_4d(&mapB, &full_bar[stage], sb + pl * (BN * 128), cb * 64, tl.n_blk * BN, pl, tap);
            }
          } else {
            const int p = tl.b * g.per_split + kb;
            const int q1 = p / g.TX, px = p - q1 * g.TX, img = q1 / g.TY, py = q1 - img * g.TY;
            // W: tl.c = tap.  W3: tl.c = kernel row ky, the T = 3 taps kx = 0..2 share this k-block's dY tile
            const int ky = MODE == CT_W3 ? tl.c : tl.c / g.k, kx0 = MODE == CT_W3 ? 0 : tl.c - ky * g.k;
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
              for (int i = 0; i < 2; ++i)  // dY patch [64 px][64 o] x 2 channel blocks
                tma_load_5d(&mapA, &full_bar[stage], sa + pl * 16384 + i * 8192, tl.a * 128 + 64 * i, px * 16, py * 4, pl, img);
#pragma unroll
              for (int tp = 0; tp < T; ++tp)
#pragma unroll
                for (int i = 0; i < BN / 64; ++i)  // shifted x patch [64 px][64 c] x BN/64 channel blocks
                  tma_load_5d(&mapB, &full_bar[stage], sb + tp * Cfg::kBBytes + pl * (BN * 128) + i * 8192, tl.n_blk * BN + 64 * i,
                              px * 16 + kx0 + tp - g.pad, py * 4 + ky - g.pad, pl, img);
            }
          }
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, BN) | (kW ? (IDESC_A_MN_MAJOR | IDESC_B_MN_MAJOR) : 0u);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (uint32_t t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      const Tile tl = decode(t);
      const int  nkb = kb_count(tl);
      if (nkb == 0) continue;  // (mode W) an empty split: neither this warp nor the epilogue touches an accumulator
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN * T);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes), sb = sa + Cfg::kABytes;
          const uint64_t d_ahi = kW ? make_smem_desc_sw128_mn(sa, 8192) : make_smem_desc_sw128(sa);
          const uint64_t d_alo = kW ? make_smem_desc_sw128_mn(sa + 16384, 8192) : make_smem_desc_sw128(sa + 16384);
          int nks = 4;
          if (MODE == CT_F) {  // channel tail: k-steps made only of zero-filled channels are skipped
            const int tap = kb / g.cblks, cb = kb - tap * g.cblks, left = g.n_contr - cb * 64;
            nks = left >= 64 ? 4 : (left + 15) >> 4;
          }
          const uint64_t step = kW ? 128 : 2;  // one UMMA_K = 16: 16 pixel rows of 128 B (MN-major) or 32 B of channels
          if (MODE == CT_W3 && BN == 64 && kWideW3) {
            // The three taps' x tiles sit kBBytes apart in the stage: read as ONE MN-major B operand of N = 192 (three 64-column
            // blocks, leading-dimension offset kBBytes) into the three adjacent accumulators. A 128x64x16 MMA with both operands
            // in shared memory is bound by the 6 KB it re-reads (~65 cycles against a 32-cycle tensor floor); 128x192x16 reads
            // 10 KB for a 96-cycle floor: a third of the instructions and about half the time per k-step.
            const uint64_t d_bhi = make_smem_desc_sw128_mn(sb, Cfg::kBBytes), d_blo = make_smem_desc_sw128_mn(sb + BN * 128, Cfg::kBBytes);
            constexpr uint32_t idesc3 = make_idesc_bf16(128, 3 * 64) | IDESC_A_MN_MAJOR | IDESC_B_MN_MAJOR;
            for (int k = 0; k < nks; ++k) {
              const uint64_t o = step * (uint64_t)k;
              umma_bf16(tmem_d, d_alo + o, d_bhi + o, idesc3, (kb | k) ? 1u : 0u);
              umma_bf16(tmem_d, d_ahi + o, d_blo + o, idesc3, 1u);
              umma_bf16(tmem_d, d_ahi + o, d_bhi + o, idesc3, 1u);
            }
          } else
#pragma unroll
          for (int tp = 0; tp < T; ++tp) {
            const uint32_t sbt = sb + tp * Cfg::kBBytes;
            const uint64_t d_bhi = kW ? make_smem_desc_sw128_mn(sbt, 8192) : make_smem_desc_sw128(sbt);
            const uint64_t d_blo = kW ? make_smem_desc_sw128_mn(sbt + BN * 128, 8192) : make_smem_desc_sw128(sbt + BN * 128);
            const uint32_t dt = tmem_d + (uint32_t)(tp * BN);
            for (int k = 0; k < nks; ++k) {
              const uint64_t o = step * (uint64_t)k;
              umma_bf16(dt, d_alo + o, d_bhi + o, idesc, (kb | k) ? 1u : 0u);
              umma_bf16(dt, d_ahi + o, d_blo + o, idesc, 1u);
              umma_bf16(dt, d_ahi + o, d_bhi + o, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&tfull_bar[acc]);
      if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3;
    int       acc = 0;
    uint32_t  acc_phase = 0;
    const bool bias_vec = (((uintptr_t)g.bias) & 15) == 0;
    for (uint32_t t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      const Tile tl = decode(t);
      const int  nkb = kb_count(tl);
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN * T) + ((uint32_t)(q * 32) << 16);
      const int m = q * 32 + lane;  // accumulator row of this thread
      float*    p0 = nullptr;       // F: address of (image, channel 0, y, x); W: address of row (first tap of the tile, o), column 0
      int64_t   cstride = 1;
      if (MODE == CT_F) {
        const int y = tl.b * 8 + (m >> 4), x = tl.a * 16 + (m & 15);
        cstride = (int64_t)g.oh * g.ow;
        if (y < g.oh && x < g.ow) p0 = g.out + (int64_t)tl.c * g.n_out * cstride + (int64_t)y * g.ow + x;
      } else {
        const int o = tl.a * 128 + m;
        if (o < g.m_rows) p0 = g.out + ((int64_t)tl.c * T * g.m_rows + o) * g.n_out;
      }
      if (nkb == 0) continue;  // W: an empty split contributes nothing (the MMA warp skips it as well)
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int tp = 0; tp < T; ++tp) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int col0 = tl.n_blk * BN + c0;
          if (col0 >= g.n_out) break;
          const int ncol = (g.n_out - col0) < 32 ? (g.n_out - col0) : 32;
          float v[32];
          tmem_ld_32x32(tmem_d + (uint32_t)(tp * BN + c0), v);
          tmem_ld_wait();
          if (tp == T - 1 && (c0 + 32 >= BN || col0 + 32 >= g.n_out)) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          if (MODE == CT_F) {
            if (g.bias) {
              if (ncol == 32 && bias_vec) {
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
            if (p0) {  // for a fixed channel the 32 lanes write two runs of 16 consecutive pixels
              float* p = p0 + (int64_t)col0 * cstride;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncol) p[(int64_t)j * cstride] = v[j];
            }
          } else if (p0) {
            float* p = p0 + (int64_t)tp * g.m_rows * g.n_out + col0;  // tmp[tap][o][c]
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) atomicAdd(p + j, v[j]);
          }
        }
      }
      if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// dW[o][c][ky][kx] = tmp[tap][o][c]
__global__ void __launch_bounds__(256) k_conv_dw_permute(const float* __restrict__ tmp, float* __restrict__ dw, int O, int C, int taps) {
  const int64_t total = (int64_t)O * C * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int64_t oc = i / taps;
    dw[i] = tmp[(int64_t)tap * O * C + oc];
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 g_enc5 = nullptr;
static int enc5() {
  if (g_enc5) return 0;
  void*                           fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return PDN_ERR_CUDA;
  }
  g_enc5 = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  return 0;
}

// 5-D map over channels-last bf16 hi/lo planes [N][2][H*W][Cp] of an activation: dims {C, W, H, plane, N}, box {64, 16, rows, 1, 1}
static int make_map_nhwc(CUtensorMap* map, const void* planes, int64_t N, int64_t Cc, int64_t Hh, int64_t Ww, int64_t Cp, int rows) {
  PDN_TRY(enc5());
  cuuint64_t dims[5] = {(cuuint64_t)Cc, (cuuint64_t)Ww, (cuuint64_t)Hh, 2, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)Cp * 2, (cuuint64_t)Ww * Cp * 2, (cuuint64_t)Hh * Ww * Cp * 2, (cuuint64_t)2 * Hh * Ww * Cp * 2};
  cuuint32_t box[5] = {64, 16, (cuuint32_t)rows, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult   r = g_enc5(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_tma: cuTensorMapEncodeTiled (NHWC planes N=%lld C=%lld H=%lld W=%lld Cp=%lld rows=%d) failed with CUresult %d", (long long)N,
              (long long)Cc, (long long)Hh, (long long)Ww, (long long)Cp, rows, (int)r);
    return PDN_ERR_CUDA;
  }
  return 0;
}

// channels-last operand planes [N][2][H*W][Cp] of an NCHW fp32 activation (transposing pack, cached by the buffer's write counter)
static int nhwc_planes(const float* act, int64_t N, int64_t Cc, int64_t Hh, int64_t Ww, long long version, Scratch* buf, PackedOperand* out) {
  const int64_t nb[3] = {1, 1, N}, bs[3] = {0, 0, Cc * Hh * Ww};
  return planes_cached(act, Hh * Ww, Cc, 1, Hh * Ww, nb, bs, version, buf, out, /*force_kmajor=*/true);
}

template <int BN, int MODE>
static int launch_ct(const CUtensorMap& mA, const CUtensorMap& mB, const CtArgs& g) {
  using Cfg = CtCfg<BN, MODE>;
  static bool attr = false;
  if (!attr) {
    PDN_CUDA(cudaFuncSetAttribute(k_conv_tma<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr = true;
  }
  const unsigned ctas = g.total_tiles < (unsigned)sm_count() ? g.total_tiles : (unsigned)sm_count();
  k_conv_tma<BN, MODE><<<ctas, 256, Cfg::kSmem, stream()>>>(mA, mB, g);
  PDN_LAUNCHED(MODE == CT_F ? "conv_tma" : "conv_tma_dw");
  return 0;
}

bool conv_tma_ok(int64_t contr_channels, int stride, int64_t N, int64_t Cc, int64_t Hh, int64_t Ww) {
  static const bool off = getenv("PDN_CONV_GATHER") != nullptr;
  return !off && stride == 1 && contr_channels >= 16 && N * Cc * Hh < 0x7fffffff && Ww < 0x7fffffff && N < 0x7fffffff;
}

// Y[n, o, y, x] (+bias) = Σ_{tap, c} act[n, c, y + sign*(ky - pad), x + sign*(kx - pad)] · wt[tap][o][c]
//   forward:       act = x  (C channels, H x W),   wt[tap][o][c] = w[o, c, ky, kx],  sign = +1, output grid oh x ow
//   backward-data: act = gy (O channels, oh x ow), wt[tap][c][o] = w[o, c, ky, kx],  sign = -1, output grid H x W
int conv_tma_forward(const float* act, int64_t N, int64_t Cc, int64_t Hh, int64_t Ww, const PackedOperand& wt, const float* bias, float* out,
                     int64_t n_out, int64_t oh, int64_t ow, int k, int pad, int sign, long long act_version) {
  Scratch       bufA;
  PackedOperand A;
  PDN_TRY(nhwc_planes(act, N, Cc, Hh, Ww, act_version, &bufA, &A));
  CtArgs g{};
  g.out = out; g.bias = bias;
  g.n_img = (int)N; g.n_contr = (int)Cc; g.n_out = (int)n_out; g.m_rows = 0;
  g.oh = (int)oh; g.ow = (int)ow;
  g.TY = (int)((oh + 7) / 8); g.TX = (int)((ow + 15) / 16);
  g.k = k; g.pad = pad; g.sign = sign;
  g.cblks = (int)((Cc + 63) / 64);
  const int BN = n_out <= 64 ? 64 : (n_out <= 128 ? 128 : 256);
  g.n_tiles_n = (int)((n_out + BN - 1) / BN); g.m_tiles = 1; g.splits = 1; g.per_split = 0;
  const int64_t total = (int64_t)g.n_tiles_n * g.TX * g.TY * N;
  PDN_CHECK(total < 0x7fffffff, "conv_tma: too many tiles");
  g.total_tiles = (unsigned)total;
  CUtensorMap mA, mB;
  PDN_TRY(make_map_nhwc(&mA, A.planes, N, Cc, Hh, Ww, A.Kp, 8));
  PDN_TRY(tc_make_map(&mB, wt.planes, wt.R, wt.K, wt.Kp, wt.nbatch, BN));
  if (BN == 64) return launch_ct<64, CT_F>(mA, mB, g);
  if (BN == 128) return launch_ct<128, CT_F>(mA, mB, g);
  return launch_ct<256, CT_F>(mA, mB, g);
}

// dw[o, c, ky, kx] = Σ_{n, y, x} gy[n, o, y, x] · x[n, c, y + ky - pad, x + kx - pad]   (stride 1)
int conv_tma_bwd_weight(const float* x, const float* gy, float* dw, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t oh, int64_t ow,
                        int k, int pad, long long x_version, long long gy_version) {
  Scratch       bufX, bufG, tmp;
  PackedOperand X, G;
  PDN_TRY(nhwc_planes(x, N, C, H, W, x_version, &bufX, &X));
  PDN_TRY(nhwc_planes(gy, N, O, oh, ow, gy_version, &bufG, &G));
  const int taps = k * k;
  PDN_TRY(tmp.alloc((size_t)taps * O * C * sizeof(float)));
  PDN_CUDA(cudaMemsetAsync(tmp.p, 0, (size_t)taps * O * C * sizeof(float), stream()));
  CtArgs g{};
  g.out = (float*)tmp.p; g.bias = nullptr;
  g.n_img = (int)N; g.n_contr = 0; g.n_out = (int)C; g.m_rows = (int)O;
  g.oh = (int)oh; g.ow = (int)ow;
  g.TY = (int)((oh + 3) / 4); g.TX = (int)((ow + 15) / 16);
  g.k = k; g.pad = pad; g.sign = 1; g.cblks = 0;
  // k = 3: the three taps of a kernel row share one dY tile per k-block (three accumulators, BN = 64): 80 KB of TMA traffic per
  // pixel patch instead of 3 x 48 KB — the one-tap form is bound by L2->SM operand delivery
  static const bool w3_off = getenv("PDN_CONV_W1") != nullptr;
  const bool w3 = k == 3 && !w3_off;
  const int BN = (w3 || C <= 64) ? 64 : 128;
  g.n_tiles_n = (int)((C + BN - 1) / BN);
  g.m_tiles = (int)((O + 127) / 128);
  const int64_t patches = N * g.TY * g.TX;
  PDN_CHECK(patches < 0x7fffffff, "conv_tma: too many patches");
  const int64_t base_tiles = (int64_t)(w3 ? k : taps) * g.m_tiles * g.n_tiles_n;
  // rounded DOWN: base_tiles x splits work items must fit ONE wave of persistent CTAs (3 x 50 = 150 items on 148 SMs made two CTAs
  // run a second item and doubled the kernel's time)
  int64_t splits = sm_count() / base_tiles;
  if (splits > patches) splits = patches;
  if (splits < 1) splits = 1;
  const int64_t per = (patches + splits - 1) / splits;
  splits = (patches + per - 1) / per;  // no empty split
  g.splits = (int)splits; g.per_split = (int)per;
  g.total_tiles = (unsigned)(base_tiles * splits);
  CUtensorMap mA, mB;
  PDN_TRY(make_map_nhwc(&mA, G.planes, N, O, oh, ow, G.Kp, 4));
  PDN_TRY(make_map_nhwc(&mB, X.planes, N, C, H, W, X.Kp, 4));
  if (w3) PDN_TRY((launch_ct<64, CT_W3>(mA, mB, g)));
  else if (BN == 64) PDN_TRY((launch_ct<64, CT_W>(mA, mB, g)));
  else PDN_TRY((launch_ct<128, CT_W>(mA, mB, g)));
  k_conv_dw_permute<<<grid_for((int64_t)O * C * taps, 256), 256, 0, stream()>>>((const float*)tmp.p, dw, (int)O, (int)C, taps);
  PDN_LAUNCHED("conv_dw_permute");
  return 0;
}

}  // namespace pdn
