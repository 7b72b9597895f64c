// rnn_persist.cu — the GRU recurrence of a WHOLE sequence as one persistent cooperative kernel (forward and BPTT).
//
// Reference: a Python loop of cells, per time step zr = sigmoid(x Wx1 + h Wh1 + b1), n = tanh(x Wx2 + (r o h) Wh2 + b2),
// h' = (1 - z) o h + z o n (pydynet/nn/modules/rnn.py:529-544, 702-708): 18 eager nodes and 4 small sgemms per step. The input
// projections of all T steps are hoisted into one large GEMM by the caller (rnn.cu); what is left is inherently sequential: two
// DEPENDENT [B, H] x [H, *] products per step. With one launch per product (rnn.cu's loop: 8 launches per step forward + backward)
// a step costs 71 us on B200, almost all of it launch / pipeline fill / drain of kernels that each run a few microseconds.
//
// Here one launch runs all T steps:
//   * every CTA owns 64 columns of a recurrent weight matrix for the whole sequence: its bf16 hi/lo planes (BF16x3 operand split,
//     fp32 accuracy) are loaded ONCE by TMA and stay in shared memory (128 KB at H = 512) as the B operand of tcgen05.mma;
//   * the batch is tiled by 128 rows (MMA M); per step a CTA streams the 128 x H activation planes (h, then r o h) through a
//     3-stage TMA ring as the A operand, accumulates in TMEM, and its 4 epilogue warps (one thread per batch row) apply the gate
//     math straight from tensor memory: phase A CTAs (Wh1 columns) write z / r and the r o h planes, phase B CTAs (Wh2 columns)
//     write n, h_t and the h planes of the next step;
//   * batch tiles never synchronise with each other (the recurrence is independent per row); inside a batch tile the two phases
//     hand over through three monotonic counters in global memory (release: threadfence + atomicAdd by one thread after the CTA's
//     stores; acquire: ld.acquire poll + fence.proxy.async before the TMA loads) — no grid-wide barrier, no kernel boundary.
// Gate functions are the reference's piecewise forms (tensor.py:1000-1015), identical to rnn.cu's.
#include "common.cuh"
#include "gemm_tc.h"
#include "tc_ptx.cuh"

#include <cuda_bf16.h>

namespace pdn {

constexpr int GP_BM = 128;    // MMA M (tcgen05 M = 64 costs the same cycles and has a different accumulator layout)
constexpr int GP_BMV = 64;    // batch rows per tile that are real: only they are loaded (TMA box of 64 rows) and written back; the MMA
                              // also multiplies whatever lies in the other half of the shared-memory tile into accumulator rows
                              // nobody reads. Half the activation stream per CTA and twice the CTAs per sequence (measured: the
                              // 256 KB stream of a full tile was 3 us of an 11 us phase)
#ifndef PDN_GP_M64
#define PDN_GP_M64 1
#endif
// tcgen05 M = 64 instead of a half-filled M = 128: with both operands in shared memory a 128x64x16 MMA is bound by the 6 KB it
// re-reads per k-step (~65 cycles), and 2 KB of that are the tile's unused rows; M = 64 reads 4 KB (measured on the plain RNN:
// 12.9 -> 11.0 us per time step). Its accumulator puts rows 16 q .. 16 q + 15 into lanes 32 q .. 32 q + 15 of TMEM lane quarter q,
// so all eight epilogue warps read (16 useful lanes each) instead of four.
constexpr bool GP_M64 = PDN_GP_M64 != 0;
constexpr int GP_BN = 64;     // weight columns per CTA (MMA N)
constexpr int GP_BK = 64;     // k-block: 64 bf16 = one 128-byte swizzle span
constexpr int GP_STAGES = 3;  // activation ring
constexpr int GP_MAXKB = 8;   // resident weight tile: up to 8 k-blocks (H <= 512)
constexpr int GP_ASTAGE = 2 * GP_BM * GP_BK * 2;  // hi + lo tiles of the activations: 32 KB of shared memory (16 KB of it loaded)
constexpr int GP_ALOAD = 2 * GP_BMV * GP_BK * 2;
constexpr int GP_WBLOCK = GP_BN * GP_BK * 2;      // one plane of one weight k-block: 8 KB

struct GruPersistArgs {
  const float *xp1, *xp2, *h0;
  float *      hs, *zr, *nn;
  __nv_bfloat16 *hP0, *hP1, *rhP;  // operand planes [2][B][Kp]
  int           T, B, H, Kp, nbt, ct1, ct2;
  unsigned int* cnt;  // [nbt][4]: 0 r columns done, 1 z columns done, 2 h done (monotonic over the sequence)
  unsigned long long* trace;  // PDN_GRU_TRACE=1: globaltimer stamps of one phase A and one phase B CTA at step 100
};

__device__ __forceinline__ void gp_stamp(const GruPersistArgs& g, int t, bool phaseA, int c, int bt, int slot) {
  if (g.trace && t == 100 && bt == 0 && c == (phaseA ? g.ct1 / 2 : 0)) {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    g.trace[(phaseA ? 0 : 16) + slot] = v;
  }
}

// The reference's piecewise forms (tensor.py:1000-1015: x > 0 ? 1/(1+e^-x) : 1 - 1/(1+e^x), and the same shape for tanh) written
// BRANCH-FREE on |x|: both pieces evaluate the same p = 1/(1 + e^-|x|), so one exponential, one reciprocal and a select per element.
// The epilogue runs one warp per scheduler with nothing to hide latency behind: the first version (two divergent pieces per element,
// full-range expf and IEEE division) spent 340 cycles per element, 11 us of a 22 us phase.
__device__ __forceinline__ float gp_sigmoid(float x) {
  const float p = __fdividef(1.f, 1.f + __expf(-fabsf(x)));
  return x > 0.f ? p : 1.f - p;
}
__device__ __forceinline__ float gp_tanh(float x) {
  const float q = __fdividef(2.f, 1.f + __expf(-2.f * fabsf(x))) - 1.f;
  return x > 0.f ? q : -q;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// spin until *p >= target; a producer that never arrives must surface as an error, never as a hung GPU (~2 s)
__device__ __forceinline__ void wait_count(const unsigned int* p, unsigned int target) {
  long long t0 = 0;
  int       spins = 0;
  while (ld_acquire_u32(p) < target) {
    if (++spins == 2048) {
      spins = 0;
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 4 consecutive values -> bf16 hi / lo planes (8 bytes each; 16 lanes of a row write 128 contiguous bytes per plane)
__device__ __forceinline__ void put_planes4(__nv_bfloat16* hi, int64_t plane_stride, const float4& a) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(a.x), h1 = __float2bfloat16_rn(a.y), h2 = __float2bfloat16_rn(a.z), h3 = __float2bfloat16_rn(a.w);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(a.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(a.y - __bfloat162float(h1));
  const __nv_bfloat16 l2 = __float2bfloat16_rn(a.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(a.w - __bfloat162float(h3));
  uint2 H, L;
  H.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16), H.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
  L.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16), L.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
  *reinterpret_cast<uint2*>(hi) = H;
  *reinterpret_cast<uint2*>(hi + plane_stride) = L;
}

//   warp 0     TMA producer (activation planes of every step; waits for the batch tile's counters)
//   warp 1     MMA issuer
//   warp 2     TMEM allocator
//   warps 4-11 epilogue: the four warps that may read TMEM lanes 0-63 (warp % 4 < 2) move the accumulator to shared memory, then all
//              eight apply the gate math ROW-MAJOR (16 lanes x float4 per row): every load and store is a whole 256-byte row segment
__global__ void __launch_bounds__(384, 1)
k_gru_persist_fwd(const __grid_constant__ CUtensorMap mapH0, const __grid_constant__ CUtensorMap mapH1, const __grid_constant__ CUtensorMap mapRH,
                  const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2, GruPersistArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t*  smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int KB = g.H / GP_BK;
  uint8_t*  wsm = smem;                                  // [KB][2 planes][64 cols][64 k] resident weights
  uint8_t*  asm_ = smem + (size_t)KB * 2 * GP_WBLOCK;    // [STAGES][hi 16 KB | lo 16 KB]
  uint64_t* full_bar = (uint64_t*)(asm_ + GP_STAGES * GP_ASTAGE);
  uint64_t* empty_bar = full_bar + GP_STAGES;
  uint64_t* wfull_bar = empty_bar + GP_STAGES;
  uint64_t* tfull_bar = wfull_bar + 1;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int  warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int  nA = g.ct1 * g.nbt;
  const bool phaseA = (int)blockIdx.x < nA;
  const int  idx = phaseA ? (int)blockIdx.x : (int)blockIdx.x - nA;
  const int  ct = phaseA ? g.ct1 : g.ct2;
  const int  c = idx % ct, bt = idx / ct;  // column tile, batch tile
  const bool rcols = phaseA && c >= g.ct1 / 2;
  unsigned int* cnt = g.cnt + bt * 4;
  const unsigned int per_r = (unsigned int)(g.ct1 / 2), per_h = (unsigned int)g.ct2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(phaseA ? &mapW1 : &mapW2);
    tma_prefetch_desc(&mapH0);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, GP_M64 ? 8 : 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    const bool leader = elect_one();
    if (leader) {  // the CTA's weight columns, once
      mbar_expect_tx(wfull_bar, (uint32_t)(KB * 2 * GP_WBLOCK));
      for (int kb = 0; kb < KB; ++kb)
        for (int pl = 0; pl < 2; ++pl)
          tma_load_4d(phaseA ? &mapW1 : &mapW2, wfull_bar, wsm + (size_t)(kb * 2 + pl) * GP_WBLOCK, kb * GP_BK, c * GP_BN, pl, 0);
    }
    int      stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < g.T; ++t) {
      // inputs of this step: phase A reads h_{t-1} (planes written by the phase B CTAs of step t-1), phase B reads r o h of step t
      if (phaseA) {
        if (t > 0) wait_count(cnt + 2, per_h * (unsigned int)t);
      } else {
        wait_count(cnt + 0, per_r * (unsigned int)(t + 1));
      }
      if (lane == 0) gp_stamp(g, t, phaseA, c, bt, 0);
      asm volatile("fence.proxy.async.global;" ::: "memory");  // other SMs' generic-proxy stores -> this SM's async-proxy (TMA) reads
      if (lane == 0) gp_stamp(g, t, phaseA, c, bt, 1);
      const CUtensorMap* mA = phaseA ? ((t & 1) ? &mapH1 : &mapH0) : &mapRH;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* st = asm_ + stage * GP_ASTAGE;
          mbar_expect_tx(&full_bar[stage], GP_ALOAD);
          tma_load_4d(mA, &full_bar[stage], st, kb * GP_BK, bt * GP_BMV, 0, 0);
          tma_load_4d(mA, &full_bar[stage], st + GP_BM * GP_BK * 2, kb * GP_BK, bt * GP_BMV, 1, 0);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) gp_stamp(g, t, phaseA, c, bt, 2);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool     leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(GP_M64 ? 64 : GP_BM, GP_BN);
    int            stage = 0;
    uint32_t       phase = 0;
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    for (int t = 0; t < g.T; ++t) {
      mbar_wait(tempty_bar, (uint32_t)(t & 1) ^ 1);  // the epilogue of the previous step has drained the accumulator
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (kb == 0 && lane == 0) gp_stamp(g, t, phaseA, c, bt, 3);
        if (leader) {
          const uint32_t sa = smem_u32(asm_ + stage * GP_ASTAGE), sb = smem_u32(wsm + (size_t)kb * 2 * GP_WBLOCK);
          const uint64_t d_ahi = make_smem_desc_sw128(sa), d_alo = make_smem_desc_sw128(sa + GP_BM * GP_BK * 2);
          const uint64_t d_bhi = make_smem_desc_sw128(sb), d_blo = make_smem_desc_sw128(sb + GP_WBLOCK);
#pragma unroll
          for (int k = 0; k < GP_BK / 16; ++k) {
            const uint64_t o = 2 * k;  // 32 bytes inside the swizzle span, in 16-byte units
            umma_bf16(tmem_base, d_alo + o, d_bhi + o, idesc, (kb | k) ? 1u : 0u);
            umma_bf16(tmem_base, d_ahi + o, d_blo + o, idesc, 1u);
            umma_bf16(tmem_base, d_ahi + o, d_bhi + o, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(tfull_bar);
      if (lane == 0) gp_stamp(g, t, phaseA, c, bt, 4);
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    // The accumulator sits in TMEM with one batch row per lane; read that way, every global access of a warp would touch 32 different
    // lines with 16 bytes each (measured: ~1 us of partial-sector transactions per stored tile, four tiles per step). So the readers
    // drop the 64 x 64 fp32 tile into shared memory (the idle activation ring) and the math runs row-major.
    const bool     reader = GP_M64 ? true : (warp & 3) < 2;   // may access TMEM lanes 32 * (warp % 4) ...
    const int      rq = GP_M64 ? (warp & 3) : (warp & 1), rc0 = ((warp - 4) >> 2) * 32;  // reader: lane quarter, first of its 32 columns
    const uint32_t taddr = tmem_base + ((uint32_t)(rq * 32) << 16) + (uint32_t)rc0;
    float4*        acc4 = reinterpret_cast<float4*>(asm_);    // [64 rows][16 float4], chunk index XOR-swizzled by (row & 7)
    const int      ew = warp - 4, cch = lane & 15;            // math: rows ew * 8 + 2 i + (lane >> 4), columns 4 * cch .. + 3
    const int64_t  H = g.H, BH = (int64_t)g.B * H;
    const int      j0 = c * GP_BN + 4 * cch;  // this thread's first column (phase A: in [0, 2H); phase B: in [0, H))
    for (int t = 0; t < g.T; ++t) {
      const float* hprev = t == 0 ? g.h0 : g.hs + (int64_t)(t - 1) * BH;
      // The operands of the gate math do not depend on the accumulator: they are requested before it is waited for (and before
      // anything is stored); the input projection even before the step's dependencies have arrived.
      float4 xr[4], zv[4], hv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t b = (int64_t)bt * GP_BMV + ew * 8 + 2 * i + (lane >> 4);
        if (b < g.B) xr[i] = __ldg(reinterpret_cast<const float4*>(phaseA ? g.xp1 + ((int64_t)t * g.B + b) * 2 * H + j0 : g.xp2 + ((int64_t)t * g.B + b) * H + j0));
      }
      if (!phaseA) {  // z of this step comes from the z-column CTAs of phase A; h_{t-1} from the previous step's phase B
        if (lane == 0) wait_count(cnt + 1, per_r * (unsigned int)(t + 1));
        __syncwarp();
      } else if (rcols && t > 0) {
        if (lane == 0) wait_count(cnt + 2, per_h * (unsigned int)t);
        __syncwarp();
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t b = (int64_t)bt * GP_BMV + ew * 8 + 2 * i + (lane >> 4);
        if (b < g.B) {
          if (rcols) hv[i] = __ldcg(reinterpret_cast<const float4*>(hprev + b * H + (j0 - (int)H)));
          else if (!phaseA) {
            zv[i] = __ldcg(reinterpret_cast<const float4*>(g.zr + ((int64_t)t * g.B + b) * 2 * H + j0));
            hv[i] = __ldcg(reinterpret_cast<const float4*>(hprev + b * H + j0));
          }
        }
      }
      if (threadIdx.x == 128) gp_stamp(g, t, phaseA, c, bt, 5);
      if (reader) {
        mbar_wait(tfull_bar, (uint32_t)(t & 1));
        tc_fence_after();
        if (threadIdx.x == 128) gp_stamp(g, t, phaseA, c, bt, 6);
        float v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        tc_fence_before();  // accumulator is in registers: hand it back
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        const int row = GP_M64 ? rq * 16 + lane : rq * 32 + lane;
        if (!GP_M64 || lane < 16) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc4[row * 16 + ((rc0 / 4 + k) ^ (row & 7))] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
      }
      epi_bar();
      if (threadIdx.x == 128) gp_stamp(g, t, phaseA, c, bt, 10);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int     row = ew * 8 + 2 * i + (lane >> 4);
        const int64_t b = (int64_t)bt * GP_BMV + row;
        if (b >= g.B) continue;
        float4 a = acc4[row * 16 + (cch ^ (row & 7))];
        if (phaseA) {
          a.x = gp_sigmoid(a.x + xr[i].x), a.y = gp_sigmoid(a.y + xr[i].y), a.z = gp_sigmoid(a.z + xr[i].z), a.w = gp_sigmoid(a.w + xr[i].w);
          *reinterpret_cast<float4*>(g.zr + ((int64_t)t * g.B + b) * 2 * H + j0) = a;
          if (rcols) {  // r o h_{t-1} -> operand planes of phase B
            a.x *= hv[i].x, a.y *= hv[i].y, a.z *= hv[i].z, a.w *= hv[i].w;
            put_planes4(g.rhP + b * g.Kp + (j0 - (int)H), (int64_t)g.B * g.Kp, a);
          }
        } else {
          const int64_t off = ((int64_t)t * g.B + b) * H + j0;
          a.x = gp_tanh(a.x + xr[i].x), a.y = gp_tanh(a.y + xr[i].y), a.z = gp_tanh(a.z + xr[i].z), a.w = gp_tanh(a.w + xr[i].w);
          *reinterpret_cast<float4*>(g.nn + off) = a;
          a.x = (1.f - zv[i].x) * hv[i].x + zv[i].x * a.x, a.y = (1.f - zv[i].y) * hv[i].y + zv[i].y * a.y;
          a.z = (1.f - zv[i].z) * hv[i].z + zv[i].z * a.z, a.w = (1.f - zv[i].w) * hv[i].w + zv[i].w * a.w;
          *reinterpret_cast<float4*>(g.hs + off) = a;
          put_planes4((((t + 1) & 1) ? g.hP1 : g.hP0) + b * g.Kp + j0, (int64_t)g.B * g.Kp, a);
        }
      }
      // this CTA's part of the step is in global memory: publish it to the batch tile's consumers
      if (threadIdx.x == 128) gp_stamp(g, t, phaseA, c, bt, 7);
      fence_proxy_async_smem();  // the scratch lives in the activation ring: generic-proxy accesses before the next TMA writes
      epi_bar();
      if (threadIdx.x == 128) {
        gp_stamp(g, t, phaseA, c, bt, 8);
        __threadfence();
        atomicAdd(cnt + (phaseA ? (rcols ? 0 : 1) : 2), 1u);
        gp_stamp(g, t, phaseA, c, bt, 9);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------------------------ LSTM forward
// lin = x Wx + h Wh + b -> f, i, o = sigmoid (that order), g = tanh; c' = f c + i g; h' = o tanh(c') (rnn.py:280-288): ONE product
// per step, so one phase per step. A CTA owns 16 hidden units x 4 gates = 64 weight columns (the rows f_j, i_j, o_j, g_j of Wh^T are
// gathered by four 16-row TMA boxes per k-block), so all four gates of a unit meet in its epilogue; the cell state of a thread's
// four units lives in registers for the whole sequence.
struct LstmPersistArgs {
  const float *xp, *h0, *c0;
  float *      hs, *cs, *gates;
  __nv_bfloat16 *hP0, *hP1;
  int           T, B, H, Kp, nbt, ctj;  // ctj = H / 16 column tiles
  unsigned int* cnt;                    // [nbt]: h planes of a step complete (monotonic)
};

__global__ void __launch_bounds__(384, 1)
k_lstm_persist_fwd(const __grid_constant__ CUtensorMap mapH0, const __grid_constant__ CUtensorMap mapH1, const __grid_constant__ CUtensorMap mapW,
                   LstmPersistArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t*  smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int KB = g.H / GP_BK;
  uint8_t*  wsm = smem;
  uint8_t*  asm_ = smem + (size_t)KB * 2 * GP_WBLOCK;
  uint64_t* full_bar = (uint64_t*)(asm_ + GP_STAGES * GP_ASTAGE);
  uint64_t* empty_bar = full_bar + GP_STAGES;
  uint64_t* wfull_bar = empty_bar + GP_STAGES;
  uint64_t* tfull_bar = wfull_bar + 1;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int c = (int)blockIdx.x % g.ctj, bt = (int)blockIdx.x / g.ctj;
  unsigned int*      cnt = g.cnt + bt;
  const unsigned int per = (unsigned int)g.ctj;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapH0);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, GP_M64 ? 8 : 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) {  // 16 rows of each gate: tile row q * 16 + jj <-> weight column q * H + 16 c + jj
      mbar_expect_tx(wfull_bar, (uint32_t)(KB * 2 * GP_WBLOCK));
      for (int kb = 0; kb < KB; ++kb)
        for (int pl = 0; pl < 2; ++pl)
          for (int q = 0; q < 4; ++q)
            tma_load_4d(&mapW, wfull_bar, wsm + (size_t)(kb * 2 + pl) * GP_WBLOCK + q * 2048, kb * GP_BK, q * g.H + 16 * c, pl, 0);
    }
    int      stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < g.T; ++t) {
      if (t > 0) wait_count(cnt, per * (unsigned int)t);
      asm volatile("fence.proxy.async.global;" ::: "memory");
      const CUtensorMap* mA = (t & 1) ? &mapH1 : &mapH0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* st = asm_ + stage * GP_ASTAGE;
          mbar_expect_tx(&full_bar[stage], GP_ALOAD);
          tma_load_4d(mA, &full_bar[stage], st, kb * GP_BK, bt * GP_BMV, 0, 0);
          tma_load_4d(mA, &full_bar[stage], st + GP_BM * GP_BK * 2, kb * GP_BK, bt * GP_BMV, 1, 0);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool     leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(GP_M64 ? 64 : GP_BM, GP_BN);
    int            stage = 0;
    uint32_t       phase = 0;
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    for (int t = 0; t < g.T; ++t) {
      mbar_wait(tempty_bar, (uint32_t)(t & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t sa = smem_u32(asm_ + stage * GP_ASTAGE), sb = smem_u32(wsm + (size_t)kb * 2 * GP_WBLOCK);
          const uint64_t d_ahi = make_smem_desc_sw128(sa), d_alo = make_smem_desc_sw128(sa + GP_BM * GP_BK * 2);
          const uint64_t d_bhi = make_smem_desc_sw128(sb), d_blo = make_smem_desc_sw128(sb + GP_WBLOCK);
#pragma unroll
          for (int k = 0; k < GP_BK / 16; ++k) {
            const uint64_t o = 2 * k;
            umma_bf16(tmem_base, d_alo + o, d_bhi + o, idesc, (kb | k) ? 1u : 0u);
            umma_bf16(tmem_base, d_ahi + o, d_blo + o, idesc, 1u);
            umma_bf16(tmem_base, d_ahi + o, d_bhi + o, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    // epilogue: the four TMEM readers drop the 64 x 64 accumulator into shared memory; thread (row, jq) then owns hidden units
    // 16 c + 4 jq .. + 3 of batch row `row`: their four gates sit in chunks q * 4 + jq of the tile row
    const bool     reader = GP_M64 ? true : (warp & 3) < 2;
    const int      rq = GP_M64 ? (warp & 3) : (warp & 1), rc0 = ((warp - 4) >> 2) * 32;
    const uint32_t taddr = tmem_base + ((uint32_t)(rq * 32) << 16) + (uint32_t)rc0;
    float4*        acc4 = reinterpret_cast<float4*>(asm_);
    const int      item = (int)threadIdx.x - 128, row = item >> 2, jq = item & 3;
    const int64_t  H = g.H, BH = (int64_t)g.B * H, b = (int64_t)bt * GP_BMV + row;
    const bool     live = b < g.B;
    const int      j = 16 * c + 4 * jq;
    float4         cst = live ? __ldg(reinterpret_cast<const float4*>(g.c0 + b * H + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < g.T; ++t) {
      float4 x[4];
      if (live) {
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = __ldg(reinterpret_cast<const float4*>(g.xp + ((int64_t)t * g.B + b) * 4 * H + q * H + j));
      }
      if (reader) {
        mbar_wait(tfull_bar, (uint32_t)(t & 1));
        tc_fence_after();
        float v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        const int r = GP_M64 ? rq * 16 + lane : rq * 32 + lane;
        if (!GP_M64 || lane < 16) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc4[r * 16 + ((rc0 / 4 + k) ^ (r & 7))] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
      }
      epi_bar();
      if (live) {
        float4 a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = acc4[row * 16 + ((q * 4 + jq) ^ (row & 7))];
        const float4 f = make_float4(gp_sigmoid(a[0].x + x[0].x), gp_sigmoid(a[0].y + x[0].y), gp_sigmoid(a[0].z + x[0].z), gp_sigmoid(a[0].w + x[0].w));
        const float4 i = make_float4(gp_sigmoid(a[1].x + x[1].x), gp_sigmoid(a[1].y + x[1].y), gp_sigmoid(a[1].z + x[1].z), gp_sigmoid(a[1].w + x[1].w));
        const float4 o = make_float4(gp_sigmoid(a[2].x + x[2].x), gp_sigmoid(a[2].y + x[2].y), gp_sigmoid(a[2].z + x[2].z), gp_sigmoid(a[2].w + x[2].w));
        const float4 gg = make_float4(gp_tanh(a[3].x + x[3].x), gp_tanh(a[3].y + x[3].y), gp_tanh(a[3].z + x[3].z), gp_tanh(a[3].w + x[3].w));
        cst = make_float4(f.x * cst.x + i.x * gg.x, f.y * cst.y + i.y * gg.y, f.z * cst.z + i.z * gg.z, f.w * cst.w + i.w * gg.w);
        const float4 h = make_float4(o.x * gp_tanh(cst.x), o.y * gp_tanh(cst.y), o.z * gp_tanh(cst.z), o.w * gp_tanh(cst.w));
        float*       gout = g.gates + ((int64_t)t * g.B + b) * 4 * H + j;
        *reinterpret_cast<float4*>(gout) = f;
        *reinterpret_cast<float4*>(gout + H) = i;
        *reinterpret_cast<float4*>(gout + 2 * H) = o;
        *reinterpret_cast<float4*>(gout + 3 * H) = gg;
        const int64_t off = ((int64_t)t * g.B + b) * H + j;
        *reinterpret_cast<float4*>(g.cs + off) = cst;
        *reinterpret_cast<float4*>(g.hs + off) = h;
        put_planes4((((t + 1) & 1) ? g.hP1 : g.hP0) + b * g.Kp + j, (int64_t)g.B * g.Kp, h);
      }
      fence_proxy_async_smem();
      epi_bar();
      if (threadIdx.x == 128) {
        __threadfence();
        atomicAdd(cnt, 1u);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------------------------ BPTT
// Per step s (T-1 ... 0), element (b, k), with hp = h_{s-1}, the saved z, r, n and the running gradient dh (rnn.py:529-544 reversed):
//   d = dh + g_hs[s];  dl2 = d z (1 - n^2) -> dxp2[s];  dl1z = d (n - hp) z (1 - z) -> dxp1[s][:, k]
//   drh = dl2 Wh2^T;   dl1r = drh hp r (1 - r) -> dxp1[s][:, H + k]
//   dh' = d (1 - z) + drh r + dl1z Wh1[:, :H]^T + dl1r Wh1[:, H:]^T
// Three CTA sets per 64-row batch tile, each CTA with 64 rows k of a weight matrix resident in shared memory (K-major over j):
//   X1  drh = dl2 . Wh2^T          epilogue: dl1r -> dxp1 + operand planes, u1 = drh r
//   X2  uz  = dl1z . Wh1[:, :H]^T  (independent of X1: runs beside it)
//   Y   ur  = dl1r . Wh1[:, H:]^T  epilogue: dh' = d (1 - z) + u1 + uz + ur, then the elementwise part of step s - 1 (dl2, dl1z -> dxp2,
//       dxp1 and the operand planes of X1 / X2); dh, d and z of a thread's elements stay in REGISTERS across the whole sequence
// Hand-over counters per batch tile: cY (planes of the next step ready), cX1, cX2. The weight gradients are single large GEMMs over all
// T*B rows after the loop (rnn.cu), as before.
struct GruBwdArgs {
  const float *g_hs, *h0, *hs, *zr, *nn;
  float *      dxp1, *dxp2, *dh0, *u1, *uz;  // u1, uz: [B][H] fp32 hand-over buffers
  __nv_bfloat16 *dl2P, *dl1zP, *dl1rP;       // operand planes [2][B][Kp]
  int           T, B, H, Kp, nbt, ctk;
  unsigned int* cnt;  // [nbt][4]: 0 cY, 1 cX1, 2 cX2
};

__global__ void __launch_bounds__(384, 1)
k_gru_persist_bwd(const __grid_constant__ CUtensorMap mapDl2, const __grid_constant__ CUtensorMap mapDl1z, const __grid_constant__ CUtensorMap mapDl1r,
                  const __grid_constant__ CUtensorMap mapW2t, const __grid_constant__ CUtensorMap mapW1t, GruBwdArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t*  smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int KB = g.H / GP_BK;
  uint8_t*  wsm = smem;
  uint8_t*  asm_ = smem + (size_t)KB * 2 * GP_WBLOCK;
  uint64_t* full_bar = (uint64_t*)(asm_ + GP_STAGES * GP_ASTAGE);
  uint64_t* empty_bar = full_bar + GP_STAGES;
  uint64_t* wfull_bar = empty_bar + GP_STAGES;
  uint64_t* tfull_bar = wfull_bar + 1;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int  warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int  per_set = g.ctk * g.nbt;
  const int  role = (int)blockIdx.x / per_set;  // 0: X1, 1: X2, 2: Y
  const int  idx = (int)blockIdx.x - role * per_set;
  const int  c = idx % g.ctk, bt = idx / g.ctk;
  unsigned int*      cnt = g.cnt + bt * 4;
  const unsigned int per = (unsigned int)g.ctk;
  const CUtensorMap* mW = role == 0 ? &mapW2t : &mapW1t;
  const CUtensorMap* mA = role == 0 ? &mapDl2 : (role == 1 ? &mapDl1z : &mapDl1r);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(mW);
    tma_prefetch_desc(mA);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, GP_M64 ? 8 : 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    const bool leader = elect_one();
    if (leader) {  // rows k of this CTA; Y contracts over the r half of Wh1's columns
      mbar_expect_tx(wfull_bar, (uint32_t)(KB * 2 * GP_WBLOCK));
      for (int kb = 0; kb < KB; ++kb)
        for (int pl = 0; pl < 2; ++pl)
          tma_load_4d(mW, wfull_bar, wsm + (size_t)(kb * 2 + pl) * GP_WBLOCK, (role == 2 ? g.H : 0) + kb * GP_BK, c * GP_BN, pl, 0);
    }
    int      stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < g.T; ++i) {
      wait_count(cnt + (role == 2 ? 1 : 0), per * (unsigned int)(i + 1));
      asm volatile("fence.proxy.async.global;" ::: "memory");
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* st = asm_ + stage * GP_ASTAGE;
          mbar_expect_tx(&full_bar[stage], GP_ALOAD);
          tma_load_4d(mA, &full_bar[stage], st, kb * GP_BK, bt * GP_BMV, 0, 0);
          tma_load_4d(mA, &full_bar[stage], st + GP_BM * GP_BK * 2, kb * GP_BK, bt * GP_BMV, 1, 0);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool     leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(GP_M64 ? 64 : GP_BM, GP_BN);
    int            stage = 0;
    uint32_t       phase = 0;
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    for (int i = 0; i < g.T; ++i) {
      mbar_wait(tempty_bar, (uint32_t)(i & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t sa = smem_u32(asm_ + stage * GP_ASTAGE), sb = smem_u32(wsm + (size_t)kb * 2 * GP_WBLOCK);
          const uint64_t d_ahi = make_smem_desc_sw128(sa), d_alo = make_smem_desc_sw128(sa + GP_BM * GP_BK * 2);
          const uint64_t d_bhi = make_smem_desc_sw128(sb), d_blo = make_smem_desc_sw128(sb + GP_WBLOCK);
#pragma unroll
          for (int k = 0; k < GP_BK / 16; ++k) {
            const uint64_t o = 2 * k;
            umma_bf16(tmem_base, d_alo + o, d_bhi + o, idesc, (kb | k) ? 1u : 0u);
            umma_bf16(tmem_base, d_ahi + o, d_blo + o, idesc, 1u);
            umma_bf16(tmem_base, d_ahi + o, d_bhi + o, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    // ===== epilogue (accumulator -> shared memory by the four TMEM readers, row-major math by all eight warps) =====
    const bool     reader = GP_M64 ? true : (warp & 3) < 2;
    const int      rq = GP_M64 ? (warp & 3) : (warp & 1), rc0 = ((warp - 4) >> 2) * 32;
    const uint32_t taddr = tmem_base + ((uint32_t)(rq * 32) << 16) + (uint32_t)rc0;
    float4*        acc4 = reinterpret_cast<float4*>(asm_);
    const int      ew = warp - 4, cch = lane & 15;
    const int64_t  H = g.H, BH = (int64_t)g.B * H, PS = (int64_t)g.B * g.Kp;
    const int      k0 = c * GP_BN + 4 * cch;  // this thread's first hidden index k
    int64_t        brow[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) brow[e] = (int64_t)bt * GP_BMV + ew * 8 + 2 * e + (lane >> 4);
    float4 dh[4], dd[4], zz[4];  // Y: running dh, d and z of the step being processed
    // elementwise part of step s for this thread's elements (Y): d = dh + g_hs[s]; dl2, dl1z -> dxp2 / dxp1 / operand planes
    auto elementwise = [&](int s) {
      const float* hprev = s == 0 ? g.h0 : g.hs + (int64_t)(s - 1) * BH;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t b = brow[e];
        if (b >= g.B) continue;
        const int64_t o1 = ((int64_t)s * g.B + b) * H + k0, o2 = ((int64_t)s * g.B + b) * 2 * H + k0;
        const float4  gg = g.g_hs ? __ldg(reinterpret_cast<const float4*>(g.g_hs + o1)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4  z = __ldg(reinterpret_cast<const float4*>(g.zr + o2)), n = __ldg(reinterpret_cast<const float4*>(g.nn + o1));
        const float4  hp = __ldg(reinterpret_cast<const float4*>(hprev + b * H + k0));
        float4        d = make_float4(dh[e].x + gg.x, dh[e].y + gg.y, dh[e].z + gg.z, dh[e].w + gg.w);
        const float4  dl2 = make_float4(d.x * z.x * (1.f - n.x * n.x), d.y * z.y * (1.f - n.y * n.y), d.z * z.z * (1.f - n.z * n.z), d.w * z.w * (1.f - n.w * n.w));
        const float4  dl1 = make_float4(d.x * (n.x - hp.x) * z.x * (1.f - z.x), d.y * (n.y - hp.y) * z.y * (1.f - z.y), d.z * (n.z - hp.z) * z.z * (1.f - z.z),
                                        d.w * (n.w - hp.w) * z.w * (1.f - z.w));
        *reinterpret_cast<float4*>(g.dxp2 + o1) = dl2;
        *reinterpret_cast<float4*>(g.dxp1 + o2) = dl1;
        put_planes4(g.dl2P + b * g.Kp + k0, PS, dl2);
        put_planes4(g.dl1zP + b * g.Kp + k0, PS, dl1);
        dd[e] = d, zz[e] = z;
      }
    };
    auto publish = [&](int which) {
      fence_proxy_async_smem();
      epi_bar();
      if (threadIdx.x == 128) {
        __threadfence();
        atomicAdd(cnt + which, 1u);
      }
    };
    if (role == 2) {  // prologue: the last time step starts from dh = 0
#pragma unroll
      for (int e = 0; e < 4; ++e) dh[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      elementwise(g.T - 1);
      publish(0);
    }
    for (int i = 0; i < g.T; ++i) {
      const int    s = g.T - 1 - i;
      const float* hprev = s == 0 ? g.h0 : g.hs + (int64_t)(s - 1) * BH;
      float4       rr[4], hp[4], u1[4], uz[4];
      if (role == 0) {  // r and h_{s-1} do not depend on the accumulator
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (brow[e] < g.B) {
            rr[e] = __ldg(reinterpret_cast<const float4*>(g.zr + ((int64_t)s * g.B + brow[e]) * 2 * H + H + k0));
            hp[e] = __ldg(reinterpret_cast<const float4*>(hprev + brow[e] * H + k0));
          }
      } else if (role == 2) {  // u1 / uz of this step come from the X1 / X2 CTAs
        if (lane == 0) {
          wait_count(cnt + 1, per * (unsigned int)(i + 1));
          wait_count(cnt + 2, per * (unsigned int)(i + 1));
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (brow[e] < g.B) {
            u1[e] = __ldcg(reinterpret_cast<const float4*>(g.u1 + brow[e] * H + k0));
            uz[e] = __ldcg(reinterpret_cast<const float4*>(g.uz + brow[e] * H + k0));
          }
      }
      if (reader) {
        mbar_wait(tfull_bar, (uint32_t)(i & 1));
        tc_fence_after();
        float v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        const int row = GP_M64 ? rq * 16 + lane : rq * 32 + lane;
        if (!GP_M64 || lane < 16) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc4[row * 16 + ((rc0 / 4 + k) ^ (row & 7))] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
      }
      epi_bar();
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int     row = ew * 8 + 2 * e + (lane >> 4);
        const int64_t b = brow[e];
        if (b >= g.B) continue;
        const float4 a = acc4[row * 16 + (cch ^ (row & 7))];
        if (role == 0) {
          const float4 r = rr[e], h = hp[e];
          const float4 dl1r = make_float4(a.x * h.x * r.x * (1.f - r.x), a.y * h.y * r.y * (1.f - r.y), a.z * h.z * r.z * (1.f - r.z), a.w * h.w * r.w * (1.f - r.w));
          *reinterpret_cast<float4*>(g.dxp1 + ((int64_t)s * g.B + b) * 2 * H + H + k0) = dl1r;
          put_planes4(g.dl1rP + b * g.Kp + k0, PS, dl1r);
          *reinterpret_cast<float4*>(g.u1 + b * H + k0) = make_float4(a.x * r.x, a.y * r.y, a.z * r.z, a.w * r.w);
        } else if (role == 1) {
          *reinterpret_cast<float4*>(g.uz + b * H + k0) = a;
        } else {
          const float4 d = dd[e], z = zz[e];
          dh[e] = make_float4(d.x * (1.f - z.x) + u1[e].x + uz[e].x + a.x, d.y * (1.f - z.y) + u1[e].y + uz[e].y + a.y, d.z * (1.f - z.z) + u1[e].z + uz[e].z + a.z,
                              d.w * (1.f - z.w) + u1[e].w + uz[e].w + a.w);
          if (s == 0) *reinterpret_cast<float4*>(g.dh0 + b * H + k0) = dh[e];
        }
      }
      if (role == 2 && s > 0) elementwise(s - 1);
      publish(role == 0 ? 1 : (role == 1 ? 2 : 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------------------------ LSTM BPTT
// Per step s (T-1 ... 0), element (b, k), saved gates f, i, o, g, cell states c_s, c_{s-1}, running dh, dc (rnn.py:280-288 reversed):
//   d = dh + g_hs[s];  tc = tanh(c_s);  dct = dc + d o (1 - tc^2)
//   dl = [dct c_{s-1} f (1 - f), dct g i (1 - i), d tc o (1 - o), dct i (1 - g^2)] -> dxp[s];  dc' = dct f;  dh' = dl . Wh^T (K = 4H)
// The contraction over 4H is cut by gate: role q (0..3) holds rows k of Wh[:, qH:(q+1)H] in shared memory and produces the partial
// product of gate q; roles 1..3 hand theirs over in fp32 (part), role 0 adds them to its own accumulator in a fixed order
// (deterministic), then runs the elementwise part of step s - 1 with dh and dc of its elements in registers.
// Counters per batch tile: [0] operand planes of the next step ready (role 0), [1] accumulators drained = planes of this step consumed
// and partials published (all four roles): role 0 overwrites the planes only past it.
struct LstmBwdArgs {
  const float *g_hs, *g_cT, *c0, *cs, *gates;
  float *      dxp, *dh0, *dc0, *part;  // part: [3][B][H]
  __nv_bfloat16* dlP;                   // operand planes [2][B][Kp4]
  int           T, B, H, Kp4, nbt, ctk;
  unsigned int* cnt;  // [nbt][16]
};

__global__ void __launch_bounds__(384, 1)
k_lstm_persist_bwd(const __grid_constant__ CUtensorMap mapDl, const __grid_constant__ CUtensorMap mapWt, LstmBwdArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t*  smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int KB = g.H / GP_BK;
  uint8_t*  wsm = smem;
  uint8_t*  asm_ = smem + (size_t)KB * 2 * GP_WBLOCK;
  uint64_t* full_bar = (uint64_t*)(asm_ + GP_STAGES * GP_ASTAGE);
  uint64_t* empty_bar = full_bar + GP_STAGES;
  uint64_t* wfull_bar = empty_bar + GP_STAGES;
  uint64_t* tfull_bar = wfull_bar + 1;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int          warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int          per_set = g.ctk * g.nbt;
  const int          role = (int)blockIdx.x / per_set;  // gate q
  const int          idx = (int)blockIdx.x - role * per_set;
  const int          c = idx % g.ctk, bt = idx / g.ctk;
  unsigned int*      cnt = g.cnt + bt * 16;
  const unsigned int per = (unsigned int)g.ctk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapWt);
    tma_prefetch_desc(&mapDl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, GP_M64 ? 8 : 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(wfull_bar, (uint32_t)(KB * 2 * GP_WBLOCK));
      for (int kb = 0; kb < KB; ++kb)
        for (int pl = 0; pl < 2; ++pl)
          tma_load_4d(&mapWt, wfull_bar, wsm + (size_t)(kb * 2 + pl) * GP_WBLOCK, role * g.H + kb * GP_BK, c * GP_BN, pl, 0);
    }
    int      stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < g.T; ++i) {
      wait_count(cnt, per * (unsigned int)(i + 1));
      asm volatile("fence.proxy.async.global;" ::: "memory");
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* st = asm_ + stage * GP_ASTAGE;
          mbar_expect_tx(&full_bar[stage], GP_ALOAD);
          tma_load_4d(&mapDl, &full_bar[stage], st, role * g.H + kb * GP_BK, bt * GP_BMV, 0, 0);
          tma_load_4d(&mapDl, &full_bar[stage], st + GP_BM * GP_BK * 2, role * g.H + kb * GP_BK, bt * GP_BMV, 1, 0);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool     leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(GP_M64 ? 64 : GP_BM, GP_BN);
    int            stage = 0;
    uint32_t       phase = 0;
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    for (int i = 0; i < g.T; ++i) {
      mbar_wait(tempty_bar, (uint32_t)(i & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t sa = smem_u32(asm_ + stage * GP_ASTAGE), sb = smem_u32(wsm + (size_t)kb * 2 * GP_WBLOCK);
          const uint64_t d_ahi = make_smem_desc_sw128(sa), d_alo = make_smem_desc_sw128(sa + GP_BM * GP_BK * 2);
          const uint64_t d_bhi = make_smem_desc_sw128(sb), d_blo = make_smem_desc_sw128(sb + GP_WBLOCK);
#pragma unroll
          for (int k = 0; k < GP_BK / 16; ++k) {
            const uint64_t o = 2 * k;
            umma_bf16(tmem_base, d_alo + o, d_bhi + o, idesc, (kb | k) ? 1u : 0u);
            umma_bf16(tmem_base, d_ahi + o, d_blo + o, idesc, 1u);
            umma_bf16(tmem_base, d_ahi + o, d_bhi + o, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    const bool     reader = GP_M64 ? true : (warp & 3) < 2;
    const int      rq = GP_M64 ? (warp & 3) : (warp & 1), rc0 = ((warp - 4) >> 2) * 32;
    const uint32_t taddr = tmem_base + ((uint32_t)(rq * 32) << 16) + (uint32_t)rc0;
    float4*        acc4 = reinterpret_cast<float4*>(asm_);
    const int      ew = warp - 4, cch = lane & 15;
    const int64_t  H = g.H, BH = (int64_t)g.B * H, PS = (int64_t)g.B * g.Kp4;
    const int      k0 = c * GP_BN + 4 * cch;
    int64_t        brow[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) brow[e] = (int64_t)bt * GP_BMV + ew * 8 + 2 * e + (lane >> 4);
    float4 dh[4], dc[4];
    auto   elementwise = [&](int s) {
      const float* cprev = s == 0 ? g.c0 : g.cs + (int64_t)(s - 1) * BH;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t b = brow[e];
        if (b >= g.B) continue;
        const int64_t o1 = ((int64_t)s * g.B + b) * H + k0, o4 = ((int64_t)s * g.B + b) * 4 * H + k0;
        const float4  gg = g.g_hs ? __ldg(reinterpret_cast<const float4*>(g.g_hs + o1)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4  f = __ldg(reinterpret_cast<const float4*>(g.gates + o4)), ii = __ldg(reinterpret_cast<const float4*>(g.gates + o4 + H));
        const float4  o = __ldg(reinterpret_cast<const float4*>(g.gates + o4 + 2 * H)), gt = __ldg(reinterpret_cast<const float4*>(g.gates + o4 + 3 * H));
        const float4  ct = __ldg(reinterpret_cast<const float4*>(g.cs + o1)), cp = __ldg(reinterpret_cast<const float4*>(cprev + b * H + k0));
        float4        v[4];
#define LSTM_BWD_LANE(m)                                                                                  \
  {                                                                                                       \
    const float d = dh[e].m + gg.m, tc = gp_tanh(ct.m), dct = dc[e].m + d * o.m * (1.f - tc * tc);        \
    v[0].m = dct * cp.m * f.m * (1.f - f.m), v[1].m = dct * gt.m * ii.m * (1.f - ii.m);                   \
    v[2].m = d * tc * o.m * (1.f - o.m), v[3].m = dct * ii.m * (1.f - gt.m * gt.m);                       \
    dc[e].m = dct * f.m;                                                                                  \
  }
        LSTM_BWD_LANE(x) LSTM_BWD_LANE(y) LSTM_BWD_LANE(z) LSTM_BWD_LANE(w)
#undef LSTM_BWD_LANE
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<float4*>(g.dxp + o4 + q * H) = v[q];
          put_planes4(g.dlP + b * g.Kp4 + q * H + k0, PS, v[q]);
        }
      }
    };
    auto publish = [&](int which) {
      fence_proxy_async_smem();
      epi_bar();
      if (threadIdx.x == 128) {
        __threadfence();
        atomicAdd(cnt + which, 1u);
      }
    };
    if (role == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dh[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        dc[e] = (g.g_cT && brow[e] < g.B) ? __ldg(reinterpret_cast<const float4*>(g.g_cT + brow[e] * H + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      elementwise(g.T - 1);
      publish(0);
    }
    for (int i = 0; i < g.T; ++i) {
      const int s = g.T - 1 - i;
      if (reader) {
        mbar_wait(tfull_bar, (uint32_t)(i & 1));
        tc_fence_after();
        float v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        const int row = GP_M64 ? rq * 16 + lane : rq * 32 + lane;
        if (!GP_M64 || lane < 16) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc4[row * 16 + ((rc0 / 4 + k) ^ (row & 7))] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
      }
      epi_bar();
      if (role == 0) {  // every CTA of this batch tile has consumed the operand planes of step s (and roles 1..3 have published)
        if (threadIdx.x == 128) atomicAdd(cnt + 1, 1u);
        if (lane == 0) wait_count(cnt + 1, 4u * per * (unsigned int)(i + 1));
        __syncwarp();
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int     row = ew * 8 + 2 * e + (lane >> 4);
        const int64_t b = brow[e];
        if (b >= g.B) continue;
        const float4 a = acc4[row * 16 + (cch ^ (row & 7))];
        if (role != 0) {
          *reinterpret_cast<float4*>(g.part + (int64_t)(role - 1) * BH + b * H + k0) = a;
        } else {
          const float4 p1 = __ldcg(reinterpret_cast<const float4*>(g.part + b * H + k0));
          const float4 p2 = __ldcg(reinterpret_cast<const float4*>(g.part + BH + b * H + k0));
          const float4 p3 = __ldcg(reinterpret_cast<const float4*>(g.part + 2 * BH + b * H + k0));
          dh[e] = make_float4(((a.x + p1.x) + p2.x) + p3.x, ((a.y + p1.y) + p2.y) + p3.y, ((a.z + p1.z) + p2.z) + p3.z, ((a.w + p1.w) + p2.w) + p3.w);
          if (s == 0) *reinterpret_cast<float4*>(g.dh0 + b * H + k0) = dh[e];
        }
      }
      if (role == 0) {
        if (s > 0) elementwise(s - 1);
        else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (brow[e] < g.B) *reinterpret_cast<float4*>(g.dc0 + brow[e] * H + k0) = dc[e];
        }
        publish(0);
      } else {
        publish(1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------------------------ plain RNN
// h_t = act(x_t Wx + b + h_{t-1} Wh), act = tanh | relu (rnn.py:46-49): one product per step and direction, one CTA set.
//   forward   CTA = (64 columns j of Wh, 64-row batch tile); epilogue: h_t -> hs + the next step's operand planes
//   BPTT      CTA = (64 rows k of Wh as K-major planes over j, batch tile): dh' = dl . Wh^T; epilogue = the element-wise part of
//             step s - 1: d = dh' + g_hs[s-1], dl = d act'(h_{s-1}) -> dxp + operand planes. The first dl (step T-1) is the prologue.
// DIR = 0 forward, 1 backward. Counters: [bt] = operand planes of the next step complete.
struct RnnPersistArgs {
  const float *xp, *hs_in, *g_hs;  // fwd: xp [T][B][H]; bwd: hs (saved), g_hs
  float *      hs, *dxp, *dh0;
  __nv_bfloat16 *P0, *P1;          // ping-pong operand planes [2][B][Kp]
  int           T, B, H, Kp, nbt, ctn, relu;
  unsigned int* cnt;
};

template <int DIR>
__global__ void __launch_bounds__(384, 1)
k_rnn_persist(const __grid_constant__ CUtensorMap mapP0, const __grid_constant__ CUtensorMap mapP1, const __grid_constant__ CUtensorMap mapW,
              RnnPersistArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t*  smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int KB = g.H / GP_BK;
  uint8_t*  wsm = smem;
  uint8_t*  asm_ = smem + (size_t)KB * 2 * GP_WBLOCK;
  uint64_t* full_bar = (uint64_t*)(asm_ + GP_STAGES * GP_ASTAGE);
  uint64_t* empty_bar = full_bar + GP_STAGES;
  uint64_t* wfull_bar = empty_bar + GP_STAGES;
  uint64_t* tfull_bar = wfull_bar + 1;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int          warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int          c = (int)blockIdx.x % g.ctn, bt = (int)blockIdx.x / g.ctn;
  unsigned int*      cnt = g.cnt + bt;
  const unsigned int per = (unsigned int)g.ctn;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapP0);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, GP_M64 ? 8 : 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // step i (0-based in processing order) reads planes P[i & 1] and its epilogue writes P[(i + 1) & 1]; forward: P0 holds h0 on
  // entry (count target per * i); backward: the prologue writes P0 (count target per * (i + 1))
  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(wfull_bar, (uint32_t)(KB * 2 * GP_WBLOCK));
      for (int kb = 0; kb < KB; ++kb)
        for (int pl = 0; pl < 2; ++pl) tma_load_4d(&mapW, wfull_bar, wsm + (size_t)(kb * 2 + pl) * GP_WBLOCK, kb * GP_BK, c * GP_BN, pl, 0);
    }
    int      stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < g.T; ++i) {
      const unsigned int target = per * (unsigned int)(DIR == 0 ? i : i + 1);
      if (target) wait_count(cnt, target);
      asm volatile("fence.proxy.async.global;" ::: "memory");
      const CUtensorMap* mA = (i & 1) ? &mapP1 : &mapP0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          uint8_t* st = asm_ + stage * GP_ASTAGE;
          mbar_expect_tx(&full_bar[stage], GP_ALOAD);
          tma_load_4d(mA, &full_bar[stage], st, kb * GP_BK, bt * GP_BMV, 0, 0);
          tma_load_4d(mA, &full_bar[stage], st + GP_BM * GP_BK * 2, kb * GP_BK, bt * GP_BMV, 1, 0);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool     leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(GP_M64 ? 64 : GP_BM, GP_BN);
    int            stage = 0;
    uint32_t       phase = 0;
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    for (int i = 0; i < g.T; ++i) {
      mbar_wait(tempty_bar, (uint32_t)(i & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t sa = smem_u32(asm_ + stage * GP_ASTAGE), sb = smem_u32(wsm + (size_t)kb * 2 * GP_WBLOCK);
          const uint64_t d_ahi = make_smem_desc_sw128(sa), d_alo = make_smem_desc_sw128(sa + GP_BM * GP_BK * 2);
          const uint64_t d_bhi = make_smem_desc_sw128(sb), d_blo = make_smem_desc_sw128(sb + GP_WBLOCK);
#pragma unroll
          for (int k = 0; k < GP_BK / 16; ++k) {
            const uint64_t o = 2 * k;
            umma_bf16(tmem_base, d_alo + o, d_bhi + o, idesc, (kb | k) ? 1u : 0u);
            umma_bf16(tmem_base, d_ahi + o, d_blo + o, idesc, 1u);
            umma_bf16(tmem_base, d_ahi + o, d_bhi + o, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    // M = 64 accumulator: 16 rows per TMEM lane quarter (rows 16 q .. 16 q + 15 in lanes 32 q .. 32 q + 15), so all eight warps read
    const bool     reader = GP_M64 ? true : (warp & 3) < 2;
    const int      rq = GP_M64 ? (warp & 3) : (warp & 1), rc0 = ((warp - 4) >> 2) * 32;
    const uint32_t taddr = tmem_base + ((uint32_t)(rq * 32) << 16) + (uint32_t)rc0;
    float4*        acc4 = reinterpret_cast<float4*>(asm_);
    const int      ew = warp - 4, cch = lane & 15;
    const int64_t  H = g.H, BH = (int64_t)g.B * H, PS = (int64_t)g.B * g.Kp;
    const int      j0 = c * GP_BN + 4 * cch;
    int64_t        brow[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) brow[e] = (int64_t)bt * GP_BMV + ew * 8 + 2 * e + (lane >> 4);
    auto publish = [&]() {
      fence_proxy_async_smem();
      epi_bar();
      if (threadIdx.x == 128) {
        __threadfence();
        atomicAdd(cnt, 1u);
      }
    };
    // BPTT element-wise part of time step s from the running gradient dh: dl = (dh + g_hs[s]) act'(h_s) -> dxp[s], planes `dst`
    auto bwd_elementwise = [&](int s, const float4* dh, __nv_bfloat16* dst) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t b = brow[e];
        if (b >= g.B) continue;
        const int64_t off = ((int64_t)s * g.B + b) * H + j0;
        const float4  gg = g.g_hs ? __ldg(reinterpret_cast<const float4*>(g.g_hs + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4  h = __ldg(reinterpret_cast<const float4*>(g.hs_in + off));
        float4        dl = make_float4(dh[e].x + gg.x, dh[e].y + gg.y, dh[e].z + gg.z, dh[e].w + gg.w);
        if (g.relu) dl.x = h.x > 0.f ? dl.x : 0.f, dl.y = h.y > 0.f ? dl.y : 0.f, dl.z = h.z > 0.f ? dl.z : 0.f, dl.w = h.w > 0.f ? dl.w : 0.f;
        else dl.x *= 1.f - h.x * h.x, dl.y *= 1.f - h.y * h.y, dl.z *= 1.f - h.z * h.z, dl.w *= 1.f - h.w * h.w;
        *reinterpret_cast<float4*>(g.dxp + off) = dl;
        put_planes4(dst + b * g.Kp + j0, PS, dl);
      }
    };
    float4 dh[4];
    if (DIR == 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e) dh[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      bwd_elementwise(g.T - 1, dh, g.P0);
      publish();
    }
    for (int i = 0; i < g.T; ++i) {
      float4 xr[4];
      if (DIR == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (brow[e] < g.B) xr[e] = __ldg(reinterpret_cast<const float4*>(g.xp + ((int64_t)i * g.B + brow[e]) * H + j0));
      }
      if (reader) {
        mbar_wait(tfull_bar, (uint32_t)(i & 1));
        tc_fence_after();
        float v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        const int row = GP_M64 ? rq * 16 + lane : rq * 32 + lane;
        if (!GP_M64 || lane < 16) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc4[row * 16 + ((rc0 / 4 + k) ^ (row & 7))] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
      }
      epi_bar();
      __nv_bfloat16* dst = ((i + 1) & 1) ? g.P1 : g.P0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int     row = ew * 8 + 2 * e + (lane >> 4);
        const int64_t b = brow[e];
        if (b >= g.B) continue;
        float4 a = acc4[row * 16 + (cch ^ (row & 7))];
        if (DIR == 0) {
          a.x += xr[e].x, a.y += xr[e].y, a.z += xr[e].z, a.w += xr[e].w;
          if (g.relu) a.x = fmaxf(a.x, 0.f), a.y = fmaxf(a.y, 0.f), a.z = fmaxf(a.z, 0.f), a.w = fmaxf(a.w, 0.f);
          else a.x = gp_tanh(a.x), a.y = gp_tanh(a.y), a.z = gp_tanh(a.z), a.w = gp_tanh(a.w);
          *reinterpret_cast<float4*>(g.hs + ((int64_t)i * g.B + b) * H + j0) = a;
          put_planes4(dst + b * g.Kp + j0, PS, a);
        } else {
          dh[e] = a;
          if (i == g.T - 1) *reinterpret_cast<float4*>(g.dh0 + b * H + j0) = a;
        }
      }
      if (DIR == 1 && i + 1 < g.T) bwd_elementwise(g.T - 2 - i, dh, dst);
      publish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 64);
}

static bool g_persist_off = getenv("PDN_GRU_PERSIST") && getenv("PDN_GRU_PERSIST")[0] == '0';

bool gru_persist_ok(int64_t T, int64_t B, int64_t H) {
  if (g_persist_off || H % GP_BK != 0 || H / GP_BK > GP_MAXKB || T < 4) return false;
  const int64_t nbt = (B + GP_BMV - 1) / GP_BMV, ctas = (3 * H / GP_BN) * nbt;
  return ctas <= sm_count();
}

// planes: hP0 holds h0 on entry ([2][B][Kp]); W1p / W2p are K-major weight planes [2][2H or H rows][Kp]
int gru_persist_forward(const float* xp1, const float* xp2, const float* h0, const PackedOperand& W1p, const PackedOperand& W2p, const PackedOperand& hP0,
                        const PackedOperand& hP1, const PackedOperand& rhP, float* hs, float* zr, float* nn, int64_t T, int64_t B, int64_t H) {
  const int nbt = (int)((B + GP_BMV - 1) / GP_BMV), ct1 = (int)(2 * H / GP_BN), ct2 = (int)(H / GP_BN);
  CUtensorMap mH0, mH1, mRH, mW1, mW2;
  PDN_TRY(tc_make_map(&mH0, hP0.planes, B, H, hP0.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&mH1, hP1.planes, B, H, hP1.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&mRH, rhP.planes, B, H, rhP.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&mW1, W1p.planes, 2 * H, H, W1p.Kp, 1, GP_BN));
  PDN_TRY(tc_make_map(&mW2, W2p.planes, H, H, W2p.Kp, 1, GP_BN));
  Scratch* cnt = new Scratch();  // freed stream-ordered below
  PDN_TRY(cnt->alloc((size_t)nbt * 4 * sizeof(unsigned int)));
  PDN_CUDA(cudaMemsetAsync(cnt->p, 0, (size_t)nbt * 4 * sizeof(unsigned int), stream()));
  GruPersistArgs g;
  g.xp1 = xp1, g.xp2 = xp2, g.h0 = h0, g.hs = hs, g.zr = zr, g.nn = nn;
  g.hP0 = (__nv_bfloat16*)hP0.planes, g.hP1 = (__nv_bfloat16*)hP1.planes, g.rhP = (__nv_bfloat16*)rhP.planes;
  g.T = (int)T, g.B = (int)B, g.H = (int)H, g.Kp = (int)hP0.Kp, g.nbt = nbt, g.ct1 = ct1, g.ct2 = ct2;
  g.cnt = (unsigned int*)cnt->p;
  g.trace = nullptr;
  static unsigned long long* trace_buf = nullptr;
  if (getenv("PDN_GRU_TRACE")) {
    if (!trace_buf) cudaMalloc(&trace_buf, 32 * 8);
    cudaMemsetAsync(trace_buf, 0, 32 * 8, stream());
    g.trace = trace_buf;
  }
  const size_t smem = (size_t)(H / GP_BK) * 2 * GP_WBLOCK + (size_t)GP_STAGES * GP_ASTAGE + 1024 + 256;
  PDN_CUDA(cudaFuncSetAttribute(k_gru_persist_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {&mH0, &mH1, &mRH, &mW1, &mW2, &g};
  PDN_CUDA(cudaLaunchCooperativeKernel((const void*)k_gru_persist_fwd, dim3((ct1 + ct2) * nbt), dim3(384), params, smem, stream()));
  PDN_LAUNCHED("gru_persist_fwd");
  if (g.trace && T > 100) {
    unsigned long long h[32];
    cudaStreamSynchronize(stream());
    cudaMemcpy(h, g.trace, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[16] = {"dep ready", "proxy fence", "TMA issued", "first tile landed", "MMAs committed", "epilogue waiting", "accumulator ready",
                             "stores issued", "epilogue warps joined", "published", "c0 tmem", "c0 math", "c0 stored", "c1 tmem", "c1 math", "c1 stored"};
    const unsigned long long t0 = h[0];
    for (int ph = 0; ph < 2; ++ph) {
      fprintf(stderr, "[gru trace] step 100, phase %c CTA (ns after phase A's inputs were ready):", ph ? 'B' : 'A');
      for (int i = 0; i < (ph ? 10 : 16); ++i) fprintf(stderr, " %s=%lld", names[i], (long long)(h[ph * 16 + i] - t0));
      fprintf(stderr, "\n");
    }
  }
  delete cnt;  // Scratch returns its block to the stream-ordered allocator: reuse is ordered after this kernel
  return 0;
}

bool rnn_persist_ok(int64_t T, int64_t B, int64_t H) {
  if (g_persist_off || H % GP_BK != 0 || H / GP_BK > GP_MAXKB || T < 4) return false;
  const int64_t nbt = (B + GP_BMV - 1) / GP_BMV;
  return (H / GP_BN) * nbt <= sm_count();
}

// dir 0: forward (Wp = K-major planes of Wh^T: rows = output columns j, contraction over the previous hidden index; P0 holds h0);
// dir 1: BPTT (Wp = K-major planes of Wh: rows = k, contraction over j; hs = the saved hidden states)
int rnn_persist_run(int dir, const float* xp, const float* hs_in, const float* g_hs, const PackedOperand& Wp, const PackedOperand& P0,
                    const PackedOperand& P1, float* hs, float* dxp, float* dh0, int64_t T, int64_t B, int64_t H, int relu) {
  const int nbt = (int)((B + GP_BMV - 1) / GP_BMV), ctn = (int)(H / GP_BN);
  CUtensorMap m0, m1, mW;
  PDN_TRY(tc_make_map(&m0, P0.planes, B, H, P0.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&m1, P1.planes, B, H, P1.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&mW, Wp.planes, H, H, Wp.Kp, 1, GP_BN));
  Scratch cnt;
  PDN_TRY(cnt.alloc((size_t)nbt * sizeof(unsigned int)));
  PDN_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)nbt * sizeof(unsigned int), stream()));
  RnnPersistArgs g;
  g.xp = xp, g.hs_in = hs_in, g.g_hs = g_hs, g.hs = hs, g.dxp = dxp, g.dh0 = dh0;
  g.P0 = (__nv_bfloat16*)P0.planes, g.P1 = (__nv_bfloat16*)P1.planes;
  g.T = (int)T, g.B = (int)B, g.H = (int)H, g.Kp = (int)P0.Kp, g.nbt = nbt, g.ctn = ctn, g.relu = relu;
  g.cnt = (unsigned int*)cnt.p;
  const size_t smem = (size_t)(H / GP_BK) * 2 * GP_WBLOCK + (size_t)GP_STAGES * GP_ASTAGE + 1024 + 256;
  const void*  fn = dir == 0 ? (const void*)k_rnn_persist<0> : (const void*)k_rnn_persist<1>;
  PDN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {&m0, &m1, &mW, &g};
  PDN_CUDA(cudaLaunchCooperativeKernel(fn, dim3(ctn * nbt), dim3(384), params, smem, stream()));
  PDN_LAUNCHED(dir == 0 ? "rnn_persist_fwd" : "rnn_persist_bwd");
  return 0;
}

bool lstm_persist_ok(int64_t T, int64_t B, int64_t H) {
  if (g_persist_off || H % GP_BK != 0 || H / GP_BK > GP_MAXKB || T < 4) return false;
  const int64_t nbt = (B + GP_BMV - 1) / GP_BMV;
  return (H / 16) * nbt <= sm_count();
}

// Wp: K-major planes of Wh as [2][4H rows (output columns)][Kp]; hP0 holds h0 on entry
int lstm_persist_forward(const float* xp, const float* h0, const float* c0, const PackedOperand& Wp, const PackedOperand& hP0, const PackedOperand& hP1,
                         float* hs, float* cs, float* gates, int64_t T, int64_t B, int64_t H) {
  const int nbt = (int)((B + GP_BMV - 1) / GP_BMV), ctj = (int)(H / 16);
  CUtensorMap mH0, mH1, mW;
  PDN_TRY(tc_make_map(&mH0, hP0.planes, B, H, hP0.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&mH1, hP1.planes, B, H, hP1.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&mW, Wp.planes, 4 * H, H, Wp.Kp, 1, 16));
  Scratch cnt;
  PDN_TRY(cnt.alloc((size_t)nbt * sizeof(unsigned int)));
  PDN_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)nbt * sizeof(unsigned int), stream()));
  LstmPersistArgs g;
  g.xp = xp, g.h0 = h0, g.c0 = c0, g.hs = hs, g.cs = cs, g.gates = gates;
  g.hP0 = (__nv_bfloat16*)hP0.planes, g.hP1 = (__nv_bfloat16*)hP1.planes;
  g.T = (int)T, g.B = (int)B, g.H = (int)H, g.Kp = (int)hP0.Kp, g.nbt = nbt, g.ctj = ctj;
  g.cnt = (unsigned int*)cnt.p;
  const size_t smem = (size_t)(H / GP_BK) * 2 * GP_WBLOCK + (size_t)GP_STAGES * GP_ASTAGE + 1024 + 256;
  PDN_CUDA(cudaFuncSetAttribute(k_lstm_persist_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {&mH0, &mH1, &mW, &g};
  PDN_CUDA(cudaLaunchCooperativeKernel((const void*)k_lstm_persist_fwd, dim3(ctj * nbt), dim3(384), params, smem, stream()));
  PDN_LAUNCHED("lstm_persist_fwd");
  return 0;
}

bool lstm_persist_bwd_ok(int64_t T, int64_t B, int64_t H) {
  if (!lstm_persist_ok(T, B, H)) return false;
  const int64_t nbt = (B + GP_BMV - 1) / GP_BMV;
  return 4 * (H / GP_BN) * nbt <= sm_count();
}

// Wt: K-major planes of Wh as [H rows j][Kp4 over the 4H gate columns]; dlP: operand planes [2][B][Kp4]; part: [3][B][H]
int lstm_persist_backward(const float* g_hs, const float* g_cT, const float* c0, const float* cs, const float* gates, const PackedOperand& Wt,
                          const PackedOperand& dlP, float* part, float* dxp, float* dh0, float* dc0, int64_t T, int64_t B, int64_t H) {
  const int nbt = (int)((B + GP_BMV - 1) / GP_BMV), ctk = (int)(H / GP_BN);
  CUtensorMap mDl, mW;
  PDN_TRY(tc_make_map(&mDl, dlP.planes, B, 4 * H, dlP.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&mW, Wt.planes, H, 4 * H, Wt.Kp, 1, GP_BN));
  Scratch cnt;
  PDN_TRY(cnt.alloc((size_t)nbt * 16 * sizeof(unsigned int)));
  PDN_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)nbt * 16 * sizeof(unsigned int), stream()));
  LstmBwdArgs g;
  g.g_hs = g_hs, g.g_cT = g_cT, g.c0 = c0, g.cs = cs, g.gates = gates;
  g.dxp = dxp, g.dh0 = dh0, g.dc0 = dc0, g.part = part, g.dlP = (__nv_bfloat16*)dlP.planes;
  g.T = (int)T, g.B = (int)B, g.H = (int)H, g.Kp4 = (int)dlP.Kp, g.nbt = nbt, g.ctk = ctk;
  g.cnt = (unsigned int*)cnt.p;
  const size_t smem = (size_t)(H / GP_BK) * 2 * GP_WBLOCK + (size_t)GP_STAGES * GP_ASTAGE + 1024 + 256;
  PDN_CUDA(cudaFuncSetAttribute(k_lstm_persist_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {&mDl, &mW, &g};
  PDN_CUDA(cudaLaunchCooperativeKernel((const void*)k_lstm_persist_bwd, dim3(4 * ctk * nbt), dim3(384), params, smem, stream()));
  PDN_LAUNCHED("lstm_persist_bwd");
  return 0;
}

// W2t: K-major planes of Wh2 as [H rows k][Kp over j]; W1t: Wh1 as [H rows k][Kp over 2H columns j]; dl2P / dl1zP / dl1rP / u1 / uz: scratch
int gru_persist_backward(const float* g_hs, const float* h0, const float* hs, const float* zr, const float* nn, const PackedOperand& W2t,
                         const PackedOperand& W1t, const PackedOperand& dl2P, const PackedOperand& dl1zP, const PackedOperand& dl1rP, float* u1, float* uz,
                         float* dxp1, float* dxp2, float* dh0, int64_t T, int64_t B, int64_t H) {
  const int nbt = (int)((B + GP_BMV - 1) / GP_BMV), ctk = (int)(H / GP_BN);
  CUtensorMap m2, m1z, m1r, mW2, mW1;
  PDN_TRY(tc_make_map(&m2, dl2P.planes, B, H, dl2P.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&m1z, dl1zP.planes, B, H, dl1zP.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&m1r, dl1rP.planes, B, H, dl1rP.Kp, 1, GP_BMV));
  PDN_TRY(tc_make_map(&mW2, W2t.planes, H, H, W2t.Kp, 1, GP_BN));
  PDN_TRY(tc_make_map(&mW1, W1t.planes, H, 2 * H, W1t.Kp, 1, GP_BN));
  Scratch cnt;
  PDN_TRY(cnt.alloc((size_t)nbt * 4 * sizeof(unsigned int)));
  PDN_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)nbt * 4 * sizeof(unsigned int), stream()));
  GruBwdArgs g;
  g.g_hs = g_hs, g.h0 = h0, g.hs = hs, g.zr = zr, g.nn = nn, g.dxp1 = dxp1, g.dxp2 = dxp2, g.dh0 = dh0, g.u1 = u1, g.uz = uz;
  g.dl2P = (__nv_bfloat16*)dl2P.planes, g.dl1zP = (__nv_bfloat16*)dl1zP.planes, g.dl1rP = (__nv_bfloat16*)dl1rP.planes;
  g.T = (int)T, g.B = (int)B, g.H = (int)H, g.Kp = (int)dl2P.Kp, g.nbt = nbt, g.ctk = ctk;
  g.cnt = (unsigned int*)cnt.p;
  const size_t smem = (size_t)(H / GP_BK) * 2 * GP_WBLOCK + (size_t)GP_STAGES * GP_ASTAGE + 1024 + 256;
  PDN_CUDA(cudaFuncSetAttribute(k_gru_persist_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {&m2, &m1z, &m1r, &mW2, &mW1, &g};
  PDN_CUDA(cudaLaunchCooperativeKernel((const void*)k_gru_persist_bwd, dim3(3 * ctk * nbt), dim3(384), params, smem, stream()));
  PDN_LAUNCHED("gru_persist_bwd");
  return 0;
}

}  // namespace pdn
