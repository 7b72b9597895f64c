// rnn.cu — GRU / LSTM whole-sequence forward and BPTT.
//
// The reference runs a Python loop of cells: 18 (GRU) / ~23 (LSTM) eager graph nodes per time step, 4 / 2 small sgemms
// each (rnn.py:268-288, 529-544, 702-708); at T=1024 that is 18k nodes and 92 s per training step on the CPU. Here:
//   * the input projections x·Wx (+b) for ALL time steps are one large tcgen05 GEMM done by the caller (hoisted),
//   * the recurrence is a C++ loop on the device stream with 4 (GRU) / 2 (LSTM) launches per step: a tcgen05 GEMM that
//     accumulates h·Wh straight onto the pre-activation buffer, and a fused gate kernel that also emits the next GEMM's
//     A operand as bf16 hi/lo planes (no separate pack pass); recurrent weights are packed once per sequence,
//   * BPTT mirrors it (2 GEMMs + 2 fused kernels per step for GRU) and the weight gradients are single large GEMMs over
//     all T·B rows after the loop.
// Gate math is the reference's: GRU zr = σ(..) with z = first half, r = second; n = tanh(x Wx2 + (r∘h) Wh2 + b2);
// h' = (1−z)∘h + z∘n. LSTM gate order f, i, o (sigmoid) then g (tanh).
#include "common.cuh"
#include "gemm_tc.h"

namespace pdn {

__device__ __forceinline__ float sigmoid_ref(float x) { return x > 0.f ? 1.f / (1.f + expf(-x)) : 1.f - 1.f / (1.f + expf(x)); }
__device__ __forceinline__ float tanh_ref(float x) { return x > 0.f ? 2.f / (1.f + expf(-2.f * x)) - 1.f : 1.f - 2.f / (1.f + expf(2.f * x)); }

__device__ __forceinline__ void put_planes(__nv_bfloat16* planes, int64_t rows, int64_t Kp, int64_t r, int64_t k, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  planes[r * Kp + k] = h;
  planes[rows * Kp + r * Kp + k] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// fp32 [rows, K] (row stride ld) -> planes [2][rows][Kp]
__global__ void __launch_bounds__(256) k_rows_to_planes(const float* __restrict__ src, int64_t ld, __nv_bfloat16* __restrict__ planes, int64_t rows,
                                                        int64_t K, int64_t Kp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * K; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / K, k = i - r * K;
    put_planes(planes, rows, Kp, r, k, src[r * ld + k]);
  }
}

// ---- GRU forward gates ---------------------------------------------------------------------------------------
// zr[b, 0:2H] holds pre-activations on entry, activations on exit; rh planes <- r ∘ h_prev
__global__ void __launch_bounds__(256) k_gru_gate1(float* __restrict__ zr, const float* __restrict__ hprev, __nv_bfloat16* __restrict__ rh_planes,
                                                   int64_t B, int64_t H, int64_t Kp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * H; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / H, j = i - b * H;
    float z = sigmoid_ref(zr[b * 2 * H + j]);
    float r = sigmoid_ref(zr[b * 2 * H + H + j]);
    zr[b * 2 * H + j] = z;
    zr[b * 2 * H + H + j] = r;
    put_planes(rh_planes, B, Kp, b, j, r * hprev[i]);
  }
}
// nn holds pre-activation on entry, n on exit; h = (1-z) hprev + z n -> hs_t and the next step's A planes
__global__ void __launch_bounds__(256) k_gru_gate2(float* __restrict__ nn, const float* __restrict__ zr, const float* __restrict__ hprev,
                                                   float* __restrict__ hout, __nv_bfloat16* __restrict__ h_planes, int64_t B, int64_t H, int64_t Kp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * H; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / H, j = i - b * H;
    float n = tanh_ref(nn[i]);
    float z = zr[b * 2 * H + j];
    float h = (1.f - z) * hprev[i] + z * n;
    nn[i] = n;
    hout[i] = h;
    put_planes(h_planes, B, Kp, b, j, h);
  }
}

// ---- GRU backward ----------------------------------------------------------------------------------------------
// dh_tot = dh + g_hs[t]; dl2 = dh_tot z (1-n^2) -> dxp2[t] + planes; dl1_z = dh_tot (n - hprev) z (1-z) -> dxp1[t, :H];
// dh <- dh_tot (1-z)
__global__ void __launch_bounds__(256) k_gru_bwd1(float* __restrict__ dh, const float* __restrict__ g_t, const float* __restrict__ zr,
                                                  const float* __restrict__ nn, const float* __restrict__ hprev, float* __restrict__ dxp1,
                                                  float* __restrict__ dxp2, __nv_bfloat16* __restrict__ dl2_planes, int64_t B, int64_t H, int64_t Kp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * H; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / H, j = i - b * H;
    float d = dh[i] + (g_t ? g_t[i] : 0.f);
    float z = zr[b * 2 * H + j], n = nn[i];
    float dl2 = d * z * (1.f - n * n);
    dxp2[i] = dl2;
    put_planes(dl2_planes, B, Kp, b, j, dl2);
    dxp1[b * 2 * H + j] = d * (n - hprev[i]) * z * (1.f - z);
    dh[i] = d * (1.f - z);
  }
}
// drh = dl2 Wh2ᵀ ; dl1_r = drh hprev r (1-r) -> dxp1[t, H:] ; dh += drh r ; planes(dl1 = dxp1[t]) for the next GEMM
__global__ void __launch_bounds__(256) k_gru_bwd2(float* __restrict__ dh, const float* __restrict__ drh, const float* __restrict__ zr,
                                                  const float* __restrict__ hprev, float* __restrict__ dxp1, __nv_bfloat16* __restrict__ dl1_planes,
                                                  int64_t B, int64_t H, int64_t Kp2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * H; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / H, j = i - b * H;
    float r = zr[b * 2 * H + H + j], d = drh[i];
    float dl1r = d * hprev[i] * r * (1.f - r);
    dxp1[b * 2 * H + H + j] = dl1r;
    dh[i] += d * r;
    put_planes(dl1_planes, B, Kp2, b, H + j, dl1r);
    put_planes(dl1_planes, B, Kp2, b, j, dxp1[b * 2 * H + j]);
  }
}
// rh[t] = r[t] ∘ hprev[t] for all t (operand of dWh2)
__global__ void __launch_bounds__(256) k_gru_rh_all(const float* __restrict__ zr, const float* __restrict__ h0, const float* __restrict__ hs,
                                                    float* __restrict__ rh, int64_t T, int64_t B, int64_t H) {
  const int64_t BH = B * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T * BH; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i / BH, e = i - t * BH, b = e / H, j = e - b * H;
    float hp = t == 0 ? h0[e] : hs[(t - 1) * BH + e];
    rh[i] = zr[(t * B + b) * 2 * H + H + j] * hp;
  }
}

// ---- plain RNN -------------------------------------------------------------------------------------------------
// h = act(lin) in place (lin = x Wx + b + h_prev Wh, accumulated by the GEMM), + the next step's operand planes
__global__ void __launch_bounds__(256) k_rnn_act(float* __restrict__ h, __nv_bfloat16* __restrict__ planes, int64_t B, int64_t H, int64_t Kp,
                                                 int relu) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < B * H; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / H, j = idx - b * H;
    const float   v = relu ? fmaxf(h[idx], 0.f) : tanh_ref(h[idx]);
    h[idx] = v;
    put_planes(planes, B, Kp, b, j, v);
  }
}
// dl = (dh + g_t) act'(h_t) -> dxp[t] and operand planes
__global__ void __launch_bounds__(256) k_rnn_bwd(const float* __restrict__ dh, const float* __restrict__ g_t, const float* __restrict__ h_t,
                                                 float* __restrict__ dxp, __nv_bfloat16* __restrict__ planes, int64_t B, int64_t H, int64_t Kp,
                                                 int relu) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < B * H; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / H, j = idx - b * H;
    const float   d = dh[idx] + (g_t ? g_t[idx] : 0.f), h = h_t[idx];
    const float   v = relu ? (h > 0.f ? d : 0.f) : d * (1.f - h * h);
    dxp[idx] = v;
    put_planes(planes, B, Kp, b, j, v);
  }
}

// ---- LSTM ------------------------------------------------------------------------------------------------------
// gates[b, 0:4H]: pre-activations in, activations (f, i, o, g) out; c = f c_prev + i g ; h = o tanh(c)
__global__ void __launch_bounds__(256) k_lstm_gate(float* __restrict__ gates, const float* __restrict__ cprev, float* __restrict__ cout,
                                                   float* __restrict__ hout, __nv_bfloat16* __restrict__ h_planes, int64_t B, int64_t H, int64_t Kp) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < B * H; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = idx / H, j = idx - b * H;
    float* gr = gates + b * 4 * H;
    float f = sigmoid_ref(gr[j]), i = sigmoid_ref(gr[H + j]), o = sigmoid_ref(gr[2 * H + j]), g = tanh_ref(gr[3 * H + j]);
    gr[j] = f; gr[H + j] = i; gr[2 * H + j] = o; gr[3 * H + j] = g;
    float c = f * cprev[idx] + i * g;
    float h = o * tanh_ref(c);
    cout[idx] = c;
    hout[idx] = h;
    put_planes(h_planes, B, Kp, b, j, h);
  }
}
__global__ void __launch_bounds__(256) k_lstm_bwd(const float* __restrict__ dh, float* __restrict__ dc, const float* __restrict__ g_t,
                                                  const float* __restrict__ gates, const float* __restrict__ c_t, const float* __restrict__ cprev,
                                                  float* __restrict__ dxp, __nv_bfloat16* __restrict__ dl_planes, int64_t B, int64_t H, int64_t Kp4) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < B * H; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = idx / H, j = idx - b * H;
    const float* gr = gates + b * 4 * H;
    float f = gr[j], i = gr[H + j], o = gr[2 * H + j], g = gr[3 * H + j];
    float d = dh[idx] + (g_t ? g_t[idx] : 0.f);
    float tc = tanh_ref(c_t[idx]);
    float dct = dc[idx] + d * o * (1.f - tc * tc);
    float v[4] = {dct * cprev[idx] * f * (1.f - f), dct * g * i * (1.f - i), d * tc * o * (1.f - o), dct * i * (1.f - g * g)};
    dc[idx] = dct * f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dxp[b * 4 * H + q * H + j] = v[q];
      put_planes(dl_planes, B, Kp4, b, q * H + j, v[q]);
    }
  }
}

static inline int64_t kpad(int64_t k) { return (k + 7) & ~(int64_t)7; }

static void tc_init(TcArgs& t, float* C, int64_t M, int64_t N, int64_t K, int accumulate) {
  for (int i = 0; i < 3; ++i) { t.nb[i] = 1; t.c_bs[i] = 0; t.a_pbs[i] = 0; t.b_pbs[i] = 0; }
  t.C = C; t.bias = nullptr; t.M = M; t.N = N; t.K = K; t.ldc = N; t.accumulate = accumulate; t.splits = 1; t.nchw_hw = 0; t.c_clear_bytes = 0; t.amax_val = nullptr; t.amax_idx = nullptr;
}

static int alloc_planes(Scratch* s, PackedOperand* op, int64_t rows, int64_t K) {
  const int64_t Kp = kpad(K);
  PDN_TRY(s->alloc((size_t)2 * rows * Kp * sizeof(__nv_bfloat16)));
  PDN_CUDA(cudaMemsetAsync(s->p, 0, (size_t)2 * rows * Kp * sizeof(__nv_bfloat16), stream()));
  op->planes = s->p; op->R = rows; op->K = K; op->Kp = Kp; op->nbatch = 1;
  op->pbs[0] = op->pbs[1] = op->pbs[2] = 0;
  return 0;
}

static const int64_t kOne[3] = {1, 1, 1}, kZero[3] = {0, 0, 0};

}  // namespace pdn

using namespace pdn;

extern "C" {

int pdn_gru_seq_fwd(const float* xp1, const float* xp2, const float* h0, const float* Wh1, const float* Wh2, float* hs, float* zr, float* nn,
                    int64_t T, int64_t B, int64_t H) {
  PDN_TRY(ensure_init());
  if (T == 0 || B == 0 || H == 0) return 0;
  const int64_t BH = B * H;
  Scratch       sW1, sW2, sH, sRH;
  PackedOperand W1, W2, Hp, RHp;
  // B operands: rows = output column n, k = hidden index: Wh[k, n]
  PDN_TRY(pack_operand_ex(Wh1, 2 * H, H, 1, 2 * H, 0, 0, kOne, kZero, &sW1, &W1));
  PDN_TRY(pack_operand_ex(Wh2, H, H, 1, H, 0, 0, kOne, kZero, &sW2, &W2));
  PDN_TRY(alloc_planes(&sH, &Hp, B, H));
  PDN_TRY(alloc_planes(&sRH, &RHp, B, H));
  const int grd = grid_for(BH, 256);
  k_rows_to_planes<<<grd, 256, 0, stream()>>>(h0, H, (__nv_bfloat16*)Hp.planes, B, H, Hp.Kp);
  PDN_LAUNCHED("rows_to_planes");
  if (gru_persist_ok(T, B, H)) {  // the whole recurrence as ONE persistent cooperative launch (rnn_persist.cu)
    Scratch       sH1;
    PackedOperand Hp1;
    PDN_TRY(alloc_planes(&sH1, &Hp1, B, H));
    return gru_persist_forward(xp1, xp2, h0, W1, W2, Hp, Hp1, RHp, hs, zr, nn, T, B, H);
  }
  PDN_CUDA(cudaMemcpyAsync(zr, xp1, (size_t)T * BH * 2 * sizeof(float), cudaMemcpyDeviceToDevice, stream()));
  PDN_CUDA(cudaMemcpyAsync(nn, xp2, (size_t)T * BH * sizeof(float), cudaMemcpyDeviceToDevice, stream()));
  TcArgs t;
  for (int64_t s = 0; s < T; ++s) {
    const float* hprev = s == 0 ? h0 : hs + (s - 1) * BH;
    tc_init(t, zr + s * BH * 2, B, 2 * H, H, 1);
    PDN_TRY(gemm_tc_packed(Hp, W1, t, 1));
    k_gru_gate1<<<grd, 256, 0, stream()>>>(zr + s * BH * 2, hprev, (__nv_bfloat16*)RHp.planes, B, H, RHp.Kp);
    PDN_LAUNCHED("gru_gate1");
    tc_init(t, nn + s * BH, B, H, H, 1);
    PDN_TRY(gemm_tc_packed(RHp, W2, t, 1));
    k_gru_gate2<<<grd, 256, 0, stream()>>>(nn + s * BH, zr + s * BH * 2, hprev, hs + s * BH, (__nv_bfloat16*)Hp.planes, B, H, Hp.Kp);
    PDN_LAUNCHED("gru_gate2");
  }
  return 0;
}

int pdn_gru_seq_bwd(const float* g_hs, const float* h0, const float* hs, const float* zr, const float* nn, const float* Wh1, const float* Wh2,
                    float* dxp1, float* dxp2, float* dh0, float* dWh1, float* dWh2, int64_t T, int64_t B, int64_t H) {
  PDN_TRY(ensure_init());
  if (T == 0 || B == 0 || H == 0) return 0;
  const int64_t BH = B * H;
  Scratch       sW1, sW2, sD2, sD1, sDrh, sRh;
  PackedOperand W1t, W2t, D2p, D1p;
  // dl2 · Wh2ᵀ : rows = j (output), k: Wh2[j, k]   |   dl1 · Wh1ᵀ : rows = j, k over 2H: Wh1[j, k]
  PDN_TRY(pack_operand_ex(Wh2, H, H, H, 1, 0, 0, kOne, kZero, &sW2, &W2t));
  PDN_TRY(pack_operand_ex(Wh1, H, 2 * H, 2 * H, 1, 0, 0, kOne, kZero, &sW1, &W1t));
  PDN_TRY(alloc_planes(&sD2, &D2p, B, H));
  if (gru_persist_ok(T, B, H)) {  // BPTT of the whole sequence as ONE persistent cooperative launch (rnn_persist.cu)
    Scratch       sZ, sR, sU1, sUz;
    PackedOperand D1z, D1r;
    PDN_TRY(alloc_planes(&sZ, &D1z, B, H));
    PDN_TRY(alloc_planes(&sR, &D1r, B, H));
    PDN_TRY(sU1.alloc((size_t)BH * sizeof(float)));
    PDN_TRY(sUz.alloc((size_t)BH * sizeof(float)));
    PDN_TRY(gru_persist_backward(g_hs, h0, hs, zr, nn, W2t, W1t, D2p, D1z, D1r, (float*)sU1.p, (float*)sUz.p, dxp1, dxp2, dh0, T, B, H));
  } else {
  PDN_TRY(alloc_planes(&sD1, &D1p, B, 2 * H));
  PDN_TRY(sDrh.alloc((size_t)BH * sizeof(float)));
  float* dh = dh0;  // running gradient wrt the hidden state lives in the dh0 output buffer
  PDN_CUDA(cudaMemsetAsync(dh, 0, (size_t)BH * sizeof(float), stream()));
  const int grd = grid_for(BH, 256);
  TcArgs t;
  for (int64_t s = T - 1; s >= 0; --s) {
    const float* hprev = s == 0 ? h0 : hs + (s - 1) * BH;
    k_gru_bwd1<<<grd, 256, 0, stream()>>>(dh, g_hs ? g_hs + s * BH : nullptr, zr + s * BH * 2, nn + s * BH, hprev, dxp1 + s * BH * 2,
                                          dxp2 + s * BH, (__nv_bfloat16*)D2p.planes, B, H, D2p.Kp);
    PDN_LAUNCHED("gru_bwd1");
    tc_init(t, (float*)sDrh.p, B, H, H, 0);
    PDN_TRY(gemm_tc_packed(D2p, W2t, t, 1));
    k_gru_bwd2<<<grd, 256, 0, stream()>>>(dh, (const float*)sDrh.p, zr + s * BH * 2, hprev, dxp1 + s * BH * 2, (__nv_bfloat16*)D1p.planes, B, H,
                                          D1p.Kp);
    PDN_LAUNCHED("gru_bwd2");
    tc_init(t, dh, B, H, 2 * H, 1);
    PDN_TRY(gemm_tc_packed(D1p, W1t, t, 1));
  }
  }
  // weight gradients: single GEMMs over all T*B rows.  dWh1 = Hprevᵀ · dxp1 with Hprev = [h0 ; hs[0..T-2]]
  if (dWh1) {
    PDN_TRY(pdn_gemm(PDN_F32, h0, dxp1, dWh1, H, 2 * H, B, 1, H, 2 * H, 1, 2 * H, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0));
    if (T > 1)
      PDN_TRY(pdn_gemm(PDN_F32, hs, dxp1 + BH * 2, dWh1, H, 2 * H, (T - 1) * B, 1, H, 2 * H, 1, 2 * H, nullptr, nullptr, nullptr, nullptr, nullptr,
                       1, 0));
  }
  if (dWh2) {
    PDN_TRY(sRh.alloc((size_t)T * BH * sizeof(float)));
    k_gru_rh_all<<<grid_for(T * BH, 256), 256, 0, stream()>>>(zr, h0, hs, (float*)sRh.p, T, B, H);
    PDN_LAUNCHED("gru_rh_all");
    PDN_TRY(pdn_gemm(PDN_F32, sRh.p, dxp2, dWh2, H, H, T * B, 1, H, H, 1, H, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0));
  }
  return 0;
}

// hs[t] = act(xp[t] + hs[t-1] Wh), hs[-1] = h0 (reference rnn.py:46-49 applied over a sequence, rnn.py:196-214)
int pdn_rnn_seq_fwd(const float* xp, const float* h0, const float* Wh, float* hs, int64_t T, int64_t B, int64_t H, int relu) {
  PDN_TRY(ensure_init());
  if (T == 0 || B == 0 || H == 0) return 0;
  const int64_t BH = B * H;
  Scratch       sW, sH;
  PackedOperand W, Hp;
  PDN_TRY(pack_operand_ex(Wh, H, H, 1, H, 0, 0, kOne, kZero, &sW, &W));  // rows = output column j, contraction over k: Wh[k, j]
  PDN_TRY(alloc_planes(&sH, &Hp, B, H));
  const int grd = grid_for(BH, 256);
  k_rows_to_planes<<<grd, 256, 0, stream()>>>(h0, H, (__nv_bfloat16*)Hp.planes, B, H, Hp.Kp);
  PDN_LAUNCHED("rows_to_planes");
  if (rnn_persist_ok(T, B, H)) {
    Scratch       sH1;
    PackedOperand Hp1;
    PDN_TRY(alloc_planes(&sH1, &Hp1, B, H));
    return rnn_persist_run(0, xp, nullptr, nullptr, W, Hp, Hp1, hs, nullptr, nullptr, T, B, H, relu);
  }
  PDN_CUDA(cudaMemcpyAsync(hs, xp, (size_t)T * BH * sizeof(float), cudaMemcpyDeviceToDevice, stream()));
  TcArgs t;
  for (int64_t s = 0; s < T; ++s) {
    tc_init(t, hs + s * BH, B, H, H, 1);
    PDN_TRY(gemm_tc_packed(Hp, W, t, 1));
    k_rnn_act<<<grd, 256, 0, stream()>>>(hs + s * BH, (__nv_bfloat16*)Hp.planes, B, H, Hp.Kp, relu);
    PDN_LAUNCHED("rnn_act");
  }
  return 0;
}

int pdn_rnn_seq_bwd(const float* g_hs, const float* h0, const float* hs, const float* Wh, float* dxp, float* dh0, float* dWh, int64_t T,
                    int64_t B, int64_t H, int relu) {
  PDN_TRY(ensure_init());
  if (T == 0 || B == 0 || H == 0) return 0;
  const int64_t BH = B * H;
  Scratch       sW, sD;
  PackedOperand Wt, Dp;
  PDN_TRY(pack_operand_ex(Wh, H, H, H, 1, 0, 0, kOne, kZero, &sW, &Wt));  // dl . Wh^T: rows = k, contraction over j: Wh[k, j]
  PDN_TRY(alloc_planes(&sD, &Dp, B, H));
  if (rnn_persist_ok(T, B, H)) {
    Scratch       sD1;
    PackedOperand Dp1;
    PDN_TRY(alloc_planes(&sD1, &Dp1, B, H));
    PDN_TRY(rnn_persist_run(1, nullptr, hs, g_hs, Wt, Dp, Dp1, nullptr, dxp, dh0, T, B, H, relu));
  } else {
    float* dh = dh0;
    PDN_CUDA(cudaMemsetAsync(dh, 0, (size_t)BH * sizeof(float), stream()));
    const int grd = grid_for(BH, 256);
    TcArgs    t;
    for (int64_t s = T - 1; s >= 0; --s) {
      k_rnn_bwd<<<grd, 256, 0, stream()>>>(dh, g_hs ? g_hs + s * BH : nullptr, hs + s * BH, dxp + s * BH, (__nv_bfloat16*)Dp.planes, B, H, Dp.Kp,
                                           relu);
      PDN_LAUNCHED("rnn_bwd");
      tc_init(t, dh, B, H, H, 0);
      PDN_TRY(gemm_tc_packed(Dp, Wt, t, 1));
    }
  }
  if (dWh) {  // dWh = Hprev^T . dxp with Hprev = [h0 ; hs[0..T-2]]
    PDN_TRY(pdn_gemm(PDN_F32, h0, dxp, dWh, H, H, B, 1, H, H, 1, H, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0));
    if (T > 1)
      PDN_TRY(pdn_gemm(PDN_F32, hs, dxp + BH, dWh, H, H, (T - 1) * B, 1, H, H, 1, H, nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0));
  }
  return 0;
}

int pdn_lstm_seq_fwd(const float* xp, const float* h0, const float* c0, const float* Wh, float* hs, float* cs, float* gates, int64_t T, int64_t B,
                     int64_t H) {
  PDN_TRY(ensure_init());
  if (T == 0 || B == 0 || H == 0) return 0;
  const int64_t BH = B * H;
  Scratch       sW, sH;
  PackedOperand W, Hp;
  PDN_TRY(pack_operand_ex(Wh, 4 * H, H, 1, 4 * H, 0, 0, kOne, kZero, &sW, &W));
  PDN_TRY(alloc_planes(&sH, &Hp, B, H));
  const int grd = grid_for(BH, 256);
  k_rows_to_planes<<<grd, 256, 0, stream()>>>(h0, H, (__nv_bfloat16*)Hp.planes, B, H, Hp.Kp);
  PDN_LAUNCHED("rows_to_planes");
  if (lstm_persist_ok(T, B, H)) {  // the whole recurrence as ONE persistent cooperative launch (rnn_persist.cu)
    Scratch       sH1;
    PackedOperand Hp1;
    PDN_TRY(alloc_planes(&sH1, &Hp1, B, H));
    return lstm_persist_forward(xp, h0, c0, W, Hp, Hp1, hs, cs, gates, T, B, H);
  }
  PDN_CUDA(cudaMemcpyAsync(gates, xp, (size_t)T * BH * 4 * sizeof(float), cudaMemcpyDeviceToDevice, stream()));
  TcArgs t;
  for (int64_t s = 0; s < T; ++s) {
    tc_init(t, gates + s * BH * 4, B, 4 * H, H, 1);
    PDN_TRY(gemm_tc_packed(Hp, W, t, 1));
    k_lstm_gate<<<grd, 256, 0, stream()>>>(gates + s * BH * 4, s == 0 ? c0 : cs + (s - 1) * BH, cs + s * BH, hs + s * BH,
                                           (__nv_bfloat16*)Hp.planes, B, H, Hp.Kp);
    PDN_LAUNCHED("lstm_gate");
  }
  return 0;
}

int pdn_lstm_seq_bwd(const float* g_hs, const float* g_cT, const float* h0, const float* c0, const float* hs, const float* cs, const float* gates,
                     const float* Wh, float* dxp, float* dh0, float* dc0, float* dWh, int64_t T, int64_t B, int64_t H) {
  PDN_TRY(ensure_init());
  if (T == 0 || B == 0 || H == 0) return 0;
  const int64_t BH = B * H;
  Scratch       sW, sD;
  PackedOperand Wt, Dp;
  // dlin · Whᵀ : rows = j (hidden), k over 4H: Wh[j, k]
  PDN_TRY(pack_operand_ex(Wh, H, 4 * H, 4 * H, 1, 0, 0, kOne, kZero, &sW, &Wt));
  PDN_TRY(alloc_planes(&sD, &Dp, B, 4 * H));
  if (lstm_persist_bwd_ok(T, B, H)) {  // BPTT of the whole sequence as ONE persistent cooperative launch (rnn_persist.cu)
    Scratch sP;
    PDN_TRY(sP.alloc((size_t)3 * BH * sizeof(float)));
    PDN_TRY(lstm_persist_backward(g_hs, g_cT, c0, cs, gates, Wt, Dp, (float*)sP.p, dxp, dh0, dc0, T, B, H));
  } else {
  float *dh = dh0, *dc = dc0;
  PDN_CUDA(cudaMemsetAsync(dh, 0, (size_t)BH * sizeof(float), stream()));
  if (g_cT) PDN_CUDA(cudaMemcpyAsync(dc, g_cT, (size_t)BH * sizeof(float), cudaMemcpyDeviceToDevice, stream()));
  else PDN_CUDA(cudaMemsetAsync(dc, 0, (size_t)BH * sizeof(float), stream()));
  const int grd = grid_for(BH, 256);
  TcArgs t;
  for (int64_t s = T - 1; s >= 0; --s) {
    k_lstm_bwd<<<grd, 256, 0, stream()>>>(dh, dc, g_hs ? g_hs + s * BH : nullptr, gates + s * BH * 4, cs + s * BH,
                                          s == 0 ? c0 : cs + (s - 1) * BH, dxp + s * BH * 4, (__nv_bfloat16*)Dp.planes, B, H, Dp.Kp);
    PDN_LAUNCHED("lstm_bwd");
    tc_init(t, dh, B, H, 4 * H, 0);
    PDN_TRY(gemm_tc_packed(Dp, Wt, t, 1));
  }
  }
  if (dWh) {
    PDN_TRY(pdn_gemm(PDN_F32, h0, dxp, dWh, H, 4 * H, B, 1, H, 4 * H, 1, 4 * H, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0));
    if (T > 1)
      PDN_TRY(pdn_gemm(PDN_F32, hs, dxp + BH * 4, dWh, H, 4 * H, (T - 1) * B, 1, H, 4 * H, 1, 4 * H, nullptr, nullptr, nullptr, nullptr, nullptr, 1,
                       0));
  }
  return 0;
}

}  // extern "C"
