// gemm_tc.h — internal interface of the tcgen05 GEMM (gemm_tc.cu) for kernels that produce pre-packed operands
// (conv.cu's im2col gather writes the bf16 hi/lo planes directly instead of materialising an fp32 column matrix).
#pragma once
#include "common.cuh"
#include <cuda.h>
namespace pdn {

// K-major bf16 planes [nbatch][2 (hi, lo)][R][Kp]; pbs = packed-batch index stride per GEMM batch dim (0 = broadcast)
// mn = 0: K-major, planes [nbatch][2][R = operand rows][Kp], K = contraction length.
// mn = 1: MN-major, planes [nbatch][2][R = contraction length][Kp], K = operand rows (M or N) — the operand's own row-major layout
//         when its rows are the unit-stride axis (W [K][N] as the B of x @ W, x [M][K] as the A of x^T @ g): no transposed copy.
struct PackedOperand {
  void*   planes;
  int64_t R, K, Kp, nbatch;
  int64_t pbs[3];
  int     mn = 0;
};

struct TcArgs {
  float* C; const float* bias;
  int64_t M, N, K, ldc;
  int64_t nb[3], c_bs[3];
  int64_t a_pbs[3], b_pbs[3];  // filled by gemm_tc_packed
  int accumulate;
  int splits;       // split-K factor; > 1 => epilogue accumulates with atomics
  int64_t nchw_hw;  // 0: C row-major [M, N] (ldc). > 0: row m = (img, pix) of an NCHW tensor: C[img][col][pix], hw = nchw_hw
  size_t c_clear_bytes;  // extent of C to clear before a split-K launch when C is not a plain [M, N] matrix
  // greedy-decode epilogue: instead of storing C, keep per (row, N-tile) the maximum of (A·B + bias) and its column
  float*     amax_val;   // [M][n_tiles] (nullptr = normal store epilogue)
  long long* amax_idx;   // [M][n_tiles]
  int a_mn = 0, b_mn = 0;  // operand orientation in shared memory (filled by gemm_tc_packed from PackedOperand::mn)
  long long* trace = nullptr;  // debug (PDN_TC_TRACE): clock64 stamps of CTA 0's pipeline stages, nullptr = off
};

int pack_operand_ex(const float* src, int64_t R, int64_t K, int64_t r_stride, int64_t k_stride, int64_t k_inner, int64_t k_outer_stride,
                    const int64_t* nb, const int64_t* bs, Scratch* buf, PackedOperand* out);
// packs `src` (logical [rows][kc] with the given element strides) in its own orientation: K-major, or MN-major when rows are unit-stride
int pack_operand_auto(const float* src, int64_t rows, int64_t kc, int64_t r_stride, int64_t k_stride, const int64_t* nb, const int64_t* bs,
                      Scratch* buf, PackedOperand* out);
// operand planes of `src` in its own orientation, served from the operand-plane cache when version >= 0 (see gemm_tc.cu)
int planes_cached(const float* src, int64_t rows, int64_t kc, int64_t r_stride, int64_t k_stride, const int64_t* nb, const int64_t* bs,
                  long long version, Scratch* buf, PackedOperand* out, bool force_kmajor = false);
// 4-D TMA map over operand planes [batch][2][R][Kp] (bf16), box = 64 (k) x box_rows x 1 x 1, 128-byte swizzle
int tc_make_map(CUtensorMap* map, const void* base, int64_t R, int64_t K, int64_t Kp, int64_t nbatch, int box_rows);
struct ConvGeom;
// TMA-tiled stride-1 convolution (conv_tma.cu)
bool conv_tma_ok(int64_t contr_channels, int stride, int64_t N, int64_t Cc, int64_t Hh, int64_t Ww);
int conv_tma_forward(const float* act, int64_t N, int64_t Cc, int64_t Hh, int64_t Ww, const PackedOperand& wt, const float* bias, float* out,
                     int64_t n_out, int64_t oh, int64_t ow, int k, int pad, int sign, long long act_version);
int conv_tma_bwd_weight(const float* x, const float* gy, float* dw, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t oh, int64_t ow,
                        int k, int pad, long long x_version, long long gy_version);
int gemm_tc_conv(const float* src, const ConvGeom& geom, int mode, int64_t Mtot, int Ktot, const PackedOperand& B, TcArgs t);
int gemm_tc_packed(const PackedOperand& A, const PackedOperand& B, TcArgs t, int splits, int* n_tiles_out = nullptr);
// persistent whole-sequence GRU recurrence (rnn_persist.cu)
bool gru_persist_ok(int64_t T, int64_t B, int64_t H);
int  gru_persist_forward(const float* xp1, const float* xp2, const float* h0, const PackedOperand& W1p, const PackedOperand& W2p, const PackedOperand& hP0,
                         const PackedOperand& hP1, const PackedOperand& rhP, float* hs, float* zr, float* nn, int64_t T, int64_t B, int64_t H);
bool rnn_persist_ok(int64_t T, int64_t B, int64_t H);
int  rnn_persist_run(int dir, const float* xp, const float* hs_in, const float* g_hs, const PackedOperand& Wp, const PackedOperand& P0,
                     const PackedOperand& P1, float* hs, float* dxp, float* dh0, int64_t T, int64_t B, int64_t H, int relu);
bool lstm_persist_ok(int64_t T, int64_t B, int64_t H);
int  lstm_persist_forward(const float* xp, const float* h0, const float* c0, const PackedOperand& Wp, const PackedOperand& hP0, const PackedOperand& hP1,
                          float* hs, float* cs, float* gates, int64_t T, int64_t B, int64_t H);
bool lstm_persist_bwd_ok(int64_t T, int64_t B, int64_t H);
int  lstm_persist_backward(const float* g_hs, const float* g_cT, const float* c0, const float* cs, const float* gates, const PackedOperand& Wt,
                           const PackedOperand& dlP, float* part, float* dxp, float* dh0, float* dc0, int64_t T, int64_t B, int64_t H);
int  gru_persist_backward(const float* g_hs, const float* h0, const float* hs, const float* zr, const float* nn, const PackedOperand& W2t,
                          const PackedOperand& W1t, const PackedOperand& dl2P, const PackedOperand& dl1zP, const PackedOperand& dl1rP, float* u1, float* uz,
                          float* dxp1, float* dxp2, float* dh0, int64_t T, int64_t B, int64_t H);

}  // namespace pdn
