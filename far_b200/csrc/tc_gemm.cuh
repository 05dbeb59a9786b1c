// tcgen05 (5th-gen tensor core) 3xTF32 GEMM engine: TMA-fed, TMEM accumulators.  See tc_gemm.cu.
#pragma once
#include "common.cuh"

namespace far {
// true when the shapes/strides satisfy the TMA + UMMA constraints of the engine
bool tc_linear_supported(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                         int M, int N);
bool tc_linear_preferred(int M, int N, int K);
// whether engine=0 (auto) should pick the tensor-core engine (env FAR_TC=0 disables)
bool tc_engine_default_on();
// whether raw-A GEMMs run their cross terms as bf16 MMAs (tc_common.cuh "cross16"); FAR_TC_CROSS=tf32 disables
namespace tc { bool tc_cross16_on(); }
size_t tc_linear_workspace_bytes(int M, int N, int K);
// exact need for these operands (no activation split buffers when TMA can read the activations in place)
size_t tc_linear_workspace_need(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, int M, int N);
int tc_linear(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
              const float* bias, const float* rowbias, int rowbias_group, float* y, int ldy, int M, int N, int act,
              int act_cols, float* workspace, size_t workspace_bytes, cudaStream_t st, const float* presplit = nullptr);

// Extended entry: fused weight blocks ([W; W2]), pre-split (and optionally per-group) B, grouped rows, and the
// linear-attention normaliser fused into the elu+1 epilogue.
struct TcLinearEx {
  const float* x1; int ldx1, K1;
  const float* x2; int ldx2, K2;
  const float* W; int ldw;          // [N1 (or N), K] weights (ignored when Whi != nullptr)
  const float* W2; int N1;          // optional second block: rows [N1, N) come from W2 [N - N1, K]
  const float* Whi; const float* Wlo;  // pre-split B: [Gb][N][K] hi / lo (tf32-truncated / remainder)
  int wlo_is_cat;                   // pre-split B in the cross16 form: Whi = raw fp32, Wlo = bf16 [lo | hi] per k-block (tc_common.cuh)
  int b_grouped;                    // B has one [N,K] matrix per group
  const float* bias;
  float* y; int ldy;
  int M, N, act, act_cols;
  int G, L;                         // G > 0: M = G * L rows in G groups (tiles do not straddle groups); G = 0: ungrouped
  const float* ksum; int ksum_rec, ksum_off; float eps;  // act == ELU1 and ksum != nullptr: y = (elu(x)+1) * Z,
                                    // Z[row, h] = 1 / (dot(y[row, 32h:32h+32], ksum[(g*N/32 + h)*ksum_rec + ksum_off : +32]) + eps)
  const float* rowbias; int rowbias_group;  // optional: + rowbias[(row / rowbias_group) * N + col]
  // LayerNorm fused across two GEMMs (tc_gemm_epi.cuh GemmArgs::ln_*): ln_stats_out [M][N/32][2] written by this GEMM's
  // epilogue (N % 32 == 0); ln_stats_in [M][K2/32][2] + gamma / beta [K2]: x2 is normalised on the fly (TS kernel only)
  float* ln_stats_out;
  const float* ln_stats_in; const float* ln_gamma; const float* ln_beta; float ln_eps;
  float* workspace; size_t workspace_bytes;
};
int tc_linear_ex(const TcLinearEx& a, cudaStream_t st);
// whether tc_linear_ex can take ln_stats_in (the TMEM-operand kernel is the active GEMM kernel)
bool tc_ln_fusion_available();
}  // namespace far
