// Batched score GEMM + row/column log-sum-exp ("pass A" of every dual-softmax on the path):
//   S[g][i][j] = scale * sum_k A_g[i][k] * B_g[j][k],   g = (batch b, head h)
//   row_lse[g][i] = log sum_j exp S,   col_lse[g][j] = log sum_i exp S
// Used by CoarseMatching (coarse_matching.py:112-118) and the FAR CrossAttention (transformer.py:275-283).
// S is never written to HBM: each 128x128 tile leaves per-row / per-column (max, sumexp) partials that a
// second tiny kernel merges in a fixed order (deterministic).
#pragma once
#include "gemm_tile.cuh"

namespace far {

struct ScoreArgs {
  const float* A; long long sAb, sAh; int lda;  // element offset of group (b,h) = b*sAb + h*sAh; row stride lda
  const float* B; long long sBb, sBh; int ldb;
  int H;        // heads per batch; groups G = nb * H
  int G;
  int L, S, K;  // rows of A, rows of B, contraction length
  float scale;
};

__host__ __device__ inline int score_tiles_i(int L) { return ceil_div(L, TBM); }
__host__ __device__ inline int score_tiles_j(int S) { return ceil_div(S, TBN); }

// workspace (floats) needed by score_lse: row partials + col partials.  The tcgen05 kernels emit two partials per
// 128-wide tile and direction (64-column / 64-row halves), so the scratch is sized for 2*JT and 2*IT.
inline size_t score_lse_scratch_floats(int G, int L, int S) {
  return (size_t)G * 2 * score_tiles_j(S) * L * 2 + (size_t)G * 4 * score_tiles_i(L) * S * 2;  // (K = 64 kernel: 4 column partials per tile)
}

// Computes row_lse [G][L] and col_lse [G][S].  `scratch` >= score_lse_scratch_floats().
// When `tcws` (>= tc_score_workspace_bytes) is given and the shape qualifies, the tile pass runs on the tcgen05
// 3xTF32 pipeline (tc_score.cu) and leaves the hi/lo-split operands in `tcws` for a following recompute pass.
// *used_tc (optional) reports which engine ran.
int score_lse(const ScoreArgs& a, float* row_lse, float* col_lse, float* scratch, cudaStream_t st,
              float* tcws = nullptr, size_t tcws_bytes = 0, int* used_tc = nullptr);

// Device helper: this thread's 8x8 scaled scores of the tile at (i0, j0) of group g.
template <bool kVec4>
__device__ __forceinline__ void score_tile(const ScoreArgs& a, int g, int i0, int j0, TileSmem& sm,
                                           float (&acc)[8][8]) {
  const int b = g / a.H, h = g % a.H;
  const float* Ap = a.A + b * a.sAb + h * a.sAh + (size_t)i0 * a.lda;
  const float* Bp = a.B + b * a.sBb + h * a.sBh + (size_t)j0 * a.ldb;
  tile_zero(acc);
  simt_tile_mma<kVec4>(Ap, a.lda, min(TBM, a.L - i0), Bp, a.ldb, min(TBN, a.S - j0), a.K, sm, acc);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] *= a.scale;
}

inline bool score_vec_ok(const ScoreArgs& a) {
  return ptr_aligned16(a.A) && ptr_aligned16(a.B) && a.lda % 4 == 0 && a.ldb % 4 == 0 && a.K % 4 == 0 &&
         a.sAb % 4 == 0 && a.sAh % 4 == 0 && a.sBb % 4 == 0 && a.sBh % 4 == 0;
}

}  // namespace far
