// tcgen05 versions of the dual-softmax score passes (see tc_score.cu).
#pragma once
#include "score.cuh"

namespace far {
bool tc_score_supported(const ScoreArgs& a);
size_t tc_score_workspace_bytes(int G, int L, int S, int K);
// `split_done` != 0: the workspace already holds the hi/lo split of exactly these operands (skip the split pass).
int tc_score_lse_partials(const ScoreArgs& a, float2* rowpart, float2* colpart, float* ws, size_t ws_bytes,
                          int split_done, cudaStream_t st);
// K = 64 (EMM) streaming variant: 2 row partials per row, 4 column partials per 128-row tile (tc_score.cu).
bool tc_lse64_supported(const ScoreArgs& a);
int tc_lse64_partials(const ScoreArgs& a, float2* rowpart, float2* colpart, float* ws, size_t ws_bytes, int split_done,
                      cudaStream_t st);
int tc_match_conf(const ScoreArgs& a, const float* rowlse, const float* collse, float2* rowmax, float* colmax,
                  float* conf_out, float* ws, size_t ws_bytes, int split_done, cudaStream_t st);
// Pointers of the dense [G][rows][K] hi/lo operand arrays inside a workspace filled by the passes above.
void tc_score_operand_ptrs(const ScoreArgs& a, float* ws, float** ahi, float** alo, float** bhi, float** blo);
}  // namespace far
