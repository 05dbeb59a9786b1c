// tcgen05 version of the FAR CrossAttention recompute pass (see tc_emm.cu).
#pragma once
#include "common.cuh"

namespace far {
bool tc_emm_supported(int N, int d);
size_t tc_emm_vt_bytes(int G, int N);
// qhi/qlo, khi/klo: dense [G][N][d] hi/lo-split operands (left in the score workspace by the LSE pass).
int tc_emm_pv(const float* qhi, const float* qlo, const float* khi, const float* klo, const float* v, long long sb,
              long long sh, int ldv, const float* pos, int Bpos, int G, int H, int N, int d, float scale,
              const float* rowlse, const float* collse, float* Fpart, float* vtws, size_t vtws_bytes, cudaStream_t st);
}  // namespace far
