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
size_t tc_linear_workspace_bytes(int M, int N, int K);
int tc_linear(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
              const float* bias, float* y, int ldy, int M, int N, int act, int act_cols, float* workspace,
              size_t workspace_bytes, cudaStream_t st);
}  // namespace far
