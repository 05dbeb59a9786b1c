// tcgen05 3xTF32 engine -- placeholder until the TMA/TMEM kernel lands (next milestone).
#include "tc_gemm.cuh"
namespace far {
bool tc_linear_supported(const float*, int, int, const float*, int, int, const float*, int, int, int) { return false; }
bool tc_engine_default_on() { return false; }
size_t tc_linear_workspace_bytes(int, int, int) { return 0; }
int tc_linear(const float*, int, int, const float*, int, int, const float*, int, const float*, float*, int, int, int,
              int, int, float*, size_t, cudaStream_t) { return FAR_ERR_ARG; }
}  // namespace far
