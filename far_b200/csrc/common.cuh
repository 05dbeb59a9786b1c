// Shared device/host helpers for libfar_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/far_sm100.h"

namespace far { extern unsigned long long g_launch_count; }

#define FAR_CHECK_LAUNCH()                                                      \
  do {                                                                          \
    ++::far::g_launch_count;                                                    \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) {                                                    \
      fprintf(stderr, "[far_sm100] %s:%d launch failed: %s\n", __FILE__, __LINE__, \
              cudaGetErrorString(_e));                                          \
      return FAR_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

#define FAR_REQUIRE(cond)                                                       \
  do {                                                                          \
    if (!(cond)) {                                                              \
      fprintf(stderr, "[far_sm100] %s:%d bad argument: %s\n", __FILE__, __LINE__, #cond); \
      return FAR_ERR_ARG;                                                       \
    }                                                                           \
  } while (0)

namespace far {

// Per-kernel device timing (far_profile_* of the C ABI).  `ProfScope s(id, flops, bytes, stream);` around a launch.
enum ProfId {
  PROF_TC_GEMM = 0, PROF_TC_SCORE, PROF_TC_EMM_PV, PROF_LA_REDUCE, PROF_LA_APPLY, PROF_LA_SMALL, PROF_LAYERNORM,
  PROF_LINEAR_SIMT, PROF_FINE_GATHER, PROF_FINE_MATCH, PROF_SPLIT, PROF_EMM_SIMT, PROF_SOLVER, PROF_FPN, PROF_ENC_FUSED,
  PROF_TC_CORRVOL, PROF_EIGHTPT, PROF_TC_FLASH_ATTN, PROF_NUM_IDS
};
extern bool g_prof_on;
void prof_begin(int id, double flops, double bytes, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(int id, double flops, double bytes, cudaStream_t s) : st(s), on(g_prof_on) {
    if (on) prof_begin(id, flops, bytes, st);
  }
  ~ProfScope() {
    if (on) prof_end(st);
  }
};

// SM count of the CURRENT device (B200: 148 = 2 dies x 74), queried once per device -- grids of the persistent kernels
// are sized from it, never from a compile-time constant.
inline int num_sms() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
#define kNumSMs (::far::num_sms())
// true the first time it is called on the current device with this flag array (cudaFuncSetAttribute is per device)
inline bool first_use_on_device(bool (&flags)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}


__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// activation codes of the C-ABI (include/far_sm100.h)
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case FAR_ACT_RELU: return fmaxf(x, 0.f);
    case FAR_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));  // exact-erf nn.GELU
    case FAR_ACT_ELU1: return x > 0.f ? x + 1.f : expm1f(x) + 1.f;                   // F.elu(x) + 1
    case FAR_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}

// online log-sum-exp pair (m = running max, s = sum of exp(x - m))
struct MS {
  float m, s;
};
__device__ __forceinline__ MS ms_init() { return MS{-INFINITY, 0.f}; }
__device__ __forceinline__ MS ms_merge(MS a, MS b) {
  float m = fmaxf(a.m, b.m);
  if (m == -INFINITY) return MS{m, 0.f};
  return MS{m, a.s * expf(a.m - m) + b.s * expf(b.m - m)};
}

}  // namespace far
