// 128x128x16 fp32 CUDA-core tile engine shared by every GEMM-shaped kernel of the first correct path
// (linear layers, dual-softmax score passes, EMM bilinear attention).  IEEE fp32 FMA accumulation:
// this is the exactness baseline the tcgen05 3xTF32 path is validated against.
//
// Operands are both "K-contiguous": A[m][k] (ld = lda), B[n][k] (ld = ldb); D[m][n] += sum_k A*B.
// 256 threads; thread (ty = t>>4, tx = t&15) owns rows {ty*4+i, 64+ty*4+i} x cols {tx*4+j, 64+tx*4+j},
// i,j in 0..3 (split micro-tile => conflict-free LDS.128 on both operands).
#pragma once
#include "common.cuh"

namespace far {

constexpr int TBM = 128;
constexpr int TBN = 128;
constexpr int TBK = 16;
constexpr int TLD = 132;  // padded leading dim (floats); 132*4 B keeps float4 alignment
constexpr int kTileThreads = 256;

struct __align__(16) TileSmem {
  float a[2][TBK][TLD];
  float b[2][TBK][TLD];
};  // 33,792 B

__device__ __forceinline__ int tile_row(int ty, int i) { return (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)); }
__device__ __forceinline__ int tile_col(int tx, int j) { return (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4)); }

template <bool kVec4>
__device__ __forceinline__ float4 tile_ld4(const float* __restrict__ base, int ld, int row, int rowsValid, int k,
                                           int K) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < rowsValid) {
    const float* p = base + (size_t)row * ld + k;
    if (kVec4) {
      if (k < K) v = __ldg(reinterpret_cast<const float4*>(p));  // K % 4 == 0 on this path
    } else {
      if (k + 0 < K) v.x = __ldg(p + 0);
      if (k + 1 < K) v.y = __ldg(p + 1);
      if (k + 2 < K) v.z = __ldg(p + 2);
      if (k + 3 < K) v.w = __ldg(p + 3);
    }
  }
  return v;
}

// acc += A[0:mValid, 0:K] * B[0:nValid, 0:K]^T for the CTA's 128x128 tile.  A / B already point at the
// tile's first row and at k = 0 of this segment.  Rows >= mValid / nValid read as zero.
// All 256 threads must call; ends with a __syncthreads() so `sm` can be reused immediately.
template <bool kVec4>
__device__ __forceinline__ void simt_tile_mma(const float* __restrict__ A, int lda, int mValid,
                                              const float* __restrict__ B, int ldb, int nValid, int K,
                                              TileSmem& sm, float (&acc)[8][8]) {
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int lr = t >> 2;        // 0..63
  const int lk = (t & 3) * 4;   // 0,4,8,12
  const int nk = (K + TBK - 1) / TBK;
  if (nk == 0) return;

  float4 ra0, ra1, rb0, rb1;
  ra0 = tile_ld4<kVec4>(A, lda, lr, mValid, lk, K);
  ra1 = tile_ld4<kVec4>(A, lda, lr + 64, mValid, lk, K);
  rb0 = tile_ld4<kVec4>(B, ldb, lr, nValid, lk, K);
  rb1 = tile_ld4<kVec4>(B, ldb, lr + 64, nValid, lk, K);

  auto stash = [&](int buf) {
    sm.a[buf][lk + 0][lr] = ra0.x; sm.a[buf][lk + 1][lr] = ra0.y;
    sm.a[buf][lk + 2][lr] = ra0.z; sm.a[buf][lk + 3][lr] = ra0.w;
    sm.a[buf][lk + 0][lr + 64] = ra1.x; sm.a[buf][lk + 1][lr + 64] = ra1.y;
    sm.a[buf][lk + 2][lr + 64] = ra1.z; sm.a[buf][lk + 3][lr + 64] = ra1.w;
    sm.b[buf][lk + 0][lr] = rb0.x; sm.b[buf][lk + 1][lr] = rb0.y;
    sm.b[buf][lk + 2][lr] = rb0.z; sm.b[buf][lk + 3][lr] = rb0.w;
    sm.b[buf][lk + 0][lr + 64] = rb1.x; sm.b[buf][lk + 1][lr + 64] = rb1.y;
    sm.b[buf][lk + 2][lr + 64] = rb1.z; sm.b[buf][lk + 3][lr + 64] = rb1.w;
  };
  stash(0);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    const bool more = (kt + 1 < nk);
    if (more) {
      const int k = (kt + 1) * TBK + lk;
      ra0 = tile_ld4<kVec4>(A, lda, lr, mValid, k, K);
      ra1 = tile_ld4<kVec4>(A, lda, lr + 64, mValid, k, K);
      rb0 = tile_ld4<kVec4>(B, ldb, lr, nValid, k, K);
      rb1 = tile_ld4<kVec4>(B, ldb, lr + 64, nValid, k, K);
    }
#pragma unroll
    for (int k = 0; k < TBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sm.b[cur][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) stash(cur ^ 1);
    __syncthreads();
  }
}

__device__ __forceinline__ void tile_zero(float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

inline bool ptr_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace far
