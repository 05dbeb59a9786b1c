#include "score.cuh"
#include "tc_score.cuh"

namespace far {

// grid (JT, IT, G); block 256.  rowpart[(g*JT + jt)*L + i] = (m, s) ; colpart[(g*IT + it)*S + j] = (m, s)
template <bool kVec4>
__global__ void __launch_bounds__(kTileThreads, 2) score_lse_kernel(ScoreArgs a, float2* __restrict__ rowpart,
                                                                    float2* __restrict__ colpart) {
  __shared__ TileSmem sm;
  __shared__ float2 colred[8][TBN];
  const int jt = blockIdx.x, it = blockIdx.y, g = blockIdx.z;
  const int JT = gridDim.x, IT = gridDim.y;
  const int i0 = it * TBM, j0 = jt * TBN;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4, warp = t >> 5, lane = t & 31;
  float acc[8][8];
  score_tile<kVec4>(a, g, i0, j0, sm, acc);

  // ---- row partials: reduce over this tile's columns
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j0 + tile_col(tx, j) < a.S) m = fmaxf(m, acc[i][j]);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j0 + tile_col(tx, j) < a.S) s += expf(acc[i][j] - m);
    MS v{m, s};
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      MS u{__shfl_xor_sync(0xffffffffu, v.m, o), __shfl_xor_sync(0xffffffffu, v.s, o)};
      v = ms_merge(v, u);
    }
    const int r = i0 + tile_row(ty, i);
    if (tx == 0 && r < a.L) rowpart[((size_t)g * JT + jt) * a.L + r] = make_float2(v.m, v.s);
  }
  // ---- column partials: reduce over this tile's rows
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i0 + tile_row(ty, i) < a.L) m = fmaxf(m, acc[i][j]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i0 + tile_row(ty, i) < a.L) s += expf(acc[i][j] - m);
    MS v{m, s};
    MS u{__shfl_xor_sync(0xffffffffu, v.m, 16), __shfl_xor_sync(0xffffffffu, v.s, 16)};
    v = ms_merge(v, u);
    if (lane < 16) colred[warp][tile_col(tx, j)] = make_float2(v.m, v.s);
  }
  __syncthreads();
  if (t < TBN) {
    MS v = ms_init();
#pragma unroll
    for (int w = 0; w < 8; ++w) v = ms_merge(v, MS{colred[w][t].x, colred[w][t].y});
    const int c = j0 + t;
    if (c < a.S) colpart[((size_t)g * IT + it) * a.S + c] = make_float2(v.m, v.s);
  }
}

// lse[g][x] = merge over tiles (fixed order) -> m + log(s)
__global__ void lse_finalize_kernel(const float2* __restrict__ part, int tiles, int len, long long total,
                                    float* __restrict__ lse, int base2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long g = idx / len;
  const int x = (int)(idx % len);
  if (base2) {  // partials are (max, sum 2^(x - max)) of log2(e)-scaled scores (tcgen05 kernels)
    float m = -INFINITY;
    for (int tl = 0; tl < tiles; ++tl) m = fmaxf(m, part[((size_t)g * tiles + tl) * len + x].x);
    float s = 0.f;
    for (int tl = 0; tl < tiles; ++tl) {
      const float2 p = part[((size_t)g * tiles + tl) * len + x];
      if (p.x > -INFINITY) s += p.y * exp2f(p.x - m);
    }
    lse[idx] = (m + log2f(s)) * 0.6931471805599453f;
    return;
  }
  MS v = ms_init();
  for (int tl = 0; tl < tiles; ++tl) {
    const float2 p = part[((size_t)g * tiles + tl) * len + x];
    v = ms_merge(v, MS{p.x, p.y});
  }
  lse[idx] = v.m + logf(v.s);
}

int score_lse(const ScoreArgs& a, float* row_lse, float* col_lse, float* scratch, cudaStream_t st, float* tcws,
              size_t tcws_bytes, int* used_tc) {
  const int IT = score_tiles_i(a.L), JT = score_tiles_j(a.S);
  float2* rowpart = reinterpret_cast<float2*>(scratch);
  const bool use_tc = tcws != nullptr && tcws_bytes >= tc_score_workspace_bytes(a.G, a.L, a.S, a.K) && tc_score_supported(a);
  if (used_tc) *used_tc = use_tc ? 1 : 0;
  const bool lse64 = use_tc && tc_lse64_supported(a);   // K = 64: query-tile-resident streaming kernel
  const int pj = lse64 ? 4 : (use_tc ? 2 * JT : JT), pi = lse64 ? 4 * IT : (use_tc ? 2 * IT : IT);  // partials per row / column
  float2* colpart = rowpart + (size_t)a.G * pj * a.L;
  if (lse64) {
    int rc = tc_lse64_partials(a, rowpart, colpart, tcws, tcws_bytes, 0, st);
    if (rc) return rc;
  } else if (use_tc) {
    int rc = tc_score_lse_partials(a, rowpart, colpart, tcws, tcws_bytes, 0, st);
    if (rc) return rc;
  } else {
    dim3 grid(JT, IT, a.G);
    if (score_vec_ok(a))
      score_lse_kernel<true><<<grid, kTileThreads, 0, st>>>(a, rowpart, colpart);
    else
      score_lse_kernel<false><<<grid, kTileThreads, 0, st>>>(a, rowpart, colpart);
    FAR_CHECK_LAUNCH();
  }
  const long long tr = (long long)a.G * a.L, tc = (long long)a.G * a.S;
  lse_finalize_kernel<<<(unsigned)ceil_div_ll(tr, 256), 256, 0, st>>>(rowpart, pj, a.L, tr, row_lse, use_tc ? 1 : 0);
  FAR_CHECK_LAUNCH();
  lse_finalize_kernel<<<(unsigned)ceil_div_ll(tc, 256), 256, 0, st>>>(colpart, pi, a.S, tc, col_lse, use_tc ? 1 : 0);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

}  // namespace far
