// CoarseMatching.forward + get_coarse_match (mp3d_loftr/src/loftr/utils/coarse_matching.py:86-265),
// dual_softmax branch, eval, no padding masks.
//
//   pass A (score.cu)        : sim tiles -> row/col log-sum-exp                  (no N x L x S tensor in HBM)
//   pass B (match_conf_kernel): recompute sim tile, conf = exp(sim-rowlse)*exp(sim-collse), per-tile row
//                              (max, lowest argmax j) and column max partials; optional dense conf output
//   decide                   : per row (b,i): merge partials, threshold (>), border frame on both grids,
//                              mutual-nearest (row max == column max of the same element), block counts
//   offsets / gather         : order-preserving compaction (ascending (b,i), what torch.where returns)
//
// Tie semantics vs the reference (:186-192): the reference keeps, per row, the LOWEST j among all entries equal
// to both the row max and their column max; we take the lowest j attaining the row max and test it against its
// column max.  These differ only if two entries of one row are bit-identical maxima and the lower one loses its
// column test -- a measure-zero event for real-valued features (documented in DESIGN.md).
#include "score.cuh"
#include "tc_score.cuh"

namespace far {

struct MatchLayout {  // byte offsets inside the workspace; the "sel" block depends on N*L only
  size_t flag, jsel, csel, bcount, boff, nblocks;
  size_t rowlse, collse, rowmax, colmax, scratch, tcws, tcws_bytes, total;
};
constexpr int kDecideThreads = 256;
static inline size_t al(size_t v) { return (v + 255) & ~size_t(255); }

static MatchLayout match_layout(int N, int L, int S, int C = 256) {
  MatchLayout m;
  const size_t R = (size_t)N * L;
  size_t off = 0;
  m.nblocks = (R + kDecideThreads - 1) / kDecideThreads;
  m.flag = off; off += al(R * 4);
  m.jsel = off; off += al(R * 4);
  m.csel = off; off += al(R * 4);
  m.bcount = off; off += al(m.nblocks * 4);
  m.boff = off; off += al(m.nblocks * 8);
  if (S > 0) {
    const int IT = score_tiles_i(L), JT = score_tiles_j(S);
    m.rowlse = off; off += al(R * 4);
    m.collse = off; off += al((size_t)N * S * 4);
    m.rowmax = off; off += al((size_t)N * 2 * JT * L * 8);  // 2x: the tcgen05 kernel emits two partials per tile
    m.colmax = off; off += al((size_t)N * 2 * IT * S * 4);
    m.scratch = off; off += al(score_lse_scratch_floats(N, L, S) * 4);
    m.tcws_bytes = tc_score_workspace_bytes(N, L, S, C);
    m.tcws = off; off += al(m.tcws_bytes);
  }
  m.total = off;
  return m;
}

struct ConfArgs {
  ScoreArgs sc;
  const float* rowlse;
  const float* collse;
  float2* rowmax;  // [(g*JT+jt)*L + i] = (conf, bits(j))
  float* colmax;   // [(g*IT+it)*S + j]
  float* conf_out; // optional dense [G][L][S]
};

template <bool kVec4>
__global__ void __launch_bounds__(kTileThreads, 2) match_conf_kernel(ConfArgs p) {
  __shared__ TileSmem sm;
  __shared__ float colred[8][TBN];
  const ScoreArgs& a = p.sc;
  const int jt = blockIdx.x, it = blockIdx.y, g = blockIdx.z;
  const int JT = gridDim.x, IT = gridDim.y;
  const int i0 = it * TBM, j0 = jt * TBN;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4, warp = t >> 5, lane = t & 31;
  float acc[8][8];
  score_tile<kVec4>(a, g, i0, j0, sm, acc);

  float rl[8], cl[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i0 + tile_row(ty, i);
    rl[i] = (r < a.L) ? p.rowlse[(size_t)g * a.L + r] : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = j0 + tile_col(tx, j);
    cl[j] = (c < a.S) ? p.collse[(size_t)g * a.S + c] : 0.f;
  }
  // conf = softmax(sim, dim=1) * softmax(sim, dim=2)   (coarse_matching.py:118)
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = acc[i][j];
      acc[i][j] = expf(s - rl[i]) * expf(s - cl[j]);
    }
  if (p.conf_out) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i0 + tile_row(ty, i);
      if (r >= a.L) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = j0 + tile_col(tx, j);
        if (c < a.S) p.conf_out[((size_t)g * a.L + r) * a.S + c] = acc[i][j];
      }
    }
  }
  // ---- row (max, lowest argmax) over this tile's columns
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float bv = -1.f;
    int bj = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // ascending column order within the thread
      const int c = j0 + tile_col(tx, j);
      if (c < a.S && acc[i][j] > bv) { bv = acc[i][j]; bj = c; }
    }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
    }
    const int r = i0 + tile_row(ty, i);
    if (tx == 0 && r < a.L) p.rowmax[((size_t)g * JT + jt) * a.L + r] = make_float2(bv, __int_as_float(bj));
  }
  // ---- column max over this tile's rows
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float bv = -1.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i0 + tile_row(ty, i) < a.L) bv = fmaxf(bv, acc[i][j]);
    bv = fmaxf(bv, __shfl_xor_sync(0xffffffffu, bv, 16));
    if (lane < 16) colred[warp][tile_col(tx, j)] = bv;
  }
  __syncthreads();
  if (t < TBN) {
    float bv = -1.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) bv = fmaxf(bv, colred[w][t]);
    const int c = j0 + t;
    if (c < a.S) p.colmax[((size_t)g * IT + it) * a.S + c] = bv;
  }
}

// one thread per row (b,i)
__global__ void __launch_bounds__(kDecideThreads) match_decide_kernel(
    const float2* __restrict__ rowmax, const float* __restrict__ colmax, int N, int L, int S, int JT, int IT,
    float thr, int border, int h0, int w0, int h1, int w1, int* __restrict__ flag, int* __restrict__ jsel,
    float* __restrict__ csel, int* __restrict__ bcount) {
  const long long R = (long long)N * L;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int f = 0;
  if (idx < R) {
    const int b = (int)(idx / L), i = (int)(idx % L);
    float bv = -1.f;
    int bj = 0x7fffffff;
    for (int jt = 0; jt < JT; ++jt) {  // ascending tiles, strict > keeps the lowest j on ties
      const float2 p = rowmax[((size_t)b * JT + jt) * L + i];
      if (p.x > bv) { bv = p.x; bj = __float_as_int(p.y); }
    }
    if (bj >= 0 && bj < S) {
      float cm = -1.f;
      for (int it = 0; it < IT; ++it) cm = fmaxf(cm, colmax[((size_t)b * IT + it) * S + bj]);
      bool ok = (bv > thr) && (bv == cm) && (bv != 0.f);  // `mconf != 0` filter of :258-262
      if (border > 0) {  // mask_border (:8-25) on both coarse grids
        const int y0 = i / w0, x0 = i % w0, y1 = bj / w1, x1 = bj % w1;
        ok = ok && y0 >= border && y0 < h0 - border && x0 >= border && x0 < w0 - border && y1 >= border &&
             y1 < h1 - border && x1 >= border && x1 < w1 - border;
      }
      f = ok ? 1 : 0;
    }
    flag[idx] = f;
    jsel[idx] = bj;
    csel[idx] = bv;
  }
  const int cnt = __syncthreads_count(f);
  if (threadIdx.x == 0) bcount[blockIdx.x] = cnt;
}

// single block: exclusive scan of block counts -> offsets, total
__global__ void match_offsets_kernel(const int* __restrict__ bcount, int nblocks, long long* __restrict__ boff,
                                     long long* __restrict__ total) {
  __shared__ long long carry;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < nblocks) ? bcount[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = (lane < (int)(blockDim.x >> 5)) ? wsum[lane] : 0;
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += u;
      }
      wsum[lane] = winc - w;  // exclusive warp offsets
    }
    __syncthreads();
    const long long excl = carry + wsum[warp] + (inc - v);
    if (i < nblocks) boff[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kDecideThreads) match_gather_kernel(
    const int* __restrict__ flag, const int* __restrict__ jsel, const float* __restrict__ csel,
    const long long* __restrict__ boff, long long R, int L, int w0, int w1, float scale0, float scale1,
    long long* __restrict__ b_ids, long long* __restrict__ i_ids, long long* __restrict__ j_ids,
    float* __restrict__ mconf, float* __restrict__ mk0, float* __restrict__ mk1) {
  __shared__ int wsum[kDecideThreads / 32];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = (idx < R) ? flag[idx] : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  const int wpre = __popc(bal & ((1u << lane) - 1));
  if (lane == 0) wsum[warp] = __popc(bal);
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += wsum[w];
  if (f) {
    const long long o = boff[blockIdx.x] + woff + wpre;
    const long long b = idx / L, i = idx % L;
    const long long j = jsel[idx];
    b_ids[o] = b;
    i_ids[o] = i;
    j_ids[o] = j;
    mconf[o] = csel[idx];
    // (i % w, i // w) * scale   (:246-254)
    mk0[o * 2 + 0] = (float)(i % w0) * scale0;
    mk0[o * 2 + 1] = (float)(i / w0) * scale0;
    mk1[o * 2 + 0] = (float)(j % w1) * scale1;
    mk1[o * 2 + 1] = (float)(j / w1) * scale1;
  }
}

}  // namespace far

using namespace far;

extern "C" size_t far_dual_softmax_match_workspace_bytes(int N, int L, int S) {
  return match_layout(N, L, S, 256).total + 256;  // sized for C <= 256 (the path's coarse width)
}

extern "C" int far_dual_softmax_match_select(const float* feat0, const float* feat1, int N, int L, int S, int C,
                                             float temperature, float thr, int border_rm, int h0c, int w0c, int h1c,
                                             int w1c, float* conf_out, long long* num_matches, int engine,
                                             float* workspace, size_t workspace_bytes, void* stream) {
  FAR_REQUIRE(num_matches != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 0 || L <= 0 || S <= 0) {
    cudaMemsetAsync(num_matches, 0, sizeof(long long), st);
    return FAR_OK;
  }
  FAR_REQUIRE(feat0 && feat1 && workspace && C > 0 && h0c * w0c == L && h1c * w1c == S && temperature > 0.f);
  FAR_REQUIRE(C <= 256);
  const MatchLayout m = match_layout(N, L, S, 256);
  if (workspace_bytes < m.total) return FAR_ERR_WORKSPACE;
  char* base = reinterpret_cast<char*>(workspace);
  float* rowlse = reinterpret_cast<float*>(base + m.rowlse);
  float* collse = reinterpret_cast<float*>(base + m.collse);

  ScoreArgs a;
  a.A = feat0; a.sAb = (long long)L * C; a.sAh = 0; a.lda = C;
  a.B = feat1; a.sBb = (long long)S * C; a.sBh = 0; a.ldb = C;
  a.H = 1; a.G = N; a.L = L; a.S = S; a.K = C;
  // sim = (f0 / sqrt(C)) . (f1 / sqrt(C)) / T   (:108-113)
  a.scale = 1.0f / ((float)C * temperature);
  float* tcws = (engine == 1) ? nullptr : reinterpret_cast<float*>(base + m.tcws);
  int used_tc = 0;
  int rc = score_lse(a, rowlse, collse, reinterpret_cast<float*>(base + m.scratch), st, tcws, m.tcws_bytes, &used_tc);
  if (rc) return rc;

  ConfArgs p;
  p.sc = a;
  p.rowlse = rowlse;
  p.collse = collse;
  p.rowmax = reinterpret_cast<float2*>(base + m.rowmax);
  p.colmax = reinterpret_cast<float*>(base + m.colmax);
  p.conf_out = conf_out;
  const int IT = score_tiles_i(L), JT = score_tiles_j(S);
  if (used_tc) {  // operands are already split in tcws by the LSE pass
    rc = tc_match_conf(a, rowlse, collse, p.rowmax, p.colmax, conf_out, tcws, m.tcws_bytes, 1, st);
    if (rc) return rc;
  } else {
    dim3 grid(JT, IT, N);
    if (score_vec_ok(a))
      match_conf_kernel<true><<<grid, kTileThreads, 0, st>>>(p);
    else
      match_conf_kernel<false><<<grid, kTileThreads, 0, st>>>(p);
    FAR_CHECK_LAUNCH();
  }

  int* flag = reinterpret_cast<int*>(base + m.flag);
  int* jsel = reinterpret_cast<int*>(base + m.jsel);
  float* csel = reinterpret_cast<float*>(base + m.csel);
  int* bcount = reinterpret_cast<int*>(base + m.bcount);
  long long* boff = reinterpret_cast<long long*>(base + m.boff);
  const int pj = used_tc ? 2 * JT : JT, pi = used_tc ? 2 * IT : IT;
  match_decide_kernel<<<(unsigned)m.nblocks, kDecideThreads, 0, st>>>(p.rowmax, p.colmax, N, L, S, pj, pi, thr,
                                                                      border_rm, h0c, w0c, h1c, w1c, flag, jsel, csel,
                                                                      bcount);
  FAR_CHECK_LAUNCH();
  match_offsets_kernel<<<1, 1024, 0, st>>>(bcount, (int)m.nblocks, boff, num_matches);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" int far_dual_softmax_match_gather(int N, int L, int w0c, int w1c, float scale0, float scale1,
                                             long long num_matches, long long* b_ids, long long* i_ids,
                                             long long* j_ids, float* mconf, float* mkpts0_c, float* mkpts1_c,
                                             const float* workspace, size_t workspace_bytes, void* stream) {
  if (num_matches <= 0 || N <= 0 || L <= 0) return FAR_OK;
  FAR_REQUIRE(b_ids && i_ids && j_ids && mconf && mkpts0_c && mkpts1_c && workspace && w0c > 0 && w1c > 0);
  const MatchLayout m = match_layout(N, L, 0);
  if (workspace_bytes < m.total) return FAR_ERR_WORKSPACE;
  const char* base = reinterpret_cast<const char*>(workspace);
  match_gather_kernel<<<(unsigned)m.nblocks, kDecideThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int*>(base + m.flag), reinterpret_cast<const int*>(base + m.jsel),
      reinterpret_cast<const float*>(base + m.csel), reinterpret_cast<const long long*>(base + m.boff),
      (long long)N * L, L, w0c, w1c, scale0, scale1, b_ids, i_ids, j_ids, mconf, mkpts0_c, mkpts1_c);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}
