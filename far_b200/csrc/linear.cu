// nn.Linear family, LayerNorm(+residual), positional-encoding flatten, FAR blend epilogue.
#include "gemm_tile.cuh"
#include "tc_gemm.cuh"

namespace far {

struct LinearArgs {
  const float* x1; int ldx1; int K1;
  const float* x2; int ldx2; int K2;
  const float* W; int ldw;
  const float* bias;
  float* y; int ldy;
  int M, N;
  int act, act_cols;
  float* ws;      // split-K partials [splits][M][N] (splits > 1)
  int splits;
  int kchunk;     // K1 range per split (multiple of TBK)
  const float* rowbias;  // optional [ceil(M/rowbias_group)][N]: y[r] += rowbias[r / rowbias_group]
  int rowbias_group;
};

template <bool kVec4>
__global__ void __launch_bounds__(kTileThreads, 2) linear_simt_kernel(LinearArgs p) {
  __shared__ TileSmem sm;
  const int n0 = blockIdx.x * TBN, m0 = blockIdx.y * TBM, z = blockIdx.z;
  const int mValid = min(TBM, p.M - m0), nValid = min(TBN, p.N - n0);
  float acc[8][8];
  tile_zero(acc);

  int kbeg = 0, klen = p.K1;
  if (p.splits > 1) {
    kbeg = z * p.kchunk;
    klen = min(p.kchunk, p.K1 - kbeg);
    if (klen < 0) klen = 0;
  }
  simt_tile_mma<kVec4>(p.x1 + (size_t)m0 * p.ldx1 + kbeg, p.ldx1, mValid, p.W + (size_t)n0 * p.ldw + kbeg, p.ldw,
                       nValid, klen, sm, acc);
  if (p.x2 != nullptr && p.K2 > 0 && (p.splits <= 1 || z == p.splits - 1))  // K-concat tail rides with the last split
    simt_tile_mma<kVec4>(p.x2 + (size_t)m0 * p.ldx2, p.ldx2, mValid, p.W + (size_t)n0 * p.ldw + p.K1, p.ldw,
                         nValid, p.K2, sm, acc);

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (p.splits > 1) {
    float* out = p.ws + (size_t)z * p.M * p.N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = m0 + tile_row(ty, i);
      if (r >= p.M) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = n0 + tile_col(tx, j);
        if (c < p.N) out[(size_t)r * p.N + c] = acc[i][j];
      }
    }
    return;
  }
  const int actc = p.act_cols < 0 ? p.N : p.act_cols;
  // Tile-uniform fast path: whole tile inside N, one activation for the whole tile, vector stores legal.  Keeps the
  // epilogue free of per-element control flow (which dominated the tcgen05 epilogue until it was removed there).
  const int act_tile = (n0 + TBN <= actc) ? p.act : ((n0 >= actc) ? FAR_ACT_NONE : -1);
  if (n0 + TBN <= p.N && act_tile >= 0 && ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15u) == 0)) {
    if (p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b = __ldg(p.bias + n0 + tile_col(tx, j));
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] += b;
      }
    }
    if (p.rowbias != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = min(m0 + tile_row(ty, i), p.M - 1);
        const float* rb = p.rowbias + (size_t)(r / p.rowbias_group) * p.N + n0;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += __ldg(rb + tile_col(tx, j));
      }
    }
    switch (act_tile) {
      case FAR_ACT_RELU:
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaxf(acc[i][j], 0.f);
        break;
      case FAR_ACT_NONE:
        break;
      default:
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = apply_act(acc[i][j], act_tile);
        break;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = m0 + tile_row(ty, i);
      if (r < p.M) {
        float* dst = p.y + (size_t)r * p.ldy + n0;
        *reinterpret_cast<float4*>(dst + tile_col(tx, 0)) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(dst + tile_col(tx, 4)) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + tile_row(ty, i);
    if (r >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c0 = n0 + tile_col(tx, jh * 4);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        float t = acc[i][jh * 4 + j];
        if (c < p.N) {
          if (p.bias) t += __ldg(p.bias + c);
          if (p.rowbias) t += __ldg(p.rowbias + (size_t)(r / p.rowbias_group) * p.N + c);
          if (c < actc) t = apply_act(t, p.act);
        }
        v[j] = t;
      }
      float* dst = p.y + (size_t)r * p.ldy + c0;
      if (c0 + 3 < p.N && ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15u) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + j < p.N) dst[j] = v[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- skinny M (<= 32) split-K
// The FAR gate / encoder MLPs are [32 pairs] x [35840 (+22)] x [512]: 73 MB of weights per call, 12 us of HBM time.  The
// 128 x 128 tile kernel above computes 128 rows for 32 (FMA-bound on padding: ~300 us per call); here the tile is
// 32 rows x 128 columns x 32 k, thread = 4 rows x 4 columns, operands transposed into shared memory with an XOR swizzle of
// the 4-column groups (conflict-free scalar stores and float4 loads), register double buffering of the next k-slab.
constexpr int SK_BM = 32, SK_BN = 128, SK_BK = 32;
struct __align__(16) SkinnySmem {
  float a[2][SK_BK][SK_BM];
  float b[2][SK_BK][SK_BN];
};  // 40 KiB

template <bool kVec4>
__device__ __forceinline__ void skinny_tile_mma(const float* __restrict__ A, int lda, int mValid,
                                                const float* __restrict__ B, int ldb, int nValid, int K,
                                                SkinnySmem& sm, float (&acc)[4][4]) {
  const int t = threadIdx.x;
  const int k4 = t & 7, lr = t >> 3;          // loader: float4 column k4 of rows lr (+32 i)
  const int tm = t >> 5, tn = t & 31;         // compute: rows 4 tm .. +3, columns 4 tn .. +3
  const int nk = (K + SK_BK - 1) / SK_BK;
  if (nk == 0) return;
  float4 ra, rb[4];
  auto fetch = [&](int kt) {
    const int k = kt * SK_BK + k4 * 4;
    ra = tile_ld4<kVec4>(A, lda, lr, mValid, k, K);
#pragma unroll
    for (int i = 0; i < 4; ++i) rb[i] = tile_ld4<kVec4>(B, ldb, lr + 32 * i, nValid, k, K);
  };
  auto stash = [&](int buf) {
    // element (row r, k) lives at [k][4 * ((r >> 2) ^ (k >> 2 & 7)) + (r & 3)]
    const int ga = ((lr >> 2) ^ k4) * 4 + (lr & 3);
    sm.a[buf][k4 * 4 + 0][ga] = ra.x; sm.a[buf][k4 * 4 + 1][ga] = ra.y;
    sm.a[buf][k4 * 4 + 2][ga] = ra.z; sm.a[buf][k4 * 4 + 3][ga] = ra.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = lr + 32 * i;
      const int gb = ((r >> 2) ^ k4) * 4 + (r & 3);
      sm.b[buf][k4 * 4 + 0][gb] = rb[i].x; sm.b[buf][k4 * 4 + 1][gb] = rb[i].y;
      sm.b[buf][k4 * 4 + 2][gb] = rb[i].z; sm.b[buf][k4 * 4 + 3][gb] = rb[i].w;
    }
  };
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    const bool more = kt + 1 < nk;
    if (more) fetch(kt + 1);
#pragma unroll
    for (int k = 0; k < SK_BK; ++k) {
      const int sw = (k >> 2) & 7;
      const float4 a4 = *reinterpret_cast<const float4*>(&sm.a[cur][k][(tm ^ sw) * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sm.b[cur][k][(tn ^ sw) * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) stash(cur ^ 1);
    __syncthreads();
  }
}

// grid (ceil(N / 128), 1, splits): partial products of split z into ws[z][M][N]; splitk_reduce_kernel finishes.
template <bool kVec4>
__global__ void __launch_bounds__(256, 3) linear_skinny_kernel(LinearArgs p) {
  __shared__ SkinnySmem sm;
  const int n0 = blockIdx.x * SK_BN, z = blockIdx.z;
  const int nValid = min(SK_BN, p.N - n0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int kbeg = z * p.kchunk;
  int klen = min(p.kchunk, p.K1 - kbeg);
  if (klen < 0) klen = 0;
  skinny_tile_mma<kVec4>(p.x1 + kbeg, p.ldx1, p.M, p.W + (size_t)n0 * p.ldw + kbeg, p.ldw, nValid, klen, sm, acc);
  if (p.x2 != nullptr && p.K2 > 0 && z == p.splits - 1)   // K-concat tail rides with the last split
    skinny_tile_mma<kVec4>(p.x2, p.ldx2, p.M, p.W + (size_t)n0 * p.ldw + p.K1, p.ldw, nValid, p.K2, sm, acc);
  const int tm = threadIdx.x >> 5, tn = threadIdx.x & 31;
  float* out = p.ws + (size_t)z * p.M * p.N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tm * 4 + i;
    if (r >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tn * 4 + j;
      if (c < p.N) out[(size_t)r * p.N + c] = acc[i][j];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, const float* __restrict__ bias,
                                     float* __restrict__ y, int ldy, int M, int N, int act, int act_cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * N) return;
  const int r = (int)(idx / N), c = (int)(idx % N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[(size_t)z * M * N + idx];  // fixed order: deterministic
  if (bias) s += bias[c];
  const int actc = act_cols < 0 ? N : act_cols;
  if (c < actc) s = apply_act(s, act);
  y[(size_t)r * ldy + c] = s;
}

static void choose_splits(int M, int N, int K, int* splits, int* kchunk) {
  const int tiles = ceil_div(M, TBM) * ceil_div(N, TBN);
  int s = 1;
  if (tiles < kNumSMs && K >= 2048) {
    s = (2 * kNumSMs) / tiles;
    const int maxs = K / 256;  // at least 256 of K per split
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    if (s > 128) s = 128;
  }
  int kc = ceil_div(ceil_div(K, s), TBK) * TBK;
  s = ceil_div(K, kc);
  *splits = s;
  *kchunk = kc;
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// one warp per row, row cached in registers (C <= 1024), two-pass mean/variance like ATen's CPU kernel.
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ pre, int pre_rows,
                                 const float* __restrict__ g, const float* __restrict__ b,
                                 const float* __restrict__ res, float* __restrict__ y, int rows, int C, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * C;
  const float* pr = pre ? pre + (size_t)(warp % pre_rows) * C : nullptr;  // broadcast table (e.g. pos_embed)
  float v[32];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + i * 32;
    v[i] = (c < C) ? xr[c] + (pr ? pr[c] : 0.f) : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + i * 32;
    const float d = (c < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + i * 32;
    if (c < C) {
      float o = (v[i] - mean) * rstd * g[c] + b[c];
      if (res) o += res[(size_t)warp * C + c];
      y[(size_t)warp * C + c] = o;
    }
  }
}

// Vectorised variant for C = 128 * NV (NV <= 4): float4 loads, NV*4 values per lane, ~40 registers => full occupancy
// so enough 512-byte row segments are in flight to cover HBM latency.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const float* __restrict__ x, const float* __restrict__ pre,
                                                            int pre_rows, const float* __restrict__ g,
                                                            const float* __restrict__ b, const float* __restrict__ res,
                                                            float* __restrict__ y, int rows, float eps) {
  constexpr int C = 128 * NV;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * C);
  const float4* pr = pre ? reinterpret_cast<const float4*>(pre + (size_t)(warp % pre_rows) * C) : nullptr;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = __ldg(xr + lane + i * 32);
    if (pr) {
      const float4 q = __ldg(pr + lane + i * 32);
      v[i].x += q.x; v[i].y += q.y; v[i].z += q.z; v[i].w += q.w;
    }
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  const float4* r4 = res ? reinterpret_cast<const float4*>(res + (size_t)warp * C) : nullptr;
  float4* y4 = reinterpret_cast<float4*>(y + (size_t)warp * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 gg = __ldg(g4 + lane + i * 32), bb = __ldg(b4 + lane + i * 32);
    float4 o;
    o.x = (v[i].x - mean) * rstd * gg.x + bb.x;
    o.y = (v[i].y - mean) * rstd * gg.y + bb.y;
    o.z = (v[i].z - mean) * rstd * gg.z + bb.z;
    o.w = (v[i].w - mean) * rstd * gg.w + bb.w;
    if (r4) {
      const float4 rr = __ldg(r4 + lane + i * 32);
      o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
    }
    y4[lane + i * 32] = o;
  }
}

static bool ln_vec_ok(const void* x, const void* pre, const void* g, const void* b, const void* res, const void* y, int C) {
  auto a16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return (C == 128 || C == 256 || C == 384 || C == 512) && a16(x) && a16(pre) && a16(g) && a16(b) && a16(res) && a16(y);
}

static void launch_layernorm(const float* x, const float* pre, int pre_rows, const float* g, const float* b,
                             const float* res, float* y, int rows, int C, float eps, cudaStream_t st) {
  const int wpb = 8;
  const int grid = ceil_div(rows, wpb);
  ProfScope prof(PROF_LAYERNORM, 8.0 * rows * C, 4.0 * rows * C * (res ? 3.0 : 2.0), st);
  if (ln_vec_ok(x, pre, g, b, res, y, C)) {
    switch (C / 128) {
      case 1: layernorm_vec_kernel<1><<<grid, wpb * 32, 0, st>>>(x, pre, pre_rows, g, b, res, y, rows, eps); return;
      case 2: layernorm_vec_kernel<2><<<grid, wpb * 32, 0, st>>>(x, pre, pre_rows, g, b, res, y, rows, eps); return;
      case 3: layernorm_vec_kernel<3><<<grid, wpb * 32, 0, st>>>(x, pre, pre_rows, g, b, res, y, rows, eps); return;
      default: layernorm_vec_kernel<4><<<grid, wpb * 32, 0, st>>>(x, pre, pre_rows, g, b, res, y, rows, eps); return;
    }
  }
  layernorm_kernel<<<grid, wpb * 32, 0, st>>>(x, pre, pre_rows, g, b, res, y, rows, C, eps);
}

// ---------------------------------------------------------------------------------------------- pos-enc + flatten
// out[n, hw, c] = feat[n,c,h,w] + pe[hw, c].  32x32 smem transpose when the input is NCHW (sw == 1);
// straight coalesced copy when it is channels_last (sc == 1).
__global__ void pos_flatten_kernel(const float* __restrict__ feat, long long sn, long long sc, long long sh,
                                   long long sw, const float* __restrict__ pe, float* __restrict__ out, int C,
                                   int H, int W) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32, HW = H * W;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const float* f = feat + (size_t)n * sn;
  if (sc == 1) {
    for (int r = ty; r < 32; r += 8) {
      const int hw = hw0 + r, c = c0 + tx;
      if (hw < HW && c < C) {
        const int h = hw / W, w = hw % W;
        out[((size_t)n * HW + hw) * C + c] = f[(size_t)h * sh + (size_t)w * sw + c] + pe[(size_t)hw * C + c];
      }
    }
    return;
  }
  for (int r = ty; r < 32; r += 8) {  // r indexes channel, tx indexes hw (contiguous when sw == 1)
    const int c = c0 + r, hw = hw0 + tx;
    float v = 0.f;
    if (c < C && hw < HW) {
      const int h = hw / W, w = hw % W;
      v = f[(size_t)c * sc + (size_t)h * sh + (size_t)w * sw];
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {  // r indexes hw, tx indexes channel
    const int hw = hw0 + r, c = c0 + tx;
    if (hw < HW && c < C) out[((size_t)n * HW + hw) * C + c] = tile[tx][r] + pe[(size_t)hw * C + c];
  }
}

// ---------------------------------------------------------------------------------------------- FAR blend
__global__ void pose_blend_mp3d_kernel(const float* __restrict__ pred, const float* __restrict__ solver, int lds,
                                       const float* __restrict__ wt, const float* __restrict__ mean,
                                       const float* __restrict__ stdv, int scale_8pt, float* __restrict__ out,
                                       int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* p = pred + (size_t)b * 9;
  const float* s = solver + (size_t)b * lds;
  float ts[3] = {s[0], s[1], s[2]};
  if (scale_8pt) {  // transformer.py:436-446
    float lu[3], ru[3], nl = 0.f, nr = 0.f;
    for (int i = 0; i < 3; ++i) {
      lu[i] = s[i] * stdv[i] + mean[i];
      ru[i] = p[i] * stdv[i] + mean[i];
      nl += lu[i] * lu[i];
      nr += ru[i] * ru[i];
    }
    nl = fminf(fmaxf(sqrtf(nl), 1e-3f), 100.f);
    nr = sqrtf(nr);
    for (int i = 0; i < 3; ++i) ts[i] = (lu[i] * nr / nl - mean[i]) / stdv[i];
  }
  const float w0 = wt[b * 2 + 0], w1 = wt[b * 2 + 1];
  float* o = out + (size_t)b * 9;
  for (int i = 0; i < 3; ++i) o[i] = w0 * p[i] + (1.f - w0) * ts[i];
  for (int i = 3; i < 9; ++i) o[i] = w1 * p[i] + (1.f - w1) * s[i];
}

// Host-side dispatcher shared with the layer composition in encoder_layer.cu.
int linear_dispatch_rb(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                       const float* bias, const float* rowbias, int rowbias_group, float* y, int ldy, int M, int N,
                       int act, int act_cols, int engine, float* workspace, size_t workspace_bytes, cudaStream_t st,
                       const float* presplit = nullptr);

// `presplit`: optional far_tc_weight_split output for W (used when the tcgen05 engine is chosen)
int linear_dispatch_ps(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                       const float* presplit, float* y, int ldy, int M, int N, int act, int act_cols, int engine,
                       float* workspace, size_t workspace_bytes, cudaStream_t st) {
  return linear_dispatch_rb(x1, ldx1, K1, x2, ldx2, K2, W, ldw, nullptr, nullptr, 1, y, ldy, M, N, act, act_cols, engine,
                            workspace, workspace_bytes, st, presplit);
}

int linear_dispatch(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                    const float* bias, float* y, int ldy, int M, int N, int act, int act_cols, int engine,
                    float* workspace, size_t workspace_bytes, cudaStream_t st) {
  return linear_dispatch_rb(x1, ldx1, K1, x2, ldx2, K2, W, ldw, bias, nullptr, 1, y, ldy, M, N, act, act_cols, engine,
                            workspace, workspace_bytes, st);
}

// `rowbias` (fine_preprocess: the per-match coarse term broadcast over the 25 window positions) is only
// implemented by the CUDA-core engine.
int linear_dispatch_rb(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                       const float* bias, const float* rowbias, int rowbias_group, float* y, int ldy, int M, int N,
                       int act, int act_cols, int engine, float* workspace, size_t workspace_bytes, cudaStream_t st,
                       const float* presplit) {
  if (M <= 0 || N <= 0) return FAR_OK;
  FAR_REQUIRE(x1 && W && y && K1 > 0 && ldx1 >= K1 && ldw >= K1 + K2 && ldy >= N);
  FAR_REQUIRE((x2 == nullptr) == (K2 == 0));

  // tcgen05 3xTF32 engine: TMA needs contiguous-K rows with 16-byte-multiple strides.
  const bool tc_ok = tc_linear_supported(x1, ldx1, K1, x2, ldx2, K2, W, ldw, M, N) && (ldy % 4 == 0) &&
                     (reinterpret_cast<uintptr_t>(y) & 15u) == 0;  // TMA-store epilogue: 16-byte aligned output rows
  const bool rb_ok = rowbias == nullptr || (N % 4 == 0 && (reinterpret_cast<uintptr_t>(rowbias) & 15u) == 0);
  if (engine == 2 && (!tc_ok || !rb_ok)) return FAR_ERR_ARG;
  const bool tc_ws_ok = workspace != nullptr &&
                        workspace_bytes >= tc_linear_workspace_need(x1, ldx1, K1, x2, ldx2, K2, M, N);
  if (rb_ok && (engine == 2 || (engine == 0 && tc_ok && tc_ws_ok && tc_linear_preferred(M, N, K1 + K2) &&
                                tc_engine_default_on()))) {
    return tc_linear(x1, ldx1, K1, x2, ldx2, K2, W, ldw, bias, rowbias, rowbias_group, y, ldy, M, N, act, act_cols,
                     workspace, workspace_bytes, st, presplit);
  }

  LinearArgs p{x1, ldx1, K1, x2, ldx2, K2, W, ldw, bias, y, ldy, M, N, act, act_cols, nullptr, 1, K1, rowbias,
               rowbias_group > 0 ? rowbias_group : 1};
  if (rowbias == nullptr) choose_splits(M, N, K1, &p.splits, &p.kchunk);
  if (p.splits > 1) {
    const size_t need = (size_t)p.splits * M * N * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) {  // fall back to no split rather than fail
      p.splits = 1;
      p.kchunk = K1;
    } else {
      p.ws = workspace;
    }
  }
  const bool vec = ptr_aligned16(x1) && ptr_aligned16(W) && (ldx1 % 4 == 0) && (ldw % 4 == 0) && (K1 % 4 == 0) &&
                   (x2 == nullptr || (ptr_aligned16(x2) && ldx2 % 4 == 0 && K2 % 4 == 0));
  // skinny problems (M <= 32 rows, long K: the 35840-wide FAR MLPs): 32-row tiles, more splits (FAR_SKINNY=0 disables)
  static const bool skinny_on = !(getenv("FAR_SKINNY") && getenv("FAR_SKINNY")[0] == '0');
  if (skinny_on && M <= SK_BM && K1 >= 2048 && rowbias == nullptr && workspace != nullptr) {
    const int ntiles = ceil_div(N, SK_BN);
    int s2 = (3 * kNumSMs) / ntiles;
    if (s2 > K1 / 256) s2 = K1 / 256;
    if (s2 > 128) s2 = 128;
    if (s2 < 2) s2 = 2;
    const int kc2 = ceil_div(ceil_div(K1, s2), SK_BK) * SK_BK;
    s2 = ceil_div(K1, kc2);
    if (s2 >= 2 && workspace_bytes >= (size_t)s2 * M * N * sizeof(float)) {
      p.splits = s2; p.kchunk = kc2; p.ws = workspace;
      dim3 g2(ntiles, 1, s2);
      ProfScope prof(PROF_LINEAR_SIMT, 2.0 * M * N * (K1 + K2), 4.0 * ((double)M * (K1 + K2) + (double)N * (K1 + K2) + (double)M * N), st);
      if (vec) linear_skinny_kernel<true><<<g2, 256, 0, st>>>(p);
      else linear_skinny_kernel<false><<<g2, 256, 0, st>>>(p);
      FAR_CHECK_LAUNCH();
      const long long tot = (long long)M * N;
      splitk_reduce_kernel<<<(unsigned)ceil_div_ll(tot, 256), 256, 0, st>>>(p.ws, p.splits, bias, y, ldy, M, N, act, act_cols);
      FAR_CHECK_LAUNCH();
      return FAR_OK;
    }
  }
  dim3 grid(ceil_div(N, TBN), ceil_div(M, TBM), p.splits);
  ProfScope prof(PROF_LINEAR_SIMT, 2.0 * M * N * (K1 + K2), 4.0 * ((double)M * (K1 + K2) + (double)N * (K1 + K2) + (double)M * N), st);
  if (vec)
    linear_simt_kernel<true><<<grid, kTileThreads, 0, st>>>(p);
  else
    linear_simt_kernel<false><<<grid, kTileThreads, 0, st>>>(p);
  FAR_CHECK_LAUNCH();
  if (p.splits > 1) {
    const long long tot = (long long)M * N;
    splitk_reduce_kernel<<<(unsigned)ceil_div_ll(tot, 256), 256, 0, st>>>(p.ws, p.splits, bias, y, ldy, M, N, act,
                                                                         act_cols);
    FAR_CHECK_LAUNCH();
  }
  return FAR_OK;
}

}  // namespace far

using namespace far;

namespace far { unsigned long long g_launch_count = 0; }

// ------------------------------------------------------------------------------------------ far_profile_*
#include <vector>
namespace far {
bool g_prof_on = false;
namespace {
struct ProfRec { int id; double flops, bytes; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_prof_recs;
const char* kProfNames[PROF_NUM_IDS] = {
    "tc_gemm_kernel", "tc_score_kernel", "tc_emm_pv_kernel", "la_reduce", "la_apply", "la_small_kernel", "layernorm",
    "linear_simt_kernel", "fine_window_gather_kernel", "fine_match_kernel", "split_kernels", "emm_simt", "solver",
    "fpn_fuse", "enc_fused_kernel", "tc_corrvol_kernel", "eightpt_kernels", "tc_flash_attn_kernel"};
void prof_clear() {
  for (auto& r : g_prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof_recs.clear();
}
}  // namespace
void prof_begin(int id, double flops, double bytes, cudaStream_t st) {
  ProfRec r{id, flops, bytes, nullptr, nullptr};
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  g_prof_recs.push_back(r);
}
void prof_end(cudaStream_t st) { cudaEventRecord(g_prof_recs.back().e1, st); }
}  // namespace far

extern "C" int far_profile_num_ids(void) { return far::PROF_NUM_IDS; }
extern "C" const char* far_profile_name(int id) { return (id >= 0 && id < far::PROF_NUM_IDS) ? far::kProfNames[id] : ""; }
extern "C" int far_profile_enable(int on) {
  far::prof_clear();
  far::g_prof_on = on != 0;
  return FAR_OK;
}
extern "C" int far_profile_read(int id, double* total_ms, unsigned long long* launches, double* flops, double* bytes) {
  if (id < 0 || id >= far::PROF_NUM_IDS || !total_ms || !launches || !flops || !bytes) return FAR_ERR_ARG;
  double ms = 0, fl = 0, by = 0;
  unsigned long long n = 0;
  for (auto& r : far::g_prof_recs) {
    if (r.id != id) continue;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return FAR_ERR_CUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) return FAR_ERR_CUDA;
    ms += t; fl += r.flops; by += r.bytes; ++n;
  }
  *total_ms = ms; *launches = n; *flops = fl; *bytes = by;
  return FAR_OK;
}

extern "C" int far_abi_version(void) { return 2; }
extern "C" unsigned long long far_launch_count(void) { return far::g_launch_count; }

extern "C" size_t far_linear_workspace_bytes(int M, int N, int K) {
  int s, kc;
  choose_splits(M, N, K, &s, &kc);
  size_t b = (s > 1) ? (size_t)s * M * N * sizeof(float) : 0;
  if (M <= SK_BM && K >= 2048) {   // skinny split-K path: up to 128 splits
    const size_t b2 = (size_t)128 * M * N * sizeof(float);
    if (b2 > b) b = b2;
  }
  size_t tcb = tc_linear_workspace_bytes(M, N, K);
  return (b > tcb ? b : tcb) + 256;
}

extern "C" int far_linear(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W,
                          int ldw, const float* bias, float* y, int ldy, int M, int N, int act, int act_cols,
                          int engine, float* workspace, size_t workspace_bytes, void* stream) {
  return linear_dispatch(x1, ldx1, K1, x2, ldx2, K2, W, ldw, bias, y, ldy, M, N, act, act_cols, engine, workspace,
                         workspace_bytes, (cudaStream_t)stream);
}

extern "C" int far_layernorm_pre(const float* x, const float* pre_add, int pre_rows, const float* gamma,
                                 const float* beta, const float* residual, float* y, int rows, int C, float eps,
                                 void* stream) {
  if (rows <= 0) return FAR_OK;
  FAR_REQUIRE(x && gamma && beta && y && C > 0 && C <= 1024 && (pre_add == nullptr || pre_rows > 0));
  launch_layernorm(x, pre_add, pre_rows, gamma, beta, residual, y, rows, C, eps, (cudaStream_t)stream);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" int far_layernorm(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                             int rows, int C, float eps, void* stream) {
  if (rows <= 0) return FAR_OK;
  FAR_REQUIRE(x && gamma && beta && y && C > 0 && C <= 1024);
  launch_layernorm(x, nullptr, 1, gamma, beta, residual, y, rows, C, eps, (cudaStream_t)stream);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" int far_pos_encode_flatten(const float* feat, long long sn, long long sc, long long sh, long long sw,
                                      const float* pe_hwc, float* out, int N, int C, int H, int W, void* stream) {
  if (N <= 0) return FAR_OK;
  FAR_REQUIRE(feat && pe_hwc && out && C > 0 && H > 0 && W > 0);
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N), block(32, 8);
  pos_flatten_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(feat, sn, sc, sh, sw, pe_hwc, out, C, H, W);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" int far_pose_blend_mp3d(const float* pred, const float* solver, int ld_solver, const float* wt,
                                   const float* mean9, const float* std9, int scale_8pt, float* out, int B,
                                   void* stream) {
  if (B <= 0) return FAR_OK;
  FAR_REQUIRE(pred && solver && wt && mean9 && std9 && out && ld_solver >= 9);
  pose_blend_mp3d_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(pred, solver, ld_solver, wt, mean9, std9,
                                                                            scale_8pt, out, B);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}
