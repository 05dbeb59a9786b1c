// FinePreprocess.forward (mp3d_loftr/src/loftr/loftr_module/fine_preprocess.py:29-59) and
// FineMatching.forward / get_fine_match (mp3d_loftr/src/loftr/utils/fine_matching.py:15-76).
#include "common.cuh"

namespace far {
int linear_dispatch(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                    const float* bias, float* y, int ldy, int M, int N, int act, int act_cols, int engine,
                    float* workspace, size_t workspace_bytes, cudaStream_t st);
int linear_dispatch_rb(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                       const float* bias, const float* rowbias, int rowbias_group, float* y, int ldy, int M, int N,
                       int act, int act_cols, int engine, float* workspace, size_t workspace_bytes, cudaStream_t st,
                       const float* presplit = nullptr);

// One warp per (side, match, window position): copies the Cf-channel vector of one fine pixel (zero outside the
// map: F.unfold padding) into win[(side*M + m)*WW + ww][:].  Replaces F.unfold + index (:40-47): reads
// M*WW*Cf*4 bytes per side instead of materialising the 61 MB/img unfold tensor.
// window of coarse cell (r,c): fine rows stride*r - W/2 + ky, cols stride*c - W/2 + kx, ww = ky*W + kx.
__global__ void __launch_bounds__(256) fine_window_gather_kernel(
    const float* __restrict__ f0, const float* __restrict__ f1, long long sn, long long sc, long long sh, long long sw,
    int Hf, int Wf, int Cf, const long long* __restrict__ b_ids, const long long* __restrict__ i_ids,
    const long long* __restrict__ j_ids, long long M, int W, int stride, int w0c, int w1c, float* __restrict__ win) {
  const int WW = W * W;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = 2 * M * WW;
  if (gw >= total) return;
  const int ww = (int)(gw % WW);
  const long long sm_ = gw / WW;
  const int side = (int)(sm_ / M);
  const long long m = sm_ % M;
  const long long b = b_ids[m];
  const long long id = side ? j_ids[m] : i_ids[m];
  const int wc = side ? w1c : w0c;
  const int r = (int)(id / wc), c = (int)(id % wc);
  const int y = r * stride - W / 2 + ww / W, x = c * stride - W / 2 + ww % W;
  const float* f = (side ? f1 : f0) + b * sn;
  float* dst = win + gw * Cf;
  const bool inside = (y >= 0 && y < Hf && x >= 0 && x < Wf);
  if (sc == 1 && (Cf & 3) == 0) {
    const float4* src = reinterpret_cast<const float4*>(f + (size_t)y * sh + (size_t)x * sw);
    for (int q = lane; q < Cf / 4; q += 32)
      reinterpret_cast<float4*>(dst)[q] = inside ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int ch = lane; ch < Cf; ch += 32)
      dst[ch] = inside ? f[(size_t)ch * sc + (size_t)y * sh + (size_t)x * sw] : 0.f;
  }
}

// rows[(side*M + m)][:] = feat_c{side}[b, id, :]
__global__ void __launch_bounds__(256) coarse_row_gather_kernel(const float* __restrict__ c0, const float* __restrict__ c1,
                                                                int L0, int L1, int Cc,
                                                                const long long* __restrict__ b_ids,
                                                                const long long* __restrict__ i_ids,
                                                                const long long* __restrict__ j_ids, long long M,
                                                                float* __restrict__ rows) {
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= 2 * M) return;
  const int side = (int)(gw / M);
  const long long m = gw % M;
  const long long b = b_ids[m];
  const float* src = side ? c1 + ((size_t)b * L1 + j_ids[m]) * Cc : c0 + ((size_t)b * L0 + i_ids[m]) * Cc;
  for (int ch = lane; ch < Cc; ch += 32) rows[gw * Cc + ch] = src[ch];
}

// One warp per match (fine_matching.py:43-57,64-76).
__global__ void __launch_bounds__(256) fine_match_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                                         long long M, int WW, int C, const float* __restrict__ mk1c,
                                                         float offset_scale, float* __restrict__ expec,
                                                         float* __restrict__ mk1f) {
  const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const int W = (int)(sqrtf((float)WW) + 0.5f);
  const float* p = f0 + ((size_t)m * WW + WW / 2) * C;  // centre token (:43)
  const float* q = f1 + (size_t)m * WW * C;
  float mine = -INFINITY;  // lane r keeps sim[r]
  for (int r = 0; r < WW; ++r) {
    float d = 0.f;
    for (int ch = lane; ch < C; ch += 32) d = fmaf(p[ch], q[(size_t)r * C + ch], d);
    d = warp_sum(d);
    if (lane == r) mine = d;
  }
  const float temp = 1.0f / sqrtf((float)C);  // softmax_temp = 1 / C**.5 (:45)
  const float z = (lane < WW) ? mine * temp : -INFINITY;
  const float mx = warp_max(z);
  const float e = (lane < WW) ? expf(z - mx) : 0.f;
  const float heat = e / warp_sum(e);
  // create_meshgrid(W, W, normalized=True): x = linspace(-1,1,W)[r % W], y = linspace(-1,1,W)[r / W]
  const float step = (W > 1) ? 2.0f / (float)(W - 1) : 0.f;
  const float gx = (lane < WW) ? -1.f + step * (float)(lane % W) : 0.f;
  const float gy = (lane < WW) ? -1.f + step * (float)(lane / W) : 0.f;
  const float cx = warp_sum(heat * gx), cy = warp_sum(heat * gy);
  const float vx = warp_sum(gx * gx * heat) - cx * cx, vy = warp_sum(gy * gy * heat) - cy * cy;
  if (lane == 0) {
    const float sd = sqrtf(fmaxf(vx, 1e-10f)) + sqrtf(fmaxf(vy, 1e-10f));  // (:53-54)
    expec[m * 3 + 0] = cx;
    expec[m * 3 + 1] = cy;
    expec[m * 3 + 2] = sd;
    mk1f[m * 2 + 0] = mk1c[m * 2 + 0] + cx * offset_scale;  // (:71)
    mk1f[m * 2 + 1] = mk1c[m * 2 + 1] + cy * offset_scale;
  }
}

static inline size_t al(size_t v) { return (v + 255) & ~size_t(255); }
struct FinePlan { size_t win, crow, cproj, cterm, lin, total; };
static FinePlan fine_plan(long long M, int WW, int Cf, int Cc) {
  FinePlan p; size_t off = 0;
  p.win = off;   off += al((size_t)2 * M * WW * Cf * 4);
  p.crow = off;  off += al((size_t)2 * M * Cc * 4);
  p.cproj = off; off += al((size_t)2 * M * Cf * 4);
  p.cterm = off; off += al((size_t)2 * M * Cf * 4);
  p.lin = off;   off += al(2 * ((size_t)Cf * Cf * 4 + 1024) + 4096);  // tcgen05 engine: hi/lo split of merge_feat's window half
  p.total = off;
  return p;
}
}  // namespace far

using namespace far;

extern "C" size_t far_fine_preprocess_workspace_bytes(long long M, int WW, int Cf, int Cc) {
  return fine_plan(M, WW, Cf, Cc).total + 256;
}

extern "C" int far_fine_preprocess(const float* feat_f0, const float* feat_f1, long long sn, long long sc,
                                   long long sh, long long sw, int Hf, int Wf, int Cf, const float* feat_c0,
                                   const float* feat_c1, int L0, int L1, int Cc, const long long* b_ids,
                                   const long long* i_ids, const long long* j_ids, long long M, int W, int stride,
                                   int w0c, int w1c, const float* down_w, const float* down_b, const float* merge_w,
                                   const float* merge_b, float* out0, float* out1, float* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (M <= 0) return FAR_OK;  // M == 0 -> empty outputs (:34-37)
  FAR_REQUIRE(feat_f0 && feat_f1 && feat_c0 && feat_c1 && b_ids && i_ids && j_ids && down_w && down_b && merge_w &&
              merge_b && out0 && out1 && workspace && W > 0 && (W & 1) && Cf % 4 == 0 && Cc % 4 == 0);
  FAR_REQUIRE(2 * M * W * W < (1LL << 31) / 32);
  const int WW = W * W;
  const FinePlan p = fine_plan(M, WW, Cf, Cc);
  if (workspace_bytes < p.total) return FAR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* base = reinterpret_cast<char*>(workspace);
  float* win = reinterpret_cast<float*>(base + p.win);
  float* crow = reinterpret_cast<float*>(base + p.crow);
  float* cproj = reinterpret_cast<float*>(base + p.cproj);
  float* cterm = reinterpret_cast<float*>(base + p.cterm);

  const long long warps = 2 * M * WW;
  {
  ProfScope prof(PROF_FINE_GATHER, 0.0, 4.0 * 2.0 * 2.0 * M * WW * Cf, st);
  fine_window_gather_kernel<<<(unsigned)ceil_div_ll(warps, 8), 256, 0, st>>>(feat_f0, feat_f1, sn, sc, sh, sw, Hf, Wf,
                                                                            Cf, b_ids, i_ids, j_ids, M, W, stride, w0c,
                                                                            w1c, win);
  }
  FAR_CHECK_LAUNCH();
  coarse_row_gather_kernel<<<(unsigned)ceil_div_ll(2 * M, 8), 256, 0, st>>>(feat_c0, feat_c1, L0, L1, Cc, b_ids, i_ids,
                                                                           j_ids, M, crow);
  FAR_CHECK_LAUNCH();
  int rc;
  // feat_c_win = down_proj(cat[feat_c0[b,i], feat_c1[b,j]])  (:50-51)
  if ((rc = linear_dispatch(crow, Cc, Cc, nullptr, 0, 0, down_w, Cc, down_b, cproj, Cf, (int)(2 * M), Cf, FAR_ACT_NONE,
                            -1, 1, nullptr, 0, st))) return rc;
  // merge_feat(cat[window, repeat(feat_c_win)]) = window . Wm[:, :Cf]^T + (feat_c_win . Wm[:, Cf:]^T + bm)  (:52-55)
  if ((rc = linear_dispatch(cproj, Cf, Cf, nullptr, 0, 0, merge_w + Cf, 2 * Cf, merge_b, cterm, Cf, (int)(2 * M), Cf,
                            FAR_ACT_NONE, -1, 1, nullptr, 0, st))) return rc;
  const int rows = (int)(M * WW);
  // the two big GEMMs ([M*WW, Cf] x [Cf, Cf] + per-match row bias) go to the tcgen05 engine when it applies
  float* lin = reinterpret_cast<float*>(base + p.lin);
  const size_t lin_bytes = p.total - p.lin;
  if ((rc = linear_dispatch_rb(win, Cf, Cf, nullptr, 0, 0, merge_w, 2 * Cf, nullptr, cterm, WW, out0, Cf, rows, Cf,
                               FAR_ACT_NONE, -1, 0, lin, lin_bytes, st))) return rc;
  return linear_dispatch_rb(win + (size_t)rows * Cf, Cf, Cf, nullptr, 0, 0, merge_w, 2 * Cf, nullptr,
                            cterm + (size_t)M * Cf, WW, out1, Cf, rows, Cf, FAR_ACT_NONE, -1, 0, lin, lin_bytes, st);
}

extern "C" int far_fine_match(const float* feat_f0, const float* feat_f1, long long M, int WW, int C,
                              const float* mkpts1_c, float offset_scale, float* expec_f, float* mkpts1_f,
                              void* stream) {
  if (M <= 0) return FAR_OK;
  FAR_REQUIRE(feat_f0 && feat_f1 && mkpts1_c && expec_f && mkpts1_f && WW > 0 && WW <= 32 && C > 0);
  ProfScope prof(PROF_FINE_MATCH, 2.0 * M * WW * C, 4.0 * 2.0 * M * WW * C, (cudaStream_t)stream);
  fine_match_kernel<<<(unsigned)ceil_div_ll(M, 8), 256, 0, (cudaStream_t)stream>>>(feat_f0, feat_f1, M, WW, C, mkpts1_c,
                                                                                  offset_scale, expec_f, mkpts1_f);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}
