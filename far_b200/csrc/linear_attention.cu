// LinearAttention.forward (mp3d_loftr/src/loftr/loftr_module/linear_attention.py:20-52).
//   Q = elu(q)+1, K = elu(k)+1, v /= S
//   KV[n,h,d,e] = sum_s K[n,s,h,d] v[n,s,h,e];  Ksum[n,h,d] = sum_s K[n,s,h,d]
//   out[n,l,h,e] = (sum_d Q[n,l,h,d] KV[d,e]) * (1 / (sum_d Q[n,l,h,d] Ksum[d] + eps)) * S
// HBM-bound (algorithmic bytes = 4 tensors of N*L*C*4); two kernels for long sequences (split reduction over
// S with fixed-order merge => deterministic), one fused warp-per-(n,h) kernel for the 5x5 fine windows.
#include "common.cuh"

namespace far {

constexpr int LA_TOK = 64;  // tokens staged per smem tile

__device__ __forceinline__ float fmap(float x, int applied) { return applied ? x : (x > 0.f ? x + 1.f : expm1f(x) + 1.f); }

// grid (N*H, splits); block 256.  Partial layout: ws[((n*H+h)*splits + z) * (D*D + D)]
template <int D>
__global__ void __launch_bounds__(256) la_reduce_kernel(const float* __restrict__ k, int ldk, const float* __restrict__ v,
                                                        int ldv, int S, int H, int applied, int chunk,
                                                        float* __restrict__ ws) {
  constexpr int PER = (D * D) / 256 > 0 ? (D * D) / 256 : 1;  // outputs per thread along e
  constexpr int ACTIVE = (D * D) / PER;                       // threads holding KV entries
  __shared__ float Ks[LA_TOK][D];
  __shared__ float Vs[LA_TOK][D];
  const int z = blockIdx.y, nh = blockIdx.x, n = nh / H, h = nh % H, t = threadIdx.x;
  const int s_beg = z * chunk, s_end = min(S, s_beg + chunk);
  const float invS_den = (float)S;
  const int d = (t * PER) / D, e0 = (t * PER) % D;
  float acc[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) acc[i] = 0.f;
  float ksum = 0.f;
  const float* kb = k + (size_t)n * S * ldk + h * D;
  const float* vb = v + (size_t)n * S * ldv + h * D;
  for (int s0 = s_beg; s0 < s_end; s0 += LA_TOK) {
    const int cnt = min(LA_TOK, s_end - s0);
    for (int idx = t; idx < LA_TOK * D; idx += 256) {
      const int r = idx / D, c = idx % D;
      float kv = 0.f, vv = 0.f;
      if (r < cnt) {
        kv = fmap(kb[(size_t)(s0 + r) * ldk + c], applied);
        vv = vb[(size_t)(s0 + r) * ldv + c] / invS_den;  // values / v_length (:44)
      }
      Ks[r][c] = kv;
      Vs[r][c] = vv;
    }
    __syncthreads();
    if (t < ACTIVE) {
      for (int r = 0; r < cnt; ++r) {
        const float kd = Ks[r][d];
#pragma unroll
        for (int i = 0; i < PER; ++i) acc[i] = fmaf(kd, Vs[r][e0 + i], acc[i]);
        if (e0 == 0) ksum += kd;
      }
    }
    __syncthreads();
  }
  float* out = ws + ((size_t)nh * gridDim.y + z) * (D * D + D);
  if (t < ACTIVE) {
#pragma unroll
    for (int i = 0; i < PER; ++i) out[d * D + e0 + i] = acc[i];
    if (e0 == 0) out[D * D + d] = ksum;
  }
}

// grid (N*H, ceil(L/LA_TOK)); block 256
template <int D>
__global__ void __launch_bounds__(256) la_apply_kernel(const float* __restrict__ q, int ldq, float* __restrict__ out,
                                                       int ldo, int L, int S, int H, int applied, int splits,
                                                       float eps, const float* __restrict__ ws) {
  __shared__ float KV[D][D + 1];
  __shared__ float Ksum[D];
  __shared__ float Qs[LA_TOK][D];
  const int nh = blockIdx.x, n = nh / H, h = nh % H, t = threadIdx.x;
  const int l0 = blockIdx.y * LA_TOK, cnt = min(LA_TOK, L - l0);
  for (int idx = t; idx < D * D + D; idx += 256) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[((size_t)nh * splits + z) * (D * D + D) + idx];
    if (idx < D * D) KV[idx / D][idx % D] = s; else Ksum[idx - D * D] = s;
  }
  const float* qb = q + ((size_t)n * L + l0) * ldq + h * D;
  for (int idx = t; idx < LA_TOK * D; idx += 256) {
    const int r = idx / D, c = idx % D;
    Qs[r][c] = (r < cnt) ? fmap(qb[(size_t)r * ldq + c], applied) : 0.f;
  }
  __syncthreads();
  constexpr int TPP = 256 / D;  // tokens per pass
  const int e = t % D;
  for (int r = t / D; r < cnt; r += TPP) {
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      const float qd = Qs[r][dd];
      num = fmaf(qd, KV[dd][e], num);
      den = fmaf(qd, Ksum[dd], den);
    }
    const float zinv = 1.f / (den + eps);
    out[((size_t)n * L + l0 + r) * ldo + h * D + e] = num * zinv * (float)S;
  }
}

// Fused small-sequence variant (fine 5x5 windows): one warp per (n,h); L,S <= LA_SMALL.  block = 8 warps.
constexpr int LA_SMALL = 25;
template <int D>
__global__ void __launch_bounds__(256) la_small_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                       int ldk, const float* __restrict__ v, int ldv,
                                                       float* __restrict__ out, int ldo, long long NH, int L, int S,
                                                       int H, int applied, float eps) {
  __shared__ float sm[8][3][LA_SMALL][D];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nh = (long long)blockIdx.x * 8 + wib;
  if (nh >= NH) return;
  const long long n = nh / H;
  const int h = (int)(nh % H);
  float(*Qs)[D] = sm[wib][0];
  float(*Ks)[D] = sm[wib][1];
  float(*Vs)[D] = sm[wib][2];
  const float* qb = q + (size_t)n * L * ldq + h * D;
  const float* kb = k + (size_t)n * S * ldk + h * D;
  const float* vb = v + (size_t)n * S * ldv + h * D;
  for (int idx = lane; idx < L * D; idx += 32) Qs[idx / D][idx % D] = fmap(qb[(size_t)(idx / D) * ldq + idx % D], applied);
  for (int idx = lane; idx < S * D; idx += 32) {
    Ks[idx / D][idx % D] = fmap(kb[(size_t)(idx / D) * ldk + idx % D], applied);
    Vs[idx / D][idx % D] = vb[(size_t)(idx / D) * ldv + idx % D] / (float)S;
  }
  __syncwarp();
  // lane owns column e = lane % D and the d's congruent to (lane / D) modulo (32 / D)
  constexpr int G = 32 / D;      // d-groups per warp (D=16 -> 2, D=32 -> 1)
  constexpr int ND = D / G;      // d's per lane
  const int e = lane % D, g = lane / D;
  float kv[ND], ks[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) { kv[i] = 0.f; ks[i] = 0.f; }
  for (int s = 0; s < S; ++s) {
    const float ve = Vs[s][e];
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      const float kd = Ks[s][g + i * G];
      kv[i] = fmaf(kd, ve, kv[i]);
      ks[i] += kd;
    }
  }
  for (int l = 0; l < L; ++l) {
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      const float qd = Qs[l][g + i * G];
      num = fmaf(qd, kv[i], num);
      den = fmaf(qd, ks[i], den);
    }
#pragma unroll
    for (int o = D; o < 32; o <<= 1) {
      num += __shfl_xor_sync(0xffffffffu, num, o);
      den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if (g == 0) out[((size_t)n * L + l) * ldo + h * D + e] = num * (1.f / (den + eps)) * (float)S;
  }
}

static int la_splits(int N, int S, int H) {
  const long long base = (long long)N * H;
  int s = (int)((4LL * kNumSMs + base - 1) / base);
  const int maxs = ceil_div(S, 2 * LA_TOK);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return s;
}

int linear_attention_dispatch(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                              int ldo, int N, int L, int S, int H, int D, float eps, int applied, float* workspace,
                              size_t workspace_bytes, cudaStream_t st) {
  if (N <= 0 || L <= 0) return FAR_OK;
  FAR_REQUIRE(q && k && v && out && S > 0 && (D == 16 || D == 32));
  if (L <= LA_SMALL && S <= LA_SMALL && D == 16) {
    const long long NH = (long long)N * H;
    const unsigned blocks = (unsigned)ceil_div_ll(NH, 8);
    la_small_kernel<16><<<blocks, 256, 0, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, NH, L, S, H, applied, eps);
    FAR_CHECK_LAUNCH();
    return FAR_OK;
  }
  const int splits = la_splits(N, S, H);
  const size_t need = (size_t)N * H * splits * (D * D + D) * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need) return FAR_ERR_WORKSPACE;
  const int chunk = ceil_div(ceil_div(S, splits), LA_TOK) * LA_TOK;
  dim3 g1(N * H, splits), g2(N * H, ceil_div(L, LA_TOK));
  if (D == 16) {
    la_reduce_kernel<16><<<g1, 256, 0, st>>>(k, ldk, v, ldv, S, H, applied, chunk, workspace);
    FAR_CHECK_LAUNCH();
    la_apply_kernel<16><<<g2, 256, 0, st>>>(q, ldq, out, ldo, L, S, H, applied, splits, eps, workspace);
  } else {
    la_reduce_kernel<32><<<g1, 256, 0, st>>>(k, ldk, v, ldv, S, H, applied, chunk, workspace);
    FAR_CHECK_LAUNCH();
    la_apply_kernel<32><<<g2, 256, 0, st>>>(q, ldq, out, ldo, L, S, H, applied, splits, eps, workspace);
  }
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

size_t linear_attention_ws_bytes(int N, int S, int H, int D) {
  return (size_t)N * H * la_splits(N, S, H) * (D * D + D) * sizeof(float) + 256;
}

}  // namespace far

extern "C" size_t far_linear_attention_workspace_bytes(int N, int S, int H, int D) {
  return far::linear_attention_ws_bytes(N, S, H, D);
}

extern "C" int far_linear_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                    float* out, int ldo, int N, int L, int S, int H, int D, float eps,
                                    int feature_map_applied, float* workspace, size_t workspace_bytes, void* stream) {
  return far::linear_attention_dispatch(q, ldq, k, ldk, v, ldv, out, ldo, N, L, S, H, D, eps, feature_map_applied,
                                        workspace, workspace_bytes, (cudaStream_t)stream);
}
