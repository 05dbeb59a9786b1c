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

// Fine-level variant with coalesced staging: one CTA per window n covers ALL 8 heads (C = 128 floats = one 512-byte row
// per token), q/k/v rows are staged with float4 loads by the whole CTA, warp h computes head h, and the result goes back
// through shared memory as full rows.  (la_small_kernel reads 64-byte head slices with scalar loads: 1.1 ms per call at
// 41.8k windows, ~1/3 of the HBM roofline.)
// The kernel is instruction-issue bound, not HBM bound (2.1 GB in 780 us = 41 % of the HBM peak with ~13 k warp
// instructions per window), so this version removes instructions: elu(x)+1 as x+1 | ex2.approx (the GEMM epilogue's form)
// instead of expm1f; Ksum accumulated by the loading threads (a thread always loads the same 4 columns) instead of 8
// redundant adds per lane per key; the normaliser Z = S / (Q.Ksum + eps) once per (token, head) instead of 8 FMAs per
// lane per token; rows padded to 132 floats so lane = token accesses spread over the banks; the output overwrites the K
// tile (dead after the KV pass), so the per-token __syncwarp is gone.
constexpr int LA_SP = 132;   // padded row pitch (floats)
__device__ __forceinline__ float fmap_fast(float x, int applied) {
  return applied ? x : (x > 0.f ? x + 1.f : exp2f(x * 1.4426950408889634f));
}
template <int D, int H>
__global__ void __launch_bounds__(32 * H) la_small_allheads_kernel(const float* __restrict__ q, int ldq,
                                                                   const float* __restrict__ k, int ldk,
                                                                   const float* __restrict__ v, int ldv,
                                                                   float* __restrict__ out, int ldo, int L, int S,
                                                                   int applied, float eps) {
  constexpr int C = D * H, NT = 32 * H, C4 = C / 4;
  static_assert(NT % C4 == 0, "a thread must always load the same column group");
  __shared__ __align__(16) float sm[3][LA_SMALL][LA_SP];
  __shared__ __align__(16) float ksp[H][C];   // per-warp partial sums of the mapped keys
  __shared__ __align__(16) float ksum_s[C];
  const long long n = blockIdx.x;
  const int t = threadIdx.x, h = t >> 5, lane = t & 31;
  const float invS = 1.f / (float)S;
  const int c4 = t % C4, rstep = NT / C4, r0 = t / C4;
  for (int r = r0; r < L; r += rstep) {
    float4 a = __ldg(reinterpret_cast<const float4*>(q + ((size_t)n * L + r) * ldq) + c4);
    a.x = fmap_fast(a.x, applied); a.y = fmap_fast(a.y, applied); a.z = fmap_fast(a.z, applied); a.w = fmap_fast(a.w, applied);
    reinterpret_cast<float4*>(&sm[0][r][0])[c4] = a;
  }
  float4 ks4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0; r < S; r += rstep) {
    float4 a = __ldg(reinterpret_cast<const float4*>(k + ((size_t)n * S + r) * ldk) + c4);
    float4 b = __ldg(reinterpret_cast<const float4*>(v + ((size_t)n * S + r) * ldv) + c4);
    a.x = fmap_fast(a.x, applied); a.y = fmap_fast(a.y, applied); a.z = fmap_fast(a.z, applied); a.w = fmap_fast(a.w, applied);
    b.x *= invS; b.y *= invS; b.z *= invS; b.w *= invS;  // values / v_length (:44)
    ks4.x += a.x; ks4.y += a.y; ks4.z += a.z; ks4.w += a.w;
    reinterpret_cast<float4*>(&sm[1][r][0])[c4] = a;
    reinterpret_cast<float4*>(&sm[2][r][0])[c4] = b;
  }
  reinterpret_cast<float4*>(&ksp[r0][0])[c4] = ks4;   // r0 = t / 32 = this thread's warp: rows r0, r0 + 8, ...
  __syncthreads();
  if (t < C) {   // Ksum[c] = sum of the 8 per-warp partials (once per CTA: the kernel is shared-memory-wavefront bound)
    float a = ksp[0][t];
#pragma unroll
    for (int w = 1; w < H; ++w) a += ksp[w][t];
    ksum_s[t] = a;
  }
  __syncthreads();
  constexpr int G = 32 / D;      // d-groups per warp (D=16 -> 2)
  constexpr int ND = D / G;      // d's per lane: the CONTIGUOUS block [g*ND, g*ND + ND) -> two broadcast LDS.128 per token
  static_assert(ND == 8 && rstep == H, "la_small_allheads_kernel is written for D = 16, H = 8");
  const int e = lane % D, g = lane / D, hb = h * D;
  // Z for token `lane` of this head: S / (Q[lane, :] . Ksum + eps)   (linear_attention.py:46-48, the S of :44 folded in)
  float z = 0.f;
  if (lane < L) {
    float den = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < D / 4; ++d4) {
      const float4 ksum = reinterpret_cast<const float4*>(&ksum_s[hb])[d4];
      const float4 qq = reinterpret_cast<const float4*>(&sm[0][lane][hb])[d4];
      den = fmaf(qq.x, ksum.x, den); den = fmaf(qq.y, ksum.y, den); den = fmaf(qq.z, ksum.z, den); den = fmaf(qq.w, ksum.w, den);
    }
    z = (float)S / (den + eps);
  }
  float kv[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) kv[i] = 0.f;
  for (int s0 = 0; s0 < S; ++s0) {
    const float ve = sm[2][s0][hb + e];
    const float4 ka = *reinterpret_cast<const float4*>(&sm[1][s0][hb + g * ND]);
    const float4 kb = *reinterpret_cast<const float4*>(&sm[1][s0][hb + g * ND + 4]);
    kv[0] = fmaf(ka.x, ve, kv[0]); kv[1] = fmaf(ka.y, ve, kv[1]); kv[2] = fmaf(ka.z, ve, kv[2]); kv[3] = fmaf(ka.w, ve, kv[3]);
    kv[4] = fmaf(kb.x, ve, kv[4]); kv[5] = fmaf(kb.y, ve, kv[5]); kv[6] = fmaf(kb.z, ve, kv[6]); kv[7] = fmaf(kb.w, ve, kv[7]);
  }
  __syncwarp();   // the K columns of this head are dead from here on (other warps only touch their own columns)
  for (int l = 0; l < L; ++l) {
    const float4 qa = *reinterpret_cast<const float4*>(&sm[0][l][hb + g * ND]);
    const float4 qb = *reinterpret_cast<const float4*>(&sm[0][l][hb + g * ND + 4]);
    float num = qa.x * kv[0];
    num = fmaf(qa.y, kv[1], num); num = fmaf(qa.z, kv[2], num); num = fmaf(qa.w, kv[3], num);
    num = fmaf(qb.x, kv[4], num); num = fmaf(qb.y, kv[5], num); num = fmaf(qb.z, kv[6], num); num = fmaf(qb.w, kv[7], num);
    num += __shfl_xor_sync(0xffffffffu, num, D);
    const float zl = __shfl_sync(0xffffffffu, z, l);
    if (g == 0) sm[1][l][hb + e] = num * zl;
  }
  __syncthreads();
  for (int r = r0; r < L; r += rstep)
    reinterpret_cast<float4*>(out + ((size_t)n * L + r) * ldo)[c4] = reinterpret_cast<const float4*>(&sm[1][r][0])[c4];
}

// ------------------------------------------------------------------------------------------------------------------
// D = 32 fast path: one CTA covers ALL heads of a token tile, so every global access is a full contiguous C-float row
// (1 KB at C = 256) instead of eight 128-byte head slices read by eight different CTAs; 16 float4 loads in flight per
// thread.  Warp w owns head w.  Same partial layout as la_reduce_kernel, so both apply kernels can consume it.
constexpr int LA2_T = 32;  // tokens per shared-memory tile

template <int H>
__global__ void __launch_bounds__(32 * H) la_reduce_allheads_kernel(const float* __restrict__ k, int ldk,
                                                                    const float* __restrict__ v, int ldv, int S,
                                                                    int applied, int chunk, float* __restrict__ ws) {
  constexpr int D = 32, C = H * D, NT = 32 * H;
  extern __shared__ __align__(16) float la_sm[];
  float(*Ks)[C] = reinterpret_cast<float(*)[C]>(la_sm);
  float(*Vs)[C] = reinterpret_cast<float(*)[C]>(la_sm + LA2_T * C);
  const int n = blockIdx.x, z = blockIdx.y, t = threadIdx.x, h = t >> 5, lane = t & 31;
  const int dq = lane >> 3, eq = lane & 7;  // this lane: d in [8 dq, 8 dq + 8), e in [4 eq, 4 eq + 4)
  const int s_beg = z * chunk, s_end = min(S, s_beg + chunk);
  const float fS = (float)S;
  float acc[8][4];
  float ksum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ksum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  const float* kb = k + (size_t)n * S * ldk;
  const float* vb = v + (size_t)n * S * ldv;
  for (int s0 = s_beg; s0 < s_end; s0 += LA2_T) {
    const int cnt = min(LA2_T, s_end - s0);
    for (int idx = t; idx < LA2_T * C / 4; idx += NT) {
      const int r = idx / (C / 4), c4 = idx % (C / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (r < cnt) {
        kv = __ldg(reinterpret_cast<const float4*>(kb + (size_t)(s0 + r) * ldk) + c4);
        vv = __ldg(reinterpret_cast<const float4*>(vb + (size_t)(s0 + r) * ldv) + c4);
        kv.x = fmap(kv.x, applied); kv.y = fmap(kv.y, applied); kv.z = fmap(kv.z, applied); kv.w = fmap(kv.w, applied);
        vv.x /= fS; vv.y /= fS; vv.z /= fS; vv.w /= fS;  // values / v_length (:44)
      }
      reinterpret_cast<float4*>(&Ks[r][0])[c4] = kv;
      reinterpret_cast<float4*>(&Vs[r][0])[c4] = vv;
    }
    __syncthreads();
    for (int r = 0; r < cnt; ++r) {
      const float4 k0 = *reinterpret_cast<const float4*>(&Ks[r][h * D + dq * 8]);
      const float4 k1 = *reinterpret_cast<const float4*>(&Ks[r][h * D + dq * 8 + 4]);
      const float4 v0 = *reinterpret_cast<const float4*>(&Vs[r][h * D + eq * 4]);
      const float kk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
      const float vv[4] = {v0.x, v0.y, v0.z, v0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ksum[i] += kk[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(kk[i], vv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  float* out = ws + ((size_t)(n * H + h) * gridDim.y + z) * (D * D + D);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    *reinterpret_cast<float4*>(&out[(dq * 8 + i) * D + eq * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (eq == 0) out[D * D + dq * 8 + i] = ksum[i];
  }
}

template <int H>
__global__ void __launch_bounds__(32 * H) la_apply_allheads_kernel(const float* __restrict__ q, int ldq,
                                                                   float* __restrict__ out, int ldo, int L, int S,
                                                                   int applied, int splits, float eps,
                                                                   const float* __restrict__ ws) {
  constexpr int D = 32, C = H * D, NT = 32 * H, KVP = D * (D + 1);
  extern __shared__ __align__(16) float la_sm[];
  float(*Qs)[C] = reinterpret_cast<float(*)[C]>(la_sm);        // [LA2_T][C]
  float* KV = la_sm + LA2_T * C;                                 // [H][D][D+1]
  float* Ksum = KV + H * KVP;                                    // [H][D]
  const int n = blockIdx.x, l0 = blockIdx.y * LA2_T, t = threadIdx.x, h = t >> 5, lane = t & 31;
  const int cnt = min(LA2_T, L - l0);
  for (int idx = t; idx < H * (D * D + D); idx += NT) {
    const int hh = idx / (D * D + D), e = idx % (D * D + D);
    float sacc = 0.f;
    for (int zz = 0; zz < splits; ++zz) sacc += ws[((size_t)(n * H + hh) * splits + zz) * (D * D + D) + e];
    if (e < D * D) KV[hh * KVP + (e / D) * (D + 1) + e % D] = sacc; else Ksum[hh * D + e - D * D] = sacc;
  }
  const float* qb = q + ((size_t)n * L + l0) * ldq;
  for (int idx = t; idx < LA2_T * C / 4; idx += NT) {
    const int r = idx / (C / 4), c4 = idx % (C / 4);
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < cnt) {
      qv = __ldg(reinterpret_cast<const float4*>(qb + (size_t)r * ldq) + c4);
      qv.x = fmap(qv.x, applied); qv.y = fmap(qv.y, applied); qv.z = fmap(qv.z, applied); qv.w = fmap(qv.w, applied);
    }
    reinterpret_cast<float4*>(&Qs[r][0])[c4] = qv;
  }
  __syncthreads();
  // lane e keeps column e of this head's KV (32 registers) and Ksum[e]; per token the Q row is read as eight broadcast
  // float4 loads (one shared-memory wavefront each) and the denominator comes from a 5-step warp reduction:
  // ~9 LDS + 32 FMA + 5 SHFL per token per warp (the previous version issued 9 LDS per 4 d and recomputed den in
  // every lane: LDS-bound at ~400 us per call).
  const float* kvh = KV + h * KVP;
  float kv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) kv[d] = kvh[d * (D + 1) + lane];
  const float ks = Ksum[h * D + lane];
  for (int r = 0; r < cnt; ++r) {
    const float* qr = &Qs[r][h * D];
    float den = qr[lane] * ks;
    float num = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < D; d4 += 4) {
      const float4 qq = *reinterpret_cast<const float4*>(qr + d4);
      num = fmaf(qq.x, kv[d4 + 0], num);
      num = fmaf(qq.y, kv[d4 + 1], num);
      num = fmaf(qq.z, kv[d4 + 2], num);
      num = fmaf(qq.w, kv[d4 + 3], num);
    }
    den = warp_sum(den);
    out[((size_t)n * L + l0 + r) * ldo + h * D + lane] = num * (1.f / (den + eps)) * (float)S;
  }
}

static inline bool ptr_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// sum of the per-split partials (fixed order) -> one [D*D + D] record per (n, h)
__global__ void la_partial_sum_kernel(const float* __restrict__ ws, int splits, int rec, long long total,
                                      float* __restrict__ outp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long nh = idx / rec;
  const int e = (int)(idx % rec);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[((size_t)nh * splits + z) * rec + e];
  outp[idx] = s;
}

static int la_splits_allheads(int N, int S) {
  // la_reduce_allheads_kernel runs 3 CTAs per SM (64 KB smem, 72 registers): size the grid N x splits to fit ONE wave
  // of 3 x 148 resident CTAs (rounding up gave 448 CTAs for N = 32: a 4-CTA second wave doubled the kernel time)
  int s = (3 * kNumSMs) / N;
  const int maxs = ceil_div(S, 4 * LA2_T);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return s;
}

static int la_splits(int N, int S, int H) {
  const long long base = (long long)N * H;
  int s = (int)((4LL * kNumSMs + base - 1) / base);
  const int maxs = ceil_div(S, 2 * LA_TOK);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return s;
}

size_t linear_attention_ws_bytes(int N, int S, int H, int D);

int linear_attention_dispatch(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                              int ldo, int N, int L, int S, int H, int D, float eps, int applied, float* workspace,
                              size_t workspace_bytes, cudaStream_t st) {
  if (N <= 0 || L <= 0) return FAR_OK;
  FAR_REQUIRE(q && k && v && out && S > 0 && (D == 16 || D == 32));
  if (L <= LA_SMALL && S <= LA_SMALL && D == 16) {
    const long long NH = (long long)N * H;
    const unsigned blocks = (unsigned)ceil_div_ll(NH, 8);
    ProfScope prof(PROF_LA_SMALL, 4.0 * NH * ((double)L + S) * 16 * 16, 4.0 * NH * 16 * (2.0 * L + 2.0 * S), st);
    if (H == 8 && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 &&
        ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
          reinterpret_cast<uintptr_t>(out)) & 15u) == 0)
      la_small_allheads_kernel<16, 8><<<(unsigned)N, 256, 0, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, L, S, applied, eps);
    else
      la_small_kernel<16><<<blocks, 256, 0, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, NH, L, S, H, applied, eps);
    FAR_CHECK_LAUNCH();
    return FAR_OK;
  }
  if (workspace == nullptr || workspace_bytes < linear_attention_ws_bytes(N, S, H, D) - 256) return FAR_ERR_WORKSPACE;
  const int splits = la_splits(N, S, H);
  if (D == 32 && H == 8 && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ptr_al16(q) && ptr_al16(k) && ptr_al16(v)) {
    constexpr int HH = 8, CC = HH * 32;
    const int sp2 = la_splits_allheads(N, S);
    const int chunk2 = ceil_div(ceil_div(S, sp2), LA2_T) * LA2_T;
    const int rec = 32 * 32 + 32;
    float* summed = workspace + (size_t)N * HH * sp2 * rec;
    const size_t sm1 = (size_t)2 * LA2_T * CC * 4, sm2 = ((size_t)LA2_T * CC + HH * (32 * 33) + HH * 32) * 4;
    static bool attr[64] = {};
    if (first_use_on_device(attr)) {
      cudaFuncSetAttribute(la_reduce_allheads_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
      cudaFuncSetAttribute(la_apply_allheads_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    }
    {
      ProfScope prof(PROF_LA_REDUCE, 2.0 * N * S * CC * 32, 4.0 * 2.0 * N * S * CC, st);
      la_reduce_allheads_kernel<HH><<<dim3(N, sp2), 32 * HH, sm1, st>>>(k, ldk, v, ldv, S, applied, chunk2, workspace);
    }
    FAR_CHECK_LAUNCH();
    const long long tot = (long long)N * HH * rec;
    la_partial_sum_kernel<<<(unsigned)ceil_div_ll(tot, 256), 256, 0, st>>>(workspace, sp2, rec, tot, summed);
    FAR_CHECK_LAUNCH();
    ProfScope prof(PROF_LA_APPLY, 2.0 * N * L * CC * 32, 4.0 * 2.0 * N * L * CC, st);
    la_apply_allheads_kernel<HH><<<dim3(N, ceil_div(L, LA2_T)), 32 * HH, sm2, st>>>(q, ldq, out, ldo, L, S, applied, 1, eps,
                                                                                    summed);
    FAR_CHECK_LAUNCH();
    return FAR_OK;
  }
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

// ---- pieces used by the fused encoder layer (encoder_layer.cu) ----------------------------------------------------
// KV / Ksum reduction only (D = 32, H = 8): returns the summed records [(n*H + h)][D*D + D] (KV[d][e] with v already
// divided by S, then Ksum[d]) inside `workspace`.
int la_reduce_summed(const float* k, int ldk, const float* v, int ldv, int N, int S, int applied, float* workspace,
                     size_t workspace_bytes, const float** summed_out, cudaStream_t st) {
  constexpr int HH = 8, CC = HH * 32;
  FAR_REQUIRE(k && v && workspace && ldk % 4 == 0 && ldv % 4 == 0 && ptr_al16(k) && ptr_al16(v));
  if (workspace_bytes < linear_attention_ws_bytes(N, S, HH, 32) - 256) return FAR_ERR_WORKSPACE;
  const int sp2 = la_splits_allheads(N, S);
  const int chunk2 = ceil_div(ceil_div(S, sp2), LA2_T) * LA2_T;
  const int rec = 32 * 32 + 32;
  float* summed = workspace + (size_t)N * HH * sp2 * rec;
  const size_t sm1 = (size_t)2 * LA2_T * CC * 4;
  static bool attr[64] = {};
  if (first_use_on_device(attr)) {
    cudaFuncSetAttribute(la_reduce_allheads_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
  }
  {
    ProfScope prof(PROF_LA_REDUCE, 2.0 * N * S * CC * 32, 4.0 * 2.0 * N * S * CC, st);
    la_reduce_allheads_kernel<HH><<<dim3(N, sp2), 32 * HH, sm1, st>>>(k, ldk, v, ldv, S, applied, chunk2, workspace);
  }
  FAR_CHECK_LAUNCH();
  const long long tot = (long long)N * HH * rec;
  la_partial_sum_kernel<<<(unsigned)ceil_div_ll(tot, 256), 256, 0, st>>>(workspace, sp2, rec, tot, summed);
  FAR_CHECK_LAUNCH();
  *summed_out = summed;
  return FAR_OK;
}

// Folds the attention apply step into the `merge` projection (transformer.py:58, linear_attention.py:47-50):
//   message[l, o] = sum_{h,e} (Z Q KV_h S)[l, h*32+e] Wm[o, h*32+e] = sum_{h,d} (Z Q)[l, h*32+d] * Bn[o, h*32+d],
//   Bn[o, h*32+d] = S * sum_e KV[n,h,d,e] * Wm[o, h*32+e]      (one [C,C] matrix per batch element n)
// written directly as the tf32 hi / lo operand pair of the tcgen05 engine.  grid (N, H), block C (thread = o).
template <bool kCat>
__global__ void __launch_bounds__(256) la_fold_merge_kernel(const float* __restrict__ summed, const float* __restrict__ Wm,
                                                            int C, float fS, float* __restrict__ bhi,
                                                            float* __restrict__ blo) {
  constexpr int D = 32, REC = D * D + D;
  __shared__ float KV[D][D + 1];
  const int n = blockIdx.x, h = blockIdx.y, H = gridDim.y, o = threadIdx.x;
  const float* rec = summed + (size_t)(n * H + h) * REC;
  for (int idx = threadIdx.x; idx < D * D; idx += blockDim.x) KV[idx / D][idx % D] = rec[idx];
  __syncthreads();
  if (o >= C) return;
  float w[D];
#pragma unroll
  for (int e = 0; e < D; e += 4) {
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(Wm + (size_t)o * C + h * D + e));
    w[e] = w4.x; w[e + 1] = w4.y; w[e + 2] = w4.z; w[e + 3] = w4.w;
  }
  float* oh = bhi + ((size_t)n * C + o) * C + h * D;
  float* ol = blo + ((size_t)n * C + o) * C + h * D;
  if (kCat) {
    // cross16 operand form (tc_common.cuh): oh = the fp32 values, ol = [32 x bf16(lo) | 32 x bf16(value)] for this 32-wide
    // k-block (row o, columns h*32 .. h*32+31)
    uint32_t* oc = reinterpret_cast<uint32_t*>(ol);
#pragma unroll 2
    for (int d = 0; d < D; d += 2) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int e = 0; e < D; ++e) { a0 = fmaf(KV[d][e], w[e], a0); a1 = fmaf(KV[d + 1][e], w[e], a1); }
      a0 *= fS; a1 *= fS;
      oh[d] = a0; oh[d + 1] = a1;
      const float l0 = a0 - __uint_as_float(__float_as_uint(a0) & 0xFFFFE000u);
      const float l1 = a1 - __uint_as_float(__float_as_uint(a1) & 0xFFFFE000u);
      uint32_t plo, phi;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(plo) : "f"(l1), "f"(l0));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(phi) : "f"(a1), "f"(a0));
      oc[d >> 1] = plo;
      oc[16 + (d >> 1)] = phi;
    }
  } else {
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < D; ++e) acc = fmaf(KV[d][e], w[e], acc);
      acc *= fS;
      const float hi = __uint_as_float(__float_as_uint(acc) & 0xFFFFE000u);
      oh[d] = hi;
      ol[d] = acc - hi;
    }
  }
}

int la_fold_merge(const float* summed, const float* Wm, int N, int C, int S, float* bhi, float* blo, bool cat,
                  cudaStream_t st) {
  FAR_REQUIRE(summed && Wm && bhi && blo && C % 32 == 0 && C <= 256 && (reinterpret_cast<uintptr_t>(Wm) & 15u) == 0);
  if (cat) la_fold_merge_kernel<true><<<dim3(N, C / 32), 256, 0, st>>>(summed, Wm, C, (float)S, bhi, blo);
  else la_fold_merge_kernel<false><<<dim3(N, C / 32), 256, 0, st>>>(summed, Wm, C, (float)S, bhi, blo);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Fused encoder-layer path (encoder_layer.cu): KV_h = sum_s K'_h[s]^T V_h[s] and Ksum_h = sum_s K'_h[s] straight from the
// [K' | V] block the projection GEMM just wrote ([N][S][2C], K' already holds elu+1), WITHOUT the v/S ... *S round trip
// of linear_attention.py:43-50 (it cancels in  Bn = S * (K'^T V / S) * Wm^T; the raw sum is the better-rounded value).
// HBM-bound (2 KB per token): a 3-stage cp.async ring keeps two 32 KB tiles in flight per CTA while the third is
// consumed; batch elements are walked in REVERSE order because the GEMM wrote them in forward order -- the tail of the
// 315 MB block is still in the 126 MB L2.  Partial sums per (n, head, split) in the la_reduce layout; the split merge
// happens inside la_fold_merge_splits_kernel (fixed order), so no separate partial-sum launch.
constexpr int LR_T = 16;        // tokens per stage
constexpr int LR_STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;   // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int H>
__global__ void __launch_bounds__(32 * H) la_reduce_kv_async_kernel(const float* __restrict__ kv, int S, int chunk,
                                                                    float* __restrict__ ws) {
  constexpr int D = 32, C = H * D, NT = 32 * H, ROW = 2 * C, ROW4 = ROW / 4;
  extern __shared__ __align__(16) float la_sm[];
  const int n = gridDim.x - 1 - blockIdx.x, z = blockIdx.y, t = threadIdx.x, h = t >> 5, lane = t & 31;
  const int dq = lane >> 3, eq = lane & 7;  // this lane: d in [8 dq, 8 dq + 8), e in [4 eq, 4 eq + 4)
  const int s_beg = z * chunk, s_end = min(S, s_beg + chunk);
  const int ntiles = (s_end - s_beg + LR_T - 1) / LR_T;
  const float* base = kv + (size_t)n * S * ROW;
  auto issue = [&](int tile) {
    if (tile < ntiles) {
      float* dst = la_sm + (size_t)(tile % LR_STAGES) * LR_T * ROW;
      const int s0 = s_beg + tile * LR_T;
#pragma unroll
      for (int i = 0; i < LR_T * ROW4 / NT; ++i) {
        const int idx = t + i * NT, r = idx / ROW4, c4 = idx % ROW4;
        const bool ok = s0 + r < s_end;
        cp_async16(dst + (size_t)r * ROW + c4 * 4, base + (size_t)(ok ? s0 + r : s_beg) * ROW + c4 * 4, ok);
      }
    }
    cp_async_commit();
  };
  float acc[8][4];
  float ksum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ksum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
#pragma unroll
  for (int st = 0; st < LR_STAGES - 1; ++st) issue(st);
  for (int tile = 0; tile < ntiles; ++tile) {
    cp_async_wait<LR_STAGES - 2>();
    __syncthreads();                       // tile `tile` landed for every thread; slot (tile-1) % STAGES is free again
    issue(tile + LR_STAGES - 1);
    const float* sm = la_sm + (size_t)(tile % LR_STAGES) * LR_T * ROW;
#pragma unroll 4
    for (int r = 0; r < LR_T; ++r) {       // rows past s_end are zero-filled: they add nothing
      const float4 k0 = *reinterpret_cast<const float4*>(&sm[r * ROW + h * D + dq * 8]);
      const float4 k1 = *reinterpret_cast<const float4*>(&sm[r * ROW + h * D + dq * 8 + 4]);
      const float4 v0 = *reinterpret_cast<const float4*>(&sm[r * ROW + C + h * D + eq * 4]);
      const float kk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
      const float vv[4] = {v0.x, v0.y, v0.z, v0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ksum[i] += kk[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(kk[i], vv[j], acc[i][j]);
      }
    }
  }
  cp_async_wait<0>();
  float* out = ws + ((size_t)(n * H + h) * gridDim.y + z) * (D * D + D);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    *reinterpret_cast<float4*>(&out[(dq * 8 + i) * D + eq * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (eq == 0) out[D * D + dq * 8 + i] = ksum[i];
  }
}

// Split merge (fixed order) + Bn = scale * KV_n W_merge^T as the tcgen05 operand pair, and the summed record
// (KV | Ksum) for the q-projection epilogue's normaliser.  grid (N, H), block C (thread = o).
template <bool kCat>
__global__ void __launch_bounds__(256) la_fold_merge_splits_kernel(const float* __restrict__ parts, int splits,
                                                                   const float* __restrict__ Wm, int C, float scale,
                                                                   float* __restrict__ summed, float* __restrict__ bhi,
                                                                   float* __restrict__ blo) {
  constexpr int D = 32, REC = D * D + D;
  __shared__ float KV[D][D + 1];
  const int n = blockIdx.x, h = blockIdx.y, H = gridDim.y, o = threadIdx.x;
  const float* src = parts + (size_t)(n * H + h) * splits * REC;
  float* rec = summed + (size_t)(n * H + h) * REC;
  for (int idx = threadIdx.x; idx < REC; idx += blockDim.x) {
    float a = 0.f;
    for (int zz = 0; zz < splits; ++zz) a += src[(size_t)zz * REC + idx];
    rec[idx] = a;
    if (idx < D * D) KV[idx / D][idx % D] = a;
  }
  __syncthreads();
  if (o >= C) return;
  float w[D];
#pragma unroll
  for (int e = 0; e < D; e += 4) {
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(Wm + (size_t)o * C + h * D + e));
    w[e] = w4.x; w[e + 1] = w4.y; w[e + 2] = w4.z; w[e + 3] = w4.w;
  }
  float* oh = bhi + ((size_t)n * C + o) * C + h * D;
  float* ol = blo + ((size_t)n * C + o) * C + h * D;
  uint32_t* oc = reinterpret_cast<uint32_t*>(ol);
#pragma unroll 2
  for (int d = 0; d < D; d += 2) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int e = 0; e < D; ++e) { a0 = fmaf(KV[d][e], w[e], a0); a1 = fmaf(KV[d + 1][e], w[e], a1); }
    a0 *= scale; a1 *= scale;
    const float h0 = __uint_as_float(__float_as_uint(a0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(a1) & 0xFFFFE000u);
    if (kCat) {   // cross16 operand form (tc_common.cuh): raw fp32 | [32 x bf16(lo) | 32 x bf16(value)]
      oh[d] = a0; oh[d + 1] = a1;
      uint32_t plo, phi;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(plo) : "f"(a1 - h1), "f"(a0 - h0));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(phi) : "f"(a1), "f"(a0));
      oc[d >> 1] = plo;
      oc[16 + (d >> 1)] = phi;
    } else {
      oh[d] = h0; oh[d + 1] = h1;
      ol[d] = a0 - h0; ol[d + 1] = a1 - h1;
    }
  }
}

// [K'|V] block -> (summed record, Bn operand pair) in two launches.  workspace: la_reduce partial layout.
int la_reduce_fold(const float* kv, int N, int S, const float* Wm, float* workspace, size_t workspace_bytes, float* bhi,
                   float* blo, bool cat, const float** summed_out, cudaStream_t st) {
  constexpr int HH = 8, CC = HH * 32, REC = 32 * 32 + 32;
  FAR_REQUIRE(kv && Wm && workspace && bhi && blo && ptr_al16(kv) && ptr_al16(Wm));
  if (workspace_bytes < linear_attention_ws_bytes(N, S, HH, 32) - 256) return FAR_ERR_WORKSPACE;
  // one wave of 2 resident CTAs per SM (96 KB of shared memory each)
  int sp = (2 * kNumSMs) / N;
  const int maxs = ceil_div(S, 4 * LR_T);
  if (sp > maxs) sp = maxs;
  if (sp < 1) sp = 1;
  const int lim = la_splits_allheads(N, S) > la_splits(N, S, HH) ? la_splits_allheads(N, S) : la_splits(N, S, HH);
  if (sp > lim) sp = lim;                  // the workspace is sized for at most `lim` partial records per (n, head)
  const int chunk = ceil_div(ceil_div(S, sp), LR_T) * LR_T;
  float* summed = workspace + (size_t)N * HH * sp * REC;
  const size_t smem = (size_t)LR_STAGES * LR_T * 2 * CC * 4;
  static bool attr[64] = {};
  if (first_use_on_device(attr))
    cudaFuncSetAttribute(la_reduce_kv_async_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  {
    ProfScope prof(PROF_LA_REDUCE, 2.0 * N * S * CC * 32, 4.0 * 2.0 * N * S * CC, st);
    la_reduce_kv_async_kernel<HH><<<dim3(N, sp), 32 * HH, smem, st>>>(kv, S, chunk, workspace);
  }
  FAR_CHECK_LAUNCH();
  if (cat) la_fold_merge_splits_kernel<true><<<dim3(N, HH), 256, 0, st>>>(workspace, sp, Wm, CC, 1.f, summed, bhi, blo);
  else la_fold_merge_splits_kernel<false><<<dim3(N, HH), 256, 0, st>>>(workspace, sp, Wm, CC, 1.f, summed, bhi, blo);
  FAR_CHECK_LAUNCH();
  *summed_out = summed;
  return FAR_OK;
}

size_t linear_attention_ws_bytes(int N, int S, int H, int D) {
  const int a = la_splits(N, S, H), b = la_splits_allheads(N, S);
  return (size_t)N * H * ((a > b ? a : b) + 1) * (D * D + D) * sizeof(float) + 256;
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
