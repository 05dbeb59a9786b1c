// FAR "essential-matrix module" CrossAttention core: dual-softmax bilinear attention.
//   mp3d   : mp3d_loftr/src/loftr/loftr_module/transformer.py:266-303   (N = 4800, 4 heads x 64)
//   8pt-ViT: interiornetStreetlearn_8ptVit/src/modules/vision_transformer.py:177-208 (N = 576, 3 heads x 64)
//
//   S_X = q_other k_X^T * scale ; P_X = softmax(S_X, -1) * softmax(S_X, -2) ; V'_X = [v_X | pos] ;
//   F_X = V'_X^T P_X V'_X   [d+6, d+6]
//
// The reference materialises S and P ([B,h,N,N] fp32: 368 MB each at N=4800, x ~4 temporaries).  Here:
//   pass A (score.cu)      row/col log-sum-exp of S from 128x128 tiles
//   pass B (emm_pv_kernel) one CTA per (128-row i-tile, batch*head): loops over the j-tiles, recomputes the S tile,
//                          P = exp(2s - rowlse_i - collse_j) staged in shared memory, T[i,:] += P V'_j ; then the
//                          partial  F_it = V'_i^T T  goes to the workspace
//   reduce                 F = sum_it F_it (fixed order)
#include "score.cuh"
#include "tc_score.cuh"
#include "tc_emm.cuh"
#include "tc_flash.cuh"
#include "tc_gemm.cuh"
#include <algorithm>

namespace far {

constexpr int EDV = 72;    // padded d+6 (d <= 64... 70 valid columns for d = 64)
constexpr int EPLD = 132;  // P tile leading dim
constexpr int ETLD = 73;   // T tile leading dim (odd: conflict-free scalar column access)

struct EmmSmem {
  TileSmem tile;
  float P[TBM][EPLD];  // later reused as T[TBM][ETLD]
  float V[TBN][EDV];
};

struct EmmArgs {
  ScoreArgs sc;         // A = q_other, B = k_X
  const float* v;       // v_X base (qkv_X + 2*C), same (b,h) strides / row stride as B
  const float* pos;     // [Bpos, N, 6]
  int Bpos;
  int d;                // head dim
  const float* rowlse;  // [G][N]
  const float* collse;  // [G][N]
  float* Fpart;         // [G][IT][dv*dv]
  float* out;           // !kBilinear: attention output [B, N, H*d]
};

__device__ __forceinline__ void load_vprime(const EmmArgs& p, int g, int r0, float (*V)[EDV]) {
  const ScoreArgs& a = p.sc;
  const int b = g / a.H, h = g % a.H;
  const float* vb = p.v + b * a.sBb + h * a.sBh;
  const float* pb = p.pos + (size_t)(p.Bpos == 1 ? 0 : b) * a.S * 6;
  for (int idx = threadIdx.x; idx < TBN * EDV; idx += kTileThreads) {
    const int r = idx / EDV, c = idx % EDV, tok = r0 + r;
    float val = 0.f;
    if (tok < a.S) {
      if (c < p.d) val = vb[(size_t)tok * a.ldb + c];
      else if (p.pos != nullptr && c < p.d + 6) val = pb[(size_t)tok * 6 + (c - p.d)];
    }
    V[r][c] = val;
  }
}

// kBilinear = true : FAR CrossAttention (P = dual softmax, output F_it = V'^T T)
// kBilinear = false: plain softmax attention of the 8pt-ViT blocks (P = exp(s - rowlse), output O = P V)
template <bool kVec4, bool kBilinear>
__global__ void __launch_bounds__(kTileThreads, 1) emm_pv_kernel(EmmArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EmmSmem& sm = *reinterpret_cast<EmmSmem*>(smem_raw);
  const ScoreArgs& a = p.sc;
  const int it = blockIdx.x, g = blockIdx.y, IT = gridDim.x;
  const int i0 = it * TBM;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int tr = t & 31, tc = t >> 5;  // PV mapping: rows {tr + 32 r}, cols tc*9 .. tc*9+8
  const int JT = score_tiles_j(a.S);

  float rl[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i0 + tile_row(ty, i);
    rl[i] = (r < a.L) ? p.rowlse[(size_t)g * a.L + r] : 0.f;
  }
  float T[4][9];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 9; ++c) T[r][c] = 0.f;

  for (int jt = 0; jt < JT; ++jt) {
    const int j0 = jt * TBN;
    float acc[8][8];
    score_tile<kVec4>(a, g, i0, j0, sm.tile, acc);  // ends with __syncthreads()
    float cl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j0 + tile_col(tx, j);
      cl[j] = (kBilinear && c < a.S) ? __ldg(p.collse + (size_t)g * a.S + c) : 0.f;
    }
    // P tile -> smem (row-major, float4 along j: conflict-free)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int lr = tile_row(ty, i);
      const bool rv = (i0 + lr) < a.L;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        float v4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = j0 + tile_col(tx, jh * 4 + j);
          const float s = acc[i][jh * 4 + j];
          const float e = kBilinear ? (s - rl[i]) + (s - cl[jh * 4 + j]) : (s - rl[i]);
          v4[j] = (rv && c < a.S) ? expf(e) : 0.f;
        }
        *reinterpret_cast<float4*>(&sm.P[lr][tile_col(tx, jh * 4)]) = make_float4(v4[0], v4[1], v4[2], v4[3]);
      }
    }
    load_vprime(p, g, j0, sm.V);
    __syncthreads();
    // T[i, c] += sum_j P[i, j] V'[j, c]
#pragma unroll 2
    for (int j = 0; j < TBN; j += 4) {
      float4 pr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) pr[r] = *reinterpret_cast<const float4*>(&sm.P[tr + 32 * r][j]);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float vv[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) vv[c] = sm.V[j + jj][tc * 9 + c];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float pv = jj == 0 ? pr[r].x : (jj == 1 ? pr[r].y : (jj == 2 ? pr[r].z : pr[r].w));
#pragma unroll
          for (int c = 0; c < 9; ++c) T[r][c] = fmaf(pv, vv[c], T[r][c]);
        }
      }
    }
    __syncthreads();  // P / V are overwritten by the next j-tile
  }

  if (!kBilinear) {  // out[b, i, h*d + c] = T[i, c]
    const int b = g / a.H, h = g % a.H, C = a.H * p.d;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = i0 + tr + 32 * r;
      if (row >= a.L) continue;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const int col = tc * 9 + c;
        if (col < p.d) p.out[((size_t)b * a.L + row) * C + h * p.d + col] = T[r][c];
      }
    }
    return;
  }
  // F_it[a, c] = sum_{i in tile} V'[i, a] T[i, c]
  float(*Ts)[ETLD] = reinterpret_cast<float(*)[ETLD]>(&sm.P[0][0]);
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 9; ++c) Ts[tr + 32 * r][tc * 9 + c] = T[r][c];
  load_vprime(p, g, i0, sm.V);  // V' rows of the i-tile (same image as the keys); rows >= N are zero
  __syncthreads();
  const int dv = p.d + 6;
  float* out = p.Fpart + ((size_t)g * IT + it) * dv * dv;
  for (int idx = t; idx < dv * dv; idx += kTileThreads) {
    const int aa = idx / dv, cc = idx % dv;
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < TBM; ++i) s = fmaf(sm.V[i][aa], Ts[i][cc], s);
    out[idx] = s;
  }
}

__global__ void emm_reduce_kernel(const float* __restrict__ Fpart, int IT, int dvdv, long long total,
                                  float* __restrict__ F) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long g = idx / dvdv;
  const int e = (int)(idx % dvdv);
  float s = 0.f;
  for (int it = 0; it < IT; ++it) s += Fpart[((size_t)g * IT + it) * dvdv + e];
  F[idx] = s;
}

static inline size_t al(size_t v) { return (v + 255) & ~size_t(255); }
struct EmmPlan { size_t rowlse, collse, fpart, scratch, tcws, tcws_bytes, vtws, vtws_bytes, total; };
static EmmPlan emm_plan(int B, int N, int h, int d) {
  EmmPlan p; size_t off = 0;
  const int G = B * h, IT = score_tiles_i(N), dv = d + 6;
  p.rowlse = off;  off += al((size_t)G * N * 4);
  p.collse = off;  off += al((size_t)G * N * 4);
  p.fpart = off;   off += al((size_t)G * IT * dv * dv * 4);
  p.scratch = off; off += al(score_lse_scratch_floats(G, N, N) * 4);
  p.tcws_bytes = tc_score_workspace_bytes(G, N, N, d);
  p.tcws = off; off += al(p.tcws_bytes);
  p.vtws_bytes = tc_emm_vt_bytes(G, N);
  p.vtws = off; off += al(p.vtws_bytes);
  p.total = off;
  return p;
}

static int emm_one_direction(const float* qkv_q, const float* qkv_kv, const float* pos, int Bpos, int B, int N, int h,
                             int d, float scale, float* F, char* base, const EmmPlan& pl, int engine, cudaStream_t st) {
  const int C = h * d, G = B * h, IT = score_tiles_i(N), dv = d + 6;
  EmmArgs p;
  ScoreArgs& a = p.sc;
  a.A = qkv_q;           a.sAb = (long long)N * 3 * C; a.sAh = d; a.lda = 3 * C;  // q = qkv[..., 0, :, :]
  a.B = qkv_kv + C;      a.sBb = (long long)N * 3 * C; a.sBh = d; a.ldb = 3 * C;  // k = qkv[..., 1, :, :]
  a.H = h; a.G = G; a.L = N; a.S = N; a.K = d; a.scale = scale;
  float* rowlse = reinterpret_cast<float*>(base + pl.rowlse);
  float* collse = reinterpret_cast<float*>(base + pl.collse);
  int used_tc = 0;
  float* tcws = engine == 1 ? nullptr : reinterpret_cast<float*>(base + pl.tcws);
  int rc = score_lse(a, rowlse, collse, reinterpret_cast<float*>(base + pl.scratch), st, tcws, pl.tcws_bytes, &used_tc);
  if (rc) return rc;
  p.v = qkv_kv + 2 * C;  // v = qkv[..., 2, :, :]
  if (used_tc && tc_emm_supported(N, d) && !getenv("FAR_EMM_SIMT")) {
    // fused tcgen05 recompute pass: S, P (in TMEM) and T = P V' all on the tensor pipe
    float *qhi, *qlo, *khi, *klo;
    tc_score_operand_ptrs(a, tcws, &qhi, &qlo, &khi, &klo);
    float* Fpart = reinterpret_cast<float*>(base + pl.fpart);
    rc = tc_emm_pv(qhi, qlo, khi, klo, p.v, a.sBb, a.sBh, a.ldb, pos, Bpos, G, h, N, d, scale, rowlse, collse, Fpart,
                   reinterpret_cast<float*>(base + pl.vtws), pl.vtws_bytes, st);
    if (rc) return rc;
    const long long total = (long long)G * dv * dv;
    emm_reduce_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(Fpart, IT, dv * dv, total, F);
    FAR_CHECK_LAUNCH();
    return FAR_OK;
  }
  p.pos = pos; p.Bpos = Bpos; p.d = d;
  p.rowlse = rowlse; p.collse = collse;
  p.Fpart = reinterpret_cast<float*>(base + pl.fpart);
  dim3 grid(IT, G);
  const size_t smem = sizeof(EmmSmem);
  p.out = nullptr;
  if (score_vec_ok(a)) {
    cudaFuncSetAttribute(emm_pv_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    emm_pv_kernel<true, true><<<grid, kTileThreads, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(emm_pv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    emm_pv_kernel<false, true><<<grid, kTileThreads, smem, st>>>(p);
  }
  FAR_CHECK_LAUNCH();
  const long long total = (long long)G * dv * dv;
  emm_reduce_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(p.Fpart, IT, dv * dv, total, F);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

}  // namespace far

using namespace far;

extern "C" size_t far_emm_bilinear_attn_workspace_bytes(int B, int Ntok, int h, int d) {
  return emm_plan(B, Ntok, h, d).total + 256;
}

extern "C" int far_emm_bilinear_attn(const float* qkv1, const float* qkv2, const float* pos, int Bpos, int B, int Ntok,
                                     int h, int d, float scale, float* F1, float* F2, int engine, float* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (B <= 0) return FAR_OK;
  FAR_REQUIRE(qkv1 && qkv2 && pos && F1 && F2 && workspace && Ntok > 0 && h > 0 && d > 0 && d + 6 <= EDV &&
              (Bpos == 1 || Bpos == B));
  const EmmPlan pl = emm_plan(B, Ntok, h, d);
  if (workspace_bytes < pl.total) return FAR_ERR_WORKSPACE;
  char* base = reinterpret_cast<char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  // attn_1 = q2 k1^T, values v1  -> fundamental_1 ; attn_2 = q1 k2^T, values v2 -> fundamental_2  (:275-292)
  int rc = emm_one_direction(qkv2, qkv1, pos, Bpos, B, Ntok, h, d, scale, F1, base, pl, engine, st);
  if (rc) return rc;
  return emm_one_direction(qkv1, qkv2, pos, Bpos, B, Ntok, h, d, scale, F2, base, pl, engine, st);
}

// timm Attention core (interiornetStreetlearn_8ptVit/src/modules/vision_transformer.py:250-257)
extern "C" size_t far_softmax_attention_workspace_bytes(int B, int Ntok, int h, int d) {
  size_t n = emm_plan(B, Ntok, h, d).total + 256;
  if (tc_flash_attention_supported(Ntok, d)) n = std::max(n, tc_flash_attention_bytes(B * h, Ntok, d));
  return n;
}

extern "C" int far_softmax_attention(const float* qkv, int B, int Ntok, int h, int d, float scale, float* out,
                                     float* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0) return FAR_OK;
  FAR_REQUIRE(qkv && out && workspace && Ntok > 0 && h > 0 && d > 0 && d <= 64);
  // tcgen05 flash kernel (tc_flash.cu) for head dims 32 / 64; FAR_TC=0 or other head dims: CUDA-core two-pass kernels
  if (tc_engine_default_on() && tc_flash_attention_supported(Ntok, d) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0))
    return tc_flash_attention(qkv, B, Ntok, h, d, scale, out, workspace, workspace_bytes, (cudaStream_t)stream);
  const EmmPlan pl = emm_plan(B, Ntok, h, d);
  if (workspace_bytes < pl.total) return FAR_ERR_WORKSPACE;
  char* base = reinterpret_cast<char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  const int C = h * d, G = B * h, IT = score_tiles_i(Ntok);
  EmmArgs p;
  ScoreArgs& a = p.sc;
  a.A = qkv;     a.sAb = (long long)Ntok * 3 * C; a.sAh = d; a.lda = 3 * C;
  a.B = qkv + C; a.sBb = (long long)Ntok * 3 * C; a.sBh = d; a.ldb = 3 * C;
  a.H = h; a.G = G; a.L = Ntok; a.S = Ntok; a.K = d; a.scale = scale;
  float* rowlse = reinterpret_cast<float*>(base + pl.rowlse);
  float* collse = reinterpret_cast<float*>(base + pl.collse);
  int rc = score_lse(a, rowlse, collse, reinterpret_cast<float*>(base + pl.scratch), st);
  if (rc) return rc;
  p.v = qkv + 2 * C; p.pos = nullptr; p.Bpos = 1; p.d = d;
  p.rowlse = rowlse; p.collse = collse; p.Fpart = nullptr; p.out = out;
  dim3 grid(IT, G);
  const size_t smem = sizeof(EmmSmem);
  if (score_vec_ok(a)) {
    cudaFuncSetAttribute(emm_pv_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    emm_pv_kernel<true, false><<<grid, kTileThreads, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(emm_pv_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    emm_pv_kernel<false, false><<<grid, kTileThreads, smem, st>>>(p);
  }
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}
