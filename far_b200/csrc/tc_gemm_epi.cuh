// Shared pieces of the tcgen05 GEMM kernels (tc_gemm.cu: SS operands; tc_gemm_ts.cu: A operand through TMEM): launch
// arguments, TMA store / prefetch wrappers and the epilogue of one 32-row x 32-column accumulator block.
#pragma once
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace far {
namespace tc {

constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;       // 320: TMA + MMA + 8 epilogue warps
constexpr int GEMM_THREADS_RAW = GEMM_THREADS + 128;    // + 4 converter warps
constexpr int STG_TILE = 32 * 128;                      // one 32-row x 128-byte staging tile per epilogue warp
constexpr size_t GEMM_SMEM = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + EPI_WARPS * STG_TILE + 256;
// tc_gemm_ts.cu (A operand through TMEM): 4 stages of (raw A | B | B) tiles
constexpr int TS_STAGES = 4;
constexpr int TS_STAGE_BYTES = 3 * TILE_BYTES;
constexpr size_t GEMM_TS_SMEM = 1024 + (size_t)TS_STAGES * TS_STAGE_BYTES + EPI_WARPS * STG_TILE + 256;
// tc_gemm_pair.cu (CTA pairs): 6 stages of (raw A | B_hi half | B_lo half)
constexpr int TP_STAGES = 6;
constexpr size_t GEMM_PAIR_SMEM = 1024 + (size_t)TP_STAGES * (TILE_BYTES + TILE_BYTES) + EPI_WARPS * STG_TILE + 256;
constexpr int ACT_ELU1Z = 100;  // elu(x)+1 followed by the per-(row, head) linear-attention normaliser (D = 32)

struct GemmArgs {
  const float* bias;
  int M, N, K;      // M = G * L
  int G, L;         // groups x rows per group (G = 1, L = M when ungrouped)
  int b_grouped;    // B operand indexed by group
  int act, act_cols;
  int kb1;          // raw-A: number of 32-wide k-blocks that come from x1 (= K1 / 32)
  const float* ksum; int ksum_rec, ksum_off, heads; float eps;  // ACT_ELU1Z: Ksum[(g*heads + h)*ksum_rec + ksum_off + d]
  const float* rowbias; int rb_group;  // + rowbias[(row / rb_group) * N + col]  (fine_preprocess: per-match coarse term)
  // LayerNorm fused across two GEMMs (encoder_layer.cu: norm1 between `message` and mlp.0): the producer's epilogue writes
  // per (row, 32-column block) the block mean and M2 = sum (v - mean)^2 (ln_out [M][N/32]); the consumer (tc_gemm_ts_kernel
  // only) combines a row's ln_chunks partials (Chan) and its converter threads apply (v - mu) rstd gamma + beta to the
  // k-blocks of the SECOND A segment before the hi/lo split.
  float2* ln_out;
  const float2* ln_in; const float* ln_gamma; const float* ln_beta; float ln_eps; int ln_chunks;
  int cross16;      // raw-A only: slot 1 = A_cat, slot 3 = B_cat, cross terms as bf16 MMAs (tc_common.cuh)
  int dbg;          // diagnostics (env FAR_TC_DBG): 1 = skip global stores, 2 = skip the epilogue body, 4 = skip MMAs
};

void launch_gemm_ts(int grid, cudaStream_t st, const CUtensorMap& mA1, const CUtensorMap& mA2, const CUtensorMap& mBhi,
                    const CUtensorMap& mBlo, const CUtensorMap& mC, const GemmArgs& p);

// CTA-pair kernel: resident clusters of 2 on this device (0: cluster launch unavailable), launch on `clusters` pairs
int gemm_pair_max_clusters();
void launch_gemm_pair(int clusters, cudaStream_t st, const CUtensorMap& mA1, const CUtensorMap& mA2, const CUtensorMap& mBhi,
                      const CUtensorMap& mBlo, const CUtensorMap& mC, const GemmArgs& p);
int gemm_pair_debug_counters(unsigned long long* out16);
int gemm_pair_debug_trace(unsigned long long* out256);
int gemm_pair_debug_cta(unsigned long long* out320);

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Epilogue of one 32 x 32 block: t = this thread's 32 accumulator values (row = TMEM lane, columns col0..col0+31) ->
// row-bias / bias / activation (optionally the linear-attention normaliser Z) -> eight conflict-free STS.128 into the
// warp's SWIZZLE_128B staging tile -> one TMA store.
__device__ __forceinline__ void epi_block(float (&t)[32], const GemmArgs& p, const CUtensorMap* mapC, int col0, int g,
                                          int r0, int quarter, int lane, int actc, float4* srow, int sx,
                                          uint32_t stg_addr, unsigned long long* trace = nullptr) {
    // one activation for the whole 32-column block, or -1: the block straddles N / act_cols (per-element path)
    const int act_here = (col0 + 32 <= actc) ? p.act : ((col0 >= actc) ? FAR_ACT_NONE : -1);
    const int grow = g * p.L + r0 + quarter * 32 + lane;  // this thread's global output row
    const bool rb_on = p.rowbias != nullptr && grow < p.M && r0 + quarter * 32 + lane < p.L;
    if (col0 + 32 <= p.N && act_here >= 0) {
      if (rb_on) {
        const float* rb = p.rowbias + (size_t)(grow / p.rb_group) * p.N + col0;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(rb + e));
          t[e] += b4.x; t[e + 1] += b4.y; t[e + 2] += b4.z; t[e + 3] += b4.w;
        }
      }
      if (p.bias != nullptr) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + e));
          t[e] += b4.x; t[e + 1] += b4.y; t[e + 2] += b4.z; t[e + 3] += b4.w;
        }
      }
      switch (act_here) {
        case FAR_ACT_RELU:
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = fmaxf(t[e], 0.f);
          break;
        case FAR_ACT_GELU:
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = 0.5f * t[e] * (1.f + erff(t[e] * 0.70710678118654752440f));
          break;
        case FAR_ACT_ELU1:
        case ACT_ELU1Z:
          // elu(x) + 1 = x + 1 (x > 0) | exp(x) (x <= 0); branch-free, ex2.approx on the clamped argument
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float ex = exp2f(fminf(t[e], 0.f) * 1.4426950408889634f);
            t[e] = t[e] > 0.f ? t[e] + 1.f : ex;
          }
          if (act_here == ACT_ELU1Z) {  // Z = 1 / (Q . Ksum + eps) for this row's head (linear_attention.py:46)
            const float* ks = p.ksum + (size_t)(g * p.heads + (col0 >> 5)) * p.ksum_rec + p.ksum_off;
            float den = 0.f;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 k4 = __ldg(reinterpret_cast<const float4*>(ks + e));
              den = fmaf(t[e], k4.x, den); den = fmaf(t[e + 1], k4.y, den);
              den = fmaf(t[e + 2], k4.z, den); den = fmaf(t[e + 3], k4.w, den);
            }
            const float z = 1.f / (den + p.eps);
#pragma unroll
            for (int e = 0; e < 32; ++e) t[e] *= z;
          }
          break;
        case FAR_ACT_SIGMOID:
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = 1.f / (1.f + __expf(-t[e]));
          break;
        default:
          break;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int col = col0 + e;
        if (col < p.N) {
          if (rb_on) t[e] += __ldg(p.rowbias + (size_t)(grow / p.rb_group) * p.N + col);
          if (p.bias) t[e] += __ldg(p.bias + col);
          if (col < actc) t[e] = apply_act(t[e], p.act == ACT_ELU1Z ? FAR_ACT_ELU1 : p.act);
        }
      }
    }
    if (p.ln_out != nullptr && col0 + 32 <= p.N && r0 + quarter * 32 + lane < p.L) {
      float sum = 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) sum += t[e];
      const float mean = sum * (1.f / 32.f);
      float m2 = 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) { const float d = t[e] - mean; m2 = fmaf(d, d, m2); }
      p.ln_out[(size_t)grow * (p.N >> 5) + (col0 >> 5)] = make_float2(mean, m2);
    }
    if (trace) trace[0] = clock64();   // bias / activation done
    // the previous TMA store of this warp must have finished READING the staging tile
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    if (trace) trace[1] = clock64();   // staging tile free
#pragma unroll
    for (int j = 0; j < 8; ++j) srow[j ^ sx] = make_float4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA store
    __syncwarp();
    if (lane == 0 && !(p.dbg & 1) && r0 + quarter * 32 < p.L) {
      tma_store_4d(mapC, stg_addr, col0, r0 + quarter * 32, g, 0);  // rows >= L / cols >= N are clipped
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (trace) trace[2] = clock64();   // staged + store issued
}

}  // namespace tc
}  // namespace far
