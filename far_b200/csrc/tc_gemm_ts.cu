// tcgen05 3xTF32 GEMM with the A operand in TENSOR MEMORY (raw fp32 activations):  C = act(A B^T + bias).
//
// Why: in tc_gemm_kernel (SS operands) every 32-wide k-block moves 176 KB through shared memory (TMA fills 48, the hi/lo
// converter 32, UMMA operand reads 80, epilogue staging 16) against 768 cycles of MMA math
// (profiles/r1_mma_rate_microbench.md).  Here the converter warps read the landed raw A tile ONCE (16 KB) and write hi and lo
// straight into TMEM with tcgen05.st; all MMAs take A from TMEM (TS form), so the port only carries the TMA fills (48),
// that one read (16), the B operand reads (48) and the epilogue staging (16) = 128 KB per k-block, and the freed 16 KB
// A_lo slot buys a 4th pipeline stage.  Measured (profiles/r3_gemm_engine_study.md): 2-6 % faster than the SS kernel back
// to back under the power cap, tensor pipe 60-74 % active under ncu (SS: 57-71 %); what is left is the L2 -> SM fill of
// 48 KB per k-block, the epilogue's interference with the next tile's MMAs and the global stores (13 %).
//
// TMEM columns (512): main0 [0,128) | cross [128,256) | main1 [256,384) | A ring: 2 stages x (hi 32 | lo 32) [384,512).
// The main accumulator is double buffered (epilogue of tile t overlaps the main loop of tile t+1); the 2^-11-times-smaller
// cross-term accumulator is single: the epilogue reads it first and hands it back before it touches main.  Even tiles run
//   [main0 | cross] += A_hi x [B_hi ; B_lo]   (one N = 256 instruction, B tiles adjacent in shared memory)
// odd tiles  [cross | main1] += A_hi x [B_lo ; B_hi]  (the producer swaps the two B slots), then  cross += A_lo x B_hi.
//
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-9 epilogue (shared with tc_gemm.cu: tc_gemm_epi.cuh), 10-13 converters
// (thread = A row = TMEM lane).
#include "tc_gemm_epi.cuh"

namespace far {
namespace tc {

constexpr uint32_t TS_CROSS = 128, TS_MAIN1 = 256, TS_ARING = 384;

// FAR_TC_DBG bit 256: per-role cycle accounting of CTA 0 (far_tc_debug_counters): where each warp role waits.
// [0] producer wait empty [1] producer total | [2] MMA wait main acc [3] wait cross [4] wait conv [5] total |
// [6] converter wait full [7] wait afree [8] total | [9] epilogue wait tfull [10] total [11] tfull -> cross handed back |
// [12] tiles [13] k-blocks per tile
__device__ unsigned long long g_ts_prof[16];
#define TS_CLK() (prof ? clock64() : 0ll)

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

__global__ void __launch_bounds__(GEMM_THREADS_RAW, 1)
tc_gemm_ts_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
                  const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                  const __grid_constant__ CUtensorMap mapC, GemmArgs p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t stg_base = base + TS_STAGES * TS_STAGE_BYTES;
  const uint32_t bar_base = stg_base + EPI_WARPS * STG_TILE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TS_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * TS_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * TS_STAGES + 2 + s); };
  auto conv_bar = [&](int s) { return bar_base + 8u * (2 * TS_STAGES + 4 + s); };    // A stage s written (4 converter warps)
  auto afree_bar = [&](int s) { return bar_base + 8u * (2 * TS_STAGES + 6 + s); };   // MMAs reading A stage s retired
  const uint32_t cross_empty = bar_base + 8u * (2 * TS_STAGES + 8);                  // epilogue has read the cross accumulator
  const uint32_t tmem_slot = bar_base + 8u * (2 * TS_STAGES + 10);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - raw));

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_pg = (p.L + BM - 1) / BM;
  const int num_tiles = p.G * tiles_pg * tiles_n;
  const int kblocks = (p.K + BK - 1) / BK;
  const bool prof = (p.dbg & 256) && blockIdx.x == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TS_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS);
      mbar_init(conv_bar(s), 4); mbar_init(afree_bar(s), 1);
    }
    mbar_init(cross_empty, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    constexpr int PF = 8;   // L2 prefetch distance of the raw A tiles, in k-blocks (see tc_gemm.cu)
    auto prefetch_a = [&](int tile_p, int kb_p) {
      if (tile_p >= num_tiles) return;
      const int tm_p = tile_p / tiles_n;
      const int g_p = tm_p / tiles_pg, r0_p = (tm_p % tiles_pg) * BM;
      if (kb_p < p.kb1) tma_prefetch_4d(&mapA1, kb_p * BK, r0_p, g_p, 0);
      else tma_prefetch_4d(&mapA2, (kb_p - p.kb1) * BK, r0_p, g_p, 0);
    };
    if (elect_one()) {
      for (int i = 0; i < PF; ++i) prefetch_a(blockIdx.x + (i / kblocks) * gridDim.x, i % kblocks);
    }
    __syncwarp();
    int odd = 0;
    long long w_empty = 0, t_begin = TS_CLK();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, odd ^= 1) {
      const int tm = tile / tiles_n, n0 = (tile % tiles_n) * BN;
      const int g = tm / tiles_pg, r0 = (tm % tiles_pg) * BM, gb = p.b_grouped ? g : 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        if (!(p.dbg & 64) && elect_one()) {
          const int ahead = kb + PF;
          prefetch_a(tile + (ahead / kblocks) * gridDim.x, ahead % kblocks);
        }
        __syncwarp();
        const long long c0 = TS_CLK();
        mbar_wait(empty_bar(stage), phase ^ 1u);
        w_empty += TS_CLK() - c0;
        if (elect_one()) {
          const uint32_t sbase = base + stage * TS_STAGE_BYTES;
          // diagnostics (FAR_TC_DBG): 8 = skip the B loads, 16 = skip the A load
          mbar_arrive_expect_tx(full_bar(stage), ((p.dbg & 8) ? 0 : 2 * TILE_BYTES) + ((p.dbg & 16) ? 0 : TILE_BYTES));
          if (!(p.dbg & 16)) {
            if (kb < p.kb1) tma_load_4d(sbase, &mapA1, full_bar(stage), kb * BK, r0, g, 0);
            else tma_load_4d(sbase, &mapA2, full_bar(stage), (kb - p.kb1) * BK, r0, g, 0);
          }
          // even tiles [B_hi ; B_lo], odd tiles [B_lo ; B_hi]: the N = 256 instruction's cross half must land on the
          // single cross accumulator that sits between main0 and main1
          if (!(p.dbg & 8)) {
            tma_load_4d(sbase + (odd ? 2 : 1) * TILE_BYTES, &mapBhi, full_bar(stage), kb * BK, n0, gb, 0);
            tma_load_4d(sbase + (odd ? 1 : 2) * TILE_BYTES, &mapBlo, full_bar(stage), kb * BK, n0, gb, 0);
          }
        }
        __syncwarp();
        if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (prof && lane == 0) { g_ts_prof[0] = w_empty; g_ts_prof[1] = TS_CLK() - t_begin; }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0, as = 0, acc = 0;
    uint32_t phase = 0, aphase = 0, acc_phase = 0, cphase = 0;
    long long w_main = 0, w_cross = 0, w_conv = 0, t_begin = TS_CLK(), ntile = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long c0 = TS_CLK();
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);   // epilogue of tile t-2 has drained this main accumulator
      const long long c1 = TS_CLK();
      mbar_wait(cross_empty, cphase ^ 1u);          // epilogue of tile t-1 has read the cross accumulator
      w_main += c1 - c0; w_cross += TS_CLK() - c1; ++ntile;
      tc_fence_after();
      const uint32_t d_merged = tmem_base + (acc ? TS_CROSS : 0u);   // [main0 | cross] or [cross | main1]
      const uint32_t d_cross = tmem_base + TS_CROSS;
      for (int kb = 0; kb < kblocks; ++kb) {
        const long long c2 = TS_CLK();
        mbar_wait(conv_bar(as), aphase);   // A stage in TMEM (the converters waited for the TMA data: B has landed too)
        w_conv += TS_CLK() - c2;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sbase = base + stage * TS_STAGE_BYTES;
          const uint64_t dB2 = make_kmajor_sw128_desc(sbase + 1 * TILE_BYTES);                // both B tiles, 256 rows
          const uint64_t dBhi = make_kmajor_sw128_desc(sbase + (acc ? 2 : 1) * TILE_BYTES);
          const uint32_t tA_hi = tmem_base + TS_ARING + (uint32_t)(as * 64), tA_lo = tA_hi + 32u;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (p.dbg & 4) break;
            const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
            const uint32_t acol = (uint32_t)(k * UMMA_K);
            umma_tf32_ts(d_merged, tA_hi + acol, dB2 + koff, kIdescTf32N2, (kb | k) ? 1u : 0u);
            umma_tf32_ts(d_cross, tA_lo + acol, dBhi + koff, kIdescTf32, 1u);
          }
          umma_commit(empty_bar(stage));   // shared-memory stage reusable
          umma_commit(afree_bar(as));      // TMEM A stage reusable
          if (kb == kblocks - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      cphase ^= 1u;
    }
    if (prof && lane == 0) {
      g_ts_prof[2] = w_main; g_ts_prof[3] = w_cross; g_ts_prof[4] = w_conv; g_ts_prof[5] = TS_CLK() - t_begin;
      g_ts_prof[12] = ntile; g_ts_prof[13] = kblocks;
    }
    (void)phase;
  } else if (warp >= 2 + EPI_WARPS) {
    // ===================== converter warps: raw A tile (shared memory) -> hi | lo in TMEM =====================
    const int quarter = warp & 3;          // the TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;   // A row of the tile = TMEM lane
    const int sw = row & 7;
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    long long w_full = 0, w_afree = 0, t_begin = TS_CLK();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      float ln_mu = 0.f, ln_rstd = 0.f;
      if (p.ln_in != nullptr) {   // this row's LayerNorm statistics from the producer's per-block partials (Chan's update)
        const int tm = tile / tiles_n;
        const int g = tm / tiles_pg, lrow = (tm % tiles_pg) * BM + row;
        if (lrow < p.L) {
          const float2* st = p.ln_in + ((size_t)g * p.L + lrow) * p.ln_chunks;
          float n = 0.f, mean = 0.f, m2 = 0.f;
          for (int c = 0; c < p.ln_chunks; ++c) {
            const float2 pc = __ldg(st + c);
            const float tot = n + 32.f, delta = pc.x - mean;
            mean = fmaf(delta, 32.f / tot, mean);
            m2 += pc.y + delta * delta * (n * 32.f / tot);
            n = tot;
          }
          ln_mu = mean;
          ln_rstd = rsqrtf(m2 / n + p.ln_eps);
        }
      }
      for (int kb = 0; kb < kblocks; ++kb) {
        const long long c0 = TS_CLK();
        mbar_wait(full_bar(stage), phase);
        w_full += TS_CLK() - c0;
        const float4* arow = reinterpret_cast<const float4*>(smem_dyn + (base + stage * TS_STAGE_BYTES - raw) + row * 128);
        uint32_t hi[32], lo[32];
        const bool ln_here = p.ln_in != nullptr && kb >= p.kb1;   // second A segment: LayerNorm applied on the fly
        const float4* lg = reinterpret_cast<const float4*>(p.ln_gamma) + (ln_here ? (kb - p.kb1) * 8 : 0);
        const float4* lb = reinterpret_cast<const float4*>(p.ln_beta) + (ln_here ? (kb - p.kb1) * 8 : 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // 16-byte chunk j of row r sits at chunk j ^ (r & 7) (SWIZZLE_128B)
          const float4 v = arow[j ^ sw];
          float x[4] = {v.x, v.y, v.z, v.w};
          if (ln_here) {
            const float4 gm = __ldg(lg + j), bt = __ldg(lb + j);
            x[0] = fmaf((x[0] - ln_mu) * ln_rstd, gm.x, bt.x); x[1] = fmaf((x[1] - ln_mu) * ln_rstd, gm.y, bt.y);
            x[2] = fmaf((x[2] - ln_mu) * ln_rstd, gm.z, bt.z); x[3] = fmaf((x[3] - ln_mu) * ln_rstd, gm.w, bt.w);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t h = __float_as_uint(x[e]) & 0xFFFFE000u;
            hi[4 * j + e] = h;
            lo[4 * j + e] = __float_as_uint(x[e] - __uint_as_float(h));
          }
        }
        const long long c1 = TS_CLK();
        mbar_wait(afree_bar(as), aphase ^ 1u);   // the MMAs that read this TMEM stage two k-blocks ago have retired
        w_afree += TS_CLK() - c1;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + TS_ARING + (uint32_t)(as * 64);
        if (!(p.dbg & 32)) {
          tmem_st32(taddr, hi);
          tmem_st32(taddr + 32u, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(conv_bar(as));
        if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
    if (prof && warp == 2 + EPI_WARPS && lane == 0) { g_ts_prof[6] = w_full; g_ts_prof[7] = w_afree; g_ts_prof[8] = TS_CLK() - t_begin; }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int chalf = ew >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int actc = p.act_cols < 0 ? p.N : p.act_cols;
    const uint32_t stg_addr = stg_base + ew * STG_TILE;
    float4* srow = reinterpret_cast<float4*>(smem_dyn + (stg_addr - raw) + lane * 128);
    const int sx = lane & 7;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    long long w_tfull = 0, w_handoff = 0, t_begin = TS_CLK();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tm = tile / tiles_n, n0 = (tile % tiles_n) * BN;
      const int g = tm / tiles_pg, r0 = (tm % tiles_pg) * BM;
      const long long c0 = TS_CLK();
      mbar_wait(tfull_bar(acc), acc_phase);
      const long long c1 = TS_CLK();
      w_tfull += c1 - c0;
      tc_fence_after();
      const bool live = !(p.dbg & 2);
      // the single cross accumulator first: both of this warp's column blocks into registers, then hand it back so the
      // next tile's MMAs can start while the main accumulator is drained
      uint32_t vs0[32], vs1[32];
      const uint32_t tcross = tmem_base + lane_off + TS_CROSS + (uint32_t)(chalf * 64);
      tmem_ld32_nowait(tcross, vs0);
      tmem_ld32_nowait(tcross + 32u, vs1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(cross_empty);
      w_handoff += TS_CLK() - c1;
      const uint32_t tmain = tmem_base + lane_off + (acc ? TS_MAIN1 : 0u) + (uint32_t)(chalf * 64);
      {
        const int col0 = n0 + chalf * 64;
        uint32_t v[32];
        tmem_ld32_nowait(tmain, v);
        tmem_ld_wait();
        if (live && col0 < p.N) {
          float t[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(v[e]) + __uint_as_float(vs0[e]);
          epi_block(t, p, &mapC, col0, g, r0, quarter, lane, actc, srow, sx, stg_addr);
        }
      }
      {
        const int col0 = n0 + chalf * 64 + 32;
        uint32_t v[32];
        tmem_ld32_nowait(tmain + 32u, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));   // main accumulator back to the MMA warp
        if (live && col0 < p.N) {
          float t[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(v[e]) + __uint_as_float(vs1[e]);
          epi_block(t, p, &mapC, col0, g, r0, quarter, lane, actc, srow, sx, stg_addr);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (prof && warp == 2 && lane == 0) { g_ts_prof[9] = w_tfull; g_ts_prof[10] = TS_CLK() - t_begin; g_ts_prof[11] = w_handoff; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

void launch_gemm_ts(int grid, cudaStream_t st, const CUtensorMap& mA1, const CUtensorMap& mA2, const CUtensorMap& mBhi,
                    const CUtensorMap& mBlo, const CUtensorMap& mC, const GemmArgs& p) {
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set))
    cudaFuncSetAttribute(tc_gemm_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_TS_SMEM);
  tc_gemm_ts_kernel<<<grid, GEMM_THREADS_RAW, GEMM_TS_SMEM, st>>>(mA1, mA2, mBhi, mBlo, mC, p);
}

}  // namespace tc
}  // namespace far

// diagnostics: the per-role cycle counters of the last tc_gemm_ts_kernel (kernel = 0) / tc_gemm_pair_kernel (1) launch
// run with FAR_TC_DBG bit 256
extern "C" int far_tc_debug_counters(int kernel, unsigned long long* out16) {
  if (kernel == 1) return far::tc::gemm_pair_debug_counters(out16);
  if (kernel == 2) return far::tc::gemm_pair_debug_trace(out16);   // 256 entries: timeline of one tile
  if (kernel == 3) return far::tc::gemm_pair_debug_cta(out16);     // 320 entries: (cycles, tiles) per CTA
  return cudaMemcpyFromSymbol(out16, far::tc::g_ts_prof, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -1;
}
