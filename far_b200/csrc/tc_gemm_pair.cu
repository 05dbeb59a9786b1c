// tcgen05 3xTF32 GEMM on CTA PAIRS (cta_group::2), A operand in tensor memory:  C = act(A B^T + bias), raw fp32 activations.
//
// Why pairs: the per-role cycle accounting of the one-CTA kernels (benchmarks/bench_gemm_roles.py, profiles/r3_gemm_roles.md)
// shows the TMA producer busy ~1030 of every ~1130 cycles per k-block and never waiting for a free stage: the kernels
// are bound by the L2 -> SM fill rate (48 KB per 128x128x32 k-block = ~45 B/clk/SM, the chip-wide L2 output cap divided
// by 148 SMs), two thirds of which is the B operand (hi and lo weight tiles re-read from L2 for every 128 rows).  A CTA pair
// (two SMs of one TPC, one 256 x 128 output tile) shares B: each CTA loads only its 64 of the 128 B rows, so a k-block
// costs 16 KB (A) + 16 KB (B halves) per SM instead of 48 KB; the tensor cores of both SMs run one M = 256 instruction
// issued by the leader CTA.
//
// Per CTA (rank r of the pair): rows [r0 + 128 r, +128) of the tile.
//   warp 0        TMA producer: raw A tile (16 KB) + its halves of B_hi / B_lo (8 KB each) into a 6-stage ring
//   warps 10-13   converters : landed A tile -> hi | lo in TMEM (tcgen05.st, thread = row = TMEM lane); arrive on the
//                              LEADER's conv barrier (remote mbarrier arrive): "A in TMEM and B in smem, in this CTA"
//   warp 1        MMA issuer (leader only): per 8-wide k-step three M256 x N128 instructions, A from TMEM, B from both
//                 CTAs' shared memory:  main += A_hi B_hi,  cross += A_hi B_lo,  cross += A_lo B_hi;
//                 tcgen05.commit multicast to both CTAs releases the smem stage / the TMEM A stage / publishes the tile
//   warps 2-9     epilogue (tc_gemm_epi.cuh), each CTA its own 128 rows; hands the accumulators back to the leader
// TMEM columns (512, same in both CTAs): main0 [0,128) | main1 [128,256) | cross [256,384) | A ring 4 x (hi 16 | lo 16):
// 16-wide sub-blocks, so three sub-blocks of MMAs (1152 cycles) cover the retire -> convert -> issue round trip of a stage.
#include "tc_gemm_epi.cuh"

namespace far {
namespace tc {

constexpr uint32_t TP_MAIN1 = 128, TP_CROSS = 256, TP_ARING = 384;
constexpr int TP_SUBK = 16;                         // K extent of one TMEM A stage (half a k-block)
constexpr int TP_ASTAGES = 4;                       // TMEM A ring: 4 x (hi 16 | lo 16) columns
constexpr int TP_BROWS = BN / 2;                    // B rows per CTA
constexpr int TP_BTILE = TP_BROWS * BK * 4;         // 8 KiB
constexpr int TP_STAGE_BYTES = TILE_BYTES + 2 * TP_BTILE;  // 32 KiB
constexpr uint32_t kIdescTf32M256 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

__device__ unsigned long long g_tp_prof[16];
// timeline (clock64) of CTA 0's 4th tile, FAR_TC_DBG bit 256: [0] MMA tile start, [1] cross_empty seen, [2 + 2 sb] conv
// seen / [3 + 2 sb] issued, per sub-block (sb < 32); [80 + 2 j] converter: afree seen / [81 + 2 j] conv arrive, per
// sub-block; [160] epilogue tfull seen, [161] cross handed back, [162] main handed back, [163] epilogue end,
// [164..167] the same for the previous tile
__device__ unsigned long long g_tp_trace[256];
// FAR_TC_DBG bit 1024: total cycles of every CTA (two clock reads per CTA, nothing in the loops): g_tp_cta[2 b] = cycles,
// [2 b + 1] = tiles of CTA b
__device__ unsigned long long g_tp_cta[2 * 160];
#define TP_TRACE(i) do { if (profh && trace_on) g_tp_trace[(i)] = clock64(); } while (0)
#define TP_CLK() (prof ? clock64() : 0ll)     // per-tile sites (bit 256)
#define TP_CLKH() (profh ? clock64() : 0ll)   // per-k-block sites (bits 256 + 2048): these perturb the loops they measure

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// (default .release.cta semantics, as CUTLASS' ClusterBarrier::arrive(cta_id): a .release.cluster arrive / .acquire.cluster
// wait makes ptxas emit cluster-scope fences that invalidate L1 -- measured 1700 cycles per converted k-block)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier that also receives arrivals from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// completion of all prior tcgen05.mma of this thread -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st32p(uint32_t taddr, const uint32_t (&v)[32]) {
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

__device__ __forceinline__ void tmem_st16p(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS_RAW, 1)
tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
                    const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                    const __grid_constant__ CUtensorMap mapC, GemmArgs p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t stg_base = base + TP_STAGES * TP_STAGE_BYTES;
  const uint32_t bar_base = stg_base + EPI_WARPS * STG_TILE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                          // own TMA data landed
  auto empty_bar = [&](int s) { return bar_base + 8u * (TP_STAGES + s); };           // leader's MMAs on stage s retired
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * TP_STAGES + s); };       // tile accumulated (multicast)
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * TP_STAGES + 2 + s); };  // leader: both epilogues drained main s
  auto conv_bar = [&](int s) { return bar_base + 8u * (2 * TP_STAGES + 4 + s); };    // leader: A stage s written in both CTAs
  auto afree_bar = [&](int s) { return bar_base + 8u * (2 * TP_STAGES + 4 + TP_ASTAGES + s); };   // MMAs reading A stage s retired (multicast)
  const uint32_t cross_empty = bar_base + 8u * (2 * TP_STAGES + 4 + 2 * TP_ASTAGES);   // leader: both epilogues read cross
  const uint32_t tmem_slot = bar_base + 8u * (2 * TP_STAGES + 6 + 2 * TP_ASTAGES);
  static_assert(8 * (2 * TP_STAGES + 7 + 2 * TP_ASTAGES) <= 256, "barrier area");
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - raw));

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_pg = (p.L + 2 * BM - 1) / (2 * BM);   // 256-row pair tiles per group
  const int num_tiles = p.G * tiles_pg * tiles_n;
  const int kblocks = (p.K + BK - 1) / BK;
  const bool prof = (p.dbg & 256) && blockIdx.x == 0;
  const bool profh = prof && (p.dbg & 2048);
  const long long cta_t0 = (p.dbg & 1024) ? clock64() : 0ll;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TP_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * EPI_WARPS); }
    for (int s = 0; s < TP_ASTAGES; ++s) { mbar_init(conv_bar(s), 8); mbar_init(afree_bar(s), 1); }
    mbar_init(cross_empty, 2 * EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the same warp of both CTAs allocates (and later frees) the pair's tensor memory
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the allocation visible in both CTAs before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    // L2 prefetch of the raw A tiles PF k-blocks ahead (diagnostics bit 8192 enables it; off by default:
    // cp.async.bulk.prefetch.tensor blocks the issuing thread ~650 cycles per k-block when the TMA queue is busy)
    constexpr int PF = 8;
    auto prefetch_a = [&](int tile_p, int kb_p) {
      if (tile_p >= num_tiles) return;
      const int tm_p = tile_p / tiles_n;
      const int g_p = tm_p / tiles_pg, r0_p = (tm_p % tiles_pg) * 2 * BM + (int)rank * BM;
      if (r0_p >= p.L) return;
      if (kb_p < p.kb1) tma_prefetch_4d(&mapA1, kb_p * BK, r0_p, g_p, 0);
      else tma_prefetch_4d(&mapA2, (kb_p - p.kb1) * BK, r0_p, g_p, 0);
    };
    long long w_empty = 0, w_pref = 0, w_issue = 0, t_begin = TP_CLKH();
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int tm = tile / tiles_n, n0 = (tile % tiles_n) * BN;
      const int g = tm / tiles_pg, r0 = (tm % tiles_pg) * 2 * BM + (int)rank * BM, gb = p.b_grouped ? g : 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        if (p.dbg & 8192) {
          if (elect_one()) {
            const int ahead = kb + PF;
            prefetch_a(tile + (ahead / kblocks) * num_pairs, ahead % kblocks);
          }
          __syncwarp();
        }
        const long long c0 = TP_CLKH();
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const long long ci = TP_CLKH();
        w_empty += ci - c0;
        if (elect_one()) {
          const uint32_t sbase = base + stage * TP_STAGE_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), TP_STAGE_BYTES);
          // rows past L (the pair tile's second half on a ragged group tail) are zero-filled by TMA
          if (kb < p.kb1) tma_load_4d(sbase, &mapA1, full_bar(stage), kb * BK, r0, g, 0);
          else tma_load_4d(sbase, &mapA2, full_bar(stage), (kb - p.kb1) * BK, r0, g, 0);
          tma_load_4d(sbase + TILE_BYTES, &mapBhi, full_bar(stage), kb * BK, n0 + (int)rank * TP_BROWS, gb, 0);
          tma_load_4d(sbase + TILE_BYTES + TP_BTILE, &mapBlo, full_bar(stage), kb * BK, n0 + (int)rank * TP_BROWS, gb, 0);
        }
        __syncwarp();
        w_issue += TP_CLKH() - ci;
        if (++stage == TP_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (prof && lane == 0) { g_tp_prof[0] = w_empty; g_tp_prof[1] = TP_CLKH() - t_begin; g_tp_prof[14] = w_pref; g_tp_prof[15] = w_issue; }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      int stage = 0, as = 0, acc = 0;
      uint32_t aphase = 0, acc_phase = 0, cphase = 0;
      long long w_main = 0, w_cross = 0, w_conv = 0, t_begin = TP_CLK(), ntile = 0;
      // MMAs of one 16-wide sub-block sb = 2 kb + h (h = which half of the 32-wide k-block): part 1 = main += A_hi B_hi,
      // part 2 = cross += A_hi B_lo + A_lo B_hi (+ the commits)
      auto issue = [&](int sb, int stage, int as, int acc, bool part1, bool part2) {
        if (elect_one()) {
          const int h = sb & 1;
          const uint32_t sbase = base + stage * TP_STAGE_BYTES;
          const uint64_t dBhi = make_kmajor_sw128_desc(sbase + TILE_BYTES);
          const uint64_t dBlo = make_kmajor_sw128_desc(sbase + TILE_BYTES + TP_BTILE);
          const uint32_t tA_hi = tmem_base + TP_ARING + (uint32_t)(as * 32), tA_lo = tA_hi + 16u;
          const uint32_t d_main = tmem_base + (acc ? TP_MAIN1 : 0u);
          const uint32_t d_cross = tmem_base + TP_CROSS;
#pragma unroll
          for (int k = 0; k < TP_SUBK / UMMA_K; ++k) {
            if (p.dbg & 4) break;
            const uint64_t koff = (uint64_t)(((h * TP_SUBK + k * UMMA_K) * 4) >> 4);   // inside the 128-byte swizzle row
            const uint32_t acol = (uint32_t)(k * UMMA_K);
            const uint32_t first = (sb | k) ? 1u : 0u;
            if (part1) umma_tf32_ts_pair(d_main, tA_hi + acol, dBhi + koff, kIdescTf32M256, first);   // main  += A_hi B_hi
            if (part2) {
              umma_tf32_ts_pair(d_cross, tA_hi + acol, dBlo + koff, kIdescTf32M256, first);           // cross += A_hi B_lo
              umma_tf32_ts_pair(d_cross, tA_lo + acol, dBhi + koff, kIdescTf32M256, 1u);              // cross += A_lo B_hi
            }
          }
          if (part2) {
            umma_commit_pair(afree_bar(as));                  // TMEM A stage reusable (both CTAs)
            if (h == 1) umma_commit_pair(empty_bar(stage));   // shared-memory stage reusable (both CTAs)
            if (sb == 2 * kblocks - 1) umma_commit_pair(tfull_bar(acc));
          }
        }
        __syncwarp();
      };
      auto wait_conv = [&](int as, uint32_t aphase) {
        const long long c2 = TP_CLKH();
        mbar_wait_cluster(conv_bar(as), aphase);   // A sub-block in TMEM and B halves in shared memory, in both CTAs
        w_conv += TP_CLKH() - c2;
        tc_fence_after();
      };
      const int nsub = 2 * kblocks;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const long long c0 = TP_CLK();
        mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u);   // both epilogues of tile t-2 drained this main accumulator
        w_main += TP_CLK() - c0; ++ntile;
        const bool trace_on = ntile == 4 && lane == 0;
        TP_TRACE(0);
        tc_fence_after();
        // The single cross accumulator is still being read by the epilogues of tile t-1 (a ~1000-cycle hand-off after the
        // last MMA retires): the first four sub-blocks (= all TMEM A stages) issue their main products first, then wait
        // for the cross accumulator and catch up with their cross products.
        const int nb = nsub < TP_ASTAGES ? nsub : TP_ASTAGES;
        int st2[TP_ASTAGES], as2[TP_ASTAGES];
        for (int i = 0; i < nb; ++i) {
          wait_conv(as, aphase);
          if (i < 32) TP_TRACE(2 + 2 * i);
          issue(i, stage, as, acc, true, false);
          st2[i] = stage; as2[i] = as;
          if (i & 1) { if (++stage == TP_STAGES) stage = 0; }
          if (++as == TP_ASTAGES) { as = 0; aphase ^= 1u; }
        }
        const long long c1 = TP_CLK();
        mbar_wait_cluster(cross_empty, cphase ^ 1u);          // both epilogues of tile t-1 read the cross accumulator
        w_cross += TP_CLK() - c1;
        TP_TRACE(1);
        tc_fence_after();
        for (int i = 0; i < nb; ++i) { issue(i, st2[i], as2[i], acc, false, true); if (i < 32) TP_TRACE(3 + 2 * i); }
        for (int sb = nb; sb < nsub; ++sb) {
          wait_conv(as, aphase);
          if (sb < 32) TP_TRACE(2 + 2 * sb);
          issue(sb, stage, as, acc, true, true);
          if (sb < 32) TP_TRACE(3 + 2 * sb);
          if (sb & 1) { if (++stage == TP_STAGES) stage = 0; }
          if (++as == TP_ASTAGES) { as = 0; aphase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        cphase ^= 1u;
      }
      if (prof && lane == 0) {
        g_tp_prof[2] = w_main; g_tp_prof[3] = w_cross; g_tp_prof[4] = w_conv; g_tp_prof[5] = TP_CLK() - t_begin;
        g_tp_prof[12] = ntile; g_tp_prof[13] = kblocks;
      }
    }
  } else if (warp >= 2 + EPI_WARPS) {
    // ===================== converter warps: raw A tile (shared memory) -> hi | lo in TMEM =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int sw = row & 7;
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    const uint32_t conv_leader = map_to_cta(conv_bar(0), 0);
    long long w_full = 0, w_afree = 0, t_begin = TP_CLKH();
    int ctile = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      ++ctile;
      const bool trace_on = ctile == 4 && warp == 2 + EPI_WARPS && lane == 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        const long long c0 = TP_CLKH();
        mbar_wait(full_bar(stage), phase);
        w_full += TP_CLKH() - c0;
        const float4* arow = reinterpret_cast<const float4*>(smem_dyn + (base + stage * TP_STAGE_BYTES - raw) + row * 128);
#pragma unroll
        for (int h = 0; h < 2; ++h) {     // two 16-wide sub-blocks per k-block, one TMEM A stage (hi 16 | lo 16) each
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {   // 16-byte chunk c of row r sits at chunk c ^ (r & 7) (SWIZZLE_128B)
            const float4 v = arow[(4 * h + j) ^ sw];
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t t = __float_as_uint(x[e]) & 0xFFFFE000u;
              hi[4 * j + e] = t;
              lo[4 * j + e] = __float_as_uint(x[e] - __uint_as_float(t));
            }
          }
          const long long c1 = TP_CLKH();
          mbar_wait(afree_bar(as), aphase ^ 1u);   // the MMAs that read this TMEM stage four sub-blocks ago have retired
          w_afree += TP_CLKH() - c1;
          if (2 * kb + h < 32) TP_TRACE(80 + 2 * (2 * kb + h));
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + TP_ARING + (uint32_t)(as * 32);
          if (!(p.dbg & 32)) {
            tmem_st16p(taddr, hi);
            tmem_st16p(taddr + 16u, lo);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(conv_leader + 8u * as);
          if (2 * kb + h < 32) TP_TRACE(81 + 2 * (2 * kb + h));
          if (++as == TP_ASTAGES) { as = 0; aphase ^= 1u; }
        }
        if (++stage == TP_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (prof && warp == 2 + EPI_WARPS && lane == 0) { g_tp_prof[6] = w_full; g_tp_prof[7] = w_afree; g_tp_prof[8] = TP_CLKH() - t_begin; }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs: each its own 128 rows) =====================
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
    const uint32_t cross_leader = map_to_cta(cross_empty, 0);
    const uint32_t tempty_leader0 = map_to_cta(tempty_bar(0), 0), tempty_leader1 = map_to_cta(tempty_bar(1), 0);
    long long w_tfull = 0, w_handoff = 0, t_begin = TP_CLK();
    int etile = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int tm = tile / tiles_n, n0 = (tile % tiles_n) * BN;
      const int g = tm / tiles_pg, r0 = (tm % tiles_pg) * 2 * BM + (int)rank * BM;
      const long long c0 = TP_CLK();
      mbar_wait(tfull_bar(acc), acc_phase);
      const long long c1 = TP_CLK();
      w_tfull += c1 - c0;
      ++etile;
      const bool trace_on = (etile == 4 || etile == 3) && warp == 2 && lane == 0;
      const int tb = etile == 4 ? 160 : 164;
      TP_TRACE(tb);
      tc_fence_after();
      const bool live = !(p.dbg & 2) && r0 < p.L;
      // the single cross accumulator first, handed back before main is touched (see tc_gemm_ts.cu)
      uint32_t vs0[32], vs1[32];
      const uint32_t tcross = tmem_base + lane_off + TP_CROSS + (uint32_t)(chalf * 64);
      if (!(p.dbg & 512)) {   // diagnostics: 512 = skip the epilogue's TMEM loads
        tmem_ld32_nowait(tcross, vs0);
        tmem_ld32_nowait(tcross + 32u, vs1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(cross_leader);
      w_handoff += TP_CLK() - c1;
      TP_TRACE(tb + 1);
      const uint32_t tmain = tmem_base + lane_off + (acc ? TP_MAIN1 : 0u) + (uint32_t)(chalf * 64);
      {
        const int col0 = n0 + chalf * 64;
        uint32_t v[32];
        if (!(p.dbg & 512)) {
          tmem_ld32_nowait(tmain, v);
          tmem_ld_wait();
        }
        TP_TRACE(tb + 4);   // first main block in registers
        if (live && col0 < p.N) {
          float t[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(v[e]) + __uint_as_float(vs0[e]);
          epi_block(t, p, &mapC, col0, g, r0, quarter, lane, actc, srow, sx, stg_addr,
                    (profh && trace_on && tb == 160) ? g_tp_trace + 170 : nullptr);
        }
      }
      {
        const int col0 = n0 + chalf * 64 + 32;
        uint32_t v[32];
        if (!(p.dbg & 512)) {
          tmem_ld32_nowait(tmain + 32u, v);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);   // main accumulator back to the leader
        TP_TRACE(tb + 2);
        if (live && col0 < p.N) {
          float t[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(v[e]) + __uint_as_float(vs1[e]);
          epi_block(t, p, &mapC, col0, g, r0, quarter, lane, actc, srow, sx, stg_addr,
                    (profh && trace_on && tb == 160) ? g_tp_trace + 174 : nullptr);
        }
      }
      TP_TRACE(tb + 3);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (prof && warp == 2 && lane == 0) { g_tp_prof[9] = w_tfull; g_tp_prof[10] = TP_CLK() - t_begin; g_tp_prof[11] = w_handoff; }
  }

  // neither CTA may exit (or free tensor memory) while the other can still address its barriers / shared memory / TMEM
  tc_fence_before();
  if ((p.dbg & 1024) && threadIdx.x == 0 && blockIdx.x < 160) {
    g_tp_cta[2 * blockIdx.x] = (unsigned long long)(clock64() - cta_t0);
    g_tp_cta[2 * blockIdx.x + 1] = (unsigned long long)((num_tiles - pair + num_pairs - 1) / num_pairs);
  }
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int gemm_pair_max_clusters() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    cudaFuncSetAttribute(tc_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_PAIR_SMEM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs, 1, 1);
    cfg.blockDim = dim3(GEMM_THREADS_RAW, 1, 1);
    cfg.dynamicSmemBytes = GEMM_PAIR_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, tc_gemm_pair_kernel, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = -1; }
    cached[dev] = n;
  }
  return cached[dev] > 0 ? cached[dev] : 0;
}

void launch_gemm_pair(int clusters, cudaStream_t st, const CUtensorMap& mA1, const CUtensorMap& mA2, const CUtensorMap& mBhi,
                      const CUtensorMap& mBlo, const CUtensorMap& mC, const GemmArgs& p) {
  tc_gemm_pair_kernel<<<2 * clusters, GEMM_THREADS_RAW, GEMM_PAIR_SMEM, st>>>(mA1, mA2, mBhi, mBlo, mC, p);
}

int gemm_pair_debug_counters(unsigned long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_tp_prof, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -1;
}
int gemm_pair_debug_cta(unsigned long long* out320) {
  return cudaMemcpyFromSymbol(out320, g_tp_cta, sizeof(unsigned long long) * 320) == cudaSuccess ? 0 : -1;
}
int gemm_pair_debug_trace(unsigned long long* out256) {
  return cudaMemcpyFromSymbol(out256, g_tp_trace, sizeof(unsigned long long) * 256) == cudaSuccess ? 0 : -1;
}

}  // namespace tc
}  // namespace far

