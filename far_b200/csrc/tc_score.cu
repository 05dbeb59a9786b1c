// Dual-softmax score passes on the tcgen05 3xTF32 pipeline (same TMA -> UMMA -> TMEM structure as tc_gemm.cu):
//   MODE_LSE  : S tile -> per-row / per-column (max, sumexp) partials           (pass A of score.cu)
//   MODE_CONF : S tile -> conf = exp(s-rowlse) exp(s-collse), per-row (max, argmax) and per-column max partials,
//               optional dense conf output                                       (pass B of coarse_match.cu)
// Outputs have exactly the layout of the CUDA-core kernels (score_lse_kernel / match_conf_kernel), so the merge /
// decide / gather kernels downstream are shared.  In the TMEM accumulator a thread owns one ROW of the tile
// (tcgen05.ld 32x32b), so row reductions are thread-local; column reductions go through a 128 x 33 shared-memory
// transpose (bank-conflict free) synchronised with a named barrier among the 128 epilogue threads.
#include "score.cuh"
#include "tc_common.cuh"
#include "tc_score.cuh"
#include "tc_gemm.cuh"

namespace far {
namespace tc {

constexpr int MODE_LSE = 0, MODE_CONF = 1;
constexpr int S_STAGES = 2;                 // 2 x 64 KiB operand ring leaves room for a full-tile transpose buffer
constexpr int S_THREADS = 64 + 256;         // producer warp, MMA warp, 8 epilogue warps
constexpr int SPITCH = BN + 1;              // 129: conflict-free row writes and column reads
constexpr size_t SCORE_SMEM = 1024 + (size_t)S_STAGES * STAGE_BYTES + 256 + (size_t)BM * SPITCH * 4 + BN * 4;
constexpr float kLog2e = 1.4426950408889634f;

struct ScoreTcArgs {
  int G, H, L, S, K;       // groups (= nb * H), heads per batch, rows of A, rows of B, contraction
  float scale2;            // scale * log2(e): scores are handled in the base-2 domain (one MUFU.EX2 per exponential)
  // MODE_LSE outputs (base-2 (max, sum 2^(x-max)) partials; 2 partial "tiles" per 128-wide tile in each direction)
  float2* rowpart;         // [(g*2JT + 2jt+half)*L + i]
  float2* colpart;         // [(g*2IT + 2it+half)*S + j]
  // MODE_CONF inputs / outputs
  const float* rowlse;     // [G][L] natural log
  const float* collse;     // [G][S]
  float2* rowmax;          // [(g*2JT + 2jt+half)*L + i] (conf, bits(j))
  float* colmax;           // [(g*2IT + 2it+half)*S + j]
  float* conf_out;         // optional [G][L][S]
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld32_nowait2(uint32_t taddr, uint32_t (&v)[32]) {
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

template <int MODE>
__global__ void __launch_bounds__(S_THREADS, 1)
tc_score_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, ScoreTcArgs p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + S_STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * S_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * S_STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S_STAGES + 4);
  unsigned char* gen = smem_dyn + (bar_base + 256 - raw);  // generic pointer to the epilogue scratch
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - raw));
  float(*scr)[SPITCH] = reinterpret_cast<float(*)[SPITCH]>(gen);
  float* cls2 = reinterpret_cast<float*>(gen + (size_t)BM * SPITCH * 4);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int IT = (p.L + BM - 1) / BM, JT = (p.S + BN - 1) / BN;
  const int num_tiles = p.G * IT * JT;
  const int kblocks = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    {  // whole warp in the loop; one elected lane issues the TMA instructions
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int jt = tile % JT, it = (tile / JT) % IT, g = tile / (JT * IT);
        const int g0 = g % p.H, g1 = g / p.H;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (elect_one()) {
            const uint32_t sbase = base + stage * STAGE_BYTES;
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_4d(sbase + 0 * TILE_BYTES, &mapAhi, full_bar(stage), kb * BK, it * BM, g0, g1);
            tma_load_4d(sbase + 1 * TILE_BYTES, &mapAlo, full_bar(stage), kb * BK, it * BM, g0, g1);
            tma_load_4d(sbase + 2 * TILE_BYTES, &mapBhi, full_bar(stage), kb * BK, jt * BN, g0, g1);
            tma_load_4d(sbase + 3 * TILE_BYTES, &mapBlo, full_bar(stage), kb * BK, jt * BN, g0, g1);
          }
          __syncwarp();
          if (++stage == S_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {  // whole warp in the loop, one elected lane issues the MMAs (tc_common.cuh: elect_one)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_main = tmem_base + (uint32_t)(acc * 2 * BN);
        const uint32_t tmem_small = tmem_main + (uint32_t)BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sbase = base + stage * STAGE_BYTES;
            const uint64_t dAhi = make_kmajor_sw128_desc(sbase + 0 * TILE_BYTES);
            const uint64_t dAlo = make_kmajor_sw128_desc(sbase + 1 * TILE_BYTES);
            const uint64_t dBhi = make_kmajor_sw128_desc(sbase + 2 * TILE_BYTES);
            const uint64_t dBlo = make_kmajor_sw128_desc(sbase + 3 * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
              umma_tf32(tmem_main, dAhi + koff, dBhi + koff, kIdescTf32N2, (kb | k) ? 1u : 0u);  // [main|cross] += Ahi [Bhi;Blo]
              umma_tf32(tmem_small, dAlo + koff, dBhi + koff, kIdescTf32, 1u);                    // cross += Alo Bhi
              (void)dBlo;
            }
            umma_commit(empty_bar(stage));
            if (kb == kblocks - 1) umma_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == S_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: 8 warps; thread = (tile row, 64-column half) =====================
    const int quarter = warp & 3;              // TMEM lane window
    const int half = (warp - 2) >> 2;          // which 64 columns of the 128-wide tile
    const int row = quarter * 32 + lane;
    const int et = (warp - 2) * 32 + lane;     // 0..255
    const int ccol = et & 127, crh = et >> 7;  // column phase: column ccol, rows [64*crh, 64*crh + 64)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int jt = tile % JT, it = (tile / JT) % IT, g = tile / (JT * IT);
      const int i0 = it * BM, j0 = jt * BN;
      const int grow = i0 + row;
      const bool rvalid = grow < p.L;
      float rl2 = 0.f;
      if (MODE == MODE_CONF) {
        rl2 = rvalid ? p.rowlse[(size_t)g * p.L + grow] * kLog2e : 0.f;
        if (et < BN) cls2[et] = (j0 + et < p.S) ? p.collse[(size_t)g * p.S + j0 + et] * kLog2e : 0.f;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      float x[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32], vs[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * BN + half * 64 + c * 32);
        tmem_ld32_nowait2(taddr, v);
        tmem_ld32_nowait2(taddr + (uint32_t)BN, vs);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 32; ++e) x[c * 32 + e] = (__uint_as_float(v[e]) + __uint_as_float(vs[e])) * p.scale2;
      }
      // the accumulator is in registers now: hand the TMEM buffer back to the MMA warp early
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      const int col0 = j0 + half * 64;
      if (MODE == MODE_LSE) {
        float m = -INFINITY;
        if (rvalid && col0 + 64 <= p.S) {  // interior: no masking
#pragma unroll
          for (int e = 0; e < 64; ++e) m = fmaxf(m, x[e]);
        } else {
#pragma unroll
          for (int e = 0; e < 64; ++e) {
            if (!(rvalid && col0 + e < p.S)) x[e] = -INFINITY;
            m = fmaxf(m, x[e]);
          }
        }
        float sum = 0.f;
        if (m > -INFINITY) {
#pragma unroll
          for (int e = 0; e < 64; ++e) sum += ex2(x[e] - m);  // 2^(-inf) = 0 for masked entries
        }
        if (rvalid) p.rowpart[((size_t)g * 2 * JT + 2 * jt + half) * p.L + grow] = make_float2(m, sum);
#pragma unroll
        for (int e = 0; e < 64; ++e) scr[row][half * 64 + e] = x[e];
        epi_bar();
        {
          float cm = -INFINITY;
#pragma unroll 16
          for (int r = 0; r < 64; ++r) cm = fmaxf(cm, scr[crh * 64 + r][ccol]);
          float cs = 0.f;
          if (cm > -INFINITY) {
#pragma unroll 16
            for (int r = 0; r < 64; ++r) cs += ex2(scr[crh * 64 + r][ccol] - cm);
          }
          const int col = j0 + ccol;
          if (col < p.S) p.colpart[((size_t)g * 2 * IT + 2 * it + crh) * p.S + col] = make_float2(cm, cs);
        }
        epi_bar();
      } else {
        epi_bar();  // cls2 staged
        float bv = -1.f;
        int bj = 0x7fffffff;
#pragma unroll
        for (int e = 0; e < 64; ++e) {
          const int col = col0 + e;
          const bool ok = rvalid && col < p.S;
          x[e] = ok ? ex2(x[e] - rl2) * ex2(x[e] - cls2[half * 64 + e]) : -1.f;   // softmax(sim,1) * softmax(sim,2)
          if (x[e] > bv) { bv = x[e]; bj = col; }   // ascending columns, strict > keeps the lowest j on ties
        }
        if (rvalid) p.rowmax[((size_t)g * 2 * JT + 2 * jt + half) * p.L + grow] = make_float2(bv, __int_as_float(bj));
        if (p.conf_out != nullptr && rvalid) {
          float* dst = p.conf_out + ((size_t)g * p.L + grow) * p.S + col0;
#pragma unroll
          for (int e = 0; e < 64; ++e)
            if (col0 + e < p.S) dst[e] = x[e];
        }
#pragma unroll
        for (int e = 0; e < 64; ++e) scr[row][half * 64 + e] = x[e];
        epi_bar();
        {
          float cm = -1.f;
#pragma unroll 16
          for (int r = 0; r < 64; ++r) cm = fmaxf(cm, scr[crh * 64 + r][ccol]);
          const int col = j0 + ccol;
          if (col < p.S) p.colmax[((size_t)g * 2 * IT + 2 * it + crh) * p.S + col] = cm;
        }
        epi_bar();
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------------------------
// LSE pass for K = 64 (the EMM head dimension): query-tile-resident streaming kernel.
// tc_score_kernel gives every 128x128 tile the whole 2-stage ring when K = 64 (2 k-blocks), so the next tile's loads
// cannot start before the current tile's MMAs retire (31 % tensor-pipe activity, ncu).  Here a CTA owns one 128-query
// tile of one (batch, head): Q (hi | lo, 64 KB) is loaded once, the 128-key K tiles stream through a 2-deep ring of
// whole tiles, S is double-buffered in TMEM (2 x (main | cross) x 128 columns).  Epilogue warps own rows: the row
// (max, sum) is carried online in registers across all key tiles (one pair per row and 64-column half at the end);
// column statistics are taken per warp (32 rows) through a warp-private 32x33 transpose -- no block-level barrier
// anywhere in the epilogue -- and written as 4 partials per query tile.
constexpr int L_EPI_WARPS = 16;             // 4 per TMEM lane quarter, one 32-column chunk of the key tile each
constexpr int L_THREADS = 64 + 32 * L_EPI_WARPS;
constexpr int L_Q_BYTES = 4 * TILE_BYTES;
constexpr int L_K_STAGE = 4 * TILE_BYTES;
constexpr size_t LSE64_SMEM = 1024 + L_Q_BYTES + 2 * L_K_STAGE + 256;

// Column sums over the 32 rows a warp owns, without shared memory: lane l holds row l's 32 column values v[0..31];
// five exchange-and-halve steps (16 + 8 + 4 + 2 + 1 shuffles) leave the sum (or max) of column l in lane l's v[0].
template <bool kMax>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      const float got = __shfl_xor_sync(0xffffffffu, send, off);
      v[i] = kMax ? fmaxf(keep, got) : keep + got;
    }
  }
  return v[0];
}

struct Lse64Args {
  int G, H, N, S;        // groups, heads per batch, query rows, key rows
  float scale2;
  float2* rowpart;       // [(g*4 + column chunk)*N + i]
  float2* colpart;       // [(g*4*IT + 4*it + quarter)*S + j]
};

__global__ void __launch_bounds__(L_THREADS, 1)
tc_lse64_kernel(const __grid_constant__ CUtensorMap mapQhi, const __grid_constant__ CUtensorMap mapQlo,
                const __grid_constant__ CUtensorMap mapKhi, const __grid_constant__ CUtensorMap mapKlo, Lse64Args p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t q_base = base, k_base = base + L_Q_BYTES;
  const uint32_t bar_base = k_base + 2 * L_K_STAGE;
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 24u + 8u * s; };
  auto tfull_bar = [&](int s) { return bar_base + 40u + 8u * s; };
  auto tempty_bar = [&](int s) { return bar_base + 56u + 8u * s; };
  const uint32_t tmem_slot = bar_base + 72u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - raw));

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int it = blockIdx.x, g = blockIdx.y, IT = gridDim.x;
  const int g0 = g % p.H, g1 = g / p.H;
  const int i0 = it * BM;
  const int JT = (p.S + BN - 1) / BN;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), L_EPI_WARPS);
    }
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
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, L_Q_BYTES);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_4d(q_base + (kb * 2 + 0) * TILE_BYTES, &mapQhi, q_full, kb * BK, i0, g0, g1);
        tma_load_4d(q_base + (kb * 2 + 1) * TILE_BYTES, &mapQlo, q_full, kb * BK, i0, g0, g1);
      }
    }
    __syncwarp();
    for (int jt = 0; jt < JT; ++jt) {
      const int st = jt & 1;
      mbar_wait(k_empty(st), (uint32_t)(((jt >> 1) & 1) ^ 1));
      if (elect_one()) {
        const uint32_t kb_s = k_base + st * L_K_STAGE;
        mbar_arrive_expect_tx(k_full(st), L_K_STAGE);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_4d(kb_s + (kb * 2 + 0) * TILE_BYTES, &mapKhi, k_full(st), kb * BK, jt * BN, g0, g1);
          tma_load_4d(kb_s + (kb * 2 + 1) * TILE_BYTES, &mapKlo, k_full(st), kb * BK, jt * BN, g0, g1);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(q_full, 0);
    for (int jt = 0; jt < JT; ++jt) {
      const int st = jt & 1;
      const uint32_t ph = (uint32_t)((jt >> 1) & 1);
      mbar_wait(k_full(st), ph);
      mbar_wait(tempty_bar(st), ph ^ 1u);   // accumulator buffer jt & 1 drained by the epilogue of tile jt - 2
      tc_fence_after();
      if (elect_one()) {
        const uint32_t kb_s = k_base + st * L_K_STAGE;
        const uint32_t t_main = tmem_base + (uint32_t)(st * 2 * BN), t_cross = t_main + (uint32_t)BN;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t dQhi = make_kmajor_sw128_desc(q_base + (kb * 2 + 0) * TILE_BYTES);
          const uint64_t dQlo = make_kmajor_sw128_desc(q_base + (kb * 2 + 1) * TILE_BYTES);
          const uint64_t dKhi = make_kmajor_sw128_desc(kb_s + (kb * 2 + 0) * TILE_BYTES);
          const uint64_t dKlo = make_kmajor_sw128_desc(kb_s + (kb * 2 + 1) * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
            umma_tf32(t_main, dQhi + koff, dKhi + koff, kIdescTf32N2, (kb | k) ? 1u : 0u);   // [main|cross] += Qhi [Khi;Klo]
            umma_tf32(t_cross, dQlo + koff, dKhi + koff, kIdescTf32, 1u);                    // cross += Qlo Khi
            (void)dKlo;
          }
        }
        umma_commit(k_empty(st));
        umma_commit(tfull_bar(st));
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue: 16 warps; thread = (query row, 32-key chunk of the key tile) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3, c = ew >> 2;      // TMEM lane quarter of this warp; its 32-column chunk
    const int row = quarter * 32 + lane, grow = i0 + row;
    const bool rvalid = grow < p.N;
    float m_run = -INFINITY, s_run = 0.f;
    for (int jt = 0; jt < JT; ++jt) {
      const int st = jt & 1;
      const int col0 = jt * BN + c * 32;
      mbar_wait(tfull_bar(st), (uint32_t)((jt >> 1) & 1));
      tc_fence_after();
      uint32_t a[32], b[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(st * 2 * BN + c * 32);
      tmem_ld32_nowait2(taddr, a);
      tmem_ld32_nowait2(taddr + (uint32_t)BN, b);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();           // this warp's part of the accumulator is in registers: hand the buffer back
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(st));
      float x[32];
      if (rvalid && col0 + 32 <= p.S) {
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = (__uint_as_float(a[e]) + __uint_as_float(b[e])) * p.scale2;
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          x[e] = (rvalid && col0 + e < p.S) ? (__uint_as_float(a[e]) + __uint_as_float(b[e])) * p.scale2 : -INFINITY;
      }
      // ---- ONE exponential per score serves the row and the column statistics: E = 2^(x - R) with R the maximum of
      // this warp's 32 x 32 chunk (warp shuffle reduce, no block barrier).  Any reference >= the data is a valid
      // log-sum-exp shift; what a shared reference can lose is a term that underflows (ex2.approx.ftz flushes below
      // 2^-126), so the shared path is taken only when the whole chunk lies within 100 binades of R (warp-uniform test;
      // otherwise, and for ragged edge chunks with -inf padding, the exact per-row / per-column maxima below).
      float c0 = fmaxf(x[0], x[4]), c1 = fmaxf(x[1], x[5]), c2 = fmaxf(x[2], x[6]), c3 = fmaxf(x[3], x[7]);
      float d0 = fminf(x[0], x[4]), d1 = fminf(x[1], x[5]), d2 = fminf(x[2], x[6]), d3 = fminf(x[3], x[7]);
#pragma unroll
      for (int e = 8; e < 32; e += 4) {
        c0 = fmaxf(c0, x[e]); c1 = fmaxf(c1, x[e + 1]); c2 = fmaxf(c2, x[e + 2]); c3 = fmaxf(c3, x[e + 3]);
        d0 = fminf(d0, x[e]); d1 = fminf(d1, x[e + 1]); d2 = fminf(d2, x[e + 2]); d3 = fminf(d3, x[e + 3]);
      }
      const float tmx = fmaxf(fmaxf(c0, c1), fmaxf(c2, c3));
      const float R = warp_max(tmx);
      const float mn = -warp_max(-fminf(fminf(d0, d1), fminf(d2, d3)));
      const int col = col0 + lane;
      float2* cp = p.colpart + ((size_t)g * 4 * IT + 4 * it + quarter) * p.S;
      if (R - mn <= 100.f) {   // (false for NaN / -inf padding)
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          x[e] = ex2(x[e] - R); x[e + 1] = ex2(x[e + 1] - R); x[e + 2] = ex2(x[e + 2] - R); x[e + 3] = ex2(x[e + 3] - R);
          a0 += x[e]; a1 += x[e + 1]; a2 += x[e + 2]; a3 += x[e + 3];
        }
        const float m_new = fmaxf(m_run, R);
        s_run = s_run * ex2(m_run - m_new) + ((a0 + a1) + (a2 + a3)) * ex2(R - m_new);
        m_run = m_new;
        const float cs = warp_transpose_reduce<false>(x, lane);   // column sums of the exponentials over the 32 rows
        if (col < p.S) cp[col] = make_float2(R, cs);
        continue;
      }
      // ---- exact path: row statistics, online across key tiles
      const float m_new = fmaxf(m_run, tmx);
      if (m_new > -INFINITY) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          a0 += ex2(x[e] - m_new); a1 += ex2(x[e + 1] - m_new); a2 += ex2(x[e + 2] - m_new); a3 += ex2(x[e + 3] - m_new);
        }
        s_run = s_run * ex2(m_run - m_new) + ((a0 + a1) + (a2 + a3));
        m_run = m_new;
      }
      // ---- column statistics of this warp's 32 rows: per-column maximum, then exponentials relative to it
      float y[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) y[e] = x[e];
      const float cmx = warp_transpose_reduce<true>(y, lane);      // lane l: max of column l over the 32 rows
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float cm = __shfl_sync(0xffffffffu, cmx, e);
        x[e] = cm > -INFINITY ? ex2(x[e] - cm) : 0.f;
      }
      const float cs = warp_transpose_reduce<false>(x, lane);
      if (col < p.S) cp[col] = make_float2(cmx, cs);
    }
    if (rvalid) p.rowpart[((size_t)g * 4 + c) * p.N + grow] = make_float2(m_run, s_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// hi/lo split of a strided operand set: copies group (b,h) rows [rows x K] into dense [G][rows][K] arrays
__global__ void split_groups_kernel(const float* __restrict__ x, long long sb, long long sh, int ld, int H, int G,
                                    int rows, int K, float* __restrict__ hi, float* __restrict__ lo) {
  const long long total = (long long)G * rows * K;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % K);
    const long long t = idx / K;
    const int r = (int)(t % rows);
    const int g = (int)(t / rows);
    const float v = x[(size_t)(g / H) * sb + (size_t)(g % H) * sh + (size_t)r * ld + k];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[idx] = h;
    lo[idx] = v - h;
  }
}
// float4 variant (K, ld, sb, sh multiples of 4 and 16-byte aligned pointers): one 16-byte load, two 16-byte stores
__global__ void __launch_bounds__(256) split_groups_vec_kernel(const float* __restrict__ x, long long sb, long long sh,
                                                               int ld, int H, int G, int rows, int K4,
                                                               float4* __restrict__ hi, float4* __restrict__ lo) {
  const long long total = (long long)G * rows * K4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k4 = (int)(idx % K4);
    const long long t = idx / K4;
    const int r = (int)(t % rows);
    const int g = (int)(t / rows);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)(g / H) * sb + (size_t)(g % H) * sh +
                                                            (size_t)r * ld) + k4);
    float4 h;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    hi[idx] = h;
    lo[idx] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  }
}
static void launch_split_groups(const float* x, long long sb, long long sh, int ld, int H, int G, int rows, int K,
                                float* hi, float* lo, cudaStream_t st) {
  const int blocks = kNumSMs * 8;
  const bool vec = (K % 4 == 0) && (ld % 4 == 0) && (sb % 4 == 0) && (sh % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15u) == 0;
  if (vec)
    split_groups_vec_kernel<<<blocks, 256, 0, st>>>(x, sb, sh, ld, H, G, rows, K / 4, reinterpret_cast<float4*>(hi),
                                                    reinterpret_cast<float4*>(lo));
  else
    split_groups_kernel<<<blocks, 256, 0, st>>>(x, sb, sh, ld, H, G, rows, K, hi, lo);
}

struct SplitOps { float *ahi, *alo, *bhi, *blo; };

static int prepare_operands(const ScoreArgs& a, float* ws, size_t ws_bytes, SplitOps* o, CUtensorMap maps[4],
                            int split_done, cudaStream_t st) {
  const size_t na = al((size_t)a.G * a.L * a.K * 4), nb = al((size_t)a.G * a.S * a.K * 4);
  if (ws == nullptr || ws_bytes < 2 * na + 2 * nb + 1024) return FAR_ERR_WORKSPACE;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  o->ahi = reinterpret_cast<float*>(base);
  o->alo = reinterpret_cast<float*>(base + na);
  o->bhi = reinterpret_cast<float*>(base + 2 * na);
  o->blo = reinterpret_cast<float*>(base + 2 * na + nb);
  if (!split_done) {
    launch_split_groups(a.A, a.sAb, a.sAh, a.lda, a.H, a.G, a.L, a.K, o->ahi, o->alo, st);
    FAR_CHECK_LAUNCH();
    launch_split_groups(a.B, a.sBb, a.sBh, a.ldb, a.H, a.G, a.S, a.K, o->bhi, o->blo, st);
    FAR_CHECK_LAUNCH();
  }
  // dense [G][rows][K]: 4-D map dims (K, rows, H, G/H)
  const int nbat = a.G / a.H;
  const bool ok = make_map4(&maps[0], o->ahi, a.K, a.L, a.K, a.H, (long long)a.L * a.K, nbat, (long long)a.H * a.L * a.K) &&
                  make_map4(&maps[1], o->alo, a.K, a.L, a.K, a.H, (long long)a.L * a.K, nbat, (long long)a.H * a.L * a.K) &&
                  make_map4(&maps[2], o->bhi, a.K, a.S, a.K, a.H, (long long)a.S * a.K, nbat, (long long)a.H * a.S * a.K) &&
                  make_map4(&maps[3], o->blo, a.K, a.S, a.K, a.H, (long long)a.S * a.K, nbat, (long long)a.H * a.S * a.K);
  return ok ? FAR_OK : FAR_ERR_CUDA;
}

static void set_attr_once() {
  static bool done[64] = {};
  if (first_use_on_device(done)) {
    cudaFuncSetAttribute(tc_score_kernel<MODE_LSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCORE_SMEM);
    cudaFuncSetAttribute(tc_score_kernel<MODE_CONF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCORE_SMEM);
  }
}

}  // namespace tc

bool tc_score_supported(const ScoreArgs& a) {
  return a.K % 4 == 0 && a.K >= 32 && a.G % a.H == 0 && tc::get_encode() != nullptr && tc_engine_default_on();
}

void tc_score_operand_ptrs(const ScoreArgs& a, float* ws, float** ahi, float** alo, float** bhi, float** blo) {
  const size_t na = tc::al((size_t)a.G * a.L * a.K * 4), nb = tc::al((size_t)a.G * a.S * a.K * 4);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  *ahi = reinterpret_cast<float*>(base);
  *alo = reinterpret_cast<float*>(base + na);
  *bhi = reinterpret_cast<float*>(base + 2 * na);
  *blo = reinterpret_cast<float*>(base + 2 * na + nb);
}

size_t tc_score_workspace_bytes(int G, int L, int S, int K) {
  return 2 * tc::al((size_t)G * L * K * 4) + 2 * tc::al((size_t)G * S * K * 4) + 4096;
}

int tc_score_lse_partials(const ScoreArgs& a, float2* rowpart, float2* colpart, float* ws, size_t ws_bytes,
                          int split_done, cudaStream_t st) {
  using namespace tc;
  SplitOps o;
  CUtensorMap maps[4];
  int rc = prepare_operands(a, ws, ws_bytes, &o, maps, split_done, st);
  if (rc) return rc;
  set_attr_once();
  ScoreTcArgs p{};
  p.G = a.G; p.H = a.H; p.L = a.L; p.S = a.S; p.K = a.K; p.scale2 = a.scale * kLog2e;
  p.rowpart = rowpart; p.colpart = colpart;
  const int tiles = a.G * ceil_div(a.L, BM) * ceil_div(a.S, BN);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  ProfScope prof(PROF_TC_SCORE, 2.0 * a.G * a.L * a.S * a.K, 4.0 * a.G * ((double)a.L + a.S) * a.K, st);
  tc_score_kernel<MODE_LSE><<<grid, S_THREADS, SCORE_SMEM, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

bool tc_lse64_supported(const ScoreArgs& a) { return tc_score_supported(a) && a.K == 64 && getenv("FAR_LSE64_OFF") == nullptr; }

// row partials: 2 per row ([(g*2 + half)*L + i]); column partials: 4 per 128-query tile ([(g*4*IT + 4*it + q)*S + j])
int tc_lse64_partials(const ScoreArgs& a, float2* rowpart, float2* colpart, float* ws, size_t ws_bytes, int split_done,
                      cudaStream_t st) {
  using namespace tc;
  SplitOps o;
  CUtensorMap maps[4];
  int rc = prepare_operands(a, ws, ws_bytes, &o, maps, split_done, st);
  if (rc) return rc;
  static bool attr[64] = {};
  if (first_use_on_device(attr)) {
    cudaFuncSetAttribute(tc_lse64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LSE64_SMEM);
  }
  Lse64Args p{a.G, a.H, a.L, a.S, a.scale * kLog2e, rowpart, colpart};
  ProfScope prof(PROF_TC_SCORE, 2.0 * a.G * a.L * a.S * a.K, 4.0 * a.G * ((double)a.L + a.S) * a.K, st);
  tc_lse64_kernel<<<dim3(ceil_div(a.L, BM), a.G), L_THREADS, LSE64_SMEM, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

int tc_match_conf(const ScoreArgs& a, const float* rowlse, const float* collse, float2* rowmax, float* colmax,
                  float* conf_out, float* ws, size_t ws_bytes, int split_done, cudaStream_t st) {
  using namespace tc;
  SplitOps o;
  CUtensorMap maps[4];
  int rc = prepare_operands(a, ws, ws_bytes, &o, maps, split_done, st);
  if (rc) return rc;
  set_attr_once();
  ScoreTcArgs p{};
  p.G = a.G; p.H = a.H; p.L = a.L; p.S = a.S; p.K = a.K; p.scale2 = a.scale * kLog2e;
  p.rowlse = rowlse; p.collse = collse; p.rowmax = rowmax; p.colmax = colmax; p.conf_out = conf_out;
  const int tiles = a.G * ceil_div(a.L, BM) * ceil_div(a.S, BN);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  ProfScope prof(PROF_TC_SCORE, 2.0 * a.G * a.L * a.S * a.K, 4.0 * a.G * ((double)a.L + a.S) * a.K, st);
  tc_score_kernel<MODE_CONF><<<grid, S_THREADS, SCORE_SMEM, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

}  // namespace far
