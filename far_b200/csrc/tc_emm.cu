// FAR CrossAttention pass B on tcgen05 (flash-style, fully fused; see emm.cu for the math and the CUDA-core version):
// one CTA per (128-query i-tile, batch*head).  For every 64-key j-tile:
//   S  = Q K^T            tcgen05.mma kind::tf32, 3xTF32, A/B from TMA-fed SWIZZLE_128B smem, accumulators in TMEM
//   P  = exp(2 S*scale - rowlse_i - collse_j)   by 4 softmax warps (thread = row): tcgen05.ld S -> exp -> hi/lo split
//        -> tcgen05.st back into TMEM (P never touches shared or global memory)
//   T += P V'             tcgen05.mma with the A operand (P) read from TMEM, B = V'^T tiles (K-major) from smem
// and finally  F_it = V'_i^T T  on the CUDA cores (70x70x128, < 5 % of the unit).
// TMEM columns: S/P buffer b: main [128b, 128b+64) cross [128b+64, 128b+128); T_main [256,336) T_cross [352,432).
#include "tc_common.cuh"
#include "tc_emm.cuh"

namespace far {
namespace tc {

constexpr int EBJ = 64;                     // keys per j-tile
constexpr int EDVP = 80;                    // d+6 = 70 padded to the MMA N granularity (16): TS MMA time scales with N
constexpr int E_Q_BYTES = 4 * TILE_BYTES;   // Q: 2 k-blocks x (hi, lo) x [128 x 32 floats]
constexpr int E_KT = EBJ * BK * 4;          // 8 KiB: one K box  [64 keys x 32 floats]
constexpr int E_VT = EDVP * BK * 4;         // 10 KiB: one V'^T box [80 channels x 32 keys]
constexpr int E_K_STAGE = 4 * E_KT;         // 2 k-blocks x (hi, lo)
constexpr int E_V_STAGE = 4 * E_VT;
constexpr int E_STAGE = E_K_STAGE + E_V_STAGE;  // 72 KiB
constexpr int E_BAR_BYTES = 1024;
constexpr size_t EMM_SMEM = 1024 + E_Q_BYTES + 2 * E_STAGE + E_BAR_BYTES;
constexpr int ETP = EDVP + 1;               // pitch of the final-stage T / V' staging tiles

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

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
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
constexpr int E_SM_WARPS = 16;              // softmax warps: 4 per TMEM lane quarter, 16 keys each (short S -> P latency)
constexpr int E_SM_THREADS = 32 * E_SM_WARPS;
constexpr int E_CW = EBJ / (E_SM_WARPS / 4);  // key columns per softmax thread (16)
constexpr int E_THREADS = 64 + E_SM_THREADS;  // producer warp, MMA warp, softmax warps
__device__ __forceinline__ void epi_bar2() { asm volatile("bar.sync 2, %0;" ::"n"(E_SM_THREADS) : "memory"); }
__device__ __forceinline__ void tmem_ld16_nw(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct EmmTcArgs {
  int G, H, N, d;            // groups, heads per batch, tokens, head dim
  float scale;
  const float* rowlse;       // [G][N]
  const float* collse;       // [G][N]
  const float* v;            // fp32 v base (qkv + 2C) for the final V'_i^T T stage
  long long sb, sh; int ldv; // its batch / head strides and row stride
  const float* pos; int Bpos;
  float* Fpart;              // [G][IT][dv*dv]
};

__global__ void __launch_bounds__(E_THREADS, 1)
tc_emm_pv_kernel(const __grid_constant__ CUtensorMap mapQhi, const __grid_constant__ CUtensorMap mapQlo,
                 const __grid_constant__ CUtensorMap mapKhi, const __grid_constant__ CUtensorMap mapKlo,
                 const __grid_constant__ CUtensorMap mapVhi, const __grid_constant__ CUtensorMap mapVlo, EmmTcArgs p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t q_base = base;
  const uint32_t st_base = base + E_Q_BYTES;
  const uint32_t bar_base = st_base + 2 * E_STAGE;
  const uint32_t q_full = bar_base + 0, t_full = bar_base + 8;
  auto k_full = [&](int s) { return bar_base + 16u + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 32u + 8u * s; };
  auto v_full = [&](int s) { return bar_base + 48u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 64u + 8u * s; };
  auto s_full = [&](int b) { return bar_base + 80u + 8u * b; };
  auto p_full = [&](int b) { return bar_base + 96u + 8u * b; };
  auto sp_empty = [&](int b) { return bar_base + 112u + 8u * b; };
  const uint32_t tmem_slot = bar_base + 128;
  unsigned char* gen_bar = smem_dyn + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_bar + 128);
  float* cls = reinterpret_cast<float*>(gen_bar + 256);          // [EBJ] column lse of the current j-tile
  unsigned char* gen_st = smem_dyn + (st_base - raw);            // final stage aliases the K / V' ring
  float(*Ts)[ETP] = reinterpret_cast<float(*)[ETP]>(gen_st);
  float(*Vs)[ETP] = reinterpret_cast<float(*)[ETP]>(gen_st + (size_t)BM * ETP * 4);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int it = blockIdx.x, g = blockIdx.y, IT = gridDim.x;
  const int g0 = g % p.H, g1 = g / p.H;
  const int i0 = it * BM;
  const int JT = (p.N + EBJ - 1) / EBJ;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1); mbar_init(t_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
      mbar_init(s_full(s), 1); mbar_init(p_full(s), E_SM_WARPS); mbar_init(sp_empty(s), 1);
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
  // TMEM columns: S/P buffer b: main [128b, 128b+64) cross [128b+64, 128b+128) -- the softmax warps overwrite S with
  // P (hi over main, lo over cross) in place, so two buffers let S(jt+1) run on the tensor pipe while the softmax of
  // tile jt and then T += P V'(jt) proceed; T_main [256,336), T_cross [352,432).
  auto tSP = [&](int b) { return tmem_base + (uint32_t)(b * 128); };
  const uint32_t tT_main = tmem_base + 256, tT_cross = tmem_base + 352;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, E_Q_BYTES);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_4d(q_base + (kb * 2 + 0) * TILE_BYTES, &mapQhi, q_full, kb * BK, i0, g0, g1);
        tma_load_4d(q_base + (kb * 2 + 1) * TILE_BYTES, &mapQlo, q_full, kb * BK, i0, g0, g1);
      }
      for (int jt = 0; jt < JT; ++jt) {  // lane 0: K tiles (freed as soon as S(jt) retires)
        const int st = jt & 1;
        const uint32_t ph = (uint32_t)((jt >> 1) & 1);
        mbar_wait(k_empty(st), ph ^ 1u);
        const uint32_t kbase = st_base + st * E_STAGE;
        mbar_arrive_expect_tx(k_full(st), E_K_STAGE);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_4d(kbase + (kb * 2 + 0) * E_KT, &mapKhi, k_full(st), kb * BK, jt * EBJ, g0, g1);
          tma_load_4d(kbase + (kb * 2 + 1) * E_KT, &mapKlo, k_full(st), kb * BK, jt * EBJ, g0, g1);
        }
      }
    } else if (lane == 1) {
      for (int jt = 0; jt < JT; ++jt) {  // lane 1: V'^T tiles (freed when T += P V'(jt) retires)
        const int st = jt & 1;
        const uint32_t ph = (uint32_t)((jt >> 1) & 1);
        mbar_wait(v_empty(st), ph ^ 1u);
        const uint32_t vbase = st_base + st * E_STAGE + E_K_STAGE;
        mbar_arrive_expect_tx(v_full(st), E_V_STAGE);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_3d(vbase + (kb * 2 + 0) * E_VT, &mapVhi, v_full(st), jt * EBJ + kb * BK, 0, g);
          tma_load_3d(vbase + (kb * 2 + 1) * E_VT, &mapVlo, v_full(st), jt * EBJ + kb * BK, 0, g);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Whole warp in the (warp-uniform) loop, barrier waits by all lanes, one elected lane issues (tc_common.cuh:
    // elect_one): UTCHMMA back to back instead of ~10 SASS instructions of ELECT/BRA/R2UR glue per 32-48-cycle MMA.
    {
      constexpr uint32_t idS = idesc_tf32(BM, EBJ), idS2 = idesc_tf32(BM, 2 * EBJ), idT = idesc_tf32(BM, EDVP);
      mbar_wait(q_full, 0);
      tc_fence_after();
      auto issue_S = [&](int jt) {   // S(jt) = Q K(jt)^T into S/P buffer jt & 1
        const int st = jt & 1;
        mbar_wait(k_full(st), (uint32_t)((jt >> 1) & 1));
        // S(jt) overwrites the S/P buffer that T += P V'(jt-2) reads.  No wait on that MMA's completion: it was issued
        // earlier by this same thread and tcgen05.mma instructions of one thread execute in issue order, so the
        // write-after-read on TMEM is ordered by the pipe itself -- and the tensor pipe never drains between tiles.
        tc_fence_after();
        if (elect_one()) {
          const uint32_t kbase = st_base + st * E_STAGE;
          const uint32_t tS_main = tSP(st), tS_cross = tSP(st) + 64;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t dQhi = make_kmajor_sw128_desc(q_base + (kb * 2 + 0) * TILE_BYTES);
            const uint64_t dQlo = make_kmajor_sw128_desc(q_base + (kb * 2 + 1) * TILE_BYTES);
            const uint64_t dKhi = make_kmajor_sw128_desc(kbase + (kb * 2 + 0) * E_KT);
            const uint64_t dKlo = make_kmajor_sw128_desc(kbase + (kb * 2 + 1) * E_KT);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
              // [S_main | S_cross] += Qhi x [Khi ; Klo] as ONE N = 128 instruction (K hi/lo tiles adjacent in shared
              // memory, S main/cross adjacent in TMEM): 64 + 48 cycles per k-step instead of 3 x 48 (SS MMAs are paced by
              // the shared-memory port, profiles/r1_mma_rate_microbench.md)
              umma_tf32(tS_main, dQhi + koff, dKhi + koff, idS2, (kb | k) ? 1u : 0u);
              umma_tf32(tS_cross, dQlo + koff, dKhi + koff, idS, 1u);
              (void)dKlo;
            }
          }
          umma_commit(s_full(st));
          umma_commit(k_empty(st));
        }
        __syncwarp();
      };
      issue_S(0);
      if (JT > 1) issue_S(1);
      for (int jt = 0; jt < JT; ++jt) {
        const int st = jt & 1;
        const uint32_t ph = (uint32_t)((jt >> 1) & 1);
        // ---- T += P V'   (A = P from TMEM, B = V'^T tile, K = 64 keys)
        mbar_wait(v_full(st), ph);
        mbar_wait(p_full(st), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t vbase = st_base + st * E_STAGE + E_K_STAGE;
          const uint32_t tP_hi = tSP(st), tP_lo = tSP(st) + 64;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t dVhi = make_kmajor_sw128_desc(vbase + (kb * 2 + 0) * E_VT);
            const uint64_t dVlo = make_kmajor_sw128_desc(vbase + (kb * 2 + 1) * E_VT);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
              const uint32_t acol = (uint32_t)(kb * BK + k * UMMA_K);
              const uint32_t first = (jt | kb | k) ? 1u : 0u;
              umma_tf32_ts(tT_cross, tP_lo + acol, dVhi + koff, idT, first);
              umma_tf32_ts(tT_cross, tP_hi + acol, dVlo + koff, idT, 1u);
              umma_tf32_ts(tT_main, tP_hi + acol, dVhi + koff, idT, first);
            }
          }
          umma_commit(v_empty(st));   // V' stage reusable
          if (jt == JT - 1) umma_commit(t_full);
        }
        __syncwarp();
        // queue S(jt+2) right behind it (same S/P buffer): the softmax warps work on tile jt+1 meanwhile
        if (jt + 2 < JT) issue_S(jt + 2);
      }
    }
  } else {
    // ===================== softmax / epilogue: 16 warps, thread = (query row, 16-key group of the j-tile) ========
    const int quarter = warp & 3;
    const int cgp = (warp - 2) >> 2;           // key-column group: columns [16*cgp, 16*cgp + 16) of the j-tile
    const int row = quarter * 32 + lane;
    const int et = (warp - 2) * 32 + lane;  // 0..255
    const int grow = i0 + row;
    const bool rvalid = grow < p.N;
    constexpr float kL2e = 1.4426950408889634f;
    const float rl2 = rvalid ? p.rowlse[(size_t)g * p.N + grow] * kL2e : 0.f;
    const float scale2x2 = 2.f * p.scale * kL2e;  // P = exp(2 s - rowlse - collse) = 2^(2 s log2e - rl2 - cl2)
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    // column lse of the j-tile, double-buffered in shared memory: tile jt+1's values are fetched from global memory
    // while tile jt is processed (the L2 latency used to sit on the per-tile critical path), one barrier per tile.
    if (et < EBJ) cls[et] = (et < p.N) ? p.collse[(size_t)g * p.N + et] * kL2e : 0.f;
    epi_bar2();
    for (int jt = 0; jt < JT; ++jt) {
      const int j0 = jt * EBJ;
      const int b = jt & 1;
      float cls_next = 0.f;
      if (et < EBJ && jt + 1 < JT && j0 + EBJ + et < p.N) cls_next = p.collse[(size_t)g * p.N + j0 + EBJ + et] * kL2e;
      const float* clsb = cls + b * EBJ;
      mbar_wait(s_full(b), (uint32_t)((jt >> 1) & 1));
      tc_fence_after();
      {
        uint32_t a[E_CW], bb[E_CW];
        const uint32_t tS = tSP(b) + lane_off + (uint32_t)(cgp * E_CW);
        tmem_ld16_nw(tS, a);
        tmem_ld16_nw(tS + 64, bb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (rvalid && j0 + EBJ <= p.N) {   // interior tile: no per-element masking
#pragma unroll
          for (int e = 0; e < E_CW; ++e) {
            const float x = __uint_as_float(a[e]) + __uint_as_float(bb[e]);
            const float pv = ex2a(fmaf(x, scale2x2, -(rl2 + clsb[cgp * E_CW + e])));
            const uint32_t h = __float_as_uint(pv) & 0xFFFFE000u;
            a[e] = h;                                             // hi
            bb[e] = __float_as_uint(pv - __uint_as_float(h));     // lo (exact)
          }
        } else {
#pragma unroll
          for (int e = 0; e < E_CW; ++e) {
            const float x = __uint_as_float(a[e]) + __uint_as_float(bb[e]);
            const bool ok = rvalid && (j0 + cgp * E_CW + e) < p.N;
            const float pv = ok ? ex2a(fmaf(x, scale2x2, -(rl2 + clsb[cgp * E_CW + e]))) : 0.f;
            const uint32_t h = __float_as_uint(pv) & 0xFFFFE000u;
            a[e] = h;
            bb[e] = __float_as_uint(pv - __uint_as_float(h));
          }
        }
        tmem_st16(tS, a);        // P_hi over S_main, P_lo over S_cross: same lanes / columns this thread just read
        tmem_st16(tS + 64, bb);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(b));
      if (et < EBJ) cls[(b ^ 1) * EBJ + et] = cls_next;   // buffer b^1 was last read in iteration jt-1
      epi_bar2();
    }
    // ---- final stage: F_it = V'_i^T T
    mbar_wait(t_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = cgp; c < EDVP / 16; c += E_SM_WARPS / 4) {
      uint32_t a[16], b[16];
      tmem_ld16_nw(tT_main + lane_off + (uint32_t)(c * 16), a);
      tmem_ld16_nw(tT_cross + lane_off + (uint32_t)(c * 16), b);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int e = 0; e < 16; ++e) Ts[row][c * 16 + e] = __uint_as_float(a[e]) + __uint_as_float(b[e]);
    }
    {
      const int b = g / p.H, h = g % p.H, dv = p.d + 6;
      const float* vb = p.v + (size_t)b * p.sb + (size_t)h * p.sh;
      const float* pb = p.pos + (size_t)(p.Bpos == 1 ? 0 : b) * p.N * 6;
      for (int idx = et; idx < BM * EDVP; idx += E_SM_THREADS) {
        const int r = idx / EDVP, c = idx % EDVP, tok = i0 + r;
        float val = 0.f;
        if (tok < p.N) {
          if (c < p.d) val = vb[(size_t)tok * p.ldv + c];
          else if (c < dv) val = pb[(size_t)tok * 6 + (c - p.d)];
        }
        Vs[r][c] = val;
      }
      epi_bar2();
      float* out = p.Fpart + ((size_t)g * IT + it) * dv * dv;
      for (int idx = et; idx < dv * dv; idx += E_SM_THREADS) {
        const int aa = idx / dv, cc = idx % dv;
        float s = 0.f;
#pragma unroll 8
        for (int i = 0; i < BM; ++i) s = fmaf(Vs[i][aa], Ts[i][cc], s);
        out[idx] = s;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// V'^T hi/lo: out[g][c][tok] for c < 96 (v channels, then the 6 positional columns, then zeros), tok < Npad.
// block (32, 8): 32-token x 32-channel tiles transposed through shared memory.
__global__ void emm_vt_kernel(const float* __restrict__ v, long long sb, long long sh, int ldv,
                              const float* __restrict__ pos, int Bpos, int H, int N, int Npad, int d,
                              float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int g = blockIdx.y, t0 = blockIdx.x * 32;
  const int b = g / H, h = g % H;
  const float* vb = v + (size_t)b * sb + (size_t)h * sh;
  const float* pb = pos + (size_t)(Bpos == 1 ? 0 : b) * N * 6;
  for (int cc = 0; cc < (EDVP + 31) / 32; ++cc) {
    for (int r = threadIdx.y; r < 32; r += 8) {  // r: token, threadIdx.x: channel (contiguous in v)
      const int tok = t0 + r, c = cc * 32 + threadIdx.x;
      float val = 0.f;
      if (tok < N) {
        if (c < d) val = vb[(size_t)tok * ldv + c];
        else if (c < d + 6) val = pb[(size_t)tok * 6 + (c - d)];
      }
      tile[r][threadIdx.x] = val;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {  // r: channel, threadIdx.x: token (contiguous in the output)
      const int c = cc * 32 + r, tok = t0 + threadIdx.x;
      if (tok < Npad && c < EDVP) {
        const float val = tile[threadIdx.x][r];
        const float hh = __uint_as_float(__float_as_uint(val) & 0xFFFFE000u);
        const size_t o = ((size_t)g * EDVP + c) * Npad + tok;
        hi[o] = hh;
        lo[o] = val - hh;
      }
    }
    __syncthreads();
  }
}

static bool make_map_vt(CUtensorMap* map, const float* ptr, int Npad, int G) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)Npad, (cuuint64_t)EDVP, (cuuint64_t)G};
  cuuint64_t gstr[2] = {(cuuint64_t)Npad * 4, (cuuint64_t)Npad * EDVP * 4};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)EDVP, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc

size_t tc_emm_vt_bytes(int G, int N) {
  const int Npad = (N + 3) & ~3;
  return 2 * tc::al((size_t)G * tc::EDVP * Npad * 4) + 2048;
}

bool tc_emm_supported(int N, int d) { return d == 64 && N >= 64 && tc::get_encode() != nullptr; }

int tc_emm_pv(const float* qhi, const float* qlo, const float* khi, const float* klo, const float* v, long long sb,
              long long sh, int ldv, const float* pos, int Bpos, int G, int H, int N, int d, float scale,
              const float* rowlse, const float* collse, float* Fpart, float* vtws, size_t vtws_bytes, cudaStream_t st) {
  using namespace tc;
  if (vtws == nullptr || vtws_bytes < tc_emm_vt_bytes(G, N)) return FAR_ERR_WORKSPACE;
  const int Npad = (N + 3) & ~3;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(vtws) + 1023) & ~uintptr_t(1023));
  float* vthi = reinterpret_cast<float*>(base);
  float* vtlo = reinterpret_cast<float*>(base + al((size_t)G * EDVP * Npad * 4));
  emm_vt_kernel<<<dim3(ceil_div(Npad, 32), G), dim3(32, 8), 0, st>>>(v, sb, sh, ldv, pos, Bpos, H, N, Npad, d, vthi, vtlo);
  FAR_CHECK_LAUNCH();
  CUtensorMap mQhi, mQlo, mKhi, mKlo, mVhi, mVlo;
  const int nbat = G / H;
  const long long gs = (long long)N * d, bs = (long long)H * N * d;
  if (!make_map4(&mQhi, qhi, d, N, d, H, gs, nbat, bs, BM) || !make_map4(&mQlo, qlo, d, N, d, H, gs, nbat, bs, BM) ||
      !make_map4(&mKhi, khi, d, N, d, H, gs, nbat, bs, EBJ) || !make_map4(&mKlo, klo, d, N, d, H, gs, nbat, bs, EBJ) ||
      !make_map_vt(&mVhi, vthi, Npad, G) || !make_map_vt(&mVlo, vtlo, Npad, G))
    return FAR_ERR_CUDA;
  static bool attr[64] = {};
  if (first_use_on_device(attr)) {
    cudaFuncSetAttribute(tc_emm_pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMM_SMEM);
  }
  EmmTcArgs p{G, H, N, d, scale, rowlse, collse, v, sb, sh, ldv, pos, Bpos, Fpart};
  ProfScope prof(PROF_TC_EMM_PV, (double)G * (2.0 * N * N * d + 2.0 * N * N * (d + 6) + 2.0 * N * (d + 6) * (d + 6)),
                 4.0 * G * (2.0 * N * d + (double)N * (d + 6)), st);
  tc_emm_pv_kernel<<<dim3(ceil_div(N, BM), G), E_THREADS, EMM_SMEM, st>>>(mQhi, mQlo, mKhi, mKlo, mVhi, mVlo, p);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

}  // namespace far
