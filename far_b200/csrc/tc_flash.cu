// Flash-style softmax(Q K^T) V on tcgen05 for the two places the path has a plain (single-direction) softmax attention:
//   * timm Attention of the 8pt-ViT blocks (interiornetStreetlearn_8ptVit/src/modules/vision_transformer.py:250-257,
//     N = 576, 3 heads x 64) and nn.MultiheadAttention of the map-free TransformerEncoder (model.py:57-61, 108 tokens,
//     8 heads x 32): far_softmax_attention  -> tc_flash_kernel<KB, DV, MODE_ATTN>
//   * the map-free correlation-volume aggregator, below                       -> tc_flash_kernel<1, 48, MODE_CORR>
// Map-free correlation-volume aggregator on tcgen05, flash style (never materialises the [HW, HW] volume):
//   mapfree_6dreg/lib/models/regression/aggregator.py:42-116  `CorrelationVolumeWarping.forward` with the shipped
//   recipe's flags (config/regression/mapfree/rot6d_trans_with_loftr.yaml:8-11: POSITION_ENCODER, MAX_SCORE_CHANNEL;
//   no dustbin / normalisation / CV layers):
//       C      = softmax_j(vol0^T vol1)                 [B, N, N], N = H*W (92 x 68 = 6256 at the 360 x 270 input)
//       vol1w  = vol1 C^T        pos = grid C^T         max_score = max_j C
//       out    = cat[vol0, vol1w, pos, max_score]       [B, 2D+3, N]
// One CTA per (128-query i-tile, batch element), D = 32 = one SWIZZLE_128B k-block.  Two sweeps over the key tiles
// inside the same CTA (the K tiles are re-streamed from L2: N*32*4*2 = 1.6 MB per batch element):
//   sweep 1   S = Q K^T (3xTF32)  ->  exact row maximum m_i (thread = row, thread-local max, no exponentials)
//   sweep 2   S again (the identical MMA sequence, so the same bits)  ->  P = 2^((s - m_i) log2 e) <= 1 by the softmax
//             warps, hi/lo split written back into TMEM over S  ->  T += P V'  with the A operand from TMEM,
//             V' = [vol1 | grid] (34 value channels padded to 48);  l_i = sum_j P in registers
//   epilogue  out = T / l ;  max_score = 1 / l   (the row maximum contributes exp(0) = 1 exactly)
// Subtracting the true row maximum is what torch.softmax does, so the result matches the reference to fp32 rounding
// without a running-max rescale of the TMEM accumulator.
// TMEM columns: S/P buffer b: main [128b, 128b+64) cross [128b+64, 128b+128);  T_main [256,304)  T_cross [320,368).
#include "tc_common.cuh"

namespace far {
namespace tc {

constexpr int CD = 32;                      // correlation volume: feature channels (ENCODER.NUM_OUT_LAYERS of the recipe)
constexpr int CBJ = 64;                     // keys per j-tile
constexpr int CDV = 48;                     // correlation volume: value channels padded to the MMA N granularity (32 + 2 grid + 14 zero)
constexpr int C_KT = CBJ * BK * 4;          // 8 KiB: one K box [64 keys x 32 floats]
constexpr int C_BAR_BYTES = 1024;
constexpr int MODE_CORR = 0, MODE_ATTN = 1;
// per-instantiation geometry: KB = 32-float k-blocks of the head dimension (1: d = 32, 2: d = 64), DV = value channels
template <int KB, int DV>
struct FlashCfg {
  static constexpr int Q_BYTES = KB * 2 * TILE_BYTES;    // per k-block: Q hi, lo [128 x 32 floats]
  static constexpr int VT = DV * BK * 4;                 // one V^T box [DV channels x 32 keys]
  static constexpr int K_STAGE = KB * 2 * C_KT;          // per k-block: hi, lo (adjacent: one N = 128 instruction)
  static constexpr int V_STAGE = 4 * VT;                 // 2 key blocks x (hi, lo)
  static constexpr int NST = (KB == 1) ? 4 : 2;          // ring depth (K and V rings are independent)
  static constexpr size_t SMEM = 1024 + Q_BYTES + NST * (K_STAGE + V_STAGE) + C_BAR_BYTES + 4 * BM * 4;
};
constexpr int C_SM_WARPS = 16;              // softmax warps: 4 per TMEM lane quarter, 16 keys each
constexpr int C_SM_THREADS = 32 * C_SM_WARPS;
constexpr int C_CW = CBJ / (C_SM_WARPS / 4);  // 16 key columns per softmax thread
constexpr int C_THREADS = 64 + C_SM_THREADS;

__host__ __device__ constexpr uint32_t c_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void c_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void c_tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void c_bar2() { asm volatile("bar.sync 2, %0;" ::"n"(C_SM_THREADS) : "memory"); }
__device__ __forceinline__ void c_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void c_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ float c_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct CorrArgs {
  int B, N, Cout;   // groups (batch, or batch * heads), tokens; MODE_CORR: output channels (2*CD + 3)
  float* out;       // MODE_CORR: [B, Cout, N];  MODE_ATTN: [B / H, N, H * DV]
  int H;            // MODE_ATTN: heads per batch element
  float scale;      // MODE_ATTN: softmax(scale * q k^T)
};

template <int KB, int DV, int MODE>
__global__ void __launch_bounds__(C_THREADS, 1)
tc_flash_kernel(const __grid_constant__ CUtensorMap mapQhi, const __grid_constant__ CUtensorMap mapQlo,
                  const __grid_constant__ CUtensorMap mapKhi, const __grid_constant__ CUtensorMap mapKlo,
                  const __grid_constant__ CUtensorMap mapVhi, const __grid_constant__ CUtensorMap mapVlo, CorrArgs p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  using Cfg = FlashCfg<KB, DV>;
  constexpr int C_NST = Cfg::NST, C_K_STAGE = Cfg::K_STAGE, C_V_STAGE = Cfg::V_STAGE, C_VT = Cfg::VT;
  const uint32_t q_base = base;
  const uint32_t k_base = base + Cfg::Q_BYTES;
  const uint32_t v_base = k_base + C_NST * C_K_STAGE;
  const uint32_t bar_base = v_base + C_NST * C_V_STAGE;
  const uint32_t q_full = bar_base + 0, t_full = bar_base + 8;
  auto k_full = [&](int s) { return bar_base + 16u + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 16u + 8u * (C_NST + s); };
  auto v_full = [&](int s) { return bar_base + 16u + 8u * (2 * C_NST + s); };
  auto v_empty = [&](int s) { return bar_base + 16u + 8u * (3 * C_NST + s); };
  auto s_full = [&](int b) { return bar_base + 16u + 8u * (4 * C_NST + b); };
  auto p_full = [&](int b) { return bar_base + 16u + 8u * (4 * C_NST + 2 + b); };
  const uint32_t tmem_slot = bar_base + 16u + 8u * (4 * C_NST + 4);
  unsigned char* gen_bar = smem_dyn + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_bar + 16 + 8 * (4 * C_NST + 4));
  float(*red)[BM] = reinterpret_cast<float(*)[BM]>(gen_bar + C_BAR_BYTES);   // [4 column groups][128 rows]

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int it = blockIdx.x, b = blockIdx.y;
  const int i0 = it * BM;
  const int JT = (p.N + CBJ - 1) / CBJ;
  const int NT = 2 * JT;   // tile visits: sweep 1 (t < JT) and sweep 2 (t >= JT)

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1); mbar_init(t_full, 1);
    for (int s = 0; s < C_NST; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(s_full(s), 1); mbar_init(p_full(s), C_SM_WARPS); }
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
  auto tSP = [&](int s) { return tmem_base + (uint32_t)(s * 128); };
  const uint32_t tT_main = tmem_base + 256, tT_cross = tmem_base + 256 + 64;   // DV <= 64

  if (warp == 0) {
    // ===================== TMA producer: lane 0 streams K tiles (both sweeps), lane 1 the V'^T tiles (sweep 2) =====
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        tma_load_4d(q_base + (kb * 2 + 0) * TILE_BYTES, &mapQhi, q_full, kb * BK, i0, 0, b);
        tma_load_4d(q_base + (kb * 2 + 1) * TILE_BYTES, &mapQlo, q_full, kb * BK, i0, 0, b);
      }
      for (int t = 0; t < NT; ++t) {
        const int jt = t < JT ? t : t - JT;
        const int st = t % C_NST;
        mbar_wait(k_empty(st), (uint32_t)(((t / C_NST) & 1) ^ 1));
        const uint32_t kb = k_base + st * C_K_STAGE;
        mbar_arrive_expect_tx(k_full(st), C_K_STAGE);
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          tma_load_4d(kb + (q * 2 + 0) * C_KT, &mapKhi, k_full(st), q * BK, jt * CBJ, 0, b);
          tma_load_4d(kb + (q * 2 + 1) * C_KT, &mapKlo, k_full(st), q * BK, jt * CBJ, 0, b);
        }
      }
    } else if (lane == 1) {
      for (int jt = 0; jt < JT; ++jt) {
        const int st = jt % C_NST;
        mbar_wait(v_empty(st), (uint32_t)(((jt / C_NST) & 1) ^ 1));
        const uint32_t vb = v_base + st * C_V_STAGE;
        mbar_arrive_expect_tx(v_full(st), C_V_STAGE);
        for (int kb = 0; kb < 2; ++kb) {
          c_tma_load_3d(vb + (kb * 2 + 0) * C_VT, &mapVhi, v_full(st), jt * CBJ + kb * BK, 0, b);
          c_tma_load_3d(vb + (kb * 2 + 1) * C_VT, &mapVlo, v_full(st), jt * CBJ + kb * BK, 0, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in the loop, one elected lane issues) =====================
    constexpr uint32_t idS = c_idesc(BM, CBJ), idS2 = c_idesc(BM, 2 * CBJ), idT = c_idesc(BM, DV);
    mbar_wait(q_full, 0);
    tc_fence_after();
    auto issue_S = [&](int t) {   // S(t) = Q K(t)^T into S/P buffer t & 1
      const int st = t % C_NST;
      mbar_wait(k_full(st), (uint32_t)((t / C_NST) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t kb = k_base + st * C_K_STAGE;
        const uint32_t tS_main = tSP(t & 1), tS_cross = tSP(t & 1) + 64;
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          const uint64_t dQhi = make_kmajor_sw128_desc(q_base + (q * 2 + 0) * TILE_BYTES);
          const uint64_t dQlo = make_kmajor_sw128_desc(q_base + (q * 2 + 1) * TILE_BYTES);
          const uint64_t dKhi = make_kmajor_sw128_desc(kb + (q * 2) * C_KT);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
            // [S_main | S_cross] += Qhi x [Khi ; Klo] as one N = 128 instruction (K hi/lo boxes adjacent in shared memory)
            umma_tf32(tS_main, dQhi + koff, dKhi + koff, idS2, (q | k) ? 1u : 0u);
            umma_tf32(tS_cross, dQlo + koff, dKhi + koff, idS, 1u);
          }
        }
        umma_commit(s_full(t & 1));
        umma_commit(k_empty(st));
      }
      __syncwarp();
    };
    issue_S(0);
    if (NT > 1) issue_S(1);
    for (int t = 0; t < NT; ++t) {
      const int sb = t & 1;
      // the softmax warps are done with S/P buffer sb: they took the row maxima (sweep 1) or wrote P (sweep 2)
      mbar_wait(p_full(sb), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      if (t >= JT) {
        const int jt = t - JT, st = jt % C_NST;
        mbar_wait(v_full(st), (uint32_t)((jt / C_NST) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t vb = v_base + st * C_V_STAGE;
          const uint32_t tP_hi = tSP(sb), tP_lo = tSP(sb) + 64;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t dVhi = make_kmajor_sw128_desc(vb + (kb * 2 + 0) * C_VT);
            const uint64_t dVlo = make_kmajor_sw128_desc(vb + (kb * 2 + 1) * C_VT);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
              const uint32_t acol = (uint32_t)(kb * BK + k * UMMA_K);
              const uint32_t first = (jt | kb | k) ? 1u : 0u;
              c_umma_ts(tT_cross, tP_lo + acol, dVhi + koff, idT, first);
              c_umma_ts(tT_cross, tP_hi + acol, dVlo + koff, idT, 1u);
              c_umma_ts(tT_main, tP_hi + acol, dVhi + koff, idT, first);
            }
          }
          umma_commit(v_empty(st));
          if (jt == JT - 1) umma_commit(t_full);
        }
        __syncwarp();
      }
      // S(t+2) goes into the same S/P buffer: ordered behind T += P V'(t) by the in-order tensor pipe
      if (t + 2 < NT) issue_S(t + 2);
    }
  } else {
    // ===================== softmax warps: thread = (query row, 16-key group of the j-tile) =====================
    const int quarter = warp & 3;
    const int cgp = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int grow = i0 + row;
    const bool rvalid = grow < p.N;
    const float kL2e = 1.4426950408889634f * (MODE == MODE_ATTN ? p.scale : 1.f);   // scale > 0: same arg max
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    // ---- sweep 1: exact row maximum
    float m = -INFINITY;
    for (int t = 0; t < JT; ++t) {
      const int sb = t & 1, j0 = t * CBJ;
      mbar_wait(s_full(sb), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      uint32_t a[C_CW], c[C_CW];
      const uint32_t tS = tSP(sb) + lane_off + (uint32_t)(cgp * C_CW);
      c_ld16(tS, a);
      c_ld16(tS + 64, c);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(sb));
      if (j0 + CBJ <= p.N) {
#pragma unroll
        for (int e = 0; e < C_CW; ++e) m = fmaxf(m, __uint_as_float(a[e]) + __uint_as_float(c[e]));
      } else {
#pragma unroll
        for (int e = 0; e < C_CW; ++e)
          if (j0 + cgp * C_CW + e < p.N) m = fmaxf(m, __uint_as_float(a[e]) + __uint_as_float(c[e]));
      }
    }
    red[cgp][row] = m;
    c_bar2();
    m = fmaxf(fmaxf(red[0][row], red[1][row]), fmaxf(red[2][row], red[3][row]));
    c_bar2();
    // ---- sweep 2: P = 2^((s - m) log2e), hi/lo split back into TMEM, row sums in registers
    float l = 0.f;
    for (int t = JT; t < NT; ++t) {
      const int sb = t & 1, j0 = (t - JT) * CBJ;
      mbar_wait(s_full(sb), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      uint32_t a[C_CW], c[C_CW];
      const uint32_t tS = tSP(sb) + lane_off + (uint32_t)(cgp * C_CW);
      c_ld16(tS, a);
      c_ld16(tS + 64, c);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const bool interior = j0 + CBJ <= p.N;
#pragma unroll
      for (int e = 0; e < C_CW; ++e) {
        const float x = __uint_as_float(a[e]) + __uint_as_float(c[e]);
        float pv = c_ex2((x - m) * kL2e);   // x - m is exact 0 at the row maximum: P = 1 there, as in torch.softmax
        if (!interior && !(j0 + cgp * C_CW + e < p.N)) pv = 0.f;
        l += pv;
        const uint32_t h = __float_as_uint(pv) & 0xFFFFE000u;
        a[e] = h;
        c[e] = __float_as_uint(pv - __uint_as_float(h));
      }
      c_st16(tS, a);
      c_st16(tS + 64, c);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(sb));
    }
    red[cgp][row] = l;
    c_bar2();
    const float inv = 1.f / (red[0][row] + red[1][row] + red[2][row] + red[3][row]);
    mbar_wait(t_full, 0);
    tc_fence_after();
    if (MODE == MODE_CORR) {
      // ---- out[b, CD + c, i] = T[i, c] / l  (c < CD + 2);  out[b, 2CD+2, i] = 1 / l   (channel-major: lanes = tokens)
      float* ob = p.out + (size_t)b * p.Cout * p.N;
      if (cgp < 3) {
        uint32_t a[16], c[16];
        c_ld16(tT_main + lane_off + (uint32_t)(cgp * 16), a);
        c_ld16(tT_cross + lane_off + (uint32_t)(cgp * 16), c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (rvalid) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int ch = cgp * 16 + e;
            if (ch < CD + 2) ob[(size_t)(CD + ch) * p.N + grow] = (__uint_as_float(a[e]) + __uint_as_float(c[e])) * inv;
          }
        }
      } else if (rvalid) {
        ob[(size_t)(2 * CD + 2) * p.N + grow] = inv;
      }
    } else {
      // ---- out[bb, i, h * DV + c] = T[i, c] / l   (token-major, heads re-interleaved: each thread owns 16-float runs)
      const int bb = b / p.H, hh = b % p.H;
      float* orow = p.out + ((size_t)bb * p.N + grow) * ((size_t)p.H * DV) + (size_t)hh * DV;
      for (int cc = cgp; cc < DV / 16; cc += 4) {
        uint32_t a[16], c[16];
        c_ld16(tT_main + lane_off + (uint32_t)(cc * 16), a);
        c_ld16(tT_cross + lane_off + (uint32_t)(cc * 16), c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (rvalid) {
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            float4 v;
            v.x = (__uint_as_float(a[e]) + __uint_as_float(c[e])) * inv;
            v.y = (__uint_as_float(a[e + 1]) + __uint_as_float(c[e + 1])) * inv;
            v.z = (__uint_as_float(a[e + 2]) + __uint_as_float(c[e + 2])) * inv;
            v.w = (__uint_as_float(a[e + 3]) + __uint_as_float(c[e + 3])) * inv;
            *reinterpret_cast<float4*>(orow + cc * 16 + e) = v;
          }
        }
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

// Operand preparation (one pass over vol0 / vol1): token-major hi/lo splits of Q = vol0^T and K = vol1^T
// [B, N, 32], channel-major hi/lo of V'^T = [vol1 ; grid ; 0] [B, 48, Npad], and the vol0 copy into out[:, :32].
// vol given with element strides (batch, channel, pixel): NCHW (sc = N, sp = 1) or channels_last (sc = 1, sp = 32).
// block (32, 8): 32-token x 32-channel tiles transposed through shared memory.
__global__ void corrvol_prep_kernel(const float* __restrict__ vol0, const float* __restrict__ vol1, long long sb,
                                    long long sc, long long sp, const float* __restrict__ grid, int N, int Npad,
                                    float* __restrict__ qhi, float* __restrict__ qlo, float* __restrict__ khi,
                                    float* __restrict__ klo, float* __restrict__ vthi, float* __restrict__ vtlo,
                                    float* __restrict__ out, int Cout) {
  __shared__ float t0[32][33], t1[32][33];   // [channel][token]
  const int b = blockIdx.y, tok0 = blockIdx.x * 32;
  const float* v0 = vol0 + (size_t)b * sb;
  const float* v1 = vol1 + (size_t)b * sb;
  const int tx = threadIdx.x;
  if (sp == 1) {   // channel-major input: threadIdx.x runs over tokens
    for (int c = threadIdx.y; c < CD; c += 8) {
      const int tok = tok0 + tx;
      t0[c][tx] = tok < N ? v0[(size_t)c * sc + tok] : 0.f;
      t1[c][tx] = tok < N ? v1[(size_t)c * sc + tok] : 0.f;
    }
  } else {         // token-major (channels_last) input: threadIdx.x runs over channels
    for (int r = threadIdx.y; r < 32; r += 8) {
      const int tok = tok0 + r;
      t0[tx][r] = tok < N ? v0[(size_t)tok * sp + (size_t)tx * sc] : 0.f;
      t1[tx][r] = tok < N ? v1[(size_t)tok * sp + (size_t)tx * sc] : 0.f;
    }
  }
  __syncthreads();
  // channel-major outputs: V'^T rows 0..31 and the vol0 copy (threadIdx.x = token)
  for (int c = threadIdx.y; c < CD; c += 8) {
    const int tok = tok0 + tx;
    if (tok < Npad) {
      const float v = t1[c][tx];
      const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      const size_t o = ((size_t)b * CDV + c) * Npad + tok;
      vthi[o] = h;
      vtlo[o] = v - h;
    }
    if (tok < N) out[((size_t)b * Cout + c) * N + tok] = t0[c][tx];
  }
  for (int c = CD + threadIdx.y; c < CDV; c += 8) {   // grid rows (u, v) then zero padding
    const int tok = tok0 + tx;
    if (tok < Npad) {
      const float v = (c < CD + 2 && tok < N) ? grid[(size_t)(c - CD) * N + tok] : 0.f;
      const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      const size_t o = ((size_t)b * CDV + c) * Npad + tok;
      vthi[o] = h;
      vtlo[o] = v - h;
    }
  }
  // token-major outputs: Q and K hi/lo (threadIdx.x = channel)
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int tok = tok0 + r;
    if (tok < N) {
      const size_t o = ((size_t)b * N + tok) * CD + tx;
      const float q = t0[tx][r], k = t1[tx][r];
      const float qh = __uint_as_float(__float_as_uint(q) & 0xFFFFE000u);
      const float kh = __uint_as_float(__float_as_uint(k) & 0xFFFFE000u);
      qhi[o] = qh; qlo[o] = q - qh;
      khi[o] = kh; klo[o] = k - kh;
    }
  }
}

static bool make_map_vt(CUtensorMap* map, const float* ptr, int Npad, int DV, int B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)Npad, (cuuint64_t)DV, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)Npad * 4, (cuuint64_t)Npad * DV * 4};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)DV, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Attention operand preparation: qkv [B, N, 3, H, d] (the fused qkv Linear's output) -> token-major hi/lo of q and k
// [G = B*H, N, d] and channel-major hi/lo of v^T [G, d, Npad].  grid (ceil(Npad/32), d/32, G), block (32, 8).
__global__ void attn_prep_kernel(const float* __restrict__ qkv, int N, int Npad, int H, int d, float* __restrict__ qhi,
                                 float* __restrict__ qlo, float* __restrict__ khi, float* __restrict__ klo,
                                 float* __restrict__ vthi, float* __restrict__ vtlo) {
  __shared__ float tv[32][33];   // [token][channel]
  const int g = blockIdx.z, bb = g / H, hh = g % H, c0 = blockIdx.y * 32, tok0 = blockIdx.x * 32;
  const int C = H * d, tx = threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int tok = tok0 + r;
    float q = 0.f, k = 0.f, v = 0.f;
    if (tok < N) {
      const float* row = qkv + ((size_t)bb * N + tok) * 3 * C + (size_t)hh * d + c0 + tx;
      q = row[0]; k = row[C]; v = row[2 * C];
      const size_t o = ((size_t)g * N + tok) * d + c0 + tx;
      const float qh = __uint_as_float(__float_as_uint(q) & 0xFFFFE000u);
      const float kh = __uint_as_float(__float_as_uint(k) & 0xFFFFE000u);
      qhi[o] = qh; qlo[o] = q - qh;
      khi[o] = kh; klo[o] = k - kh;
    }
    tv[r][tx] = v;
  }
  __syncthreads();
  for (int c = threadIdx.y; c < 32; c += 8) {   // threadIdx.x = token: contiguous in v^T
    const int tok = tok0 + tx;
    if (tok < Npad) {
      const float v = tv[tx][c];
      const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      const size_t o = ((size_t)g * d + c0 + c) * Npad + tok;
      vthi[o] = h;
      vtlo[o] = v - h;
    }
  }
}

template <int KB, int DV, int MODE>
static int launch_flash(const float* qhi, const float* qlo, const float* khi, const float* klo, const float* vthi,
                        const float* vtlo, int G, int N, int Npad, const CorrArgs& p, cudaStream_t st) {
  constexpr int d = KB * BK;
  CUtensorMap mQhi, mQlo, mKhi, mKlo, mVhi, mVlo;
  const long long bs = (long long)N * d;
  if (!make_map4(&mQhi, qhi, d, N, d, 1, bs, G, bs, BM) || !make_map4(&mQlo, qlo, d, N, d, 1, bs, G, bs, BM) ||
      !make_map4(&mKhi, khi, d, N, d, 1, bs, G, bs, CBJ) || !make_map4(&mKlo, klo, d, N, d, 1, bs, G, bs, CBJ) ||
      !make_map_vt(&mVhi, vthi, Npad, DV, G) || !make_map_vt(&mVlo, vtlo, Npad, DV, G))
    return FAR_ERR_CUDA;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set))
    cudaFuncSetAttribute(tc_flash_kernel<KB, DV, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)FlashCfg<KB, DV>::SMEM);
  tc_flash_kernel<KB, DV, MODE><<<dim3(ceil_div(N, BM), G), C_THREADS, FlashCfg<KB, DV>::SMEM, st>>>(mQhi, mQlo, mKhi, mKlo,
                                                                                                   mVhi, mVlo, p);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

}  // namespace tc

// ---- softmax attention (far_softmax_attention, emm.cu) on the flash kernel: d = 32 or 64 -----------------------------
bool tc_flash_attention_supported(int N, int d) { return (d == 32 || d == 64) && N >= 1 && tc::get_encode() != nullptr; }

size_t tc_flash_attention_bytes(int G, int N, int d) {
  const int Npad = (N + 3) & ~3;
  return 4 * tc::al((size_t)G * N * d * 4) + 2 * tc::al((size_t)G * d * Npad * 4) + 2048;
}

int tc_flash_attention(const float* qkv, int B, int N, int H, int d, float scale, float* out, float* workspace,
                       size_t workspace_bytes, cudaStream_t st) {
  using namespace tc;
  const int G = B * H, Npad = (N + 3) & ~3;
  if (workspace_bytes < tc_flash_attention_bytes(G, N, d)) return FAR_ERR_WORKSPACE;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  const size_t qb = al((size_t)G * N * d * 4), vb = al((size_t)G * d * Npad * 4);
  float* qhi = reinterpret_cast<float*>(base);
  float* qlo = reinterpret_cast<float*>(base + qb);
  float* khi = reinterpret_cast<float*>(base + 2 * qb);
  float* klo = reinterpret_cast<float*>(base + 3 * qb);
  float* vthi = reinterpret_cast<float*>(base + 4 * qb);
  float* vtlo = reinterpret_cast<float*>(base + 4 * qb + vb);
  attn_prep_kernel<<<dim3(ceil_div(Npad, 32), d / 32, G), dim3(32, 8), 0, st>>>(qkv, N, Npad, H, d, qhi, qlo, khi, klo,
                                                                                vthi, vtlo);
  FAR_CHECK_LAUNCH();
  CorrArgs p{G, N, 0, out, H, scale};
  // algorithmic work: S = q k^T twice (row-maximum sweep + recompute) + P V; bytes: qkv in, out
  ProfScope prof(PROF_TC_FLASH_ATTN, (double)G * (2.0 * 2.0 * N * (double)N * d + 2.0 * N * (double)N * d),
                 4.0 * G * 4.0 * (double)N * d, st);
  if (d == 64) return launch_flash<2, 64, MODE_ATTN>(qhi, qlo, khi, klo, vthi, vtlo, G, N, Npad, p, st);
  return launch_flash<1, 32, MODE_ATTN>(qhi, qlo, khi, klo, vthi, vtlo, G, N, Npad, p, st);
}

}  // namespace far

using namespace far;
using namespace far::tc;

extern "C" size_t far_corr_volume_warp_workspace_bytes(int B, int N) {
  const int Npad = (N + 3) & ~3;
  return 4 * al((size_t)B * N * CD * 4) + 2 * al((size_t)B * CDV * Npad * 4) + 2048;
}

extern "C" int far_corr_volume_warp(const float* vol0, const float* vol1, long long sb, long long sc, long long sp,
                                    const float* grid, int B, int N, int D, float* out, float* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (B <= 0 || N <= 0) return FAR_OK;
  FAR_REQUIRE(vol0 && vol1 && grid && out && workspace);
  FAR_REQUIRE(D == CD);   // ENCODER.NUM_OUT_LAYERS = 32 in the FAR map-free recipe (one SWIZZLE_128B k-block)
  FAR_REQUIRE(sp == 1 || sc == 1);
  if (workspace_bytes < far_corr_volume_warp_workspace_bytes(B, N)) return FAR_ERR_WORKSPACE;
  if (get_encode() == nullptr) return FAR_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const int Npad = (N + 3) & ~3;
  const int Cout = 2 * CD + 3;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  const size_t qb = al((size_t)B * N * CD * 4), vb = al((size_t)B * CDV * Npad * 4);
  float* qhi = reinterpret_cast<float*>(base);
  float* qlo = reinterpret_cast<float*>(base + qb);
  float* khi = reinterpret_cast<float*>(base + 2 * qb);
  float* klo = reinterpret_cast<float*>(base + 3 * qb);
  float* vthi = reinterpret_cast<float*>(base + 4 * qb);
  float* vtlo = reinterpret_cast<float*>(base + 4 * qb + vb);
  corrvol_prep_kernel<<<dim3(ceil_div(Npad, 32), B), dim3(32, 8), 0, st>>>(vol0, vol1, sb, sc, sp, grid, N, Npad, qhi, qlo,
                                                                           khi, klo, vthi, vtlo, out, Cout);
  FAR_CHECK_LAUNCH();
  CorrArgs p{B, N, Cout, out, 1, 1.f};
  // algorithmic work: S twice (row-max sweep + recompute) + P V';  bytes: vol0, vol1 in, 2D+3 channels out
  ProfScope prof(PROF_TC_CORRVOL, (double)B * (2.0 * 2.0 * N * (double)N * CD + 2.0 * N * (double)N * (CD + 2)),
                 4.0 * B * ((double)N * 2 * CD + (double)N * Cout), st);
  return launch_flash<1, CDV, MODE_CORR>(qhi, qlo, khi, klo, vthi, vtlo, B, N, Npad, p, st);
}
