// tcgen05 3xTF32 GEMM engine for sm_100a:  C[M,N] = act( A[M,K] * B[N,K]^T + bias ),  fp32 in / fp32 out.
//
// Why 3xTF32: every reference GEMM on the path is IEEE fp32 and "bit-exact match indices" needs ~fp32 accuracy;
// tcgen05 has no fp32 MMA kind.  Each operand is split  x = hi + lo  (hi = x with the low 13 mantissa bits cleared,
// lo = x - hi exactly) and the product is accumulated as  lo*hi + hi*lo + hi*hi  in the fp32 TMEM accumulator
// (error ~2^-21 relative per product; SURVEY.md 7 measured an identical match set with this scheme).
//
// Structure (one persistent CTA per SM; 320 threads, 448 with the in-kernel operand split):
//   warp 0       TMA producer : cp.async.bulk.tensor (SWIZZLE_128B, 128-row x 32-float boxes) of A (raw fp32 or
//                               pre-split hi/lo), Bhi, Blo into a 3-stage shared-memory ring, mbarrier complete_tx
//   warp 1       MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M128 x N128 x K8), 12 per
//                               stage, into one of two (main | cross-term) TMEM accumulator pairs; tcgen05.commit
//                               releases the smem stage / publishes the accumulator
//   warps 2-9    epilogue     : two warps per TMEM lane quarter (64 columns each): tcgen05.ld 32x32b.x32 of both
//                               accumulators, bias + activation (branch-free; optionally the linear-attention
//                               normaliser Z), eight conflict-free STS.128 into a SWIZZLE_128B staging tile, then ONE
//                               TMA store (cp.async.bulk.tensor ... global.shared::cta) per 32x32 block.  Overlaps the
//                               next tile's main loop through the second accumulator pair.
//   warps 10-13  converters   : (raw-A mode) split each landed fp32 A tile into hi (in place) and lo (second slot).
// Grouped mode: rows are G groups of L rows (tiles never straddle a group; TMA zero-fills / clips the ragged last
// tile of each group) and B may be a per-group [G][N][K] operand -- used to fold the linear-attention apply step into
// the `merge` projection (encoder_layer.cu).
#include "tc_gemm_epi.cuh"

namespace far {
namespace tc {

// kRawA = false: A arrives pre-split (mapAhi / mapAlo).
// kRawA = true : mapAhi / mapAlo are the ORIGINAL fp32 activation tensors x1 [M,K1] / x2 [M,K2] (k-blocks below
//                K1/32 come from x1, the rest from x2); the converter warps split each landed tile in shared memory
//                (hi in place, lo into the second slot; the swizzle is a permutation of 16-byte chunks, so an
//                elementwise pass at identical offsets preserves the UMMA layout) and publish it to the MMA warp through
//                a third barrier after fence.proxy.async.  No split pass over the activations in HBM, half the A bytes.
template <bool kRawA>
__global__ void __launch_bounds__(kRawA ? GEMM_THREADS_RAW : GEMM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
               const __grid_constant__ CUtensorMap mapC, GemmArgs p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t stg_base = base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = stg_base + EPI_WARPS * STG_TILE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  auto conv_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - raw));

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_pg = (p.L + BM - 1) / BM;  // m-tiles per group
  const int num_tiles = p.G * tiles_pg * tiles_n;
  const int kblocks = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(conv_bar(s), 4); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is warp-wide; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {  // whole warp in the loop; one elected lane issues the TMA instructions
      int stage = 0;
      uint32_t phase = 0;
      // L2 prefetch of the raw activation tiles PF k-blocks ahead of the shared-memory ring (cp.async.bulk.prefetch):
      // a stage can only be refilled once its MMAs retire, so with 3 stages the HBM latency of A sat on the
      // stage-recycling chain; with the tile already in L2 the refill is an L2 hit.
      constexpr int PF = 6;
      auto prefetch_a = [&](int tile_p, int kb_p) {
        if (tile_p >= num_tiles) return;
        const int tm_p = tile_p / tiles_n;
        const int g_p = tm_p / tiles_pg, r0_p = (tm_p % tiles_pg) * BM;
        if (kb_p < p.kb1) tma_prefetch_4d(&mapAhi, kb_p * BK, r0_p, g_p, 0);
        else tma_prefetch_4d(&mapAlo, (kb_p - p.kb1) * BK, r0_p, g_p, 0);
      };
      if (kRawA && !(p.dbg & 64) && elect_one()) {
        for (int i = 0; i < PF; ++i) prefetch_a(blockIdx.x + (i / kblocks) * gridDim.x, i % kblocks);
      }
      __syncwarp();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tm = tile / tiles_n, n0 = (tile % tiles_n) * BN;
        const int g = tm / tiles_pg, r0 = (tm % tiles_pg) * BM, gb = p.b_grouped ? g : 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (kRawA && !(p.dbg & 64)) {   // issued before the (possibly long) wait for a free stage
            if (elect_one()) {
              const int ahead = kb + PF;
              prefetch_a(tile + (ahead / kblocks) * gridDim.x, ahead % kblocks);
            }
            __syncwarp();
          }
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sbase = base + stage * STAGE_BYTES;
          if (elect_one()) {
          if (kRawA && (p.dbg & 24)) {  // diagnostics: 8 = skip the B loads, 16 = skip the A load
            mbar_arrive_expect_tx(full_bar(stage), ((p.dbg & 8) ? 0 : 2 * TILE_BYTES) + ((p.dbg & 16) ? 0 : TILE_BYTES));
            if (!(p.dbg & 16)) tma_load_4d(sbase + 0 * TILE_BYTES, &mapAhi, full_bar(stage), kb * BK, r0, g, 0);
            if (!(p.dbg & 8)) {
              tma_load_4d(sbase + 2 * TILE_BYTES, &mapBhi, full_bar(stage), kb * BK, n0, gb, 0);
              tma_load_4d(sbase + 3 * TILE_BYTES, &mapBlo, full_bar(stage), kb * BK, n0, gb, 0);
            }
          } else if (kRawA) {
            mbar_arrive_expect_tx(full_bar(stage), 3 * TILE_BYTES);
            if (kb < p.kb1) tma_load_4d(sbase + 0 * TILE_BYTES, &mapAhi, full_bar(stage), kb * BK, r0, g, 0);
            else tma_load_4d(sbase + 0 * TILE_BYTES, &mapAlo, full_bar(stage), (kb - p.kb1) * BK, r0, g, 0);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_4d(sbase + 0 * TILE_BYTES, &mapAhi, full_bar(stage), kb * BK, r0, g, 0);
            tma_load_4d(sbase + 1 * TILE_BYTES, &mapAlo, full_bar(stage), kb * BK, r0, g, 0);
          }
          if (!(kRawA && (p.dbg & 24))) {
            tma_load_4d(sbase + 2 * TILE_BYTES, &mapBhi, full_bar(stage), kb * BK, n0, gb, 0);
            tma_load_4d(sbase + 3 * TILE_BYTES, &mapBlo, full_bar(stage), kb * BK, n0, gb, 0);
          }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (warp-uniform) loop and waits on the barriers; one elected lane issues the MMAs and
    // the commits, so descriptors and loop state stay in uniform registers.
    {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        // Tensor-core accumulation into fp32 rounds toward zero at every MMA (measured: a systematic ~K/8 * 3 * 2^-24
        // relative shrink when all three product groups share one accumulator).  The 2^-11-times-smaller cross terms
        // therefore get their own accumulator: the main one sees K/8 accumulation steps instead of 3K/8.
        const uint32_t tmem_main = tmem_base + (uint32_t)(acc * 2 * BN);
        const uint32_t tmem_small = tmem_main + (uint32_t)BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(kRawA ? conv_bar(stage) : full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sbase = base + stage * STAGE_BYTES;
            const uint64_t dAhi = make_kmajor_sw128_desc(sbase + 0 * TILE_BYTES);
            const uint64_t dAlo = make_kmajor_sw128_desc(sbase + 1 * TILE_BYTES);
            const uint64_t dBhi = make_kmajor_sw128_desc(sbase + 2 * TILE_BYTES);
            const uint64_t dBlo = make_kmajor_sw128_desc(sbase + 3 * TILE_BYTES);
            if (kRawA && p.cross16) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                if (p.dbg & 4) break;
                const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);  // 32 bytes per step: 8 tf32 or 16 bf16 of K
                umma_tf32(tmem_main, dAhi + koff, dBhi + koff, kIdescTf32, (kb | k) ? 1u : 0u);   // main  += A B      (tf32)
                umma_bf16(tmem_small, dAlo + koff, dBlo + koff, kIdescBf16, (kb | k) ? 1u : 0u);  // cross += A_cat B_cat^T (bf16)
              }
            } else {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                if (p.dbg & 4) break;
                const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);  // advance inside the 128-byte swizzle row
                umma_tf32(tmem_main, dAhi + koff, dBhi + koff, kIdescTf32N2, (kb | k) ? 1u : 0u);  // [main|cross] += Ahi [Bhi;Blo]
                umma_tf32(tmem_small, dAlo + koff, dBhi + koff, kIdescTf32, 1u);                    // cross += Alo Bhi
              }
            }
            umma_commit(empty_bar(stage));  // smem stage reusable once these MMAs retire
            if (kb == kblocks - 1) umma_commit(tfull_bar(acc));  // accumulator complete
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 2 + EPI_WARPS) {
    // ===================== converter warps: x -> (hi, lo) in shared memory (raw-A mode only) =====================
    if (kRawA) {
      const int ct = threadIdx.x - (2 + EPI_WARPS) * 32;  // 0..127
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          unsigned char* sa = smem_dyn + (base + stage * STAGE_BYTES - raw);
          float4* a4 = reinterpret_cast<float4*>(sa);
          float4* l4 = reinterpret_cast<float4*>(sa + TILE_BYTES);
          if (p.cross16) {
            // A_cat row r = [32 x bf16(x) | 32 x bf16(x - trunc_tf32(x))]: the fp32 chunk pair (2j, 2j+1) of a row becomes
            // bf16 chunk j (hi half) and chunk 4+j (lo half); SWIZZLE_128B stores 16-byte chunk c of row r at c ^ (r & 7).
            uint4* c4 = reinterpret_cast<uint4*>(sa + TILE_BYTES);
#pragma unroll
            for (int i = 0; i < ((p.dbg & 32) ? 0 : 4); ++i) {
              const int q = ct + i * 128, r = q >> 2, j = q & 3, sw = r & 7;
              const float4 u = a4[r * 8 + ((2 * j) ^ sw)], w = a4[r * 8 + ((2 * j + 1) ^ sw)];
              uint4 hi, lo;
              hi.x = pack_bf16x2(u.x, u.y); hi.y = pack_bf16x2(u.z, u.w); hi.z = pack_bf16x2(w.x, w.y); hi.w = pack_bf16x2(w.z, w.w);
              lo.x = pack_bf16x2(tf32_lo(u.x), tf32_lo(u.y)); lo.y = pack_bf16x2(tf32_lo(u.z), tf32_lo(u.w));
              lo.z = pack_bf16x2(tf32_lo(w.x), tf32_lo(w.y)); lo.w = pack_bf16x2(tf32_lo(w.z), tf32_lo(w.w));
              c4[r * 8 + (j ^ sw)] = hi;
              c4[r * 8 + ((4 + j) ^ sw)] = lo;
            }
          } else
#pragma unroll
          for (int i = 0; i < ((p.dbg & 32) ? 0 : TILE_BYTES / 16 / 128); ++i) {  // dbg 32: skip the split itself
            const float4 v = a4[ct + i * 128];
            float4 h;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            // kind::tf32 reads the top 19 bits of each fp32 operand word (the low 13 mantissa bits are ignored), so the
            // raw tile already IS the hi operand: only lo = x - trunc(x) is written (one third less converter traffic
            // through the shared-memory port that paces the MMAs).  FAR_TC_DBG bit 128 restores the explicit hi store.
            if (p.dbg & 128) a4[ct + i * 128] = h;
            l4[ct + i * 128] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA reads
          __syncwarp();
          if (lane == 0) mbar_arrive(conv_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;   // TMEM lane window of this warp: lanes [32*quarter, 32*quarter + 32)
    const int chalf = ew >> 2;      // columns [64*chalf, 64*chalf + 64) of the 128-wide tile
    int acc = 0;
    uint32_t acc_phase = 0;
    const int actc = p.act_cols < 0 ? p.N : p.act_cols;
    const uint32_t stg_addr = stg_base + ew * STG_TILE;
    unsigned char* stg = smem_dyn + (stg_addr - raw);
    // swizzled 16-byte chunk slots of this thread's staging row (SWIZZLE_128B: chunk ^= row & 7)
    float4* srow = reinterpret_cast<float4*>(stg + lane * 128);
    const int sx = lane & 7;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tm = tile / tiles_n, n0 = (tile % tiles_n) * BN;
      const int g = tm / tiles_pg, r0 = (tm % tiles_pg) * BM;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const bool live = !(p.dbg & 2);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = chalf * 2 + cc;
        const int col0 = n0 + c * 32;
        uint32_t v[32], vs[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * BN + c * 32);
        tmem_ld32_nowait(taddr, v);
        tmem_ld32_nowait(taddr + (uint32_t)BN, vs);
        tmem_ld_wait();
        if (cc == 1) {  // both column blocks are in registers: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        if (!live || col0 >= p.N) continue;
        float t[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(v[e]) + __uint_as_float(vs[e]);
        epi_block(t, p, &mapC, col0, g, r0, quarter, lane, actc, srow, sx, stg_addr);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

__global__ void split_cat_kernel(const float* __restrict__ W, int ld, int K, long long rows, float* __restrict__ hi,
                                 uint32_t* __restrict__ cat) {
  const long long total = rows * K;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / K;
    const int c = (int)(idx % K), w = c & 31, k0 = (c & ~31) + 2 * (w & 15);
    const float* row = W + r * ld;
    hi[idx] = row[c];
    const float e0 = row[k0], e1 = row[k0 + 1];
    cat[idx] = (w < 16) ? pack_bf16x2(tf32_lo(e0), tf32_lo(e1)) : pack_bf16x2(e0, e1);
  }
}

static int g_cross16 = -1;
bool tc_cross16_on() {
  if (g_cross16 < 0) {
    // default: full 3xTF32 (tf32 cross terms).  Measured on B200 (profiles/r2_cross16_ab.md): the bf16 cross terms cut
    // the tensor cycles by a third but the kernel is paced by the shared-memory port, not the tensor pipe (245 vs 250 us
    // per launch), so the 4x looser rounding buys nothing yet.  FAR_TC_CROSS=bf16 (or 1) / far_tc_set_cross16(1) enables.
    const char* e = getenv("FAR_TC_CROSS");
    g_cross16 = (e && (e[0] == 'b' || e[0] == '1')) ? 1 : 0;
  }
  return g_cross16 == 1;
}
int tc_cross16_set(int on) {
  const int prev = tc_cross16_on() ? 1 : 0;
  g_cross16 = on ? 1 : 0;
  return prev;
}

// x = hi + lo with hi = tf32-truncated x.  Handles the [x1 | x2] concatenation: dst row stride = K1 + K2.
__global__ void split_tf32_kernel(const float* __restrict__ x1, int ld1, int K1, const float* __restrict__ x2, int ld2,
                                  int K2, long long rows, float* __restrict__ hi, float* __restrict__ lo) {
  const int K = K1 + K2;
  const long long total = rows * K;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / K;
    const int c = (int)(idx % K);
    const float v = (c < K1) ? x1[r * ld1 + c] : x2[r * ld2 + (c - K1)];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[idx] = h;
    lo[idx] = v - h;
  }
}

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried[64] = {};
  if (first_use_on_device(tried)) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 row-major [rows, K] tensor, box = 128 rows x 32 floats, SWIZZLE_128B, zero fill out of bounds.
bool make_map(CUtensorMap* map, const float* ptr, long long rows, int K) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_map4(CUtensorMap* map, const float* ptr, int K, long long rows, long long ld, int g0, long long s0, int g1,
               long long s1, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)g0, (cuuint64_t)g1};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 4, (cuuint64_t)s0 * 4, (cuuint64_t)s1 * 4};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}



}  // namespace tc

bool tc_engine_default_on() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("FAR_TC");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

bool tc_linear_supported(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                         int M, int N) {
  (void)x1; (void)ldx1; (void)x2; (void)ldx2; (void)W;
  const int K = K1 + K2;
  (void)M; (void)N;
  if (K % 4 != 0 || K < 32 || ldw < K) return false;          // TMA: 16-byte row pitch (the weight split densifies W)
  return tc::get_encode() != nullptr;
}

bool tc_linear_preferred(int M, int N, int K) {
  // small / skinny problems stay on the CUDA-core engine (split-K, HBM-bound on the weights)
  return (long long)M * N >= 128LL * 128 * 32 && K <= 4096;
}

size_t tc_linear_workspace_bytes(int M, int N, int K) {
  return 2 * tc::al((size_t)M * K * 4) + 2 * tc::al((size_t)N * K * 4) + 2048;
}

static bool raw_a_ok(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2) {
  // raw-A path: TMA reads the activations in place (contiguous rows, 16-byte pitch, whole 32-wide k-blocks per source)
  static const bool raw_off = getenv("FAR_TC_PRESPLIT") != nullptr;
  return !raw_off && ldx1 == K1 && K1 % tc::BK == 0 && (reinterpret_cast<uintptr_t>(x1) & 15u) == 0 &&
         (x2 == nullptr || (ldx2 == K2 && (reinterpret_cast<uintptr_t>(x2) & 15u) == 0));
}

// exact workspace need of tc_linear for these operands (the activation split buffers are only needed off the raw-A path)
size_t tc_linear_workspace_need(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, int M, int N) {
  const int K = K1 + K2;
  return (raw_a_ok(x1, ldx1, K1, x2, ldx2, K2) ? 0 : 2 * tc::al((size_t)M * K * 4)) + 2 * tc::al((size_t)N * K * 4) + 2048;
}

static int tc_ts_mode() {
  static const int m = getenv("FAR_TC_TS") ? atoi(getenv("FAR_TC_TS")) : 1;
  return m;
}
bool tc_ln_fusion_available() {
  static const bool off = getenv("FAR_LN_FUSION") && getenv("FAR_LN_FUSION")[0] == '0';
  return !off && tc_ts_mode() == 1 && !tc::tc_cross16_on();
}

int tc_linear_ex(const TcLinearEx& a, cudaStream_t st) {
  using namespace tc;
  const int K = a.K1 + a.K2, M = a.M, N = a.N;
  const int G = a.G > 0 ? a.G : 1, L = a.G > 0 ? a.L : M;
  if (G * L != M) return FAR_ERR_ARG;
  // the epilogue writes through a TMA store: 16-byte aligned rows
  if ((a.ldy & 3) != 0 || (reinterpret_cast<uintptr_t>(a.y) & 15u) != 0) return FAR_ERR_ARG;
  const bool presplitB = a.Whi != nullptr;
  const bool rawA = raw_a_ok(a.x1, a.ldx1, a.K1, a.x2, a.ldx2, a.K2);
  const size_t need = (rawA ? 0 : 2 * al((size_t)M * K * 4)) + (presplitB ? 0 : 2 * al((size_t)N * K * 4)) + 2048;
  if ((need > 2048) && (a.workspace == nullptr || a.workspace_bytes < need)) return FAR_ERR_WORKSPACE;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(a.workspace) + 1023) & ~uintptr_t(1023));
  float* xhi = reinterpret_cast<float*>(base);
  float* xlo = reinterpret_cast<float*>(base + al((size_t)M * K * 4));
  char* wbase = base + (rawA ? 0 : 2 * al((size_t)M * K * 4));
  const float* whi = a.Whi;
  const float* wlo = a.Wlo;
  const int sblocks = kNumSMs * 8;
  if (!rawA) {
    split_tf32_kernel<<<sblocks, 256, 0, st>>>(a.x1, a.ldx1, a.K1, a.x2, a.ldx2, a.K2, M, xhi, xlo);
    FAR_CHECK_LAUNCH();
  }
  // cross16 (tc_common.cuh): raw-A path with whole 32-float k-blocks; a pre-split B must then be in the (raw, cat) form
  const bool cross16 = rawA && K % BK == 0 && (presplitB ? a.wlo_is_cat != 0 : tc_cross16_on());
  if (presplitB && a.wlo_is_cat && !cross16) return FAR_ERR_ARG;
  if (!presplitB) {
    float* h = reinterpret_cast<float*>(wbase);
    float* l = reinterpret_cast<float*>(wbase + al((size_t)N * K * 4));
    const int N1 = a.W2 ? a.N1 : N;  // rows [0,N1) from W, rows [N1,N) from W2 (fused projections, e.g. [Wk; Wv])
    if (cross16) {
      split_cat_kernel<<<sblocks, 256, 0, st>>>(a.W, a.ldw, K, N1, h, reinterpret_cast<uint32_t*>(l));
      FAR_CHECK_LAUNCH();
      if (a.W2) {
        split_cat_kernel<<<sblocks, 256, 0, st>>>(a.W2, a.ldw, K, N - N1, h + (size_t)N1 * K,
                                                  reinterpret_cast<uint32_t*>(l) + (size_t)N1 * K);
        FAR_CHECK_LAUNCH();
      }
    } else {
      split_tf32_kernel<<<sblocks, 256, 0, st>>>(a.W, a.ldw, K, nullptr, 0, 0, N1, h, l);
      FAR_CHECK_LAUNCH();
      if (a.W2) {
        split_tf32_kernel<<<sblocks, 256, 0, st>>>(a.W2, a.ldw, K, nullptr, 0, 0, N - N1, h + (size_t)N1 * K, l + (size_t)N1 * K);
        FAR_CHECK_LAUNCH();
      }
    }
    whi = h; wlo = l;
  }
  const int Gb = a.b_grouped ? G : 1;
  CUtensorMap mAhi, mAlo, mBhi, mBlo, mC;
  bool ok = make_map4(&mBhi, whi, K, N, K, Gb, (long long)N * K, 1, (long long)Gb * N * K) &&
            make_map4(&mBlo, wlo, K, N, K, Gb, (long long)N * K, 1, (long long)Gb * N * K);
  if (rawA) {
    ok = ok && make_map4(&mAhi, a.x1, a.K1, L, a.K1, G, (long long)L * a.K1, 1, (long long)M * a.K1);
    if (a.x2) ok = ok && make_map4(&mAlo, a.x2, a.K2, L, a.K2, G, (long long)L * a.K2, 1, (long long)M * a.K2);
    else mAlo = mAhi;
  } else {
    ok = ok && make_map4(&mAhi, xhi, K, L, K, G, (long long)L * K, 1, (long long)M * K) &&
         make_map4(&mAlo, xlo, K, L, K, G, (long long)L * K, 1, (long long)M * K);
  }
  // output: dims (N, L, G, 1), 32 x 32 boxes, SWIZZLE_128B staging
  ok = ok && make_map4(&mC, a.y, N, L, a.ldy, G, (long long)L * a.ldy, 1, (long long)M * a.ldy, 32);
  if (!ok) return FAR_ERR_CUDA;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
  }
  static const int dbg = getenv("FAR_TC_DBG") ? atoi(getenv("FAR_TC_DBG")) : 0;
  GemmArgs p{};
  p.bias = a.bias; p.M = M; p.N = N; p.K = K; p.G = G; p.L = L; p.b_grouped = a.b_grouped ? 1 : 0;
  p.act = a.act; p.act_cols = a.act_cols; p.kb1 = a.K1 / BK;
  if (a.ksum != nullptr) {
    if (a.act != FAR_ACT_ELU1 || N % 32 != 0 || (a.act_cols >= 0 && a.act_cols != N)) return FAR_ERR_ARG;
    p.act = ACT_ELU1Z; p.ksum = a.ksum; p.ksum_rec = a.ksum_rec; p.ksum_off = a.ksum_off; p.heads = N / 32; p.eps = a.eps;
  }
  if (a.rowbias != nullptr) {
    if (a.rowbias_group <= 0 || N % 4 != 0 || (reinterpret_cast<uintptr_t>(a.rowbias) & 15u) != 0) return FAR_ERR_ARG;
    p.rowbias = a.rowbias; p.rb_group = a.rowbias_group;
  }
  p.dbg = dbg;
  p.cross16 = cross16 ? 1 : 0;
  const int tiles = G * ceil_div(L, BM) * ceil_div(N, BN);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  ProfScope prof(PROF_TC_GEMM, 2.0 * M * N * K, 4.0 * ((double)M * K + (double)Gb * N * K + (double)M * N), st);
  // raw fp32 activations: the A operand goes through TMEM.  FAR_TC_TS = 1 (default): one CTA per 128 x 128 tile
  // (tc_gemm_ts.cu); 2: CTA pairs sharing B (tc_gemm_pair.cu: 13 % faster than mode 1 at ~1.1 GHz SM clock, 15-30 % slower at
  // the 1.5-1.7 GHz of the power-capped step); 0: the all-shared-memory kernel below
  const int ts_mode = tc_ts_mode();
  if (a.ln_stats_in != nullptr &&
      !(rawA && !cross16 && ts_mode == 1 && a.x2 != nullptr && a.K2 % 32 == 0 && a.ln_gamma && a.ln_beta))
    return FAR_ERR_ARG;   // only the TMEM-operand kernel normalises its second A segment
  if (a.ln_stats_out != nullptr && N % 32 != 0) return FAR_ERR_ARG;
  p.ln_out = reinterpret_cast<float2*>(a.ln_stats_out);
  p.ln_in = reinterpret_cast<const float2*>(a.ln_stats_in);
  p.ln_gamma = a.ln_gamma; p.ln_beta = a.ln_beta; p.ln_eps = a.ln_eps; p.ln_chunks = a.K2 / 32;
  const int pair_clusters = (rawA && !cross16 && ts_mode >= 2) ? gemm_pair_max_clusters() : 0;
  if (pair_clusters > 0) {
    CUtensorMap mBhi2, mBlo2;   // 64-row boxes: each CTA of a pair loads its half of the 128 B rows
    if (!make_map4(&mBhi2, whi, K, N, K, Gb, (long long)N * K, 1, (long long)Gb * N * K, BN / 2) ||
        !make_map4(&mBlo2, wlo, K, N, K, Gb, (long long)N * K, 1, (long long)Gb * N * K, BN / 2))
      return FAR_ERR_CUDA;
    const int ptiles = G * ceil_div(L, 2 * BM) * ceil_div(N, BN);
    launch_gemm_pair(ptiles < pair_clusters ? ptiles : pair_clusters, st, mAhi, mAlo, mBhi2, mBlo2, mC, p);
  } else if (rawA && !cross16 && ts_mode >= 1)
    launch_gemm_ts(grid, st, mAhi, mAlo, mBhi, mBlo, mC, p);
  else if (rawA)
    tc_gemm_kernel<true><<<grid, GEMM_THREADS_RAW, GEMM_SMEM, st>>>(mAhi, mAlo, mBhi, mBlo, mC, p);
  else
    tc_gemm_kernel<false><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(mAhi, mAlo, mBhi, mBlo, mC, p);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

int tc_linear(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
              const float* bias, const float* rowbias, int rowbias_group, float* y, int ldy, int M, int N, int act,
              int act_cols, float* workspace, size_t workspace_bytes, cudaStream_t st, const float* presplit) {
  if (workspace == nullptr || workspace_bytes < tc_linear_workspace_need(x1, ldx1, K1, x2, ldx2, K2, M, N))
    return FAR_ERR_WORKSPACE;
  TcLinearEx a{};
  if (presplit != nullptr) {   // far_tc_weight_split output: tf32 hi at 0, lo at al(N*K*4)
    a.Whi = presplit;
    a.Wlo = reinterpret_cast<const float*>(reinterpret_cast<const char*>(presplit) + tc::al((size_t)N * (K1 + K2) * 4));
  }
  a.x1 = x1; a.ldx1 = ldx1; a.K1 = K1; a.x2 = x2; a.ldx2 = ldx2; a.K2 = K2;
  a.W = W; a.ldw = ldw; a.bias = bias; a.y = y; a.ldy = ldy; a.M = M; a.N = N; a.act = act; a.act_cols = act_cols;
  a.rowbias = rowbias; a.rowbias_group = rowbias_group;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  return tc_linear_ex(a, st);
}

}  // namespace far

extern "C" int far_tc_set_cross16(int on) { return far::tc::tc_cross16_set(on); }

// ---- cached weight operands: the static weights of a module are split once, not on every call (114 split launches per
// 32-pair FAR step otherwise) ----------------------------------------------------------------------------------------------
extern "C" size_t far_tc_weight_split_bytes(int N, int K) { return 2 * far::tc::al((size_t)N * K * 4) + 1024; }

extern "C" int far_tc_weight_split(const float* W, int ldw, int N, int K, float* out, size_t out_bytes, void* stream) {
  using namespace far;
  using namespace far::tc;
  FAR_REQUIRE(W && out && N > 0 && K > 0 && ldw >= K && (reinterpret_cast<uintptr_t>(out) & 1023u) == 0);
  if (out_bytes < far_tc_weight_split_bytes(N, K)) return FAR_ERR_WORKSPACE;
  float* hi = out;
  float* lo = reinterpret_cast<float*>(reinterpret_cast<char*>(out) + al((size_t)N * K * 4));
  split_tf32_kernel<<<kNumSMs * 8, 256, 0, (cudaStream_t)stream>>>(W, ldw, K, nullptr, 0, 0, N, hi, lo);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}
