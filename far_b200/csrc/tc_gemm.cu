// tcgen05 3xTF32 GEMM engine for sm_100a:  C[M,N] = act( A[M,K] * B[N,K]^T + bias ),  fp32 in / fp32 out.
//
// Why 3xTF32: every reference GEMM on the path is IEEE fp32 and "bit-exact match indices" needs ~fp32 accuracy;
// tcgen05 has no fp32 MMA kind.  Each operand is split  x = hi + lo  (hi = x with the low 13 mantissa bits cleared,
// lo = x - hi exactly) and the product is accumulated as  lo*hi + hi*lo + hi*hi  in the fp32 TMEM accumulator
// (error ~2^-21 relative per product; SURVEY.md 7 measured an identical match set with this scheme).
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0     TMA producer : cp.async.bulk.tensor (SWIZZLE_128B, 128-row x 32-float boxes) of Ahi,Alo,Bhi,Blo into
//                             a 3-stage shared-memory ring, mbarrier complete_tx
//   warp 1     MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M128 x N128 x K8), 12 per
//                             stage, accumulating into one of two 128-column TMEM accumulators; tcgen05.commit
//                             releases the smem stage / publishes the accumulator
//   warps 2-5  epilogue     : tcgen05.ld (32 lanes x 32 columns per warp), bias + activation, vectorised global store;
//                             overlaps the next tile's main loop through the second accumulator
// The hi/lo split of both operands is a separate elementwise kernel into the caller's workspace (no in-kernel
// conversion => the main loop is pure TMA -> UMMA).
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace far {
namespace tc {

struct GemmArgs {
  float* C; int ldc;
  const float* bias;
  int M, N, K;
  int act, act_cols;
  int kb1;  // kRawA: number of 32-wide k-blocks that come from x1 (= K1 / 32)
  int dbg;  // diagnostics (env FAR_TC_DBG): 1 = skip global stores, 2 = skip the whole epilogue body, 4 = skip MMAs
};

// kRawA = false: A arrives pre-split (mapAhi / mapAlo).
// kRawA = true : mapAhi / mapAlo are the ORIGINAL fp32 activation tensors x1 [M,K1] / x2 [M,K2] (k-blocks below
//                K1/32 come from x1, the rest from x2); four extra "converter" warps split each landed tile in shared
//                memory (hi in place, lo into the second slot; the swizzle is a permutation of 16-byte chunks, so an
//                elementwise pass at identical offsets preserves the UMMA layout) and publish it to the MMA warp through
//                a third barrier after fence.proxy.async.  No split pass over the activations in HBM, half the A bytes.
constexpr int GEMM_THREADS_RAW = NUM_THREADS + 128;

template <bool kRawA>
__global__ void __launch_bounds__(kRawA ? GEMM_THREADS_RAW : NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, GemmArgs p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  auto conv_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
  const int num_tiles = tiles_m * tiles_n;
  const int kblocks = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(conv_bar(s), 4); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sbase = base + stage * STAGE_BYTES;
          if (kRawA) {
            mbar_arrive_expect_tx(full_bar(stage), 3 * TILE_BYTES);
            if (kb < p.kb1) tma_load_2d(sbase + 0 * TILE_BYTES, &mapAhi, full_bar(stage), kb * BK, m0);
            else tma_load_2d(sbase + 0 * TILE_BYTES, &mapAlo, full_bar(stage), (kb - p.kb1) * BK, m0);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_2d(sbase + 0 * TILE_BYTES, &mapAhi, full_bar(stage), kb * BK, m0);
            tma_load_2d(sbase + 1 * TILE_BYTES, &mapAlo, full_bar(stage), kb * BK, m0);
          }
          tma_load_2d(sbase + 2 * TILE_BYTES, &mapBhi, full_bar(stage), kb * BK, n0);
          tma_load_2d(sbase + 3 * TILE_BYTES, &mapBlo, full_bar(stage), kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
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
          const uint32_t sbase = base + stage * STAGE_BYTES;
          const uint64_t dAhi = make_kmajor_sw128_desc(sbase + 0 * TILE_BYTES);
          const uint64_t dAlo = make_kmajor_sw128_desc(sbase + 1 * TILE_BYTES);
          const uint64_t dBhi = make_kmajor_sw128_desc(sbase + 2 * TILE_BYTES);
          const uint64_t dBlo = make_kmajor_sw128_desc(sbase + 3 * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < ((p.dbg & 4) ? 0 : BK / UMMA_K); ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);  // advance inside the 128-byte swizzle row
            umma_tf32(tmem_small, dAlo + koff, dBhi + koff, kIdescTf32, (kb | k) ? 1u : 0u);
            umma_tf32(tmem_small, dAhi + koff, dBlo + koff, kIdescTf32, 1u);
            umma_tf32(tmem_main, dAhi + koff, dBhi + koff, kIdescTf32, (kb | k) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // smem stage reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));      // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (kRawA && warp >= 6) {
    // ===================== converter warps (6..9): x -> (hi, lo) in shared memory =====================
    const int ct = threadIdx.x - 6 * 32;  // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        unsigned char* sa = smem_dyn + (base + stage * STAGE_BYTES - raw);
        float4* a4 = reinterpret_cast<float4*>(sa);
        float4* l4 = reinterpret_cast<float4*>(sa + TILE_BYTES);
#pragma unroll
        for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
          const float4 v = a4[ct + i * 128];
          float4 h;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          a4[ct + i * 128] = h;
          l4[ct + i * 128] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA reads
        __syncwarp();
        if (lane == 0) mbar_arrive(conv_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane window of this warp: lanes [32*quarter, 32*quarter + 32)
    int acc = 0;
    uint32_t acc_phase = 0;
    const int actc = p.act_cols < 0 ? p.N : p.act_cols;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
    float* stg = reinterpret_cast<float*>(smem_dyn + (bar_base + 256 - raw)) + (warp - 2) * (32 * 33);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < ((p.dbg & 2) ? 0 : BN / 32); ++c) {
        uint32_t v[32], vs[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * BN + c * 32);
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + (uint32_t)BN, vs);
        const int col0 = n0 + c * 32;
        // bias / activation on this thread's row segment, then a 32x32 transpose through shared memory so that every
        // global store instruction writes four full 128-byte row segments (row-per-thread 16-byte stores measured
        // ~20 us per tile: partial-sector writes).
        __syncwarp();
        // Branch-free fast path (whole chunk inside N, one activation for the whole chunk): per-element control flow
        // here cost ~20 us per tile with one epilogue warp per scheduler (measured with FAR_TC_DBG).
        const int act_here = (col0 + 32 <= actc) ? p.act : ((col0 >= actc) ? FAR_ACT_NONE : -1);
        if (col0 + 32 <= p.N && act_here >= 0) {
          float t[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(v[e]) + __uint_as_float(vs[e]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int e = 0; e < 32; ++e) t[e] += __ldg(p.bias + col0 + e);
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
#pragma unroll
              for (int e = 0; e < 32; ++e) t[e] = t[e] > 0.f ? t[e] + 1.f : expm1f(t[e]) + 1.f;
              break;
            case FAR_ACT_SIGMOID:
#pragma unroll
              for (int e = 0; e < 32; ++e) t[e] = 1.f / (1.f + expf(-t[e]));
              break;
            default:
              break;
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) stg[lane * 33 + e] = t[e];
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int col = col0 + e;
            float t = __uint_as_float(v[e]) + __uint_as_float(vs[e]);
            if (col < p.N) {
              if (p.bias) t += __ldg(p.bias + col);
              if (col < actc) t = apply_act(t, p.act);
            }
            stg[lane * 33 + e] = t;
          }
        }
        __syncwarp();
        const int c4 = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3);
          const int grow = m0 + quarter * 32 + r;
          const int col = col0 + c4 * 4;
          const float* src = stg + r * 33 + c4 * 4;
          if (grow < p.M && col < p.N && !(p.dbg & 1)) {
            float* dst = p.C + (size_t)grow * p.ldc + col;
            if (vec_ok && col + 3 < p.N) {
              *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[1], src[2], src[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (col + e < p.N) dst[e] = src[e];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
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
  static bool tried = false;
  if (!tried) {
    tried = true;
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
  if (K % 4 != 0 || K < 32 || ldw != K) return false;         // TMA: 16-byte row pitch; weights contiguous
  return tc::get_encode() != nullptr;
}

bool tc_linear_preferred(int M, int N, int K) {
  // small / skinny problems stay on the CUDA-core engine (split-K, HBM-bound on the weights)
  return (long long)M * N >= 128LL * 128 * 32 && K <= 4096;
}

size_t tc_linear_workspace_bytes(int M, int N, int K) {
  return 2 * tc::al((size_t)M * K * 4) + 2 * tc::al((size_t)N * K * 4) + 2048;
}

int tc_linear(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
              const float* bias, float* y, int ldy, int M, int N, int act, int act_cols, float* workspace,
              size_t workspace_bytes, cudaStream_t st) {
  using namespace tc;
  const int K = K1 + K2;
  if (workspace == nullptr || workspace_bytes < tc_linear_workspace_bytes(M, N, K)) return FAR_ERR_WORKSPACE;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  float* xhi = reinterpret_cast<float*>(base);
  float* xlo = reinterpret_cast<float*>(base + al((size_t)M * K * 4));
  float* whi = reinterpret_cast<float*>(base + 2 * al((size_t)M * K * 4));
  float* wlo = reinterpret_cast<float*>(base + 2 * al((size_t)M * K * 4) + al((size_t)N * K * 4));
  const int sblocks = kNumSMs * 8;
  // raw-A path: TMA reads the activations in place (contiguous rows, 16-byte pitch, whole 32-wide k-blocks per source)
  static const bool raw_off = getenv("FAR_TC_PRESPLIT") != nullptr;
  const bool rawA = !raw_off && ldx1 == K1 && K1 % BK == 0 && (reinterpret_cast<uintptr_t>(x1) & 15u) == 0 &&
                    (x2 == nullptr || (ldx2 == K2 && (reinterpret_cast<uintptr_t>(x2) & 15u) == 0));
  if (!rawA) {
    split_tf32_kernel<<<sblocks, 256, 0, st>>>(x1, ldx1, K1, x2, ldx2, K2, M, xhi, xlo);
    FAR_CHECK_LAUNCH();
  }
  split_tf32_kernel<<<sblocks, 256, 0, st>>>(W, ldw, K, nullptr, 0, 0, N, whi, wlo);
  FAR_CHECK_LAUNCH();
  CUtensorMap mAhi, mAlo, mBhi, mBlo;
  bool ok = make_map(&mBhi, whi, N, K) && make_map(&mBlo, wlo, N, K);
  if (rawA) ok = ok && make_map(&mAhi, x1, M, K1) && make_map(&mAlo, x2 ? x2 : x1, M, x2 ? K2 : K1);
  else ok = ok && make_map(&mAhi, xhi, M, K) && make_map(&mAlo, xlo, M, K);
  if (!ok) return FAR_ERR_CUDA;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    attr_set = true;
  }
  static const int dbg = getenv("FAR_TC_DBG") ? atoi(getenv("FAR_TC_DBG")) : 0;
  GemmArgs p{y, ldy, bias, M, N, K, act, act_cols, K1 / BK, dbg};
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  ProfScope prof(PROF_TC_GEMM, 2.0 * M * N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N), st);
  if (rawA)
    tc_gemm_kernel<true><<<grid, GEMM_THREADS_RAW, SMEM_BYTES, st>>>(mAhi, mAlo, mBhi, mBlo, p);
  else
    tc_gemm_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mAhi, mAlo, mBhi, mBlo, p);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

}  // namespace far
