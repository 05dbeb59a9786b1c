// Shared pieces of the tcgen05 3xTF32 kernels (tc_gemm.cu, tc_score.cu): PTX wrappers (mbarrier, TMA, tcgen05),
// UMMA descriptors, tensor-map encoding.  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace far {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;  // BK floats = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 8;                   // tf32: 32 bytes of K per MMA
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;     // 16 KiB per operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES; // Ahi, Alo, Bhi, Blo
constexpr int TMEM_COLS = 4 * BN;           // two accumulator buffers x (main hi*hi | small cross terms) = 512 columns

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// One elected lane of a converged warp (CUTLASS' elect_one_sync): the predicate comes from ELECT, so the compiler keeps
// the guarded tcgen05 / TMA instructions and their descriptor arithmetic on the uniform datapath instead of wrapping
// every instruction issued under `if (lane == 0)` in an ELECT / BRA.U.ANY loop with R2UR moves (~10 SASS instructions
// per UTCHMMA: the single issuing thread could not keep 32-64-cycle MMAs back to back).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// warp index as a warp-uniform value (shuffle broadcast: the compiler's uniformity analysis accepts it)
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows x 128 B) >> 4
// in [32,46), version = 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format TF32 (2) [7,10)/[10,13), K-major A and B,
// N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
// Same with N = 2*BN: one instruction computes  [main | cross] += A_hi x [B_hi ; B_lo]  (the B_hi and B_lo tiles are
// adjacent in shared memory, the main and cross accumulators adjacent in TMEM).  With  cross += A_lo x B_hi  that is 2
// instructions and 20 KB of operand fetch per k-step instead of 3 and 24 KB -- the shared-memory port paces these MMAs.
constexpr uint32_t kIdescTf32N2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);


// ---- "cross16": the two 2^-11-times-smaller cross products  A_hi B_lo + A_lo B_hi  of the split-operand scheme only
// need ~8 significant bits each (their sum is a 2^-11 correction to A_hi B_hi), so they run as ONE bf16 MMA over a
// K-concatenated operand pair:  A_cat = [bf16(A) | bf16(A_lo)],  B_cat = [bf16(B_lo) | bf16(B)]  -- 64 bf16 = one 128-byte
// SWIZZLE_128B row per 32-float k-block, so tile shapes, TMA boxes and descriptors are unchanged.  Per k-block: 4 tf32
// MMAs (main) + 4 bf16 MMAs (cross, K = 16 each) = 512 tensor cycles and 64 KB of operand fetch, instead of 768 cycles
// and 80 KB with tf32 cross terms.  Relative accuracy 2^-19 per product (bf16 round-to-nearest of the correction terms)
// instead of 2^-21: held to the same 2e-5 + 1e-5 |ref| tolerance vs fp64 in tests/test_gpu_tcgen05.py.
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
// two floats -> packed bf16x2 (round to nearest even); e0 lands in the low half (lower address)
__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  return d;
}
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// whether the raw-A GEMM path uses bf16 cross terms (env FAR_TC_CROSS=tf32 restores full 3xTF32)
bool tc_cross16_on();
// B operand of the cross16 scheme from row-major weights: hi = W (raw copy; kind::tf32 ignores the low 13 bits), cat =
// per 32-float k-block [16 words: 32 x bf16(lo) | 16 words: 32 x bf16(W)].  K % 32 == 0.
__global__ void split_cat_kernel(const float* __restrict__ W, int ld, int K, long long rows, float* __restrict__ hi,
                                 uint32_t* __restrict__ cat);

// x = hi + lo with hi = tf32-truncated x.  Handles the [x1 | x2] concatenation: dst row stride = K1 + K2.
__global__ void split_tf32_kernel(const float* __restrict__ x1, int ld1, int K1, const float* __restrict__ x2, int ld2,
                                  int K2, long long rows, float* __restrict__ hi, float* __restrict__ lo);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();
// 2-D fp32 row-major [rows, K] tensor, box = 128 rows x 32 floats, SWIZZLE_128B, zero fill out of bounds.
bool make_map(CUtensorMap* map, const float* ptr, long long rows, int K);
// 4-D fp32 tensor: dims (K, rows, g0, g1) with byte strides (rows: ld*4, g0: s0*4, g1: s1*4); same box (32 x box_rows).
bool make_map4(CUtensorMap* map, const float* ptr, int K, long long rows, long long ld, int g0, long long s0, int g1,
               long long s1, int box_rows = BM);

static inline size_t al(size_t v) { return (v + 1023) & ~size_t(1023); }

}  // namespace tc
}  // namespace far
