// tcgen05.mma issue/execute rate on sm_100a for the shapes the far_b200 kernels use: kind::tf32, M = 128,
// N in {64, 96, 128, 256}, A from shared memory (SS) or from TMEM (TS).  One CTA per SM, one elected thread issues
// `iters` back-to-back MMAs on garbage operands (the tensor pipe does not care), commits, waits; cycles / MMA from
// clock64.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu ; run: ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = (uint64_t)((a & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* cycles) {
  extern __shared__ unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      const uint64_t da = desc_sw128(base), db = desc_sw128(base + 32768);
      const uint32_t id = idesc(128, N);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (TS) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                       ::"r"(tmem), "r"(tmem + 256), "l"(db), "r"(id), "r"(1u), "r"(0u) : "memory");
        } else {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                       ::"r"(tmem), "l"(da), "l"(db), "r"(id), "r"(1u), "r"(0u) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      } while (!done);
      t1 = clock64();
      if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int N, bool TS>
static void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 4096;
  const size_t smem = 1024 + 32768 + 65536;
  cudaFuncSetAttribute(mma_rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<N, TS><<<148, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("{\"mma\": \"%s\", \"M\": 128, \"N\": %d, \"K\": 8, \"cycles_per_mma\": %.1f, \"formula_M128xN/256\": %.1f, \"status\": \"%s\"}\n",
         name, N, (double)c / iters, 128.0 * N / 256.0, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, false>("tf32 SS");
  run<96, false>("tf32 SS");
  run<128, false>("tf32 SS");
  run<256, false>("tf32 SS");
  run<64, true>("tf32 TS (A in TMEM)");
  run<96, true>("tf32 TS (A in TMEM)");
  run<128, true>("tf32 TS (A in TMEM)");
  return 0;
}
