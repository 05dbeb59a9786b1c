// tc_flash.cu: flash-style softmax attention on tcgen05 (d = 32 / 64), shared with the map-free correlation volume.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace far {
bool tc_flash_attention_supported(int N, int d);
size_t tc_flash_attention_bytes(int G, int N, int d);
// qkv [B, N, 3, H, d] -> out [B, N, H*d] = softmax(scale q k^T) v, heads re-interleaved
int tc_flash_attention(const float* qkv, int B, int N, int H, int d, float scale, float* out, float* workspace,
                       size_t workspace_bytes, cudaStream_t st);
}  // namespace far
