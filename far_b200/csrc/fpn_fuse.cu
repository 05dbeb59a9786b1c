// FPN top-down glue of the ResNet-FPN backbone, fused for NHWC (channels_last) feature maps:
//   out = skip + bilinear_upsample_2x(low, align_corners=True)     (resnet_fpn.py:106-112: F.interpolate + add)
//   x   = leaky_relu(x * scale[c] + shift[c])                      (resnet_fpn.py:84-95: BatchNorm2d(eval) + LeakyReLU)
// Both are pure HBM streams (the 1/2-resolution maps of a 32-pair batch are 2.5-3.9 GB); torch runs them as
// upsample_bilinear2d_nhwc + add (3 passes, ~7.5 ms) and bn + leaky_relu (2 passes) -- here they are one pass each.
#include "common.cuh"

namespace far {

// One thread: one output pixel x 4 channels.  low:[N,Hin,Win,C], skip/out:[N,2Hin,2Win,C], C % 4 == 0.
__global__ void __launch_bounds__(256) upsample2x_add_nhwc_kernel(const float4* __restrict__ low,
                                                                   const float4* __restrict__ skip,
                                                                   float4* __restrict__ out, int N, int Hin, int Win,
                                                                   int C4, float sy_scale, float sx_scale) {
  const int Hout = 2 * Hin, Wout = 2 * Win;
  const long long total = (long long)N * Hout * Wout * C4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C4);
    long long pix = idx / C4;
    const int x = (int)(pix % Wout);
    pix /= Wout;
    const int y = (int)(pix % Hout);
    const int n = (int)(pix / Hout);
    // area_pixel_compute_source_index(align_corners=True): src = dst * (in - 1) / (out - 1)
    const float fy = sy_scale * (float)y, fx = sx_scale * (float)x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hin - 1), x1 = x0 + (x0 < Win - 1);
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0;
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float4* base = low + (size_t)n * Hin * Win * C4 + c;
    const float4 v00 = __ldg(base + ((size_t)y0 * Win + x0) * C4);
    const float4 v01 = __ldg(base + ((size_t)y0 * Win + x1) * C4);
    const float4 v10 = __ldg(base + ((size_t)y1 * Win + x0) * C4);
    const float4 v11 = __ldg(base + ((size_t)y1 * Win + x1) * C4);
    float4 r;
    r.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
    r.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
    r.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
    r.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
    if (skip != nullptr) {
      const float4 s = __ldcs(skip + idx);   // streamed once
      r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
    }
    __stcs(out + idx, r);
  }
}

// In place: x[p, c] = act(x[p, c] * scale[c] + shift[c]); act: ReLU or LeakyReLU(slope).  C % 4 == 0.
__global__ void __launch_bounds__(256) scale_shift_act_nhwc_kernel(float4* __restrict__ x,
                                                                    const float4* __restrict__ scale,
                                                                    const float4* __restrict__ shift, long long total,
                                                                    int C4, float slope) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C4);
    float4 v = x[idx];
    const float4 b = __ldg(shift + c);
    if (scale != nullptr) {
      const float4 a = __ldg(scale + c);
      v.x = fmaf(v.x, a.x, b.x); v.y = fmaf(v.y, a.y, b.y); v.z = fmaf(v.z, a.z, b.z); v.w = fmaf(v.w, a.w, b.w);
    } else {
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    v.x = v.x > 0.f ? v.x : v.x * slope;
    v.y = v.y > 0.f ? v.y : v.y * slope;
    v.z = v.z > 0.f ? v.z : v.z * slope;
    v.w = v.w > 0.f ? v.w : v.w * slope;
    x[idx] = v;
  }
}

}  // namespace far

using namespace far;

extern "C" int far_upsample2x_add_nhwc(const float* low, const float* skip, float* out, int N, int Hin, int Win, int C,
                                       void* stream) {
  if (N <= 0 || Hin <= 0 || Win <= 0 || C <= 0) return FAR_OK;
  if (low == nullptr || out == nullptr || (C & 3) != 0) return FAR_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(low) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(skip)) & 15u)
    return FAR_ERR_ARG;
  const int Hout = 2 * Hin, Wout = 2 * Win;
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const float sx = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  const long long total = (long long)N * Hout * Wout * (C / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  upsample2x_add_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(low), reinterpret_cast<const float4*>(skip), reinterpret_cast<float4*>(out), N,
      Hin, Win, C / 4, sy, sx);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" int far_scale_shift_act_nhwc(float* x, const float* scale, const float* shift, long long pixels, int C,
                                        float negative_slope, void* stream) {
  if (pixels <= 0 || C <= 0) return FAR_OK;
  if (x == nullptr || shift == nullptr || (C & 3) != 0) return FAR_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15u)
    return FAR_ERR_ARG;
  const long long total = pixels * (C / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  scale_shift_act_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(scale), reinterpret_cast<const float4*>(shift),
      total, C / 4, negative_slope);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}
