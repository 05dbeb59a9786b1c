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
__global__ void __launch_bounds__(256) scale_shift_act_nhwc_kernel(const float4* x, float4* y,
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
    y[idx] = v;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Backbone stem: conv1 (7x7, stride 2, padding 3, ONE input channel -> 128) + folded BatchNorm + ReLU
// (resnet_fpn.py:52-54,80: `x0 = relu(bn1(conv1(x)))`), written NHWC.  cuDNN has no tensor-core engine for C_in = 1:
// it ran a generic fp32 kernel (3.2 ms for 64 images) plus separate bias and ReLU passes over the 2.5 GB output.
// Weights in REGISTERS, inputs by broadcast: a lane owns one output channel (its 49 taps live in registers), a warp
// owns 32 channels of one output row and walks it 8 pixels at a time; per kernel row the 21 input values those 8
// pixels need are six warp-uniform LDS.128 (one wavefront each) feeding 56 FMAs per lane -- shared-memory traffic is
// ~0.1 wavefront per FMA instruction (a version with the weights in shared memory was LSU-bound at 0.4).  Each store
// instruction writes 128 contiguous bytes of an NHWC pixel.
constexpr int ST_TH = 2, ST_TW = 64, ST_C = 128, ST_K = 7, ST_PX = 8;
constexpr int ST_PH = 2 * ST_TH + 5, ST_PW = 2 * ST_TW + 5, ST_PP = 2 * ST_TW + 8 + 8;  // 9 x 133 patch, pitch 144

// wt: weights pre-transposed to [49][128] (tap-major: coalesced per-lane weight loads).
// grid (x tiles, y chunks, N); block 256 = 2 rows x 4 channel quarters; a CTA walks `tiles_per_cta` 2-row tiles down.
__global__ void __launch_bounds__(256, 2) stem_conv7x7s2_relu_kernel(const float* __restrict__ x,
                                                                      const float* __restrict__ wt,
                                                                      const float* __restrict__ bias,
                                                                      float* __restrict__ y, int H, int W, int OH,
                                                                      int OW, int tiles_per_cta) {
  __shared__ __align__(16) float patch[ST_PH][ST_PP];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int ch = (warp & 3) * 32 + lane, row = warp >> 2;
  const int n = blockIdx.z, ox0 = blockIdx.x * ST_TW;
  float wreg[ST_K * ST_K];
#pragma unroll
  for (int i = 0; i < ST_K * ST_K; ++i) wreg[i] = __ldg(wt + i * ST_C + ch);
  const float b = __ldg(bias + ch);
  const float* xin = x + (size_t)n * H * W;
  const int ix0 = 2 * ox0 - 3;
  for (int ty = 0; ty < tiles_per_cta; ++ty) {
    const int oy0 = (blockIdx.y * tiles_per_cta + ty) * ST_TH;
    if (oy0 >= OH) break;
    const int iy0 = 2 * oy0 - 3;
    __syncthreads();  // previous tile's reads of `patch` are done
    for (int idx = t; idx < ST_PH * ST_PP; idx += 256) {
      const int r = idx / ST_PP, c = idx % ST_PP;
      const int iy = iy0 + r, ix = ix0 + c;
      patch[r][c] = (c < ST_PW + 3 && iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xin + (size_t)iy * W + ix) : 0.f;
    }
    __syncthreads();
    const int oy = oy0 + row;
#pragma unroll 1
    for (int xc = 0; xc < ST_TW / ST_PX; ++xc) {
      float acc[ST_PX];
#pragma unroll
      for (int i = 0; i < ST_PX; ++i) acc[i] = b;
#pragma unroll
      for (int ky = 0; ky < ST_K; ++ky) {
        // inputs of 8 consecutive output pixels for this kernel row: patch columns [16 xc, 16 xc + 21)
        const float4* pr = reinterpret_cast<const float4*>(&patch[2 * row + ky][2 * ST_PX * xc]);
        float in[24];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const float4 v = pr[q];
          in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int kx = 0; kx < ST_K; ++kx)
#pragma unroll
          for (int i = 0; i < ST_PX; ++i) acc[i] = fmaf(in[2 * i + kx], wreg[ky * ST_K + kx], acc[i]);
      }
      if (oy < OH) {
#pragma unroll
        for (int i = 0; i < ST_PX; ++i) {
          const int ox = ox0 + xc * ST_PX + i;
          if (ox < OW) __stcs(y + (((size_t)n * OH + oy) * OW + ox) * ST_C + ch, fmaxf(acc[i], 0.f));
        }
      }
    }
  }
}

// w [Cout][49] -> wt [49][Cout]
__global__ void stem_weight_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < Cout * 49) wt[(idx % 49) * Cout + idx / 49] = w[idx];
}

// 8pt-ViT input preprocessing (interiornetStreetlearn_8ptVit/src/model.py:131-141): BGR -> RGB, / 255, ImageNet
// mean / std, F.interpolate(size = out, mode = 'nearest').  Nearest-neighbour resampling commutes with per-pixel
// arithmetic, so only the out x out sampled pixels are read (224^2 of 480 x 640: 6x fewer bytes than normalising the full
// image first) and the arithmetic on them is the reference's, in its order: ((x / 255) - mean) / std with IEEE
// divisions -> bit-identical to the eager sequence.  in [n, 3, H, W] (BGR, 0..255), out [n, 3, OH, OW] (RGB) NCHW.
__global__ void __launch_bounds__(256) vit_preprocess_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                             long long total, int H, int W, int OH, int OW,
                                                             float sy, float sx, float m0, float m1, float m2,
                                                             float s0, float s1, float s2) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % OW);
    const int oy = (int)((idx / OW) % OH);
    const int c = (int)((idx / ((long long)OW * OH)) % 3);
    const long long n = idx / ((long long)OW * OH * 3);
    // ATen upsample_nearest: src = min(floor(dst * scale), in - 1), scale = (float)in / out
    const int iy = min((int)floorf((float)oy * sy), H - 1), ix = min((int)floorf((float)ox * sx), W - 1);
    const float x = img[((n * 3 + (2 - c)) * H + iy) * (long long)W + ix];
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    out[idx] = __fdiv_rn(__fsub_rn(__fdiv_rn(x, 255.0f), mean), sd);
  }
}

}  // namespace far

using namespace far;

extern "C" int far_vit_preprocess(const float* images, float* out, long long n, int H, int W, int OH, int OW,
                                  const float* mean3, const float* std3, void* stream) {
  if (n <= 0) return FAR_OK;
  if (images == nullptr || out == nullptr || mean3 == nullptr || std3 == nullptr || H < 1 || W < 1 || OH < 1 || OW < 1)
    return FAR_ERR_ARG;
  const long long total = n * 3 * OH * OW;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  vit_preprocess_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      images, out, total, H, W, OH, OW, (float)H / (float)OH, (float)W / (float)OW, mean3[0], mean3[1], mean3[2],
      std3[0], std3[1], std3[2]);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}


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
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(scale),
      reinterpret_cast<const float4*>(shift), total, C / 4, negative_slope);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" int far_scale_shift_act_nhwc_out(const float* x, float* y, const float* scale, const float* shift,
                                            long long pixels, int C, float negative_slope, void* stream) {
  if (pixels <= 0 || C <= 0) return FAR_OK;
  if (x == nullptr || y == nullptr || shift == nullptr || (C & 3) != 0) return FAR_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(scale) |
       reinterpret_cast<uintptr_t>(shift)) & 15u)
    return FAR_ERR_ARG;
  const long long total = pixels * (C / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  scale_shift_act_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(scale),
      reinterpret_cast<const float4*>(shift), total, C / 4, negative_slope);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" size_t far_stem_conv_workspace_bytes(int Cout) { return (size_t)Cout * 49 * sizeof(float) + 256; }

extern "C" int far_stem_conv7x7s2_relu_nhwc(const float* x, const float* w, const float* bias, float* y, int N, int H,
                                            int W, int Cout, float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return FAR_OK;
  if (x == nullptr || w == nullptr || bias == nullptr || y == nullptr || H < 1 || W < 1 || Cout != ST_C) return FAR_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(workspace)) & 15u)
    return FAR_ERR_ARG;
  if (workspace == nullptr || workspace_bytes < far_stem_conv_workspace_bytes(Cout)) return FAR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  stem_weight_transpose_kernel<<<ceil_div(Cout * 49, 256), 256, 0, st>>>(w, workspace, Cout);
  FAR_CHECK_LAUNCH();
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;   // floor((H + 2*3 - 7) / 2) + 1
  const int ytiles = ceil_div(OH, ST_TH), xtiles = ceil_div(OW, ST_TW);
  // enough CTAs for ~8 waves of 2 CTAs/SM, each walking several tiles with the weights resident
  int tiles_per_cta = 1;
  while (tiles_per_cta < ytiles && (long long)xtiles * ceil_div(ytiles, tiles_per_cta * 2) * N >= 16LL * kNumSMs) tiles_per_cta *= 2;
  dim3 grid(xtiles, ceil_div(ytiles, tiles_per_cta), N);
  ProfScope prof(PROF_FPN, 2.0 * N * OH * OW * ST_C * 49, 4.0 * ((double)N * H * W + (double)N * OH * OW * ST_C), st);
  stem_conv7x7s2_relu_kernel<<<grid, 256, 0, st>>>(x, workspace, bias, y, H, W, OH, OW, tiles_per_cta);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}
